// Test infrastructure: pins bnpp_b200/host/uai_parse.hpp (the buffer-based UAI reader) against a
// literal restatement of the reference's token-at-a-time reader (code/io.cpp:14-100: `ifstream >>
// string`, a '#' token comments out the rest of its line, stoi / stod).  No device code.
//   uai_parse_check file...   prints per file: FAST <ms> SLOW <ms> SAME | FALLBACK | DIFF
// exit code 1 on any DIFF.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "uai_parse.hpp"

namespace {

bool next_token(std::ifstream &in, std::string &tok)
{
    while (in >> tok) {
        if (tok[0] != '#') return true;
        std::getline(in, tok);
    }
    return false;
}

// -> false when the reference would have thrown or run off the end of the file
bool slow_parse(const std::string &path, bn::uai::Parsed &out)
{
    std::ifstream in(path);
    if (!in.is_open()) return false;
    std::string tok;
    try {
        if (!next_token(in, tok)) return false;
        out.type = tok;
        if (!next_token(in, tok)) return false;
        const unsigned n = (unsigned)std::stoi(tok);
        out.card.resize(n);
        for (unsigned i = 0; i < n; ++i) {
            if (!next_token(in, tok)) return false;
            out.card[i] = (unsigned)std::stoi(tok);
        }
        if (!next_token(in, tok)) return false;
        const unsigned m = (unsigned)std::stoi(tok);
        out.scopes.resize(m);
        for (unsigned f = 0; f < m; ++f) {
            if (!next_token(in, tok)) return false;
            const unsigned w = (unsigned)std::stoi(tok);
            for (unsigned j = 0; j < w; ++j) {
                if (!next_token(in, tok)) return false;
                out.scopes[f].push_back((unsigned)std::stoi(tok));
            }
        }
        out.values.resize(m);
        out.partition.assign(m, 0.0);
        for (unsigned f = 0; f < m; ++f) {
            if (!next_token(in, tok)) return false;
            const unsigned size = (unsigned)std::stoi(tok);
            double z = 0;
            for (unsigned j = 0; j < size; ++j) {
                if (!next_token(in, tok)) return false;
                out.values[f].push_back(std::stod(tok));
                z += out.values[f].back();
            }
            out.partition[f] = z;
        }
    } catch (...) {
        return false;
    }
    return true;
}

bool same(const bn::uai::Parsed &a, const bn::uai::Parsed &b)
{
    if (a.type != b.type || a.card != b.card || a.scopes != b.scopes || a.values.size() != b.values.size()) return false;
    for (size_t f = 0; f < a.values.size(); ++f) {
        if (a.values[f].size() != b.values[f].size()) return false;
        if (!a.values[f].empty() && memcmp(a.values[f].data(), b.values[f].data(), 8 * a.values[f].size()) != 0) return false;
        if (memcmp(&a.partition[f], &b.partition[f], 8) != 0) return false;
    }
    return true;
}

double ms_since(std::chrono::steady_clock::time_point t0)
{
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}

}  // namespace

int main(int argc, char **argv)
{
    int bad = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string path = argv[i];
        bn::uai::Parsed fast, slow;
        auto t0 = std::chrono::steady_clock::now();
        std::string buf;
        const bool okf = bn::uai::slurp(path, buf) && bn::uai::parse_model(buf.data(), buf.size(), fast);
        const double tf = ms_since(t0);
        t0 = std::chrono::steady_clock::now();
        const bool oks = slow_parse(path, slow);
        const double ts = ms_since(t0);
        const char *verdict;
        if (!okf) verdict = "FALLBACK";                       // the product re-reads with the slow reader: nothing to compare
        else if (oks && same(fast, slow)) verdict = "SAME";
        else {
            verdict = "DIFF";
            ++bad;
        }
        printf("%s FAST %.3f SLOW %.3f %s %s\n", path.c_str(), tf, ts, verdict, oks ? "slow-ok" : "slow-failed");
    }
    return bad ? 1 : 0;
}
