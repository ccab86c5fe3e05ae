// UAI model text -> plain host structures, from a buffer holding the whole file.
//
// Replaces, for well-formed files, the token-at-a-time `ifstream >> string` + `stoi` / `stod`
// reader of the reference (code/io.cpp:14-100; Diabetes / Mildew / Barley are 1.2-1.7 MB of text):
// one read of the file, tokens scanned in place, numbers through std::from_chars (correctly
// rounded, so every value is bit-identical to what strtod -- hence stod -- returns).
//
// It is a FAST PATH, not a second dialect: anything it is not sure the reference reads the same
// way -- a token that is not a plain non-negative decimal integer where one is expected, a number
// from_chars does not consume to the last character, '+' signs, hex floats, overflow, results in
// the subnormal range (stod throws there), a truncated file, a scope id that names no variable,
// an unknown header -- makes parse_model() return false, and the caller falls back to the
// ifstream reader, which behaves exactly like the reference (messages and failures included).
// Header-only and free of device code: tests/harness/uai_parse_check.cpp pins it on the CPU.
#pragma once
#include <cfloat>
#include <charconv>
#include <climits>
#include <cmath>
#include <cstddef>
#include <fstream>
#include <string>
#include <vector>

namespace bn {
namespace uai {

struct Parsed {
    std::string type;                               // "BAYES" or "MARKOV"
    std::vector<unsigned> card;                     // per variable id
    std::vector<std::vector<unsigned>> scopes;      // per factor, file order
    std::vector<std::vector<double>> values;        // per factor, file order (last scope variable fastest)
    std::vector<double> partition;                  // per factor: sum of its values in file order (code/io.cpp:93-96)
};

class Tokens {
public:
    Tokens(const char *data, size_t n) : p_(data), end_(data + n) {}

    // next whitespace-separated token; a token that starts with '#' comments out the rest of its line
    bool next(const char *&b, const char *&e)
    {
        for (;;) {
            while (p_ < end_ && is_space(*p_)) ++p_;
            if (p_ == end_) return false;
            b = p_;
            while (p_ < end_ && !is_space(*p_)) ++p_;
            e = p_;
            if (*b != '#') return true;
            while (p_ < end_ && *p_ != '\n') ++p_;      // getline: up to and including the newline
            if (p_ < end_) ++p_;
        }
    }

    bool next_unsigned(unsigned &out)
    {
        const char *b, *e;
        if (!next(b, e)) return false;
        unsigned long long v = 0;
        for (const char *c = b; c < e; ++c) {
            if (*c < '0' || *c > '9') return false;
            v = v * 10 + (unsigned)(*c - '0');
            if (v > (unsigned long long)INT_MAX) return false;      // stoi would throw
        }
        out = (unsigned)v;
        return true;
    }

    bool next_double(double &out)
    {
        const char *b, *e;
        if (!next(b, e)) return false;
        if (*b == '+') return false;
        const std::from_chars_result r = std::from_chars(b, e, out, std::chars_format::general);
        if (r.ec != std::errc() || r.ptr != e) return false;
        if (!std::isfinite(out)) return false;
        if (out == 0.0 || std::fabs(out) < DBL_MIN) {
            // an exact zero is fine; a non-zero literal that lands at or below the subnormal range is where
            // strtod reports ERANGE and stod throws
            for (const char *c = b; c < e && *c != 'e' && *c != 'E'; ++c)
                if (*c >= '1' && *c <= '9') return false;
        }
        return true;
    }

private:
    static bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; }
    const char *p_, *end_;
};

inline bool parse_model(const char *data, size_t n, Parsed &out)
{
    Tokens t(data, n);
    const char *b, *e;
    if (!t.next(b, e)) return false;
    out.type.assign(b, e);
    if (out.type != "BAYES" && out.type != "MARKOV") return false;
    unsigned nvars = 0;
    if (!t.next_unsigned(nvars)) return false;
    out.card.resize(nvars);
    for (unsigned i = 0; i < nvars; ++i)
        if (!t.next_unsigned(out.card[i])) return false;
    unsigned nfac = 0;
    if (!t.next_unsigned(nfac)) return false;
    out.scopes.resize(nfac);
    for (unsigned f = 0; f < nfac; ++f) {
        unsigned w = 0;
        if (!t.next_unsigned(w)) return false;
        out.scopes[f].resize(w);
        for (unsigned j = 0; j < w; ++j) {
            if (!t.next_unsigned(out.scopes[f][j])) return false;
            if (out.scopes[f][j] >= nvars) return false;
        }
    }
    out.values.resize(nfac);
    out.partition.assign(nfac, 0.0);
    for (unsigned f = 0; f < nfac; ++f) {
        unsigned size = 0;
        if (!t.next_unsigned(size)) return false;
        std::vector<double> &v = out.values[f];
        v.resize(size);
        double z = 0;
        for (unsigned j = 0; j < size; ++j) {
            if (!t.next_double(v[j])) return false;
            z += v[j];
        }
        out.partition[f] = z;
    }
    return true;
}

inline bool slurp(const std::string &filename, std::string &buf)
{
    std::ifstream in(filename, std::ios::binary);
    if (!in.is_open()) return false;
    in.seekg(0, std::ios::end);
    const std::streamoff n = in.tellg();
    if (n < 0) return false;
    in.seekg(0, std::ios::beg);
    buf.resize((size_t)n);
    if (n) in.read(&buf[0], n);
    return (std::streamoff)in.gcount() == n;
}

}  // namespace uai
}  // namespace bn
