// TEST INFRASTRUCTURE: the UAI solution / evidence writers of bnpp_b200/host/io.cpp without a device.  Writes a PR
// file, an evidence file and reads the evidence back through read_uai_evidence; prints what it wrote.
// (write_uai_mar needs bn::Factor values, i.e. a device: it is covered by tests/test_gpu_host.py.)
#include <cstdio>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <unordered_map>

namespace bn {
int write_uai_pr(const std::string &filename, double partition);
int write_uai_evidence(const std::string &filename, const std::unordered_map<unsigned,unsigned> &evidence);
int read_uai_evidence(std::string &filename, std::unordered_map<unsigned,unsigned> &evidence);
}

static std::string slurp(const std::string &p)
{
    std::ifstream in(p);
    std::stringstream ss;
    ss << in.rdbuf();
    return ss.str();
}

int main(int argc, char **argv)
{
    const std::string dir = argc > 1 ? argv[1] : "/tmp";
    std::string pr = dir + "/w.PR", ev = dir + "/w.evid";
    if (bn::write_uai_pr(pr, 776008564204256.38)) return 1;
    std::unordered_map<unsigned,unsigned> e{{4, 1}, {0, 1}, {5, 1}}, back;
    if (bn::write_uai_evidence(ev, e)) return 2;
    if (bn::read_uai_evidence(ev, back)) return 3;
    std::cout << slurp(pr) << "--\n" << slurp(ev) << "--\n" << (back == e ? "roundtrip ok" : "roundtrip MISMATCH") << "\n";
    return bn::write_uai_pr("/nonexistent-dir/x.PR", 1.0) == -1 ? 0 : 4;
}
