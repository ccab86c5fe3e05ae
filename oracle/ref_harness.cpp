// TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.
//
// Scriptable driver that LINKS THE UNMODIFIED REFERENCE OBJECTS (compiled in
// place from /root/reference/code by oracle/Makefile into oracle/_ref/) and
// prints 17-significant-digit results.  It is the 1e-9 oracle for the GPU path:
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may execute it.
//
// The reference CLIs print 6-7 digits only (code/bn.cpp:233, code/factor.cpp:316),
// so results are taken through the library API instead:
//   bn::BN::partition / marginals / query_ve   code/model.cpp:204-346
//   bn::Graph::ordering / order_width          code/graph.cpp:41-237
//   bn::FactorGraph::update / marginal         code/graph.cpp:298-403
//   bn::Factor::{product,divide,sum_out,conditioning,normalize}  code/factor.cpp:117-255
//
// Commands are read one per line from stdin (or the file given as argv[1]):
//   model <path>                 load BAYES or MARKOV uai file as a bn::BN
//   evid <path>                  load evidence file (code/io.cpp:157-180)
//   evidset k id val ...         set evidence inline;  "evidset 0" clears
//   opt [mf|wmf|md|bb|ve|sp|v]*  set option flags (cleared first)
//   pr                           -> PR <Z> <uptime_ms>
//   mar                          -> MAR <n> / M <id> <size> <v...>
//   order                        -> ORDER <width> <n> <ids...>   (non-evidence vars, conditioned factors)
//   widths                       -> WIDTHS orig md mf wmf  (as the `width` prompt command, code/bn.cpp:418-481)
//   bp <max> <eps> [cond]        -> BP <sweeps> / M ...  (FactorGraph on raw or conditioned factors)
//   queryve t1,t2 [| e1,e2]      -> FACTOR dump of the query result
//   vars n c0 c1 ...             define free-standing variables for op tests
//   factor <name> w id.. v..     define a factor (values in row-major, last var fastest)
//   randfactor <name> seed w id..  U(0.1,1) values from mt19937_64(seed)
//   product a b out | divide a b out | sumout a var out | cond a out k id val.. | normalize a out
//   max a | min a                -> SCALAR v
//   dump a                       -> FACTOR <w> <ids..> <size> <partition> / V <values...>
//   digest a                     -> DIGEST <size> <partition> <sum of v[i]*(1+(i%7))>
//   time <cmd...>                run cmd, then print TIME <ms>
//   drop a                       free a factor
#include "io.hh"
#include "utils.hh"
#include "model.hh"
#include "graph.hh"

#include <chrono>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <map>
#include <random>
#include <sstream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace bn {
// external-linkage helpers of code/io.cpp:43-100 (not declared in io.hh)
std::string read_file_header(std::ifstream &input_file);
void read_variables(std::ifstream &input_file, std::vector<Variable*> &variables);
void read_factors(std::ifstream &input_file, std::vector<Variable*> &variables, std::vector<Factor*> &factors);
}

using namespace bn;
using namespace std;

static BN *g_model = nullptr;
static unordered_map<unsigned,unsigned> g_evid;
static unordered_map<string,bool> g_opt;
static vector<Variable*> g_vars;            // free-standing variables for op tests
static map<string, Factor*> g_fac;

static void print_factor(const Factor &f)
{
    printf("FACTOR %u", f.width());
    for (unsigned i = 0; i < f.width(); ++i) printf(" %u", f.domain()[i]->id());
    printf(" %u %.17g\nV", f.size(), f.partition());
    for (unsigned i = 0; i < f.size(); ++i) printf(" %.17g", f[i]);
    printf("\n");
}

static void print_marginals(const vector<const Factor*> &marg)
{
    printf("MAR %zu\n", marg.size());
    for (size_t i = 0; i < marg.size(); ++i) {
        const Factor *pf = marg[i];
        printf("M %zu %u", i, pf->size());
        for (unsigned j = 0; j < pf->size(); ++j) printf(" %.17g", (*pf)[j]);
        printf("\n");
    }
}

static void load_model(string path)
{
    ifstream in(path);
    if (!in.is_open()) { printf("ERROR cannot open %s\n", path.c_str()); return; }
    vector<Variable*> variables;
    vector<Factor*> factors;
    string type = read_file_header(in);
    read_variables(in, variables);
    read_factors(in, variables, factors);
    delete g_model;
    g_model = new BN(path, variables, factors);
    g_evid.clear();
    printf("MODEL %s %zu %zu\n", type.c_str(), variables.size(), factors.size());
}

static void set_opts(istringstream &ss)
{
    g_opt.clear();
    string t;
    while (ss >> t) {
        if (t == "mf") g_opt["min-fill"] = true;
        else if (t == "wmf") g_opt["weighted-min-fill"] = true;
        else if (t == "md") g_opt["min-degree"] = true;
        else if (t == "bb") g_opt["bayes-ball"] = true;
        else if (t == "ve") g_opt["variable-elimination"] = true;
        else if (t == "sp") g_opt["sum-product"] = true;
        else if (t == "v") g_opt["verbose"] = true;
    }
}

static vector<const Factor*> conditioned_factors()
{
    vector<const Factor*> out;
    for (auto pf : g_model->factors()) out.push_back(new Factor(pf->conditioning(g_evid)));
    return out;
}

static const Variable *op_var(unsigned id)
{
    return g_vars.at(id);
}

static bool run(const string &line);

static bool run(const string &line)
{
    istringstream ss(line);
    string cmd;
    if (!(ss >> cmd) || cmd[0] == '#') return true;

    if (cmd == "quit") return false;
    else if (cmd == "time") {
        string rest;
        getline(ss, rest);
        auto t0 = chrono::steady_clock::now();
        run(rest);
        auto t1 = chrono::steady_clock::now();
        printf("TIME %.6f\n", chrono::duration<double, milli>(t1 - t0).count());
    }
    else if (cmd == "model") { string p; ss >> p; load_model(p); }
    else if (cmd == "evid") {
        string p; ss >> p;
        g_evid.clear();
        int rc = read_uai_evidence(p, g_evid);
        printf("EVID %d %zu\n", rc, g_evid.size());
    }
    else if (cmd == "evidset") {
        unsigned k; ss >> k;
        g_evid.clear();
        for (unsigned i = 0; i < k; ++i) { unsigned id, val; ss >> id >> val; g_evid[id] = val; }
        printf("EVID 0 %zu\n", g_evid.size());
    }
    else if (cmd == "opt") { set_opts(ss); printf("OPT %zu\n", g_opt.size()); }
    else if (cmd == "pr") {
        double uptime = 0;
        double z = g_model->partition(g_evid, g_opt, uptime);
        printf("PR %.17g %.6f\n", z, uptime);
    }
    else if (cmd == "mar") {
        double uptime = 0;
        vector<const Factor*> marg = g_model->marginals(g_evid, g_opt, uptime);
        print_marginals(marg);
        printf("UPTIME %.6f\n", uptime);
        for (auto pf : marg) delete pf;
    }
    else if (cmd == "order") {
        // same inputs as BN::partition hands to variable_elimination (code/model.cpp:277-287,360-365)
        vector<const Variable*> variables;
        for (auto pv : g_model->variables())
            if (g_evid.find(pv->id()) == g_evid.end()) variables.push_back(pv);
        vector<const Factor*> factors = conditioned_factors();
        vector<const Variable*> model_variables(g_model->variables().begin(), g_model->variables().end());
        Graph g(model_variables, factors);
        unsigned width = 0;
        auto t0 = chrono::steady_clock::now();
        vector<unsigned> ids = g.ordering(variables, width, g_opt);
        auto t1 = chrono::steady_clock::now();
        printf("ORDER %u %zu", width, ids.size());
        for (auto id : ids) printf(" %u", id);
        printf("\nORDER_MS %.6f\n", chrono::duration<double, milli>(t1 - t0).count());
        for (auto pf : factors) delete pf;
    }
    else if (cmd == "widths") {
        vector<const Variable*> vars(g_model->variables().begin(), g_model->variables().end());
        vector<const Factor*> factors(g_model->factors().begin(), g_model->factors().end());
        Graph g(vars, factors);
        unsigned w0 = g.order_width(vars), wmd, wmf, wwmf;
        unordered_map<string,bool> o;
        o["min-degree"] = true; g.ordering(vars, wmd, o);
        o.clear(); o["min-fill"] = true; g.ordering(vars, wmf, o);
        o.clear(); o["weighted-min-fill"] = true; g.ordering(vars, wwmf, o);
        printf("WIDTHS %u %u %u %u\n", w0, wmd, wmf, wwmf);
    }
    else if (cmd == "bp") {
        unsigned maxit; double eps; string mode;
        ss >> maxit >> eps >> mode;
        vector<const Variable*> variables(g_model->variables().begin(), g_model->variables().end());
        vector<const Factor*> factors;
        bool cond = (mode == "cond");
        if (cond) factors = conditioned_factors();
        else factors.assign(g_model->factors().begin(), g_model->factors().end());
        auto t0 = chrono::steady_clock::now();
        vector<const Factor*> marg;
        unsigned sweeps;
        {
            FactorGraph fg(variables, factors);
            sweeps = fg.update(maxit, eps);
            for (auto pv : variables) {
                // an observed variable has no edges once factors are conditioned
                if (cond && g_evid.count(pv->id())) marg.push_back(new Factor(1.0));
                else marg.push_back(new Factor(fg.marginal(pv)));
            }
        }
        auto t1 = chrono::steady_clock::now();
        printf("BP %u\n", sweeps);
        print_marginals(marg);
        printf("UPTIME %.6f\n", chrono::duration<double, milli>(t1 - t0).count());
        for (auto pf : marg) delete pf;
        if (cond) for (auto pf : factors) delete pf;
    }
    else if (cmd == "queryve") {
        string rest; getline(ss, rest);
        string t, e;
        size_t bar = rest.find('|');
        t = rest.substr(0, bar);
        if (bar != string::npos) e = rest.substr(bar + 1);
        auto strip = [](string s) { string o; for (char c : s) if (!isspace((unsigned char)c)) o += c; return o; };
        t = strip(t); e = strip(e);
        unordered_set<const Variable*> tv, ev;
        parse_vars_set(g_model, t, tv);
        if (!e.empty()) parse_vars_set(g_model, e, ev);
        double uptime = 0;
        Factor q = g_model->query_ve(tv, ev, g_opt, uptime);
        print_factor(q);
        printf("UPTIME %.6f\n", uptime);
    }
    else if (cmd == "vars") {
        unsigned n; ss >> n;
        for (auto pv : g_vars) delete pv;
        g_vars.clear();
        for (unsigned i = 0; i < n; ++i) { unsigned c; ss >> c; g_vars.push_back(new Variable(i, c)); }
        printf("VARS %u\n", n);
    }
    else if (cmd == "factor" || cmd == "randfactor") {
        string name; ss >> name;
        unsigned long long seed = 0;
        if (cmd == "randfactor") ss >> seed;
        unsigned w; ss >> w;
        vector<const Variable*> scope;
        for (unsigned i = 0; i < w; ++i) { unsigned id; ss >> id; scope.push_back(op_var(id)); }
        Domain *d = new Domain(scope);
        vector<double> values;
        values.reserve(d->size());
        double z = 0;
        if (cmd == "factor") {
            for (unsigned i = 0; i < d->size(); ++i) { double v; ss >> v; values.push_back(v); z += v; }
        } else {
            mt19937_64 gen(seed);
            for (unsigned i = 0; i < d->size(); ++i) {
                double v = 0.1 + 0.9 * ((gen() >> 11) * (1.0 / 9007199254740992.0));
                values.push_back(v); z += v;
            }
        }
        delete g_fac[name];
        g_fac[name] = new Factor(d, values, z);
        printf("OK %s %u\n", name.c_str(), d->size());
    }
    else if (cmd == "product" || cmd == "divide") {
        string a, b, o; ss >> a >> b >> o;
        Factor *r = new Factor(cmd == "product" ? g_fac.at(a)->product(*g_fac.at(b)) : g_fac.at(a)->divide(*g_fac.at(b)));
        delete g_fac[o]; g_fac[o] = r;
        printf("OK %s %u\n", o.c_str(), r->size());
    }
    else if (cmd == "sumout") {
        string a, o; unsigned v; ss >> a >> v >> o;
        Factor *r = new Factor(g_fac.at(a)->sum_out(op_var(v)));
        delete g_fac[o]; g_fac[o] = r;
        printf("OK %s %u\n", o.c_str(), r->size());
    }
    else if (cmd == "cond") {
        string a, o; unsigned k; ss >> a >> o >> k;
        unordered_map<unsigned,unsigned> ev;
        for (unsigned i = 0; i < k; ++i) { unsigned id, val; ss >> id >> val; ev[id] = val; }
        Factor *r = new Factor(g_fac.at(a)->conditioning(ev));
        delete g_fac[o]; g_fac[o] = r;
        printf("OK %s %u\n", o.c_str(), r->size());
    }
    else if (cmd == "normalize") {
        string a, o; ss >> a >> o;
        Factor *r = new Factor(g_fac.at(a)->normalize());
        delete g_fac[o]; g_fac[o] = r;
        printf("OK %s %u\n", o.c_str(), r->size());
    }
    else if (cmd == "max") { string a; ss >> a; printf("SCALAR %.17g\n", g_fac.at(a)->max()); }
    else if (cmd == "min") { string a; ss >> a; printf("SCALAR %.17g\n", g_fac.at(a)->min()); }
    else if (cmd == "dump") { string a; ss >> a; print_factor(*g_fac.at(a)); }
    else if (cmd == "digest") {
        string a; ss >> a;
        const Factor &f = *g_fac.at(a);
        double s = 0;
        for (unsigned i = 0; i < f.size(); ++i) s += f[i] * (1 + (i % 7));
        printf("DIGEST %u %.17g %.17g\n", f.size(), f.partition(), s);
    }
    else if (cmd == "drop") { string a; ss >> a; delete g_fac[a]; g_fac.erase(a); printf("OK\n"); }
    else printf("ERROR unknown command %s\n", cmd.c_str());
    fflush(stdout);
    return true;
}

int main(int argc, char *argv[])
{
    istream *in = &cin;
    ifstream file;
    if (argc > 1) { file.open(argv[1]); in = &file; }
    string line;
    while (getline(*in, line)) {
        if (!run(line)) break;
    }
    return 0;
}
