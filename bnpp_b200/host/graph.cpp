#include "graph.hh"
#include "runtime.hh"
#include "../csrc/elim_order.hpp"

#include <iostream>

namespace bn {

// ---- Graph: thin wrapper over bnpp::InteractionGraph (csrc/elim_order.hpp) -------------------

static bnpp::InteractionGraph *build(const std::vector<const Variable*> &variables, const std::vector<const Factor*> &factors)
{
    std::vector<std::vector<unsigned>> scopes;
    scopes.reserve(factors.size());
    for (const Factor *pf : factors) {
        const std::vector<uint32_t> &ids = pf->domain().ids();
        scopes.emplace_back(ids.begin(), ids.end());
    }
    std::vector<unsigned> card;
    card.reserve(variables.size());
    for (const Variable *pv : variables) card.push_back(pv->size());   // indexed by position == id, as `_variables.at(id)`
    return new bnpp::InteractionGraph(scopes, card);
}

Graph::Graph(const std::vector<const Variable*> &variables, const std::vector<const Factor*> &factors)
    : _variables(variables), _g(build(variables, factors))
{
}

Graph::Graph(const Graph &g) : _variables(g._variables), _g(new bnpp::InteractionGraph(*g._g)) {}

Graph::~Graph() { delete _g; }

std::unordered_set<unsigned> Graph::neighbors(unsigned id) const { return _g->neighbors(id); }

bool Graph::connected(unsigned id1, unsigned id2) const { return _g->connected(id1, id2); }

std::vector<unsigned> Graph::ordering(const std::vector<const Variable*> &variables, unsigned &width,
                                      std::unordered_map<std::string,bool> &options) const
{
    std::vector<unsigned> ids;
    ids.reserve(variables.size());
    for (const Variable *pv : variables) ids.push_back(pv->id());
    // precedence of the flags as in code/graph.cpp:62-70 (operator[] inserts missing keys, as there)
    bnpp::Heuristic h = bnpp::H_MIN_FILL;
    if (options["min-degree"]) h = bnpp::H_MIN_DEGREE;
    else if (options["weighted-min-fill"]) h = bnpp::H_WEIGHTED_MIN_FILL;
    return _g->ordering(ids, h, width);
}

unsigned Graph::min_fill(const std::unordered_set<unsigned> &vars) const { return _g->pick_min_fill(vars); }
unsigned Graph::weighted_min_fill(const std::unordered_set<unsigned> &vars) const { return _g->pick_weighted_min_fill(vars); }
unsigned Graph::min_degree(const std::unordered_set<unsigned> &vars) const { return _g->pick_min_degree(vars); }

unsigned Graph::order_width(const std::vector<const Variable*> &variables) const
{
    std::vector<unsigned> ids;
    for (const Variable *pv : variables) ids.push_back(pv->id());
    return _g->order_width(ids);
}

std::ostream &operator<<(std::ostream &os, const Graph &g)
{
    os << "Graph:" << std::endl;
    for (const auto &node : g._g->adjacency()) {
        os << node.first << " :";
        for (unsigned nb : node.second) os << " " << nb;
        os << std::endl;
    }
    return os << std::endl;
}

// ---- FactorGraph: all messages of a phase in one launch (csrc/sumproduct.cu) -------------------

FactorGraph::FactorGraph(const std::vector<const Variable*> &variables, const std::vector<const Factor*> &factors)
    : _variables(variables), _factors(factors), _fg(nullptr), _marg_valid(false)
{
    std::vector<uint32_t> card;
    for (const Variable *pv : variables) card.push_back(pv->size());
    std::vector<int32_t> foff(1, 0);
    std::vector<uint32_t> fscope;
    std::vector<uint64_t> toff;
    std::vector<double> tab;
    for (const Factor *pf : factors) {
        const Domain &d = pf->domain();
        fscope.insert(fscope.end(), d.ids().begin(), d.ids().end());
        foff.push_back((int32_t)fscope.size());
        toff.push_back(tab.size());
        for (unsigned i = 0; i < pf->size(); ++i) tab.push_back((*pf)[i]);
    }
    _marg_off.assign(1, 0);
    for (uint32_t c : card) _marg_off.push_back(_marg_off.back() + c);
    uint32_t dummy_scope = 0;
    uint64_t dummy_off = 0;
    double dummy_tab = 0.0;
    gpu::check(bnpp_fg_create(gpu::ctx(), (int)card.size(), card.data(), (int)factors.size(), foff.data(),
                              fscope.empty() ? &dummy_scope : fscope.data(), toff.empty() ? &dummy_off : toff.data(),
                              tab.empty() ? &dummy_tab : tab.data(), &_fg),
               "bnpp_fg_create");
}

FactorGraph::FactorGraph(FactorGraph &&g)
    : _variables(std::move(g._variables)), _factors(std::move(g._factors)), _fg(g._fg), _marg(std::move(g._marg)),
      _marg_valid(g._marg_valid), _marg_off(std::move(g._marg_off))
{
    g._fg = nullptr;
}

FactorGraph::~FactorGraph() { bnpp_fg_destroy(_fg); }

unsigned FactorGraph::update(unsigned max, double epsilon)
{
    uint32_t sweeps = 0;
    gpu::check(bnpp_fg_update(_fg, max, epsilon, &sweeps), "bnpp_fg_update");
    _marg_valid = false;
    return sweeps;
}

Factor FactorGraph::marginal(const Variable *v) const
{
    if (!_marg_valid) {
        _marg.assign(_marg_off.back() ? _marg_off.back() : 1, 0.0);
        gpu::check(bnpp_fg_marginals(_fg, _marg.data()), "bnpp_fg_marginals");
        _marg_valid = true;
    }
    std::vector<const Variable*> sc(1, v);
    const unsigned o = _marg_off[v->id()];
    std::vector<double> values(_marg.begin() + o, _marg.begin() + o + v->size());
    return Factor(new Domain(sc), values, 1.0);   // normalised, partition 1 (code/factor.cpp:252)
}

}  // namespace bn
