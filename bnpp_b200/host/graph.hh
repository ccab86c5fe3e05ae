// bn::Graph (interaction graph + elimination-order heuristics, host) and bn::FactorGraph
// (loopy sum-product, device).  Public API of reference code/graph.hh:13-55.
#ifndef BNPP_HOST_GRAPH_HH
#define BNPP_HOST_GRAPH_HH

#include "variable.hh"
#include "factor.hh"

#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

struct bnpp_fg;

namespace bnpp { class InteractionGraph; }

namespace bn {

class Graph {
public:
    Graph(const std::vector<const Variable*> &variables, const std::vector<const Factor*> &factors);
    Graph(const Graph &g);
    ~Graph();

    std::unordered_set<unsigned> neighbors(unsigned id) const;
    bool connected(unsigned id1, unsigned id2) const;

    std::vector<unsigned> ordering(
        const std::vector<const Variable*> &variables,
        unsigned &width,
        std::unordered_map<std::string,bool> &options) const;

    unsigned min_fill(const std::unordered_set<unsigned> &vars) const;
    unsigned weighted_min_fill(const std::unordered_set<unsigned> &vars) const;
    unsigned min_degree(const std::unordered_set<unsigned> &vars) const;

    unsigned order_width(const std::vector<const Variable*> &variables) const;

    friend std::ostream &operator<<(std::ostream &os, const Graph &g);

private:
    const std::vector<const Variable*> _variables;
    bnpp::InteractionGraph *_g;     // same containers and mutation order as the reference
};

class FactorGraph {
public:
    FactorGraph(const std::vector<const Variable*> &variables, const std::vector<const Factor*> &factors);
    FactorGraph(FactorGraph &&g);
    ~FactorGraph();

    unsigned update(unsigned max, double epsilon);
    Factor marginal(const Variable *v) const;

private:
    FactorGraph(const FactorGraph &);
    std::vector<const Variable*> _variables;
    std::vector<const Factor*> _factors;
    bnpp_fg *_fg;                   // messages and edge tables live on the device
    mutable std::vector<double> _marg;
    mutable bool _marg_valid;
    std::vector<unsigned> _marg_off;
};

}  // namespace bn

#endif
