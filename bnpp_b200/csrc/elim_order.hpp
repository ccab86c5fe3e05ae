// Elimination-order heuristics on the interaction graph (host side).
//
// Replaces bn::Graph's constructor, ordering(), min_fill(), weighted_min_fill(),
// min_degree() and order_width() (reference code/graph.cpp:9-237).  north_star keeps
// this on the host and demands BIT-EXACT orders.  The reference breaks ties by the
// iteration order of libstdc++ `std::unordered_set<unsigned>`, so this
// file deliberately uses the same containers and performs insertions, erasures and
// scans in the same sequence; only then do `*begin()`, the strict `<` scans and the
// fill-in edge insertions see the same element order.  Header-only so the C-ABI
// library (bnpp_elim_order) and the bn::Graph wrapper share one implementation.
#pragma once
#include <cstdint>
#include <unordered_map>
#include <unordered_set>
#include <vector>

namespace bnpp {

enum Heuristic { H_MIN_FILL = 0, H_WEIGHTED_MIN_FILL = 1, H_MIN_DEGREE = 2 };

class InteractionGraph {
public:
    typedef std::unordered_set<unsigned> NodeSet;
    typedef std::unordered_map<unsigned, NodeSet> Adjacency;

    InteractionGraph() {}

    // scopes[f] = variable ids of factor f in scope order; card[id] = cardinality
    InteractionGraph(const std::vector<std::vector<unsigned>> &scopes, const std::vector<unsigned> &card) : card_(card)
    {
        // every scope variable becomes a node first (code/graph.cpp:11-18) ...
        for (const auto &sc : scopes)
            for (unsigned v : sc) adj_[v] = NodeSet();
        // ... then each factor contributes a clique, pairs (i, j > i) in scope order (code/graph.cpp:20-34)
        for (const auto &sc : scopes) {
            const size_t w = sc.size();
            for (size_t i = 0; i + 1 < w; ++i)
                for (size_t j = i + 1; j < w; ++j) {
                    adj_[sc[i]].insert(sc[j]);
                    adj_[sc[j]].insert(sc[i]);
                }
        }
    }

    const Adjacency &adjacency() const { return adj_; }
    Adjacency &adjacency() { return adj_; }

    bool has(unsigned id) const { return adj_.find(id) != adj_.end(); }

    // The reference dereferences find() unchecked and crashes on a node it does not know
    // (SURVEY A.2 i); here an unknown node simply has no neighbours.
    const NodeSet &neighbors(unsigned id) const
    {
        static const NodeSet none;
        Adjacency::const_iterator it = adj_.find(id);
        return it == adj_.end() ? none : it->second;
    }

    bool connected(unsigned a, unsigned b) const
    {
        Adjacency::const_iterator it = adj_.find(a);
        return it != adj_.end() && it->second.count(b) != 0;
    }

    // number (or cardinality-weighted number) of edges elimination of `id` would add
    unsigned fill_in(unsigned id, bool weighted) const
    {
        const NodeSet &nb = neighbors(id);
        unsigned score = 0;
        for (unsigned a : nb)
            for (unsigned b : nb)
                if (a < b && !connected(a, b)) score += weighted ? card_.at(a) * card_.at(b) : 1u;
        return score;
    }

    // code/graph.cpp:103-120
    unsigned pick_min_degree(const NodeSet &cand) const
    {
        unsigned best = *cand.begin();
        unsigned best_deg = (unsigned)adj_.size() + 1;
        for (unsigned id : cand) {
            const unsigned deg = (unsigned)neighbors(id).size();
            if (deg < best_deg) {
                best = id;
                best_deg = deg;
            }
        }
        return best;
    }

    // code/graph.cpp:122-153: the best score is seeded with |nodes|+1, NOT with the first candidate's score
    unsigned pick_min_fill(const NodeSet &cand) const
    {
        unsigned best = *cand.begin();
        unsigned best_fill = (unsigned)adj_.size() + 1;
        for (unsigned id : cand) {
            const unsigned f = fill_in(id, false);
            if (f < best_fill || (f == best_fill && neighbors(id).size() < neighbors(best).size())) {
                best = id;
                best_fill = f;
            }
        }
        return best;
    }

    // code/graph.cpp:155-195: seeded with the first candidate's real score
    unsigned pick_weighted_min_fill(const NodeSet &cand) const
    {
        unsigned best = *cand.begin();
        unsigned best_fill = fill_in(best, true);
        for (unsigned id : cand) {
            const unsigned f = fill_in(id, true);
            if (f < best_fill || (f == best_fill && neighbors(id).size() < neighbors(best).size())) {
                best = id;
                best_fill = f;
            }
        }
        return best;
    }

    // Removes `id`, connecting its neighbours pairwise; returns its degree at that moment.
    // Sequence of mutations as in code/graph.cpp:73-96.
    unsigned eliminate(unsigned id, bool erase_node_first)
    {
        const NodeSet nb = neighbors(id);   // a copy iterates exactly like the original
        if (erase_node_first) adj_.erase(id);          // order_width() erases the node first (code/graph.cpp:217-218)
        for (unsigned a : nb) adj_[a].erase(id);
        for (unsigned a : nb)
            for (unsigned b : nb)
                if (a != b && !connected(a, b)) {
                    adj_[a].insert(b);
                    adj_[b].insert(a);
                }
        if (!erase_node_first) adj_.erase(id);
        return (unsigned)nb.size();
    }

    // Graph::ordering, code/graph.cpp:41-101.  `vars` in the caller's order.
    std::vector<unsigned> ordering(const std::vector<unsigned> &vars, Heuristic h, unsigned &width) const
    {
        InteractionGraph g(*this);
        NodeSet cand;
        for (unsigned v : vars) cand.insert(v);
        std::vector<unsigned> order;
        order.reserve(vars.size());
        width = 0;
        while (!cand.empty()) {
            unsigned next;
            if (h == H_MIN_DEGREE) next = g.pick_min_degree(cand);
            else if (h == H_WEIGHTED_MIN_FILL) next = g.pick_weighted_min_fill(cand);
            else next = g.pick_min_fill(cand);
            order.push_back(next);
            const unsigned deg = g.eliminate(next, false);
            if (deg > width) width = deg;
            cand.erase(next);
        }
        return order;
    }

    // Graph::order_width, code/graph.cpp:197-237
    unsigned order_width(const std::vector<unsigned> &order) const
    {
        InteractionGraph g(*this);
        unsigned width = 0;
        for (unsigned v : order) {
            const unsigned deg = g.eliminate(v, true);
            if (deg > width) width = deg;
        }
        return width;
    }

private:
    std::vector<unsigned> card_;
    Adjacency adj_;
};

}  // namespace bnpp
