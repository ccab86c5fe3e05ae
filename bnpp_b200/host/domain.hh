// bn::Domain -- ordered scope of a potential table.  Public API of reference
// code/domain.hh:11-45; the layout rule is the reference's: row-major, LAST scope
// variable fastest (code/domain.cpp:20-24).
//
// In this build a Domain is host-side metadata only: it describes a device table to the
// C ABI (`bnpp_scope`).  The per-entry odometer / hash-lookup functions of the reference
// (next_valuation*, position_*) are kept for API compatibility but nothing on the hot
// path calls them -- index arithmetic happens inside the CUDA kernels.
#ifndef BNPP_HOST_DOMAIN_HH
#define BNPP_HOST_DOMAIN_HH

#include "variable.hh"

#include <cstdint>
#include <unordered_map>
#include <vector>

namespace bn {

class Domain {
public:
    Domain();
    Domain(std::vector<const Variable*> scope);
    Domain(const Domain &d);
    Domain(const Domain &d1, const Domain &d2);                 // union: d1, then d2-only variables in d2 order
    Domain(const Domain &d, const Variable *v);                 // d minus v
    Domain(const Domain &d, const std::unordered_map<unsigned,unsigned> &evidence);   // d minus observed

    std::vector<const Variable*> scope() const { return _scope; }
    unsigned width() const { return (unsigned)_scope.size(); }
    unsigned size() const { return _size; }

    const Variable *operator[](unsigned i) const;

    bool in_scope(const Variable *v) const;
    bool in_scope(unsigned id) const;

    void next_valuation(std::vector<unsigned> &valuation) const;
    void next_valuation_with_evidence(std::vector<unsigned> &valuation, const std::unordered_map<unsigned,unsigned> &evidence) const;
    void update_valuation_with_evidence(std::vector<unsigned> &valuation, const std::unordered_map<unsigned,unsigned> &evidence) const;

    unsigned position_valuation(std::vector<unsigned> valuation) const;
    unsigned position_consistent_valuation(std::vector<unsigned> valuation, const Domain &domain) const;
    unsigned position_consistent_valuation(std::vector<unsigned> valuation, const Domain &domain, const Variable *v, unsigned value) const;

    friend std::ostream &operator<<(std::ostream &o, const Domain &v);

    // ---- additions for the C ABI (not in the reference) ----
    const std::vector<uint32_t> &ids() const { return _ids; }
    const std::vector<uint32_t> &cards() const { return _cards; }
    unsigned stride(unsigned i) const { return _stride[i]; }
    int index_of(unsigned id) const;

private:
    void finish();

    std::vector<const Variable*> _scope;
    std::vector<uint32_t> _ids, _cards;
    std::vector<unsigned> _stride;
    unsigned _size;
};

}  // namespace bn

#endif
