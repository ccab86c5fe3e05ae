#include "domain.hh"

#include <iostream>

namespace bn {

// strides and size, reference code/domain.cpp:15-26.  Sizes are the reference's 32-bit
// `unsigned`; a table of 2^32 entries or more is refused at this edge.
void Domain::finish()
{
    const size_t w = _scope.size();
    _ids.resize(w);
    _cards.resize(w);
    _stride.resize(w);
    uint64_t n = 1;
    for (size_t i = w; i-- > 0;) {
        _ids[i] = _scope[i]->id();
        _cards[i] = _scope[i]->size();
        _stride[i] = (unsigned)n;
        n *= _cards[i];
        if (n >> 32) throw "Domain: table would have 2^32 entries or more.";
    }
    _size = (unsigned)n;
}

Domain::Domain() : _size(1) {}

Domain::Domain(std::vector<const Variable*> scope) : _scope(scope) { finish(); }

Domain::Domain(const Domain &d) : _scope(d._scope) { finish(); }

Domain::Domain(const Domain &d1, const Domain &d2) : _scope(d1._scope)
{
    for (const Variable *v : d2._scope)
        if (!d1.in_scope(v)) _scope.push_back(v);
    finish();
}

Domain::Domain(const Domain &d, const Variable *v)
{
    for (const Variable *u : d._scope)
        if (u != v) _scope.push_back(u);
    finish();
}

Domain::Domain(const Domain &d, const std::unordered_map<unsigned,unsigned> &evidence)
{
    for (const Variable *u : d._scope)
        if (evidence.find(u->id()) == evidence.end()) _scope.push_back(u);
    finish();
}

const Variable *Domain::operator[](unsigned i) const
{
    if (i >= _scope.size()) throw "Domain::operator[unsigned i]: Index out of range!";   // code/domain.cpp:96
    return _scope[i];
}

int Domain::index_of(unsigned id) const
{
    for (size_t i = 0; i < _ids.size(); ++i)
        if (_ids[i] == id) return (int)i;
    return -1;
}

bool Domain::in_scope(const Variable *v) const { return index_of(v->id()) >= 0; }
bool Domain::in_scope(unsigned id) const { return index_of(id) >= 0; }

// mixed-radix +1, last digit fastest (code/domain.cpp:113-123)
void Domain::next_valuation(std::vector<unsigned> &valuation) const
{
    for (size_t j = valuation.size(); j-- > 0;) {
        if (valuation[j] + 1 < _cards[j]) {
            ++valuation[j];
            return;
        }
        valuation[j] = 0;
    }
}

// the same with observed digits pinned (code/domain.cpp:125-136)
void Domain::next_valuation_with_evidence(std::vector<unsigned> &valuation, const std::unordered_map<unsigned,unsigned> &evidence) const
{
    for (size_t j = valuation.size(); j-- > 0;) {
        if (evidence.count(_ids[j])) continue;
        if (valuation[j] + 1 < _cards[j]) {
            ++valuation[j];
            return;
        }
        valuation[j] = 0;
    }
}

void Domain::update_valuation_with_evidence(std::vector<unsigned> &valuation, const std::unordered_map<unsigned,unsigned> &evidence) const
{
    for (size_t j = 0; j < _ids.size(); ++j) {
        auto it = evidence.find(_ids[j]);
        if (it != evidence.end()) valuation[j] = it->second;
    }
}

unsigned Domain::position_valuation(std::vector<unsigned> valuation) const
{
    unsigned pos = 0;
    for (size_t j = 0; j < _ids.size(); ++j) pos += valuation[j] * _stride[j];
    return pos;
}

// an axis the other domain does not have contributes nothing (code/domain.cpp:162-179)
unsigned Domain::position_consistent_valuation(std::vector<unsigned> valuation, const Domain &domain) const
{
    unsigned pos = 0;
    for (size_t j = 0; j < _ids.size(); ++j) {
        const int k = domain.index_of(_ids[j]);
        if (k >= 0) pos += _stride[j] * valuation[k];
    }
    return pos;
}

unsigned Domain::position_consistent_valuation(std::vector<unsigned> valuation, const Domain &domain, const Variable *v, unsigned value) const
{
    unsigned pos = position_consistent_valuation(valuation, domain);
    const int k = index_of(v->id());
    if (k >= 0) pos += _stride[k] * value;
    return pos;
}

std::ostream &operator<<(std::ostream &o, const Domain &d)
{
    o << "Domain{";
    for (size_t i = 0; i < d._ids.size(); ++i) o << (i ? ", " : "") << d._ids[i];
    return o << "}";
}

}  // namespace bn
