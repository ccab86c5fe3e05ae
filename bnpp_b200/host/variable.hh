// bn::Variable -- value type (id, cardinality).  API of reference code/variable.hh:8-20.
#ifndef BNPP_HOST_VARIABLE_HH
#define BNPP_HOST_VARIABLE_HH

#include <ostream>

namespace bn {

class Variable {
public:
    Variable(unsigned id, unsigned size) : _id(id), _size(size) {}

    unsigned id() const { return _id; }
    unsigned size() const { return _size; }

    friend std::ostream &operator<<(std::ostream &o, const Variable &v);

private:
    unsigned _id, _size;
};

}  // namespace bn

#endif
