#include "variable.hh"

namespace bn {

// same text as the reference (code/variable.cpp:12-17): the CLIs' -v output depends on it
std::ostream &operator<<(std::ostream &o, const Variable &v)
{
    return o << "Variable(id:" << v._id << ", size:" << v._size << ")";
}

}  // namespace bn
