// API of reference code/utils.hh:11-12.
#ifndef BNPP_HOST_UTILS_HH
#define BNPP_HOST_UTILS_HH

#include "model.hh"

#include <string>
#include <unordered_set>

namespace bn {

// "1,2,3" -> variables of the model; -1 when the text is not a comma-separated id list
int parse_vars_set(const Model *model, const std::string s, std::unordered_set<const Variable*> &vars_set);

}  // namespace bn

#endif
