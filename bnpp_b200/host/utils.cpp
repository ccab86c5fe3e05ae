#include "utils.hh"

#include <cctype>

namespace bn {

int parse_vars_set(const Model *model, const std::string s, std::unordered_set<const Variable*> &vars_set)
{
    // grammar of code/utils.cpp:8-27: [0-9]+(,[0-9]+)*
    if (s.empty() || s.back() == ',' || s.front() == ',') return -1;
    for (size_t i = 0; i < s.size(); ++i) {
        if (s[i] == ',') {
            if (s[i - 1] == ',') return -1;
        } else if (!std::isdigit((unsigned char)s[i])) {
            return -1;
        }
    }
    size_t pos = 0;
    while (pos < s.size()) {
        size_t comma = s.find(',', pos);
        if (comma == std::string::npos) comma = s.size();
        vars_set.insert(model->variables()[std::stoi(s.substr(pos, comma - pos))]);
        pos = comma + 1;
    }
    return 0;
}

}  // namespace bn
