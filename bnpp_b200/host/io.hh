// UAI-format readers.  API of reference code/io.hh:11-18 (host glue; parsing stays on the host).
#ifndef BNPP_HOST_IO_HH
#define BNPP_HOST_IO_HH

#include <fstream>
#include <string>
#include <unordered_map>

#include "model.hh"

namespace bn {

int read_uai_model(std::string &filename, BN **model);
int read_uai_model(std::string &filename, MN **model);
int read_uai_evidence(std::string &filename, std::unordered_map<unsigned,unsigned> &evidence);

// helpers with external linkage in the reference (code/io.cpp:43-100); harnesses use them to
// load MARKOV files as a bn::BN
std::string read_file_header(std::ifstream &input_file);
void read_variables(std::ifstream &input_file, std::vector<Variable*> &variables);
void read_factors(std::ifstream &input_file, std::vector<Variable*> &variables, std::vector<Factor*> &factors);

}  // namespace bn

#endif
