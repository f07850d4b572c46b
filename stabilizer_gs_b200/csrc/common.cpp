// common.cpp — thread-local last-error string behind sgs_last_error().
#include "common.hpp"

namespace sgs {

static thread_local std::string g_last_error;

void set_last_error(const std::string &msg) { g_last_error = msg; }
const char *get_last_error() { return g_last_error.c_str(); }

}  // namespace sgs
