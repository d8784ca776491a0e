#include "common.cuh"

namespace ttb {

namespace {
thread_local std::string g_last_error;
}

void set_last_error(const std::string& msg) { g_last_error = msg; }
const char* last_error_cstr() { return g_last_error.c_str(); }

}  // namespace ttb
