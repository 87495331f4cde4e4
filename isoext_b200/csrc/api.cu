// api.cu -- error plumbing and library identification of the C-ABI (include/isoext_b200.h).
#include "common.cuh"

namespace isx {
std::string &last_error() {
    static thread_local std::string e;
    return e;
}
int fail(int code, const std::string &msg) {
    last_error() = msg;
    return code;
}
}   // namespace isx

extern "C" {
const char *isoext_last_error(void) { return isx::last_error().c_str(); }
const char *isoext_build_info(void) { return "isoext_b200 sm_100a " __DATE__; }
int isoext_abi_version(void) { return 1; }
}
