// Version / error reporting of the C ABI (include/pwc_b200.h).
#include "common.cuh"
#include <cstdarg>

namespace pwc {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace pwc

extern "C" int pwc_version(void) { return PWC_ABI_VERSION; }
extern "C" const char* pwc_last_error(void) { return pwc::g_err; }
