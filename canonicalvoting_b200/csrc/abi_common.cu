// canonicalvoting_b200/csrc/abi_common.cu -- version + error plumbing of the C ABI.
#include <stdarg.h>

#include "common.cuh"

namespace cvb200 {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace cvb200

extern "C" int cvb200_abi_version(void) { return CVB200_ABI_VERSION; }
extern "C" const char *cvb200_last_error(void) { return cvb200::g_err; }
