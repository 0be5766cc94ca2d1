// error.cu -- thread-local last-error string + version entry points.
#include <stdarg.h>

#include "common.cuh"

namespace fnx {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace fnx

extern "C" {
int fnx_abi_version(void) { return FNX_ABI_VERSION; }
const char *fnx_last_error(void) { return fnx::g_err; }
const char *fnx_build_arch(void) { return "sm_100a"; }
}
