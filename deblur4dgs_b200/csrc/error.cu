// error.cu -- thread-local error string + version of the C ABI.
#include <stdarg.h>

#include "common.cuh"

namespace d4 {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace d4

extern "C" int d4_version(void) { return D4_ABI_VERSION; }
extern "C" const char *d4_last_error(void) { return d4::g_err; }
