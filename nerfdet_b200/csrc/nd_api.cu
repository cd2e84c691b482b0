// Library-wide entry points of the C ABI: version and per-thread error string.
#include <stdarg.h>

#include "nd_common.cuh"

namespace nd {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

}  // namespace nd

extern "C" {

int nd_version(void) { return ND_VERSION; }

const char *nd_last_error_string(void) { return nd::g_err; }

}  // extern "C"
