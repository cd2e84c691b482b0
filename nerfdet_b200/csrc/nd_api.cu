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

void set_lift_trace(int *buf);

}  // namespace nd

extern "C" {

// Development hook (not part of include/nerfdet_lift.h): device buffer of (warps + 1) * 256 * 4 int32 that CTA 0 of
// k_lift_planes fills with per-warp per-stage clock stamps; NULL switches tracing off.  Used by tools/lift_trace.py.
void nd_debug_set_trace(void *device_buffer) { nd::set_lift_trace(reinterpret_cast<int *>(device_buffer)); }


int nd_version(void) { return ND_VERSION; }

const char *nd_last_error_string(void) { return nd::g_err; }

}  // extern "C"
