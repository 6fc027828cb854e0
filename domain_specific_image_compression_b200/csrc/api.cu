// Library-wide entry points: version and thread-local error text.
#include <stdarg.h>

#include "common.cuh"

namespace sic {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace sic

extern "C" int sic_version(void) { return SIC_VERSION; }
extern "C" const char *sic_last_error(void) { return sic::g_err; }
