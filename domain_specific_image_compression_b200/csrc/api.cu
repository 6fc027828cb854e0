// Library-wide entry points: version and thread-local error text.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace sic {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
namespace {
struct Entry { const char *name; const void *fn; };
Entry *entries(int **count) {           // function-local statics: safe to use from other translation units' static initialisers
    static Entry table[64];
    static int n = 0;
    *count = &n;
    return table;
}
}  // namespace
void register_kernel(const char *name, const void *fn) {
    int *n;
    Entry *t = entries(&n);
    if (*n < 64) t[(*n)++] = Entry{name, fn};
}
}  // namespace sic

extern "C" int sic_kernel_count(void) {
    int *n;
    sic::entries(&n);
    return *n;
}
extern "C" const char *sic_kernel_name(int i) {
    int *n;
    sic::Entry *t = sic::entries(&n);
    return (i >= 0 && i < *n) ? t[i].name : nullptr;
}
extern "C" int sic_kernel_registers(const char *name) {
    int *n;
    sic::Entry *t = sic::entries(&n);
    for (int i = 0; i < *n; ++i) {
        if (name && strcmp(t[i].name, name) == 0) {
            cudaFuncAttributes a;
            cudaError_t e = cudaFuncGetAttributes(&a, t[i].fn);
            if (e != cudaSuccess) {
                sic::set_error("sic_kernel_registers(%s): %s", name, cudaGetErrorString(e));
                cudaGetLastError();
                return SIC_E_UNSUPPORTED;
            }
            return a.numRegs;
        }
    }
    sic::set_error("sic_kernel_registers: no kernel named '%s'", name ? name : "(null)");
    return SIC_E_BADARG;
}

extern "C" int sic_version(void) { return SIC_VERSION; }
extern "C" const char *sic_last_error(void) { return sic::g_err; }
