// Error reporting and ABI bookkeeping for libdqomap_b200.so.
#include "common.cuh"
#include <stdarg.h>
#include <stdio.h>

namespace dqo {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
} // namespace dqo

extern "C" int dqo_abi_version(void) { return DQO_ABI_VERSION; }
extern "C" const char *dqo_last_error(void) { return dqo::g_err; }
