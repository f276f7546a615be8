#include "common.cuh"
#include <cstdarg>
#include <cstdlib>

namespace pd {
namespace {
thread_local char g_err[1024] = "";
}
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char* last_error() { return g_err; }
bool pdl_enabled() {
    static const bool on = getenv("PD_NO_PDL") == nullptr;
    return on;
}
}  // namespace pd
