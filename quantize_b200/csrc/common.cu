// qb200 C-ABI: version, error string, launch counter.
#include <stdlib.h>
#include "common.cuh"

namespace qb200 {

static thread_local char g_err[512] = "";
static thread_local uint64_t g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches += (uint64_t)n; }

bool pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("QB200_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}

}  // namespace qb200

extern "C" {

int qb200_version(void) { return QB200_VERSION; }

const char* qb200_last_error(void) { return qb200::g_err; }

uint64_t qb200_launch_count(void) { return qb200::g_launches; }

void qb200_launch_count_reset(void) { qb200::g_launches = 0; }

}  // extern "C"
