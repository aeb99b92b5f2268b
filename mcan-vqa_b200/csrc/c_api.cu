// Library-level entry points of the C ABI (include/mcan_b200.h): version, error string, SM count.
#include <atomic>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/mcan_b200.h"
#include "common.cuh"

namespace mcan {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

int device_num_sms();

static std::atomic<int> g_pdl{1};
bool pdl_enabled() { return g_pdl.load(std::memory_order_relaxed) != 0; }

}  // namespace mcan

extern "C" int mcan_set_pdl(int enabled) {
    mcan::g_pdl.store(enabled ? 1 : 0, std::memory_order_relaxed);
    return 0;
}

extern "C" int mcan_version(void) { return MCAN_B200_ABI_VERSION; }
extern "C" const char* mcan_last_error(void) { return mcan::g_last_error; }
extern "C" int mcan_num_sms(void) { return mcan::device_num_sms(); }
