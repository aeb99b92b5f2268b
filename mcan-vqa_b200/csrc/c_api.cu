// Library-level entry points of the C ABI (include/mcan_b200.h): version, error string, SM count.
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

}  // namespace mcan

extern "C" int mcan_version(void) { return MCAN_B200_ABI_VERSION; }
extern "C" const char* mcan_last_error(void) { return mcan::g_last_error; }
extern "C" int mcan_num_sms(void) { return mcan::device_num_sms(); }
