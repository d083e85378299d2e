// host_util.cpp -- error string + build info of libfluidstep_b200.so
#include "host_util.h"

#include <stdio.h>

#include "../../include/fluidstep.h"

#include <atomic>

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void fnx_count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int fnx_set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

extern "C" {
const char* fnx_last_error(void) { return g_err; }
const char* fnx_build_info(void) { return "libfluidstep_b200 sm_100a " __DATE__ " " __TIME__; }
int fnx_abi_version(void) { return 1; }
long long fnx_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
}
