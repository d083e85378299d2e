// host_util.h -- error reporting shared by the translation units of libfluidstep_b200.so
#pragma once
#include <stdarg.h>

// records a printf-formatted message for fnx_last_error() and returns `code`
int fnx_set_error(int code, const char* fmt, ...);

// kernel-launch counter (bench.py reports it as gpu_launches); bumped by every launch site
void fnx_count_launches(int n);
