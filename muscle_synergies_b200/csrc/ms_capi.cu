// Process-wide bookkeeping of the C ABI (include/ms_b200.h).
#include <stdio.h>

#include "ms_common.cuh"

int64_t g_ms_launches = 0;                      // bumped atomically: entry points may run on several host threads
thread_local char g_ms_last_error[256] = "";    // per calling thread, like errno

extern "C" const char* ms_last_cuda_error(void) { return g_ms_last_error; }
extern "C" const char* ms_version(void) { return "muscle_synergies_b200 0.1 (sm_100a)"; }
extern "C" int64_t ms_launch_count(void) { return __atomic_load_n(&g_ms_launches, __ATOMIC_RELAXED); }
