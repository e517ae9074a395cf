// Process-wide bookkeeping of the C ABI (include/ms_b200.h).
#include <stdio.h>

#include "ms_common.cuh"

int64_t g_ms_launches = 0;
char g_ms_last_error[256] = "";

extern "C" const char* ms_last_cuda_error(void) { return g_ms_last_error; }
extern "C" const char* ms_version(void) { return "muscle_synergies_b200 0.1 (sm_100a)"; }
extern "C" int64_t ms_launch_count(void) { return g_ms_launches; }
