#include <atomic>
#include <stdio.h>
#include <string.h>

#include "../../include/serl_b200.h"
#include "common.cuh"

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int serl_fail(int code, const char* msg)
{
    snprintf(g_err, sizeof(g_err), "%s", msg);
    return code;
}

int serl_fail_cuda(cudaError_t e, const char* where)
{
    snprintf(g_err, sizeof(g_err), "%s: %s", where, cudaGetErrorString(e));
    return SERL_ERR_CUDA;
}

void serl_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

extern "C" int64_t serl_launch_count(void) { return (int64_t)g_launches.load(); }
extern "C" const char* serl_last_error(void) { return g_err; }
