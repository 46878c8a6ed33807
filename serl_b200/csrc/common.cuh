// Shared host-side helpers of the C-ABI: error reporting and launch accounting.
#pragma once
#include <cuda_runtime.h>

int serl_fail(int code, const char* msg);
int serl_fail_cuda(cudaError_t e, const char* where);
void serl_count_launch();
