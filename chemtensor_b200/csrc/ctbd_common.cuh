/*
 * ctbd_common.cuh -- shared state and helpers of the thin C-ABI CUDA layer (include/ctb_device.h).
 *
 * One process drives one B200: a single non-blocking stream carries every kernel and copy, device
 * memory comes from the stream-ordered pool (cudaMallocAsync) so the thousands of short-lived
 * temporaries of a DMRG sweep never hit cudaMalloc/cudaFree, and every launch is counted.
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "ctb_device.h"

namespace ctbd {

struct Runtime
{
	bool ready = false;
	int device = 0;
	int sm_count = 0;
	int smem_optin = 0;
	cudaStream_t stream = nullptr;
	long long launches = 0;
	long long bytes_in_use = 0;
	char err[512] = "";
};

Runtime& rt();

int fail(const char* what, cudaError_t e, const char* file, int line);
int fail_msg(const char* msg);

#define CTBD_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { return ::ctbd::fail(#call, e_, __FILE__, __LINE__); } } while (0)
#define CTBD_REQUIRE_INIT() do { if (!::ctbd::rt().ready) { int rc_ = ctbd_init(-1); if (rc_ < 0) { return rc_; } } } while (0)
#define CTBD_LAUNCH_CHECK() do { ::ctbd::rt().launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { return ::ctbd::fail("kernel launch", e_, __FILE__, __LINE__); } } while (0)

/* upload a host array into pool memory on the layer's stream (the host array may be freed on return) */
int upload(const void* host, size_t bytes, void** dev);

static inline __host__ __device__ int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

} // namespace ctbd
