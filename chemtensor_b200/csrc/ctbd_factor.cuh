/*
 * ctbd_factor.cuh -- element-type helpers shared by the batched factorization kernels (ctbd_factor.cu, ctbd_svd_bj.cu).
 */
#pragma once

#include <float.h>
#include "ctbd_common.cuh"

namespace ctbd {

/* ---- minimal complex helpers so that one template covers double and double2 ---- */
__device__ __forceinline__ double  cj(double a)  { return a; }
__device__ __forceinline__ double2 cj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ double  mul(double a, double b)   { return a * b; }
__device__ __forceinline__ double2 mul(double2 a, double2 b) { return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ double  smul(double s, double a)  { return s * a; }
__device__ __forceinline__ double2 smul(double s, double2 a) { return make_double2(s * a.x, s * a.y); }
__device__ __forceinline__ double  add(double a, double b)   { return a + b; }
__device__ __forceinline__ double2 add(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double  sub(double a, double b)   { return a - b; }
__device__ __forceinline__ double2 sub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double  abs2(double a)  { return a * a; }
__device__ __forceinline__ double  abs2(double2 a) { return a.x * a.x + a.y * a.y; }
__device__ __forceinline__ double  re(double a)  { return a; }
__device__ __forceinline__ double  re(double2 a) { return a.x; }
__device__ __forceinline__ double  im(double)  { return 0.0; }
__device__ __forceinline__ double  im(double2 a) { return a.y; }
template <typename T> __device__ __forceinline__ T from_real(double r);
template <> __device__ __forceinline__ double  from_real<double>(double r)  { return r; }
template <> __device__ __forceinline__ double2 from_real<double2>(double r) { return make_double2(r, 0.0); }
__device__ __forceinline__ double  shfl_xor(double v, int o)  { return __shfl_xor_sync(0xffffffffu, v, o); }
__device__ __forceinline__ double2 shfl_xor(double2 v, int o) { return make_double2(__shfl_xor_sync(0xffffffffu, v.x, o), __shfl_xor_sync(0xffffffffu, v.y, o)); }

__device__ __forceinline__ double absmax_of(double a)  { return fabs(a); }
__device__ __forceinline__ double absmax_of(double2 a) { return fmax(fabs(a.x), fabs(a.y)); }

/* power-of-two factor that brings a block with largest entry 'amax' to O(1), as LAPACK's ?lascl-based drivers do:
 * the factorizations square entries (norms, Gram entries), which would under- or overflow for |a| ~ 1e+-160 */
__device__ __forceinline__ double pow2_scale(double amax)
{
	if (!(amax > 0.0) || !isfinite(amax)) { return 1.0; }
	int e; frexp(amax, &e);
	return ldexp(1.0, -e);
}

/* block-wide maximum; result valid in all threads; 'red' has blockDim.x / 32 entries */
__device__ __forceinline__ double block_max(double v, double* red)
{
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) { v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o)); }
	__syncthreads();
	if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = v; }
	__syncthreads();
	double r = red[0];
	for (int w = 1; w < (int)(blockDim.x >> 5); w++) { r = fmax(r, red[w]); }
	__syncthreads();
	return r;
}

/* one CTA per matrix: scale[b] = power-of-two normalisation of block b (m x n entries at a_off) */
template <typename T>
static __global__ void __launch_bounds__(256) absmax_scale_kernel(const int64_t* __restrict__ a_off, const int64_t* __restrict__ numel, const T* __restrict__ A, double* __restrict__ scale)
{
	__shared__ double red[8];
	const T* a = A + a_off[blockIdx.x];
	double v = 0;
	for (int64_t e = threadIdx.x; e < numel[blockIdx.x]; e += blockDim.x) { v = fmax(v, absmax_of(a[e])); }
	const double amax = block_max(v, red);
	if (threadIdx.x == 0) { scale[blockIdx.x] = pow2_scale(amax); }
}


/* block-Jacobi SVD of the blocks that do not fit into shared memory (ctbd_svd_bj.cu); returns 0, or < 0 on failure */
template <typename T>
int svd_bj_impl(int nmat, const ctbd_mat_desc* descs, const void* A, void* U, void* Vh, double* S);

} // namespace ctbd
