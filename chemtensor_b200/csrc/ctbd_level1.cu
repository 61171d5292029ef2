/*
 * ctbd_level1.cu -- fused level-1 kernels of the Lanczos iteration.
 *
 * Replaces cblas_dnrm2 / dscal / ddot / zdotc and the two hand-written update loops of the reference's
 * lanczos_iteration_d/z (src/util/krylov.c:31-71, :103-150) and the Ritz-vector GEMM (:242, :335).
 * Scalars (alpha_j, beta_j, norms) are produced and consumed on the device.  Reductions are two-level
 * and deterministic: warp shuffle -> shared memory -> one partial per CTA, and the CTA that draws the
 * last ticket adds the partials in index order.  All kernels are HBM-bound streaming passes.
 */
#include "ctbd_common.cuh"

namespace ctbd {

static constexpr int RED_THREADS = 256;
static constexpr int RED_MAX_BLOCKS = 1184;   /* 8 CTAs per SM on 148 SMs */

struct RedScratch { double* partial = nullptr; unsigned int* ticket = nullptr; };
static RedScratch g_red;

static int ensure_scratch()
{
	if (g_red.partial != nullptr) { return 0; }
	void* p = nullptr;
	if (ctbd_malloc(&p, (size_t)(2 * RED_MAX_BLOCKS) * sizeof(double) + 64) < 0) { return -1; }
	g_red.partial = (double*)p;
	g_red.ticket = (unsigned int*)(g_red.partial + 2 * RED_MAX_BLOCKS);
	return 0;
}

__device__ __forceinline__ double warp_sum(double v)
{
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) { v += __shfl_xor_sync(0xffffffffu, v, o); }
	return v;
}

/* block-level sum of (a, b); valid in thread 0 */
__device__ __forceinline__ void block_sum2(double& a, double& b)
{
	__shared__ double sa[RED_THREADS / 32], sb[RED_THREADS / 32];
	a = warp_sum(a); b = warp_sum(b);
	const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
	if (l == 0) { sa[w] = a; sb[w] = b; }
	__syncthreads();
	if (w == 0) {
		a = (l < RED_THREADS / 32) ? sa[l] : 0.0;
		b = (l < RED_THREADS / 32) ? sb[l] : 0.0;
		a = warp_sum(a); b = warp_sum(b);
	}
}

/* finish: last CTA adds the per-CTA partials in index order; mode_sqrt -> out[0] = sqrt(sum) */
__device__ __forceinline__ void finish_reduce(double a, double b, double* partial, unsigned int* ticket, double* out, int nout, int mode_sqrt)
{
	__shared__ bool last;
	if (threadIdx.x == 0) {
		partial[2 * blockIdx.x] = a; partial[2 * blockIdx.x + 1] = b;
		__threadfence();
		const unsigned int t = atomicAdd(ticket, 1u);
		last = (t == gridDim.x - 1);
	}
	__syncthreads();
	if (last) {
		__threadfence();
		double sa = 0, sb = 0;
		for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) { sa += partial[2 * i]; sb += partial[2 * i + 1]; }
		__syncthreads();
		block_sum2(sa, sb);
		if (threadIdx.x == 0) {
			out[0] = mode_sqrt ? sqrt(sa) : sa;
			if (nout > 1) { out[1] = sb; }
			*ticket = 0;
		}
	}
}

/* out = (Re, Im) of sum conj(x_i) y_i; n counts elements */
template <bool CPLX>
__global__ void __launch_bounds__(RED_THREADS) dotc_kernel(int64_t n, const double* __restrict__ x, const double* __restrict__ y,
	double* partial, unsigned int* ticket, double* out)
{
	double re = 0, im = 0;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	if (CPLX) {
		const double2* x2 = reinterpret_cast<const double2*>(x);
		const double2* y2 = reinterpret_cast<const double2*>(y);
		for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
			const double2 a = x2[i], b = y2[i];
			re += a.x * b.x + a.y * b.y;
			im += a.x * b.y - a.y * b.x;
		}
	}
	else {
		for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) { re += x[i] * y[i]; }
	}
	block_sum2(re, im);
	finish_reduce(re, im, partial, ticket, out, 2, 0);
}

/* out[0] = sqrt(sum x_i^2) over nd doubles */
__global__ void __launch_bounds__(RED_THREADS) nrm2_kernel(int64_t nd, const double* __restrict__ x, double* partial, unsigned int* ticket, double* out)
{
	double s = 0, z = 0;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += stride) { const double v = x[i]; s += v * v; }
	block_sum2(s, z);
	finish_reduce(s, z, partial, ticket, out, 1, 1);
}

/* w -= alpha vj + beta vjm1 (real scalars), out[0] = ||w|| afterwards; nd doubles */
__global__ void __launch_bounds__(RED_THREADS) lanczos_update_kernel(int64_t nd, double* __restrict__ w, const double* __restrict__ vj, const double* __restrict__ vjm1,
	const double* __restrict__ alpha, const double* __restrict__ beta_prev, double* partial, unsigned int* ticket, double* out)
{
	const double a = alpha[0];
	const double b = (vjm1 != nullptr) ? beta_prev[0] : 0.0;
	double s = 0, z = 0;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	if (vjm1 != nullptr) {
		for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += stride) {
			const double v = w[i] - (a * vj[i] + b * vjm1[i]);
			w[i] = v; s += v * v;
		}
	}
	else {
		for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += stride) {
			const double v = w[i] - a * vj[i];
			w[i] = v; s += v * v;
		}
	}
	block_sum2(s, z);
	finish_reduce(s, z, partial, ticket, out, 1, 1);
}

__global__ void __launch_bounds__(256) rscale_kernel(int64_t nd, const double* __restrict__ x, const double* __restrict__ s, int divide, double* __restrict__ y)
{
	const double f = s[0];
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	if (divide) { for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += stride) { y[i] = x[i] / f; } }
	else        { for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += stride) { y[i] = x[i] * f; } }
}

__global__ void __launch_bounds__(256) scale_kernel(int64_t nd, double* __restrict__ x, double alpha)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += stride) { x[i] *= alpha; }
}

__global__ void __launch_bounds__(256) zscale_kernel(int64_t n, double2* __restrict__ x, double re, double im)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const double2 v = x[i];
		x[i] = make_double2(v.x * re - v.y * im, v.x * im + v.y * re);
	}
}

static constexpr int LINCOMB_MAX = 32;
struct LincombCoef { double c[LINCOMB_MAX]; };

/* out (+)= sum_j c[j] V[j*ldv + i]; nd / ldv count doubles */
__global__ void __launch_bounds__(256) lincomb_kernel(int64_t nd, const double* __restrict__ V, int64_t ldv, int m, const LincombCoef coef, int accumulate, double* __restrict__ out)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += stride) {
		double s = accumulate ? out[i] : 0.0;
		for (int j = 0; j < m; j++) { s += coef.c[j] * V[(int64_t)j * ldv + i]; }
		out[i] = s;
	}
}

static inline int stream_blocks(int64_t n, int threads, int per_sm)
{
	int64_t b = ceil_div(n, threads);
	const int64_t maxb = (int64_t)rt().sm_count * per_sm;
	if (b > maxb) { b = maxb; }
	if (b < 1) { b = 1; }
	return (int)b;
}

static inline int red_blocks(int64_t n)
{
	int b = stream_blocks(n, RED_THREADS, 8);
	return b > RED_MAX_BLOCKS ? RED_MAX_BLOCKS : b;
}

static inline int64_t ndoubles(int dtype, int64_t n) { return dtype == CTBD_C128 ? 2 * n : n; }

} // namespace ctbd

using namespace ctbd;

extern "C" {

int ctbd_dotc(int dtype, int64_t n, const void* x, const void* y, double* out_dev)
{
	CTBD_REQUIRE_INIT();
	if (ensure_scratch() < 0) { return -1; }
	if (dtype == CTBD_C128) {
		dotc_kernel<true><<<red_blocks(n), RED_THREADS, 0, rt().stream>>>(n, (const double*)x, (const double*)y, g_red.partial, g_red.ticket, out_dev);
	}
	else {
		dotc_kernel<false><<<red_blocks(n), RED_THREADS, 0, rt().stream>>>(n, (const double*)x, (const double*)y, g_red.partial, g_red.ticket, out_dev);
	}
	CTBD_LAUNCH_CHECK();
	return 0;
}

int ctbd_nrm2(int dtype, int64_t n, const void* x, double* out_dev)
{
	CTBD_REQUIRE_INIT();
	if (ensure_scratch() < 0) { return -1; }
	const int64_t nd = ndoubles(dtype, n);
	nrm2_kernel<<<red_blocks(nd), RED_THREADS, 0, rt().stream>>>(nd, (const double*)x, g_red.partial, g_red.ticket, out_dev);
	CTBD_LAUNCH_CHECK();
	return 0;
}

int ctbd_rscale(int dtype, int64_t n, const void* x, const double* s_dev, int divide, void* y)
{
	CTBD_REQUIRE_INIT();
	const int64_t nd = ndoubles(dtype, n);
	rscale_kernel<<<stream_blocks(nd, 256, 8), 256, 0, rt().stream>>>(nd, (const double*)x, s_dev, divide, (double*)y);
	CTBD_LAUNCH_CHECK();
	return 0;
}

int ctbd_lanczos_update(int dtype, int64_t n, void* w, const void* vj, const void* vjm1,
	const double* alpha_dev, const double* beta_prev_dev, double* out_dev)
{
	CTBD_REQUIRE_INIT();
	if (ensure_scratch() < 0) { return -1; }
	const int64_t nd = ndoubles(dtype, n);
	lanczos_update_kernel<<<red_blocks(nd), RED_THREADS, 0, rt().stream>>>(nd, (double*)w, (const double*)vj, (const double*)vjm1,
		alpha_dev, beta_prev_dev, g_red.partial, g_red.ticket, out_dev);
	CTBD_LAUNCH_CHECK();
	return 0;
}

int ctbd_lincomb(int dtype, int64_t n, const void* V, int64_t ldv, int m, const double* coef_host, void* out)
{
	CTBD_REQUIRE_INIT();
	const int64_t nd = ndoubles(dtype, n);
	const int64_t ldd = ndoubles(dtype, ldv);
	if (m <= 0) { return ctbd_memset_zero(out, (size_t)nd * sizeof(double)); }
	for (int j0 = 0; j0 < m; j0 += LINCOMB_MAX)
	{
		const int mc = (m - j0 < LINCOMB_MAX) ? m - j0 : LINCOMB_MAX;
		LincombCoef c;
		for (int j = 0; j < LINCOMB_MAX; j++) { c.c[j] = (j < mc) ? coef_host[j0 + j] : 0.0; }
		lincomb_kernel<<<stream_blocks(nd, 256, 8), 256, 0, rt().stream>>>(nd, (const double*)V + (int64_t)j0 * ldd, ldd, mc, c, j0 > 0, (double*)out);
		CTBD_LAUNCH_CHECK();
	}
	return 0;
}

int ctbd_zscale_host(int dtype, int64_t n, void* x, double re, double im)
{
	CTBD_REQUIRE_INIT();
	if (dtype != CTBD_C128) { return ctbd_scale_host(dtype, n, x, re); }
	zscale_kernel<<<stream_blocks(n, 256, 8), 256, 0, rt().stream>>>(n, (double2*)x, re, im);
	CTBD_LAUNCH_CHECK();
	return 0;
}

int ctbd_scale_host(int dtype, int64_t n, void* x, double alpha)
{
	CTBD_REQUIRE_INIT();
	const int64_t nd = ndoubles(dtype, n);
	scale_kernel<<<stream_blocks(nd, 256, 8), 256, 0, rt().stream>>>(nd, (double*)x, alpha);
	CTBD_LAUNCH_CHECK();
	return 0;
}

} // extern "C"
