/*
 * ctbd_truncate.cu -- singular-value selection of a split on the device.
 *
 * Device form of the reference's retained_bond_indices (src/algorithm/truncation.c:110-223; rule restated in SURVEY.md 9.7):
 * sort sigma ascending, square, optionally divide by the total, running sum from the smallest, zero the n - max_vdim smallest sums,
 * keep index i iff its running sum > tol; norm (BLAS nrm2 scaling) and von Neumann entropy (:13-27) of the retained values.
 * The singular values never leave the device: what crosses to the host is the RESULT the host's tensor metadata needs anyway (the
 * ascending list of retained indices and three scalars), and the retained values go straight into the scaling kernel.
 *
 *   rank kernel   : rank of every value in the ascending order, equal values ordered by index (one of the orders the reference's
 *                   qsort may produce, and the order the host restatement uses) -- all pairs compared, n <= a few 10^4
 *   select kernel : the floating-point sums run SEQUENTIALLY in one thread with round-to-nearest adds, multiplies and divisions and
 *                   no contraction, i.e. in exactly the order and arithmetic of the reference's loops, so that the comparison
 *                   "running sum > tol" decides bit for bit as on the host; flags and the compaction are integer work.
 */
#include <vector>
#include <float.h>
#include "ctbd_common.cuh"

namespace ctbd {

__global__ void __launch_bounds__(256) trunc_rank_kernel(int64_t n, const double* __restrict__ sigma, double* __restrict__ sorted_v, int64_t* __restrict__ sorted_i)
{
	__shared__ double tile[256];
	const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
	const double vi = (i < n) ? sigma[i] : 0.0;
	int64_t rank = 0;
	for (int64_t j0 = 0; j0 < n; j0 += 256)
	{
		__syncthreads();
		tile[threadIdx.x] = (j0 + threadIdx.x < n) ? sigma[j0 + threadIdx.x] : 0.0;
		__syncthreads();
		const int m = (int)min((int64_t)256, n - j0);
		for (int k = 0; k < m; k++) {
			const double vj = tile[k];
			rank += (vj < vi || (vj == vi && j0 + k < i)) ? 1 : 0;
		}
	}
	if (i < n) { sorted_v[rank] = vi; sorted_i[rank] = i; }
}

/* out: [0] number retained, [1] norm_sigma, [2] entropy, [3] tol_eff (doubles), then the retained indices as int64 */
__global__ void __launch_bounds__(256) trunc_select_kernel(int64_t n, const double* __restrict__ sigma, double* __restrict__ sorted_v, const int64_t* __restrict__ sorted_i,
	double tol, int relative, int64_t max_vdim, int renormalize, int* __restrict__ keep, double* __restrict__ out, double* __restrict__ s_ret)
{
	__shared__ double s_toleff;
	__shared__ int s_zero;
	int64_t* ind = reinterpret_cast<int64_t*>(out + 4);
	if (threadIdx.x == 0)
	{
		double sqsum = 0.0;
		for (int64_t r = 0; r < n; r++) {
			const double sq = __dmul_rn(sorted_v[r], sorted_v[r]);
			sorted_v[r] = sq;
			sqsum = __dadd_rn(sqsum, sq);
		}
		s_zero = (sqsum == 0.0) ? 1 : 0;
		double tol_eff = tol;
		if (sqsum != 0.0)
		{
			double acc = 0.0;
			for (int64_t r = 0; r < n; r++) {
				const double x = relative ? __ddiv_rn(sorted_v[r], sqsum) : sorted_v[r];
				acc = (r == 0) ? x : __dadd_rn(x, acc);
				sorted_v[r] = acc;
			}
			if (max_vdim < n) {
				tol_eff = fmax(tol, sorted_v[n - max_vdim - 1]);
				for (int64_t r = 0; r < n - max_vdim; r++) { sorted_v[r] = 0.0; }
			}
		}
		s_toleff = tol_eff;
	}
	__syncthreads();
	if (s_zero) {
		if (threadIdx.x == 0) { out[0] = 0.0; out[1] = 0.0; out[2] = 0.0; out[3] = s_toleff; s_ret[0] = 0.0; }      /* dummy bond: one zero value */
		return;
	}
	for (int64_t r = threadIdx.x; r < n; r += blockDim.x) { keep[sorted_i[r]] = (sorted_v[r] > tol) ? 1 : 0; }
	__syncthreads();
	if (threadIdx.x == 0)
	{
		/* ascending index list; norm of the retained values with the scaling of BLAS nrm2; norm of all values (renormalisation) */
		int64_t num = 0;
		double scale = 0.0, ssq = 1.0, nrm_all = 0.0;
		for (int64_t i = 0; i < n; i++)
		{
			const double v = sigma[i];
			nrm_all = __dadd_rn(nrm_all, __dmul_rn(v, v));
			if (!keep[i]) { continue; }
			ind[num++] = i;
			const double a = fabs(v);
			if (a > 0.0) {
				if (scale < a) { const double q = __ddiv_rn(scale, a); ssq = __dadd_rn(1.0, __dmul_rn(__dmul_rn(ssq, q), q)); scale = a; }
				else { const double q = __ddiv_rn(a, scale); ssq = __dadd_rn(ssq, __dmul_rn(q, q)); }
			}
		}
		const double norm_sigma = __dmul_rn(scale, __dsqrt_rn(ssq));
		double entropy = 0.0;
		const double resc = (renormalize && num > 0) ? __ddiv_rn(__dsqrt_rn(nrm_all), norm_sigma) : 1.0;
		for (int64_t k = 0; k < num; k++)
		{
			const double v = sigma[ind[k]];
			const double p = __ddiv_rn(v, norm_sigma);
			if (p > 0.0) { const double sq = __dmul_rn(p, p); entropy = __dadd_rn(entropy, -__dmul_rn(sq, log(sq))); }
			s_ret[k] = renormalize ? __dmul_rn(v, resc) : v;
		}
		if (num == 0) { s_ret[0] = 0.0; }
		out[0] = (double)num; out[1] = (num > 0) ? norm_sigma : 0.0; out[2] = (num > 0) ? entropy : 0.0; out[3] = s_toleff;
	}
}

} // namespace ctbd

using namespace ctbd;

extern "C" int ctbd_truncate_select(int64_t n, const double* S_dev, double tol, int relative, int64_t max_vdim, int renormalize,
	int64_t* nret, int64_t* ind_host, double* info3_host, double* s_ret_dev)
{
	CTBD_REQUIRE_INIT();
	*nret = 0; info3_host[0] = 0; info3_host[1] = 0; info3_host[2] = tol;
	if (n <= 0) { return 0; }
	void *d_sv = nullptr, *d_si = nullptr, *d_keep = nullptr, *d_out = nullptr;
	int rc = ctbd_malloc_noinit(&d_sv, (size_t)n * sizeof(double));
	if (rc == 0) { rc = ctbd_malloc_noinit(&d_si, (size_t)n * sizeof(int64_t)); }
	if (rc == 0) { rc = ctbd_malloc(&d_keep, (size_t)n * sizeof(int)); }
	if (rc == 0) { rc = ctbd_malloc(&d_out, (size_t)(n + 4) * sizeof(double)); }
	if (rc == 0) {
		trunc_rank_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, rt().stream>>>(n, S_dev, (double*)d_sv, (int64_t*)d_si);
		rt().launches++;
		trunc_select_kernel<<<1, 256, 0, rt().stream>>>(n, S_dev, (double*)d_sv, (const int64_t*)d_si, tol, relative, max_vdim, renormalize, (int*)d_keep, (double*)d_out, s_ret_dev);
		rt().launches++;
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess) { rc = fail("truncation kernels", e, __FILE__, __LINE__); }
	}
	if (rc == 0)
	{
		/* the result: at most min(n, max_vdim) indices; one copy */
		const int64_t cap = (max_vdim < n && max_vdim >= 0) ? max_vdim : n;
		std::vector<double> h((size_t)cap + 4);
		rc = ctbd_d2h(h.data(), d_out, (size_t)(cap + 4) * sizeof(double));
		if (rc == 0) {
			*nret = (int64_t)h[0];
			info3_host[0] = h[1]; info3_host[1] = h[2]; info3_host[2] = h[3];
			if (*nret > cap) { rc = fail_msg("truncation: more retained indices than max_vdim"); }
			else { memcpy(ind_host, h.data() + 4, (size_t)(*nret) * sizeof(int64_t)); }
		}
	}
	ctbd_free(d_out); ctbd_free(d_keep); ctbd_free(d_si); ctbd_free(d_sv);
	return rc;
}
