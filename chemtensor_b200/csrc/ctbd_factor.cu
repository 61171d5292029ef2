/*
 * ctbd_factor.cu -- batched dense factorizations of the sector blocks, entirely on the device.
 *
 * Replaces the per-block LAPACK calls of the reference (src/tensor/dense_tensor.c):
 *   ?gesvd            (:3538 dense_tensor_svd_fill)  -> one-sided Jacobi (Hestenes) SVD
 *   ?geqrf + ?orgqr   (:2253 dense_tensor_qr_fill)   -> Householder QR, LAPACK sign convention
 *   ?gerqf + ?orgrq   (:2680 dense_tensor_rq_fill)   -> the same QR applied to (E A E)^H
 * All blocks of a block-sparse matrix are processed by the same launches (work items = blocks x row pairs).
 *
 * SVD: the rows of G = A (m <= n) or A^H (m > n) are orthogonalised by plane rotations in a round-robin
 * (tournament) ordering, R/2 independent row pairs per round, one warp per pair; the rotations are
 * accumulated in W (G = W G0).  Singular values are the final row norms (sorted descending per block as
 * LAPACK does), the singular vectors the normalised rows and the rows of W.  One-sided Jacobi delivers
 * singular vectors orthogonal to working precision independent of the conditioning of the block.
 */
#include <vector>
#include <algorithm>
#include <float.h>
#include <cooperative_groups.h>
#include "ctbd_factor.cuh"

namespace cg = cooperative_groups;

namespace ctbd {

/* ============================================================================================== */
/* SVD                                                                                              */
/* ============================================================================================== */

struct SvdMat
{
	int64_t a_off, u_off, vh_off, s_off;
	int64_t g_off;        /* element offset of the work matrix [G | W] (R x (C + R), row-major) */
	int32_t m, n, R, C;
	int32_t pair_begin;   /* first work item (row pair slot) of this matrix in a round */
	int32_t npair;        /* pair slots per round: ceil(R / 2) */
};
/* note: blocks are normalised by a power of two on load (scale[b]) and the singular values scaled back */

template <typename T>
__global__ void __launch_bounds__(256) svd_init_kernel(int nmat, const SvdMat* __restrict__ mats, const double* __restrict__ scale, const T* __restrict__ A, T* __restrict__ G)
{
	const SvdMat mt = mats[blockIdx.y];
	const double sc = scale[blockIdx.y];
	const int ld = mt.C + mt.R;
	const int64_t total = (int64_t)mt.R * ld;
	const bool wide = (mt.m <= mt.n);
	T* g = G + mt.g_off;
	const T* a = A + mt.a_off;
	for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
	{
		const int i = (int)(e / ld), k = (int)(e % ld);
		T v;
		if (k < mt.C) { v = smul(sc, wide ? a[(int64_t)i * mt.n + k] : cj(a[(int64_t)k * mt.n + i])); }
		else { v = from_real<T>((k - mt.C) == i ? 1.0 : 0.0); }
		g[e] = v;
	}
}

/* one round of the tournament: warp w handles pair slot w of its matrix */
template <typename T>
__global__ void __launch_bounds__(256) svd_round_kernel(int nmat, const SvdMat* __restrict__ mats, int total_pairs, int round, double tol,
	T* __restrict__ G, int* __restrict__ rot_count, const int* __restrict__ done)
{
	const int gw = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
	const int lane = threadIdx.x & 31;
	if (gw >= total_pairs) { return; }
	/* matrix owning this pair slot */
	int lo = 0, hi = nmat - 1;
	while (lo < hi) {
		const int mid = (lo + hi + 1) >> 1;
		if (mats[mid].pair_begin <= gw) { lo = mid; } else { hi = mid - 1; }
	}
	if (done[lo]) { return; }
	const SvdMat mt = mats[lo];
	const int R = mt.R;
	if (R < 2) { return; }
	const int N = R + (R & 1);           /* players incl. a dummy for odd R */
	const int r = round % (N - 1);
	const int i = gw - mt.pair_begin;    /* 0 .. N/2 - 1 */
	int p, q;
	if (i == 0) { p = N - 1; q = r; }
	else { p = (r + i) % (N - 1); q = (r - i + (N - 1)) % (N - 1); }
	if (p >= R || q >= R) { return; }
	if (p > q) { const int t = p; p = q; q = t; }

	const int ld = mt.C + R;
	T* x = G + mt.g_off + (int64_t)p * ld;
	T* y = G + mt.g_off + (int64_t)q * ld;
	double alpha = 0, beta = 0;
	T gamma = from_real<T>(0.0);
	for (int k = lane; k < mt.C; k += 32) {
		const T a = x[k], b = y[k];
		alpha += abs2(a); beta += abs2(b);
		gamma = add(gamma, mul(a, cj(b)));
	}
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		alpha += __shfl_xor_sync(0xffffffffu, alpha, o);
		beta  += __shfl_xor_sync(0xffffffffu, beta, o);
		gamma = add(gamma, shfl_xor(gamma, o));
	}
	const double ag = sqrt(abs2(gamma));
	/* threshold of LAPACK's ?gesvj: sqrt(length) * eps, relative to the row norms */
	if (ag == 0.0 || ag <= tol * sqrt((double)mt.C) * sqrt(alpha * beta)) { return; }
	if (lane == 0) { atomicAdd(&rot_count[lo], 1); }
	const T ph = smul(1.0 / ag, gamma);
	const double zeta = (beta - alpha) / (2.0 * ag);
	const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
	const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
	for (int k = lane; k < ld; k += 32) {
		const T a = x[k], b = mul(ph, y[k]);
		x[k] = sub(smul(c, a), smul(s, b));
		y[k] = add(smul(s, a), smul(c, b));
	}
}

/* ---- blocks beyond shared memory: cooperative block Jacobi ----
 * The rows of [G | W] are cut into row blocks of b rows (b chosen so that two blocks fit into shared memory).
 * One persistent cooperative grid runs the whole iteration: per sweep, first every row block is orthogonalised
 * internally, then the block pairs of a tournament over the blocks are processed, one pair per CTA: the CTA pulls the
 * 2b rows from L2 into shared memory, rotates all b x b cross pairs there (one warp per pair, b pairs in flight,
 * cached squared norms) and writes the rows back; a grid barrier separates the rounds.  Compared with one pair per
 * launch this cuts both the number of grid-wide steps and the L2 traffic by the factor b. */
struct SvdBlkMat
{
	int64_t g_off;
	int32_t R, C, b, nb;       /* nb row blocks of b rows (the last may be shorter) */
	int32_t blk_begin;         /* first item of this matrix in the intra-block phase */
	int32_t pair_begin;        /* first item of this matrix in a tournament round (nbp / 2 items) */
};

/* one plane rotation of rows x, y (length ld, dot product over the first C entries); returns true if rotated */
template <typename T>
__device__ __forceinline__ bool jacobi_rotate_rows(T* x, T* y, double* sqx, double* sqy, int C, int ld, double thresh, int lane)
{
	T gamma = from_real<T>(0.0);
	for (int k = lane; k < C; k += 32) { gamma = add(gamma, mul(x[k], cj(y[k]))); }
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) { gamma = add(gamma, shfl_xor(gamma, o)); }
	const double alpha = *sqx, beta = *sqy;
	const double ag = sqrt(abs2(gamma));
	if (ag == 0.0 || ag <= thresh * sqrt(alpha * beta)) { return false; }
	const T ph = smul(1.0 / ag, gamma);
	const double zeta = (beta - alpha) / (2.0 * ag);
	const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
	const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
	for (int k = lane; k < ld; k += 32) {
		const T xa = x[k], yb = mul(ph, y[k]);
		x[k] = sub(smul(c, xa), smul(s, yb));
		y[k] = add(smul(s, xa), smul(c, yb));
	}
	__syncwarp();
	if (lane == 0) {
		*sqx = fmax(alpha - t * ag, 0.0);
		*sqy = beta + t * ag;
	}
	return true;
}

static constexpr int SVD_BLK_THREADS = 512;

template <typename T>
__global__ void __launch_bounds__(SVD_BLK_THREADS, 1) svd_block_kernel(int nmat, const SvdBlkMat* __restrict__ mats, int items_intra, int items_round, int max_rounds,
	double tol, int max_sweeps, T* __restrict__ G, int* __restrict__ rot /* [max_sweeps][nmat] */)
{
	cg::grid_group grid = cg::this_grid();
	extern __shared__ __align__(16) unsigned char svd_blk_raw[];
	__shared__ double sq[64];
	__shared__ int s_rot;
	T* rows = reinterpret_cast<T*>(svd_blk_raw);
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = SVD_BLK_THREADS / 32;

	auto find = [&](int item, bool intra) {
		int lo = 0, hi = nmat - 1;
		while (lo < hi) {
			const int mid = (lo + hi + 1) >> 1;
			const int beg = intra ? mats[mid].blk_begin : mats[mid].pair_begin;
			if (beg <= item) { lo = mid; } else { hi = mid - 1; }
		}
		return lo;
	};
	/* copy rows [r0, r0 + nr) of a matrix to / from shared memory slot 'slot0'; squared norms on load */
	auto load_rows = [&](const SvdBlkMat& mt, int r0, int nr, int slot0) {
		const int ld = mt.C + mt.R;
		const T* src = G + mt.g_off + (int64_t)r0 * ld;
		T* dst = rows + (size_t)slot0 * ld;
		for (int e = tid; e < nr * ld; e += SVD_BLK_THREADS) { dst[e] = __ldcg(src + e); }
	};
	auto store_rows = [&](const SvdBlkMat& mt, int r0, int nr, int slot0) {
		const int ld = mt.C + mt.R;
		T* dst = G + mt.g_off + (int64_t)r0 * ld;
		const T* src = rows + (size_t)slot0 * ld;
		for (int e = tid; e < nr * ld; e += SVD_BLK_THREADS) { __stcg(dst + e, src[e]); }
	};
	auto norms = [&](const SvdBlkMat& mt, int nr) {
		const int ld = mt.C + mt.R;
		for (int i = warp; i < nr; i += nwarp) {
			double s = 0;
			for (int k = lane; k < mt.C; k += 32) { s += abs2(rows[(size_t)i * ld + k]); }
			#pragma unroll
			for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); }
			if (lane == 0) { sq[i] = s; }
		}
	};

	for (int sweep = 0; sweep < max_sweeps; sweep++)
	{
		int* rot_now = rot + (size_t)sweep * nmat;
		const int* rot_prev = rot + (size_t)(sweep > 0 ? sweep - 1 : 0) * nmat;
		if (sweep > 0) {
			bool any = false;
			for (int mm = 0; mm < nmat; mm++) { any = any || (__ldcg(rot_prev + mm) != 0); }
			if (!any) { break; }     /* uniform over the grid: every CTA reads the same counters after the barrier */
		}

		/* ---- phase 0: pairs inside each row block ---- */
		for (int item = blockIdx.x; item < items_intra; item += gridDim.x)
		{
			const int mi = find(item, true);
			const SvdBlkMat mt = mats[mi];
			if (sweep > 0 && __ldcg(rot_prev + mi) == 0) { continue; }
			const int I = item - mt.blk_begin;
			const int r0 = I * mt.b, nr = min(mt.b, mt.R - r0);
			if (nr < 2) { continue; }
			const int ld = mt.C + mt.R;
			const double thresh = tol * sqrt((double)mt.C);
			if (tid == 0) { s_rot = 0; }
			load_rows(mt, r0, nr, 0);
			__syncthreads();
			norms(mt, nr);
			__syncthreads();
			const int N = nr + (nr & 1);
			for (int r = 0; r < N - 1; r++) {
				for (int i = warp; i < N / 2; i += nwarp) {
					int p, q;
					if (i == 0) { p = N - 1; q = r; }
					else { p = (r + i) % (N - 1); q = (r - i + (N - 1)) % (N - 1); }
					if (p >= nr || q >= nr) { continue; }
					if (p > q) { const int t = p; p = q; q = t; }
					if (jacobi_rotate_rows<T>(rows + (size_t)p * ld, rows + (size_t)q * ld, &sq[p], &sq[q], mt.C, ld, thresh, lane) && lane == 0) { atomicAdd(&s_rot, 1); }
				}
				__syncthreads();
			}
			store_rows(mt, r0, nr, 0);
			if (tid == 0 && s_rot > 0) { atomicAdd(&rot_now[mi], s_rot); }
			__syncthreads();
		}
		grid.sync();

		/* ---- tournament over the row blocks ---- */
		for (int round = 0; round < max_rounds; round++)
		{
			for (int item = blockIdx.x; item < items_round; item += gridDim.x)
			{
				const int mi = find(item, false);
				const SvdBlkMat mt = mats[mi];
				if (mt.nb < 2 || (sweep > 0 && __ldcg(rot_prev + mi) == 0)) { continue; }
				const int N = mt.nb + (mt.nb & 1);
				const int r = round % (N - 1);
				const int i = item - mt.pair_begin;
				int I, J;
				if (i == 0) { I = N - 1; J = r; }
				else { I = (r + i) % (N - 1); J = (r - i + (N - 1)) % (N - 1); }
				if (I >= mt.nb || J >= mt.nb) { continue; }
				if (I > J) { const int t = I; I = J; J = t; }
				const int ri = I * mt.b, ni = min(mt.b, mt.R - ri);
				const int rj = J * mt.b, nj = min(mt.b, mt.R - rj);
				const int ld = mt.C + mt.R;
				const double thresh = tol * sqrt((double)mt.C);
				if (tid == 0) { s_rot = 0; }
				load_rows(mt, ri, ni, 0);
				load_rows(mt, rj, nj, ni);
				__syncthreads();
				norms(mt, ni + nj);
				__syncthreads();
				const int bm = max(ni, nj);
				for (int rr = 0; rr < bm; rr++) {
					for (int x = warp; x < bm; x += nwarp) {
						const int y = (x + rr) % bm;
						if (x >= ni || y >= nj) { continue; }
						if (jacobi_rotate_rows<T>(rows + (size_t)x * ld, rows + (size_t)(ni + y) * ld, &sq[x], &sq[ni + y], mt.C, ld, thresh, lane) && lane == 0) { atomicAdd(&s_rot, 1); }
					}
					__syncthreads();
				}
				store_rows(mt, ri, ni, 0);
				store_rows(mt, rj, nj, ni);
				if (tid == 0 && s_rot > 0) { atomicAdd(&rot_now[mi], s_rot); }
				__syncthreads();
			}
			grid.sync();
		}
	}
}

/* after a window of rounds: a matrix without any rotation in a window that covered at least one full sweep is converged */
__global__ void svd_mark_kernel(int nmat, const SvdMat* __restrict__ mats, int window, int* rot_count, int* done, int* pending)
{
	const int b = blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nmat) { return; }
	if (!done[b]) {
		const int R = mats[b].R;
		const int N = R + (R & 1);
		if (R < 2 || (rot_count[b] == 0 && window >= N - 1)) { done[b] = 1; }
		else { atomicAdd(pending, 1); }
	}
	rot_count[b] = 0;
}

/* singular values (row norms, sorted descending) and vectors; one CTA per matrix */
template <typename T>
__global__ void __launch_bounds__(256) svd_finish_kernel(const SvdMat* __restrict__ mats, const double* __restrict__ scale, const T* __restrict__ G, double* __restrict__ sig_work, double* __restrict__ wn_work, int* __restrict__ ord_work,
	T* __restrict__ U, T* __restrict__ Vh, double* __restrict__ S)
{
	const SvdMat mt = mats[blockIdx.x];
	const double unscale = 1.0 / scale[blockIdx.x];
	const int R = mt.R, C = mt.C, ld = C + R;
	const T* g = G + mt.g_off;
	double* sig = sig_work + mt.s_off;
	double* wn = wn_work + mt.s_off;      /* 1 / norm of the accumulated-rotation rows: removes the O(#rotations) eps drift of their length */
	int* ord = ord_work + mt.s_off;
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
	for (int i = warp; i < R; i += nwarp) {
		double s = 0, w = 0;
		for (int k = lane; k < C; k += 32) { s += abs2(g[(int64_t)i * ld + k]); }
		for (int k = lane; k < R; k += 32) { w += abs2(g[(int64_t)i * ld + C + k]); }
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); w += __shfl_xor_sync(0xffffffffu, w, o); }
		if (lane == 0) { sig[i] = sqrt(s); wn[i] = (w > 0) ? 1.0 / sqrt(w) : 1.0; }
	}
	__syncthreads();
	/* rank sort, descending, ties by index */
	for (int i = threadIdx.x; i < R; i += blockDim.x) {
		const double si = sig[i];
		int rank = 0;
		for (int j = 0; j < R; j++) { const double sj = sig[j]; rank += (sj > si || (sj == si && j < i)) ? 1 : 0; }
		ord[rank] = i;
	}
	__syncthreads();
	const bool wide = (mt.m <= mt.n);
	const int m = mt.m, n = mt.n;
	T* u = U + mt.u_off;
	T* vh = Vh + mt.vh_off;
	for (int r = threadIdx.x; r < R; r += blockDim.x) { S[mt.s_off + r] = unscale * sig[ord[r]]; }
	/* Vh: R x n row-major, coalesced along the row */
	for (int64_t e = threadIdx.x; e < (int64_t)R * n; e += blockDim.x) {
		const int r = (int)(e / n), k = (int)(e % n);
		const int i = ord[r];
		if (wide) { const double s = sig[i]; vh[e] = smul(s > 0 ? 1.0 / s : 0.0, g[(int64_t)i * ld + k]); }
		else      { vh[e] = smul(wn[i], g[(int64_t)i * ld + C + k]); }
	}
	/* U: m x R row-major; U[k][r] */
	for (int64_t e = threadIdx.x; e < (int64_t)m * R; e += blockDim.x) {
		const int k = (int)(e / R), r = (int)(e % R);
		const int i = ord[r];
		if (wide) { u[e] = smul(wn[i], cj(g[(int64_t)i * ld + C + k])); }
		else      { const double s = sig[i]; u[e] = smul(s > 0 ? 1.0 / s : 0.0, cj(g[(int64_t)i * ld + k])); }
	}
}

/* ---- blocks that fit into shared memory: the whole Jacobi iteration in ONE CTA, no global round trips ----
 * [G | W] (R x (C + R)) lives in shared memory; warps take the R/2 disjoint row pairs of a tournament round, one block
 * barrier per round.  Squared row norms are cached and updated by the rotation formulas (recomputed once per sweep), so a
 * pair costs one dot product instead of three. */
static constexpr int SVD_SMEM_THREADS = 512;
static constexpr size_t SVD_SMEM_LIMIT = 220 * 1024;

template <typename T>
__global__ void __launch_bounds__(SVD_SMEM_THREADS) svd_smem_kernel(const SvdMat* __restrict__ mats, const int* __restrict__ sel, double tol, int max_sweeps,
	const T* __restrict__ A, T* __restrict__ U, T* __restrict__ Vh, double* __restrict__ S, int* __restrict__ status)
{
	extern __shared__ __align__(16) unsigned char svd_smem_raw[];
	const SvdMat mt = mats[sel[blockIdx.x]];
	const int R = mt.R, C = mt.C, ld = C + R, m = mt.m, n = mt.n;
	const bool wide = (m <= n);
	T* g = reinterpret_cast<T*>(svd_smem_raw);
	double* sq = reinterpret_cast<double*>(g + (size_t)R * ld);
	double* wn = sq + R;
	int* ord = reinterpret_cast<int*>(wn + R);
	__shared__ int s_rot;
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = SVD_SMEM_THREADS / 32;
	const T* a = A + mt.a_off;

	__shared__ double s_red[SVD_SMEM_THREADS / 32];
	double vmax = 0;
	for (int e = tid; e < m * n; e += SVD_SMEM_THREADS) { vmax = fmax(vmax, absmax_of(a[e])); }
	const double sc = pow2_scale(block_max(vmax, s_red));
	for (int e = tid; e < R * ld; e += SVD_SMEM_THREADS) {
		const int i = e / ld, k = e % ld;
		T v;
		if (k < C) { v = smul(sc, wide ? a[(int64_t)i * n + k] : cj(a[(int64_t)k * n + i])); }
		else { v = from_real<T>((k - C) == i ? 1.0 : 0.0); }
		g[e] = v;
	}
	__syncthreads();

	const int N = R + (R & 1);
	const double thresh = tol * sqrt((double)C);
	bool conv = (R < 2);
	for (int sweep = 0; sweep < max_sweeps && R >= 2; sweep++)
	{
		if (tid == 0) { s_rot = 0; }
		for (int i = warp; i < R; i += nwarp) {
			double s = 0;
			for (int k = lane; k < C; k += 32) { s += abs2(g[i * ld + k]); }
			#pragma unroll
			for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); }
			if (lane == 0) { sq[i] = s; }
		}
		__syncthreads();
		for (int r = 0; r < N - 1; r++)
		{
			for (int i = warp; i < N / 2; i += nwarp)
			{
				int p, q;
				if (i == 0) { p = N - 1; q = r; }
				else { p = (r + i) % (N - 1); q = (r - i + (N - 1)) % (N - 1); }
				if (p >= R || q >= R) { continue; }
				if (p > q) { const int t = p; p = q; q = t; }
				T* x = g + p * ld;
				T* y = g + q * ld;
				T gamma = from_real<T>(0.0);
				for (int k = lane; k < C; k += 32) { gamma = add(gamma, mul(x[k], cj(y[k]))); }
				#pragma unroll
				for (int o = 16; o > 0; o >>= 1) { gamma = add(gamma, shfl_xor(gamma, o)); }
				const double alpha = sq[p], beta = sq[q];
				const double ag = sqrt(abs2(gamma));
				if (ag == 0.0 || ag <= thresh * sqrt(alpha * beta)) { continue; }
				const T ph = smul(1.0 / ag, gamma);
				const double zeta = (beta - alpha) / (2.0 * ag);
				const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
				const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
				for (int k = lane; k < ld; k += 32) {
					const T xa = x[k], yb = mul(ph, y[k]);
					x[k] = sub(smul(c, xa), smul(s, yb));
					y[k] = add(smul(s, xa), smul(c, yb));
				}
				if (lane == 0) {
					sq[p] = fmax(alpha - t * ag, 0.0);
					sq[q] = beta + t * ag;
					atomicAdd(&s_rot, 1);
				}
			}
			__syncthreads();
		}
		const int nrot = s_rot;
		__syncthreads();
		if (nrot == 0) { conv = true; break; }
	}
	/* rotations still pending after the last sweep: the factors are not an SVD to working precision -- reported to the host, which
	 * returns < 0 like the reference does when LAPACK ?gesvd fails to converge (dense_tensor.c:3636-3671) */
	if (!conv && tid == 0) { atomicExch(status, 1); }

	/* singular values = exact final row norms, sorted descending (ties by index); the accumulated-rotation rows are re-normalised */
	for (int i = warp; i < R; i += nwarp) {
		double s = 0, w = 0;
		for (int k = lane; k < C; k += 32) { s += abs2(g[i * ld + k]); }
		for (int k = lane; k < R; k += 32) { w += abs2(g[i * ld + C + k]); }
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); w += __shfl_xor_sync(0xffffffffu, w, o); }
		if (lane == 0) { sq[i] = sqrt(s); wn[i] = (w > 0) ? 1.0 / sqrt(w) : 1.0; }
	}
	__syncthreads();
	for (int i = tid; i < R; i += SVD_SMEM_THREADS) {
		const double si = sq[i];
		int rank = 0;
		for (int j = 0; j < R; j++) { const double sj = sq[j]; rank += (sj > si || (sj == si && j < i)) ? 1 : 0; }
		ord[rank] = i;
	}
	__syncthreads();
	T* u = U + mt.u_off;
	T* vh = Vh + mt.vh_off;
	for (int r = tid; r < R; r += SVD_SMEM_THREADS) { S[mt.s_off + r] = sq[ord[r]] / sc; }
	for (int e = tid; e < R * n; e += SVD_SMEM_THREADS) {
		const int r = e / n, k = e % n;
		const int i = ord[r];
		if (wide) { const double s = sq[i]; vh[e] = smul(s > 0 ? 1.0 / s : 0.0, g[i * ld + k]); }
		else      { vh[e] = smul(wn[i], g[i * ld + C + k]); }
	}
	for (int e = tid; e < m * R; e += SVD_SMEM_THREADS) {
		const int k = e / R, r = e % R;
		const int i = ord[r];
		if (wide) { u[e] = smul(wn[i], cj(g[i * ld + C + k])); }
		else      { const double s = sq[i]; u[e] = smul(s > 0 ? 1.0 / s : 0.0, cj(g[i * ld + k])); }
	}
}

/* ---- blocks beyond shared memory: work matrices [G | W] in global memory, three phases ----
 * setup (scaled copy of every block into its work matrix), iterate (block-Jacobi tournament until no rotation is left),
 * finish (sort, normalise, write U, Vh, S).  The phases are separate so that a host-driven GEMM stage (ctbd_svdws_*) can work on
 * the same work matrices between setup and iterate. */
template <typename T>
struct SvdBig
{
	int nmat = 0;
	std::vector<SvdMat> mats;
	void *d_mats = nullptr, *d_G = nullptr, *d_int = nullptr, *d_sig = nullptr, *d_scale = nullptr;
	int64_t g_total = 0, smax = 0;
	int total_pairs = 0, Rmax = 0;
};

template <typename T>
static void svd_big_free(SvdBig<T>* st)
{
	if (st == nullptr) { return; }
	ctbd_free(st->d_scale); ctbd_free(st->d_sig); ctbd_free(st->d_int); ctbd_free(st->d_G); ctbd_free(st->d_mats);
	delete st;
}

template <typename T>
static SvdBig<T>* svd_big_setup(int nmat, const ctbd_mat_desc* descs, const void* A)
{
	SvdBig<T>* st = new SvdBig<T>();
	st->nmat = nmat;
	st->mats.resize(nmat);
	std::vector<SvdMat>& mats = st->mats;
	for (int b = 0; b < nmat; b++)
	{
		SvdMat& mt = mats[b];
		mt.a_off = descs[b].a_off; mt.u_off = descs[b].o0_off; mt.vh_off = descs[b].o1_off; mt.s_off = descs[b].s_off;
		mt.m = descs[b].m; mt.n = descs[b].n;
		mt.R = std::min(mt.m, mt.n); mt.C = std::max(mt.m, mt.n);
		mt.g_off = st->g_total; st->g_total += (int64_t)mt.R * (mt.C + mt.R);
		mt.pair_begin = st->total_pairs; mt.npair = (mt.R + 1) / 2; st->total_pairs += mt.npair;
		st->Rmax = std::max(st->Rmax, mt.R);
		st->smax = std::max(st->smax, mt.s_off + mt.R);
	}
	const int64_t smax = st->smax;
	bool ok = upload(mats.data(), (size_t)nmat * sizeof(SvdMat), &st->d_mats) == 0;
	ok = ok && ctbd_malloc(&st->d_G, (size_t)st->g_total * sizeof(T)) == 0;
	/* ints: rot_count[nmat], done[nmat], pending[1], ord[smax] */
	ok = ok && ctbd_malloc(&st->d_int, (size_t)(2 * nmat + 1 + smax) * sizeof(int)) == 0;
	ok = ok && ctbd_malloc(&st->d_sig, (size_t)2 * smax * sizeof(double)) == 0;
	ok = ok && ctbd_malloc(&st->d_scale, (size_t)nmat * sizeof(double)) == 0;
	int64_t maxel = 0;
	for (int b = 0; b < nmat; b++) { maxel = std::max(maxel, (int64_t)mats[b].R * (mats[b].C + mats[b].R)); }
	void *d_aoff = nullptr, *d_numel = nullptr;
	if (ok)
	{
		std::vector<int64_t> aoff(nmat), numel(nmat);
		for (int b = 0; b < nmat; b++) { aoff[b] = mats[b].a_off; numel[b] = (int64_t)mats[b].m * mats[b].n; }
		ok = ok && upload(aoff.data(), (size_t)nmat * sizeof(int64_t), &d_aoff) == 0;
		ok = ok && upload(numel.data(), (size_t)nmat * sizeof(int64_t), &d_numel) == 0;
	}
	if (ok)
	{
		absmax_scale_kernel<T><<<nmat, 256, 0, rt().stream>>>((const int64_t*)d_aoff, (const int64_t*)d_numel, (const T*)A, (double*)st->d_scale);
		rt().launches++;
		dim3 grid((unsigned)std::min<int64_t>(ceil_div(maxel, 256), 64), (unsigned)nmat);
		svd_init_kernel<T><<<grid, 256, 0, rt().stream>>>(nmat, (const SvdMat*)st->d_mats, (const double*)st->d_scale, (const T*)A, (T*)st->d_G);
		rt().launches++;
		if (cudaGetLastError() != cudaSuccess) { ok = false; }
	}
	ctbd_free(d_numel); ctbd_free(d_aoff);
	if (!ok) { svd_big_free<T>(st); return nullptr; }
	return st;
}

template <typename T>
static int svd_big_iterate(SvdBig<T>* st)
{
	const int nmat = st->nmat;
	const std::vector<SvdMat>& mats = st->mats;
	void* d_mats = st->d_mats; void* d_G = st->d_G;
	int* rot_count = (int*)st->d_int; int* done = rot_count + nmat; int* pending = done + nmat;
	const int Rmax = st->Rmax, total_pairs = st->total_pairs;
	int rc = 0;
	/* rows per block such that two row blocks fit into shared memory */
	const size_t blk_budget = 208 * 1024;
	bool use_block = (getenv("CTB_SVD_TOURNAMENT") == nullptr);
	std::vector<SvdBlkMat> bm(nmat);
	int items_intra = 0, items_round = 0, max_rounds = 0; size_t blk_smem = 0;
	/* rows per block: at most 32 (and never more than fit into shared memory).  Smaller blocks mean more row-block pairs per round
	 * (R / 2b per matrix -- with b = 32 a 500-row block keeps only 8 CTAs busy) at the price of more grid-wide rounds; measured on the
	 * Fermi-Hubbard L=32 D=1024 sweep: SVD phase 5.66 s with the cap of 32, 4.07 s with CTB_SVD_BLOCK_ROWS=8, same energies
	 * (profiles/r1_sweep_phases.json).  The default stays at 32 until the full GPU suite has run with the smaller cap.
	 * Tuning knob CTB_SVD_BLOCK_ROWS=<2..32>; "auto" picks the largest power of two that still gives every SM a pair. */
	int rows_cap = 32;
	{
		const char* env = getenv("CTB_SVD_BLOCK_ROWS");
		if (env != nullptr && strcmp(env, "auto") == 0) {
			rows_cap = 32;
			int64_t rows_total = 0;
			for (int b = 0; b < nmat; b++) { rows_total += mats[b].R; }
			while (rows_cap > 4 && rows_total / (2 * rows_cap) < rt().sm_count) { rows_cap /= 2; }
		}
		else if (env != nullptr && atoi(env) >= 2 && atoi(env) <= 32) { rows_cap = atoi(env); }
	}
	for (int b = 0; b < nmat && use_block; b++)
	{
		const size_t ld = (size_t)(mats[b].C + mats[b].R);
		int rows_fit = (int)(blk_budget / (2 * ld * sizeof(T)));
		if (rows_fit < 1) { use_block = false; break; }
		if (rows_fit > rows_cap) { rows_fit = rows_cap; }
		SvdBlkMat& x = bm[b];
		x.g_off = mats[b].g_off; x.R = mats[b].R; x.C = mats[b].C; x.b = rows_fit;
		x.nb = (int)ceil_div(x.R, x.b);
		x.blk_begin = items_intra; items_intra += x.nb;
		const int nbp = x.nb + (x.nb & 1);
		x.pair_begin = items_round; items_round += nbp / 2;
		max_rounds = std::max(max_rounds, x.nb >= 2 ? nbp - 1 : 0);
		blk_smem = std::max(blk_smem, (size_t)2 * x.b * ld * sizeof(T));
	}
	if (Rmax >= 2 && use_block)
	{
		const int max_sweeps = 40;
		void *d_bm = nullptr, *d_rot = nullptr;
		if (upload(bm.data(), (size_t)nmat * sizeof(SvdBlkMat), &d_bm) < 0) { return -1; }
		if (ctbd_malloc(&d_rot, (size_t)max_sweeps * nmat * sizeof(int)) < 0) { return -1; }
		static bool attr_done = false;
		if (!attr_done) {
			CTBD_CUDA(cudaFuncSetAttribute(svd_block_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(216 * 1024)));
			attr_done = true;
		}
		int occ = 0;
		CTBD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, svd_block_kernel<T>, SVD_BLK_THREADS, blk_smem));
		if (occ < 1) { occ = 1; }
		int grid = std::min(rt().sm_count * occ, std::max(items_intra, items_round));
		if (grid < 1) { grid = 1; }
		int a_nmat = nmat, a_mr = max_rounds, a_ms = max_sweeps; double a_tol = DBL_EPSILON;
		const SvdBlkMat* a_mats = (const SvdBlkMat*)d_bm; T* a_G = (T*)d_G; int* a_rot = (int*)d_rot;
		void* kargs[] = { &a_nmat, &a_mats, &items_intra, &items_round, &a_mr, &a_tol, &a_ms, &a_G, &a_rot };
		CTBD_CUDA(cudaLaunchCooperativeKernel((void*)svd_block_kernel<T>, dim3(grid), dim3(SVD_BLK_THREADS), kargs, blk_smem, rt().stream));
		rt().launches++;
		ctbd_free(d_rot); ctbd_free(d_bm);
	}
	else if (Rmax >= 2)
	{
		const int Nmax = Rmax + (Rmax & 1);
		const int window = Nmax - 1;            /* rounds per global sweep */
		const int max_sweeps = 40;
		const int blocks = (int)ceil_div((int64_t)total_pairs * 32, 256);
		int round = 0;
		for (int sweep = 0; sweep < max_sweeps; sweep++)
		{
			for (int r = 0; r < window; r++, round++) {
				svd_round_kernel<T><<<blocks, 256, 0, rt().stream>>>(nmat, (const SvdMat*)d_mats, total_pairs, round, DBL_EPSILON, (T*)d_G, rot_count, done);
				CTBD_LAUNCH_CHECK();
			}
			CTBD_CUDA(cudaMemsetAsync(pending, 0, sizeof(int), rt().stream));
			svd_mark_kernel<<<(nmat + 127) / 128, 128, 0, rt().stream>>>(nmat, (const SvdMat*)d_mats, window, rot_count, done, pending);
			CTBD_LAUNCH_CHECK();
			int h_pending = 0;
			if (ctbd_d2h(&h_pending, pending, sizeof(int)) < 0) { rc = -1; break; }
			if (h_pending == 0) { break; }
			if (sweep == max_sweeps - 1) { (void)fail_msg("batched SVD: rotations still pending after 40 tournament sweeps (result kept)"); }
		}
	}
	return rc;
}

template <typename T>
static int svd_big_finish(SvdBig<T>* st, void* U, void* Vh, double* S)
{
	const int nmat = st->nmat;
	int* ord = (int*)st->d_int + 2 * nmat + 1;
	svd_finish_kernel<T><<<nmat, 256, 0, rt().stream>>>((const SvdMat*)st->d_mats, (const double*)st->d_scale, (const T*)st->d_G, (double*)st->d_sig, (double*)st->d_sig + st->smax, ord, (T*)U, (T*)Vh, S);
	rt().launches++;
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { return fail("svd_finish_kernel", e, __FILE__, __LINE__); }
	return 0;
}

template <typename T>
static int svd_batched_impl(int nmat_all, const ctbd_mat_desc* descs_all, const void* A, void* U, void* Vh, double* S)
{
	if (nmat_all == 0) { return 0; }
	/* blocks whose [G | W] fits into shared memory take the single-CTA path, the rest the global tournament below */
	std::vector<SvdMat> small_mats;
	std::vector<ctbd_mat_desc> big;
	size_t smem_max = 0;
	void* d_status = nullptr;
	/* test knobs: CTB_SVD_NO_SMEM=1 / CTB_SVD_SMEM_LIMIT=<bytes> send smaller blocks down the big-block path */
	size_t smem_limit = SVD_SMEM_LIMIT;
	if (getenv("CTB_SVD_NO_SMEM") != nullptr) { smem_limit = 0; }
	else if (getenv("CTB_SVD_SMEM_LIMIT") != nullptr) { smem_limit = std::min((size_t)atoll(getenv("CTB_SVD_SMEM_LIMIT")), SVD_SMEM_LIMIT); }
	for (int b = 0; b < nmat_all; b++)
	{
		const int R = std::min(descs_all[b].m, descs_all[b].n), C = std::max(descs_all[b].m, descs_all[b].n);
		const size_t need = (size_t)R * (C + R) * sizeof(T) + (size_t)R * (2 * sizeof(double) + sizeof(int)) + 16;
		if (need <= smem_limit) {
			SvdMat mt; memset(&mt, 0, sizeof(mt));
			mt.a_off = descs_all[b].a_off; mt.u_off = descs_all[b].o0_off; mt.vh_off = descs_all[b].o1_off; mt.s_off = descs_all[b].s_off;
			mt.m = descs_all[b].m; mt.n = descs_all[b].n; mt.R = R; mt.C = C;
			small_mats.push_back(mt);
			smem_max = std::max(smem_max, need);
		}
		else { big.push_back(descs_all[b]); }
	}
	if (!small_mats.empty())
	{
		/* heaviest first */
		std::stable_sort(small_mats.begin(), small_mats.end(), [](const SvdMat& x, const SvdMat& y) {
			return (double)x.R * x.R * (x.C + x.R) > (double)y.R * y.R * (y.C + y.R); });
		std::vector<int> sel(small_mats.size());
		for (size_t i = 0; i < sel.size(); i++) { sel[i] = (int)i; }
		void *d_small = nullptr, *d_sel = nullptr;
		if (upload(small_mats.data(), small_mats.size() * sizeof(SvdMat), &d_small) < 0) { return -1; }
		if (upload(sel.data(), sel.size() * sizeof(int), &d_sel) < 0) { return -1; }
		static bool attr_done = false;
		if (!attr_done) {
			CTBD_CUDA(cudaFuncSetAttribute(svd_smem_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SVD_SMEM_LIMIT));
			attr_done = true;
		}
		if (ctbd_malloc(&d_status, sizeof(int)) < 0) { ctbd_free(d_sel); ctbd_free(d_small); return -1; }
		svd_smem_kernel<T><<<(int)small_mats.size(), SVD_SMEM_THREADS, smem_max, rt().stream>>>((const SvdMat*)d_small, (const int*)d_sel, DBL_EPSILON, 40,
			(const T*)A, (T*)U, (T*)Vh, S, (int*)d_status);
		CTBD_LAUNCH_CHECK();
		ctbd_free(d_sel); ctbd_free(d_small);
	}
	int rc = 0;
	if (!big.empty())
	{
		/* default: QR-preconditioned block Jacobi on the tensor pipe (ctbd_svd_bj.cu); CTB_SVD_BJ=0 keeps the scalar tournament below */
		const char* env = getenv("CTB_SVD_BJ");
		if (env == nullptr || atoi(env) != 0) { rc = svd_bj_impl<T>((int)big.size(), big.data(), A, U, Vh, S); }
		else {
			SvdBig<T>* st = svd_big_setup<T>((int)big.size(), big.data(), A);
			if (st == nullptr) { rc = -1; }
			else {
				rc = svd_big_iterate<T>(st);
				if (rc == 0) { rc = svd_big_finish<T>(st, U, Vh, S); }
				svd_big_free<T>(st);
			}
		}
	}
	if (d_status != nullptr)
	{
		/* convergence word of the single-CTA path (one 4-byte copy; the split that follows synchronises anyway) */
		int h_status = 0;
		if (ctbd_d2h(&h_status, d_status, sizeof(int)) < 0) { rc = -1; }
		ctbd_free(d_status);
		if (rc == 0 && h_status != 0)
		{
			/* Rotations were still pending after the last sweep.  This happens for numerically rank-deficient or very strongly graded
			 * blocks (a random start state, singular values 30 decades apart): the part of the factorisation that belongs to
			 * singular values above ~1e-13 of the largest one converges within a few sweeps, what keeps rotating are the rows at
			 * rounding level (NumPy model of this loop: 26 sweeps for R = 100 over 30 decades).  The factors reconstruct the block to
			 * working precision either way, so the call succeeds; the condition is recorded in ctbd_last_error() and reported once
			 * (the reference would return -1 only if LAPACK ?gesvd itself failed, dense_tensor.c:3636-3671). */
			static bool warned = false;
			(void)fail_msg("batched SVD: rotations at rounding level still pending after 40 sweeps of the single-CTA one-sided Jacobi (result kept)");
			if (!warned) { warned = true; fprintf(stderr, "chemtensor_b200: warning: %s\n", ctbd_last_error()); }
		}
	}
	return rc;
}

/* ============================================================================================== */
/* QR / RQ                                                                                          */
/* ============================================================================================== */

struct QrMat
{
	int64_t a_off, o0_off, o1_off;
	int64_t x_off;      /* work matrix X (rows x cols) */
	int64_t q_off;      /* work matrix Q (rows x k) */
	int32_t m, n, rows, cols, k;
};

static constexpr int QR_THREADS = 512;
static constexpr int QR_TX = 32;                     /* threads along a row (columns of the trailing matrix) */
static constexpr int QR_TY = QR_THREADS / QR_TX;     /* row groups */

/* Householder QR of X (rows x cols) in place, one CTA per matrix; Q (rows x k) formed explicitly.
 * Sign convention of LAPACK ?larfg: beta = -sign(Re x0) ||x||, H = I - tau v v^H, v_0 = 1. */
template <typename T>
__global__ void __launch_bounds__(QR_THREADS) qr_kernel(const QrMat* __restrict__ mats, int rq, const T* __restrict__ A, T* __restrict__ Xw, T* __restrict__ Qw, T* __restrict__ tauw,
	T* __restrict__ O0, T* __restrict__ O1)
{
	const QrMat mt = mats[blockIdx.x];
	const int rows = mt.rows, cols = mt.cols, k = mt.k, m = mt.m, n = mt.n;
	T* X = Xw + mt.x_off;
	T* Q = Qw + mt.q_off;
	T* tau = tauw + mt.q_off;   /* k <= rows*k entries available: reuse the offset space of Q's first row block in a separate buffer */
	const T* a = A + mt.a_off;
	const int tid = threadIdx.x;
	const int tx = tid % QR_TX, ty = tid / QR_TX;

	__shared__ double red[QR_THREADS / 32];
	__shared__ T sw[QR_TY][QR_TX + 1];
	__shared__ T s_tau;
	__shared__ double s_beta;

	/* load: X = A, or for RQ X = (E A E)^H, i.e. X[i][j] = conj(A[m-1-j][n-1-i]); normalised by a power of two */
	double vmax = 0;
	for (int64_t e = tid; e < (int64_t)rows * cols; e += QR_THREADS) { vmax = fmax(vmax, absmax_of(a[e])); }
	const double sc = pow2_scale(block_max(vmax, red));
	const double unsc = 1.0 / sc;
	for (int64_t e = tid; e < (int64_t)rows * cols; e += QR_THREADS) {
		const int i = (int)(e / cols), j = (int)(e % cols);
		X[e] = smul(sc, rq ? cj(a[(int64_t)(m - 1 - j) * n + (n - 1 - i)]) : a[e]);
	}
	__syncthreads();

	for (int j = 0; j < k; j++)
	{
		/* squared norm of X[j+1:, j] */
		double part = 0;
		for (int i = j + 1 + tid; i < rows; i += QR_THREADS) { part += abs2(X[(int64_t)i * cols + j]); }
		#pragma unroll
		for (int o = 16; o > 0; o >>= 1) { part += __shfl_xor_sync(0xffffffffu, part, o); }
		if ((tid & 31) == 0) { red[tid >> 5] = part; }
		__syncthreads();
		if (tid == 0)
		{
			double xn = 0;
			for (int w = 0; w < QR_THREADS / 32; w++) { xn += red[w]; }
			const T x0 = X[(int64_t)j * cols + j];
			if (xn == 0.0 && im(x0) == 0.0) {
				s_tau = from_real<T>(0.0);
				s_beta = re(x0);
			}
			else {
				const double nrm = sqrt(abs2(x0) + xn);
				const double beta = -(re(x0) >= 0.0 ? 1.0 : -1.0) * nrm;
				/* tau = (beta - x0) / beta */
				s_tau = smul(1.0 / beta, sub(from_real<T>(beta), x0));
				s_beta = beta;
			}
		}
		__syncthreads();
		const T tj = s_tau;
		if (tid == 0) { tau[j] = tj; }
		if (abs2(tj) != 0.0)
		{
			/* v = x / (x0 - beta) below the diagonal */
			const T x0 = X[(int64_t)j * cols + j];
			const T d = sub(x0, from_real<T>(s_beta));
			const double dmx = absmax_of(d);
			const T ds = smul(1.0 / dmx, d);
			const T scal = smul(1.0 / (abs2(ds) * dmx), cj(ds));     /* 1 / (x0 - beta), safe against underflow of |d|^2 */
			__syncthreads();
			for (int i = j + 1 + tid; i < rows; i += QR_THREADS) { X[(int64_t)i * cols + j] = mul(X[(int64_t)i * cols + j], scal); }
			if (tid == 0) { X[(int64_t)j * cols + j] = from_real<T>(s_beta); }
			__syncthreads();
			/* trailing update: X[j:, c] -= v * (conj(tau) * v^H X[j:, c]) for c > j, QR_TX columns at a time */
			const T ctj = cj(tj);
			for (int c0 = j + 1; c0 < cols; c0 += QR_TX)
			{
				const int c = c0 + tx;
				T acc = from_real<T>(0.0);
				if (c < cols) {
					for (int i = j + ty; i < rows; i += QR_TY) {
						const T v = (i == j) ? from_real<T>(1.0) : X[(int64_t)i * cols + j];
						acc = add(acc, mul(cj(v), X[(int64_t)i * cols + c]));
					}
				}
				sw[ty][tx] = acc;
				__syncthreads();
				if (ty == 0) {
					T s = sw[0][tx];
					for (int r = 1; r < QR_TY; r++) { s = add(s, sw[r][tx]); }
					sw[0][tx] = mul(ctj, s);
				}
				__syncthreads();
				if (c < cols) {
					const T s = sw[0][tx];
					for (int i = j + ty; i < rows; i += QR_TY) {
						const T v = (i == j) ? from_real<T>(1.0) : X[(int64_t)i * cols + j];
						X[(int64_t)i * cols + c] = sub(X[(int64_t)i * cols + c], mul(v, s));
					}
				}
				__syncthreads();
			}
		}
		else {
			if (tid == 0) { X[(int64_t)j * cols + j] = from_real<T>(s_beta); }
			__syncthreads();
		}
	}

	/* Q = H_0 H_1 ... H_{k-1} [I; 0] (rows x k), backward accumulation */
	for (int64_t e = tid; e < (int64_t)rows * k; e += QR_THREADS) {
		const int i = (int)(e / k), c = (int)(e % k);
		Q[e] = from_real<T>(i == c ? 1.0 : 0.0);
	}
	__syncthreads();
	for (int j = k - 1; j >= 0; j--)
	{
		const T tj = tau[j];
		if (abs2(tj) == 0.0) { continue; }
		for (int c0 = j; c0 < k; c0 += QR_TX)
		{
			const int c = c0 + tx;
			T acc = from_real<T>(0.0);
			if (c < k) {
				for (int i = j + ty; i < rows; i += QR_TY) {
					const T v = (i == j) ? from_real<T>(1.0) : X[(int64_t)i * cols + j];
					acc = add(acc, mul(cj(v), Q[(int64_t)i * k + c]));
				}
			}
			sw[ty][tx] = acc;
			__syncthreads();
			if (ty == 0) {
				T s = sw[0][tx];
				for (int r = 1; r < QR_TY; r++) { s = add(s, sw[r][tx]); }
				sw[0][tx] = mul(tj, s);
			}
			__syncthreads();
			if (c < k) {
				const T s = sw[0][tx];
				for (int i = j + ty; i < rows; i += QR_TY) {
					const T v = (i == j) ? from_real<T>(1.0) : X[(int64_t)i * cols + j];
					Q[(int64_t)i * k + c] = sub(Q[(int64_t)i * k + c], mul(v, s));
				}
			}
			__syncthreads();
		}
	}

	/* outputs */
	T* o0 = O0 + mt.o0_off;
	T* o1 = O1 + mt.o1_off;
	if (!rq)
	{
		for (int64_t e = tid; e < (int64_t)m * k; e += QR_THREADS) { o0[e] = Q[e]; }
		for (int64_t e = tid; e < (int64_t)k * n; e += QR_THREADS) {
			const int i = (int)(e / n), c = (int)(e % n);
			o1[e] = (c >= i) ? smul(unsc, X[(int64_t)i * cols + c]) : from_real<T>(0.0);
		}
	}
	else
	{
		/* R[i][j] = conj(Rt[k-1-j][m-1-i]) (m x k),  Q[i][j] = conj(Qt[n-1-j][k-1-i]) (k x n) */
		for (int64_t e = tid; e < (int64_t)m * k; e += QR_THREADS) {
			const int i = (int)(e / k), j = (int)(e % k);
			const int ri = k - 1 - j, rc = m - 1 - i;
			o0[e] = (rc >= ri) ? smul(unsc, cj(X[(int64_t)ri * cols + rc])) : from_real<T>(0.0);
		}
		for (int64_t e = tid; e < (int64_t)k * n; e += QR_THREADS) {
			const int i = (int)(e / n), j = (int)(e % n);
			o1[e] = cj(Q[(int64_t)(n - 1 - j) * k + (k - 1 - i)]);
		}
	}
}

template <typename T>
static int qr_batched_impl(int rq, int nmat, const ctbd_mat_desc* descs, const void* A, void* O0, void* O1)
{
	if (nmat == 0) { return 0; }
	std::vector<QrMat> mats(nmat);
	int64_t x_total = 0, q_total = 0;
	for (int b = 0; b < nmat; b++)
	{
		QrMat& mt = mats[b];
		mt.a_off = descs[b].a_off; mt.o0_off = descs[b].o0_off; mt.o1_off = descs[b].o1_off;
		mt.m = descs[b].m; mt.n = descs[b].n;
		mt.k = std::min(mt.m, mt.n);
		mt.rows = rq ? mt.n : mt.m; mt.cols = rq ? mt.m : mt.n;
		mt.x_off = x_total; x_total += (int64_t)mt.rows * mt.cols;
		mt.q_off = q_total; q_total += (int64_t)mt.rows * mt.k;
	}
	void *d_mats = nullptr, *d_X = nullptr, *d_Q = nullptr, *d_tau = nullptr;
	if (upload(mats.data(), (size_t)nmat * sizeof(QrMat), &d_mats) < 0) { return -1; }
	if (ctbd_malloc(&d_X, (size_t)x_total * sizeof(T)) < 0) { return -1; }
	if (ctbd_malloc(&d_Q, (size_t)q_total * sizeof(T)) < 0) { return -1; }
	if (ctbd_malloc(&d_tau, (size_t)q_total * sizeof(T)) < 0) { return -1; }
	qr_kernel<T><<<nmat, QR_THREADS, 0, rt().stream>>>((const QrMat*)d_mats, rq, (const T*)A, (T*)d_X, (T*)d_Q, (T*)d_tau, (T*)O0, (T*)O1);
	rt().launches++;
	int rc = 0;
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { rc = fail("qr_kernel", e, __FILE__, __LINE__); }
	ctbd_free(d_tau); ctbd_free(d_Q); ctbd_free(d_X); ctbd_free(d_mats);
	return rc;
}

/* ---- convergence measure of the GEMM-driven block-Jacobi stage (ctbd_svdws_*): over a list of small Gram matrices ----
 * out[1] = max(out[1], max_i G_ii);  out[0] = max(out[0], max_{i != j} |G_ij|^2 / max(G_ii G_jj, (floor_rel out[1])^2)).
 * Non-negative doubles order like their bit patterns, so the running maxima are atomicMax on the bits. */
__device__ __forceinline__ void atomic_max_nonneg(double* addr, double v)
{
	atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(v));
}

template <typename T>
__global__ void __launch_bounds__(256) gram_diagmax_kernel(const int64_t* __restrict__ off, const int32_t* __restrict__ dim, const T* __restrict__ G, double* __restrict__ out)
{
	__shared__ double red[8];
	const T* g = G + off[blockIdx.x];
	const int n = dim[blockIdx.x];
	double v = 0;
	for (int i = threadIdx.x; i < n; i += blockDim.x) { v = fmax(v, sqrt(abs2(g[(int64_t)i * n + i]))); }
	const double m = block_max(v, red);
	if (threadIdx.x == 0) { atomic_max_nonneg(out + 1, m); }
}

template <typename T>
__global__ void __launch_bounds__(256) gram_offdiag_kernel(const int64_t* __restrict__ off, const int32_t* __restrict__ dim, const T* __restrict__ G, double floor_rel, double* __restrict__ out)
{
	__shared__ double red[8];
	const T* g = G + off[blockIdx.x];
	const int n = dim[blockIdx.x];
	const double fl = floor_rel * out[1];
	const double fl2 = fl * fl;
	double v = 0;
	for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
		const int i = e / n, j = e % n;
		if (i == j) { continue; }
		const double dii = sqrt(abs2(g[(int64_t)i * n + i])), djj = sqrt(abs2(g[(int64_t)j * n + j]));
		const double den = fmax(dii * djj, fl2);
		if (den > 0) { v = fmax(v, abs2(g[e]) / den); }
	}
	const double m = block_max(v, red);
	if (threadIdx.x == 0) { atomic_max_nonneg(out, m); }
}

template <typename T>
static int svdws_finish_impl(SvdBig<T>* st, const void* G_cur, int polish, void* U, void* Vh, double* S)
{
	int rc = 0;
	if (G_cur != nullptr && G_cur != st->d_G) {
		cudaError_t e = cudaMemcpyAsync(st->d_G, G_cur, (size_t)st->g_total * sizeof(T), cudaMemcpyDeviceToDevice, rt().stream);
		if (e != cudaSuccess) { rc = fail("cudaMemcpyAsync (SVD workspace)", e, __FILE__, __LINE__); }
	}
	if (rc == 0 && polish) { rc = svd_big_iterate<T>(st); }
	if (rc == 0) { rc = svd_big_finish<T>(st, U, Vh, S); }
	svd_big_free<T>(st);
	return rc;
}

struct SvdWs { int dtype; void* st; };

} // namespace ctbd

using namespace ctbd;

extern "C" {

int ctbd_svd_batched(int dtype, int nmat, const struct ctbd_mat_desc* descs_host, const void* A, void* U, void* Vh, double* S_dev)
{
	CTBD_REQUIRE_INIT();
	if (dtype == CTBD_F64)  { return svd_batched_impl<double>(nmat, descs_host, A, U, Vh, S_dev); }
	if (dtype == CTBD_C128) { return svd_batched_impl<double2>(nmat, descs_host, A, U, Vh, S_dev); }
	return fail_msg("batched SVD: unsupported dtype");
}

int ctbd_qr_batched(int dtype, int rq, int nmat, const struct ctbd_mat_desc* descs_host, const void* A, void* O0, void* O1)
{
	CTBD_REQUIRE_INIT();
	if (dtype == CTBD_F64)  { return qr_batched_impl<double>(rq, nmat, descs_host, A, O0, O1); }
	if (dtype == CTBD_C128) { return qr_batched_impl<double2>(rq, nmat, descs_host, A, O0, O1); }
	return fail_msg("batched QR: unsupported dtype");
}

/* ---- work matrices of the big-block SVD exposed to a host-driven stage (see include/ctb_device.h) ---- */

int ctbd_svdws_create(int dtype, int nmat, const struct ctbd_mat_desc* descs_host, const void* A, void** ws, void** G, int64_t* g_total)
{
	CTBD_REQUIRE_INIT();
	if (nmat <= 0) { return fail_msg("SVD workspace: no blocks"); }
	SvdWs* w = new SvdWs();
	w->dtype = dtype; w->st = nullptr;
	if (dtype == CTBD_F64) {
		SvdBig<double>* st = svd_big_setup<double>(nmat, descs_host, A);
		if (st != nullptr) { w->st = st; *G = st->d_G; *g_total = st->g_total; }
	}
	else if (dtype == CTBD_C128) {
		SvdBig<double2>* st = svd_big_setup<double2>(nmat, descs_host, A);
		if (st != nullptr) { w->st = st; *G = st->d_G; *g_total = st->g_total; }
	}
	if (w->st == nullptr) { delete w; return fail_msg("SVD workspace: setup failed"); }
	*ws = w;
	return 0;
}

int ctbd_svdws_finish(void* ws, const void* G_cur, int polish, void* U, void* Vh, double* S_dev)
{
	SvdWs* w = (SvdWs*)ws;
	if (w == nullptr) { return 0; }
	int rc;
	if (w->dtype == CTBD_F64) { rc = svdws_finish_impl<double>((SvdBig<double>*)w->st, G_cur, polish, U, Vh, S_dev); }
	else                      { rc = svdws_finish_impl<double2>((SvdBig<double2>*)w->st, G_cur, polish, U, Vh, S_dev); }
	delete w;
	return rc;
}

int ctbd_gram_offdiag(int dtype, int ngram, const int64_t* off_dev, const int32_t* dim_dev, const void* G, double floor_rel, double* out_dev)
{
	CTBD_REQUIRE_INIT();
	if (ngram <= 0) { return 0; }
	if (dtype == CTBD_F64) {
		gram_diagmax_kernel<double><<<ngram, 256, 0, rt().stream>>>(off_dev, dim_dev, (const double*)G, out_dev);
		gram_offdiag_kernel<double><<<ngram, 256, 0, rt().stream>>>(off_dev, dim_dev, (const double*)G, floor_rel, out_dev);
	}
	else if (dtype == CTBD_C128) {
		gram_diagmax_kernel<double2><<<ngram, 256, 0, rt().stream>>>(off_dev, dim_dev, (const double2*)G, out_dev);
		gram_offdiag_kernel<double2><<<ngram, 256, 0, rt().stream>>>(off_dev, dim_dev, (const double2*)G, floor_rel, out_dev);
	}
	else { return fail_msg("gram_offdiag: unsupported dtype"); }
	rt().launches += 2;
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) { return fail("gram_offdiag", e, __FILE__, __LINE__); }
	return 0;
}

} // extern "C"
