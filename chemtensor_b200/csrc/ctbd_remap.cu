/*
 * ctbd_remap.cu -- logical-index remaps between packed block-sparse layouts resident on the device.
 *
 * One gather kernel covers the structural primitives of the reference that move entries between
 * block structures (src/tensor/block_sparse_tensor.c): _transpose :785 (incl. the conjugating
 * variant :880), _flatten_axes :950, _split_axis :1123, _slice :1446 and
 * _multiply_pointwise_vector :1654.  Every destination entry computes its logical multi-index,
 * maps it to the source logical multi-index of the operation and reads the source entry through
 * the source sector tables; writes are fully coalesced (destination order), reads are coalesced
 * along the innermost source run.  HBM-bound: 2 x sizeof(T) bytes per stored entry.
 */
#include <vector>
#include <algorithm>
#include <stdlib.h>
#include "ctbd_common.cuh"

namespace ctbd {

struct LayoutDev
{
	int ndim, dtype;
	int64_t dim[CTBD_MAXDIM];
	int nsec[CTBD_MAXDIM];
	const int32_t* sec_of[CTBD_MAXDIM];
	const int32_t* pos_of[CTBD_MAXDIM];
	const int32_t* secstart[CTBD_MAXDIM];
	const int32_t* log_of[CTBD_MAXDIM];
	int64_t ngrid;
	const int64_t* grid_off;
	int nblk;
	const int64_t* blk_grid;
	const int64_t* blk_off;
	int64_t nstore;
};

struct Layout
{
	LayoutDev d;
	void* arena;    /* one device allocation holding all tables */
	int64_t max_block;   /* entries of the largest stored block, largest logical dimension: the fast kernel works in 32 bits */
	int64_t max_dim;
};

struct RemapParams
{
	int op, i_ax, conj, scale_ax;
	int perm[CTBD_MAXDIM];
	const int64_t* ind;
	const double* scale;
};

template <typename T> __device__ __forceinline__ T conj_if(T v, int) { return v; }
template <> __device__ __forceinline__ double2 conj_if<double2>(double2 v, int c) { if (c) { v.y = -v.y; } return v; }
__device__ __forceinline__ double  scale_by(double v, double s)  { return v * s; }
__device__ __forceinline__ double2 scale_by(double2 v, double s) { return make_double2(v.x * s, v.y * s); }
template <typename T> __device__ __forceinline__ T zero_of();
template <> __device__ __forceinline__ double zero_of<double>() { return 0.0; }
template <> __device__ __forceinline__ double2 zero_of<double2>() { return make_double2(0.0, 0.0); }

template <typename T>
__global__ void __launch_bounds__(256) remap_kernel(const LayoutDev D, const LayoutDev S, const RemapParams p, T* __restrict__ dst, const T* __restrict__ src)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < D.nstore; e += stride)
	{
		/* stored block containing entry e */
		int lo = 0, hi = D.nblk - 1;
		while (lo < hi) {
			const int mid = (lo + hi + 1) >> 1;
			if (D.blk_off[mid] <= e) { lo = mid; } else { hi = mid - 1; }
		}
		int64_t cell = D.blk_grid[lo];
		int64_t r = e - D.blk_off[lo];
		int sec[CTBD_MAXDIM];
		#pragma unroll
		for (int i = CTBD_MAXDIM - 1; i >= 0; i--) {
			if (i < D.ndim) { sec[i] = (int)(cell % D.nsec[i]); cell /= D.nsec[i]; }
		}
		int64_t ld[CTBD_MAXDIM], ls[CTBD_MAXDIM];
		#pragma unroll
		for (int i = CTBD_MAXDIM - 1; i >= 0; i--) {
			if (i < D.ndim) {
				const int s0 = D.secstart[i][sec[i]];
				const int bd = D.secstart[i][sec[i] + 1] - s0;
				const int pos = (int)(r % bd); r /= bd;
				ld[i] = D.log_of[i][s0 + pos];
			}
		}
		switch (p.op)
		{
			case CTBD_REMAP_TRANSPOSE:
				#pragma unroll
				for (int i = 0; i < CTBD_MAXDIM; i++) { if (i < D.ndim) { ls[p.perm[i]] = ld[i]; } }
				break;
			case CTBD_REMAP_FLATTEN:
				#pragma unroll
				for (int i = 0; i < CTBD_MAXDIM; i++) {
					if (i < D.ndim) {
						if (i < p.i_ax) { ls[i] = ld[i]; }
						else if (i == p.i_ax) { ls[i] = ld[i] / S.dim[i + 1]; ls[i + 1] = ld[i] % S.dim[i + 1]; }
						else { ls[i + 1] = ld[i]; }
					}
				}
				break;
			case CTBD_REMAP_SPLIT:
				#pragma unroll
				for (int i = 0; i < CTBD_MAXDIM; i++) {
					if (i < D.ndim) {
						if (i < p.i_ax) { ls[i] = ld[i]; }
						else if (i == p.i_ax) { ls[i] = ld[i] * D.dim[i + 1]; }
						else if (i == p.i_ax + 1) { ls[i - 1] += ld[i]; }
						else { ls[i - 1] = ld[i]; }
					}
				}
				break;
			case CTBD_REMAP_SLICE:
				#pragma unroll
				for (int i = 0; i < CTBD_MAXDIM; i++) { if (i < D.ndim) { ls[i] = (i == p.i_ax) ? p.ind[ld[i]] : ld[i]; } }
				break;
			default:
				#pragma unroll
				for (int i = 0; i < CTBD_MAXDIM; i++) { if (i < D.ndim) { ls[i] = ld[i]; } }
				break;
		}
		int64_t scell = 0, soff = 0;
		#pragma unroll
		for (int i = 0; i < CTBD_MAXDIM; i++) {
			if (i < S.ndim) {
				const int s = S.sec_of[i][ls[i]];
				scell = scell * S.nsec[i] + s;
				soff = soff * (S.secstart[i][s + 1] - S.secstart[i][s]) + S.pos_of[i][ls[i]];
			}
		}
		const int64_t sbase = S.grid_off[scell];
		T v = (sbase >= 0) ? src[sbase + soff] : zero_of<T>();
		v = conj_if<T>(v, p.conj);
		if (p.scale_ax >= 0) { v = scale_by(v, p.scale[ld[p.scale_ax]]); }
		dst[e] = v;
	}
}

template <int ND> __device__ __forceinline__ int pick(const int (&a)[ND], int idx)
{
	int v = a[0];
	#pragma unroll
	for (int i = 1; i < ND; i++) { if (i == idx) { v = a[i]; } }
	return v;
}

/* Fast gather form of the same remap (round 2): the per-element work of remap_kernel above (binary search over the block list, 64-bit
 * div/mod chains, ~5 dependent table loads per axis) is amortised.
 *   - one block search per WARP chunk of 32 x K consecutive destination entries (uniform loads), afterwards the block index only moves
 *     forward; all in-block arithmetic is 32-bit (the host checks that no block exceeds 2^31 entries, else the kernel above runs);
 *   - the decode of the outer destination axes is cached per destination ROW (fixed outer positions, innermost axis running);
 *   - the sector tables of a source axis are consulted only when the logical index on that axis changed, the block-grid lookup only
 *     when the source cell changed: along a row that is one load of log_of, one of sec_of / pos_of and the entry itself;
 *   - writes are coalesced (lane = consecutive destination entry), reads along every innermost source run; the K loads of a thread
 *     are issued back to back before the K stores.
 * ND = compile-time bound on the number of axes of both layouts (4 or 8): keeps the per-axis caches in registers. */
template <typename T, int ND, int K>
__global__ void __launch_bounds__(256) remap_fast_kernel(const LayoutDev D, const LayoutDev S, const RemapParams p, T* __restrict__ dst, const T* __restrict__ src)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	constexpr int64_t CH = 256 * K;
	const int dn = D.ndim, sn = S.ndim;
	const int last = dn - 1;
	for (int64_t cbase = (int64_t)blockIdx.x * CH; cbase < D.nstore; cbase += (int64_t)gridDim.x * CH)
	{
		const int64_t wbase = cbase + (int64_t)warp * (32 * K);
		if (wbase >= D.nstore) { continue; }
		int b;
		{
			int lo = 0, hi = D.nblk - 1;
			while (lo < hi) {
				const int mid = (lo + hi + 1) >> 1;
				if (D.blk_off[mid] <= wbase) { lo = mid; } else { hi = mid - 1; }
			}
			b = lo;
		}
		int cur_b = -1;
		int64_t boff = 0, bend = D.blk_off[b + 1];
		int d_s0[ND], d_bd[ND], d_ld[ND];      /* destination block: first slot of the sector in log_of, block extent; logical index of the row */
		uint32_t cur_row = 0xFFFFFFFFu;
		int64_t ind_c = 0;                       /* SLICE: ind[ld[i_ax]] of the row (or of the entry when i_ax is the innermost axis) */
		int s_ls[ND], s_sec[ND], s_pos[ND], s_bd[ND];
		#pragma unroll
		for (int j = 0; j < ND; j++) { s_ls[j] = -1; s_sec[j] = -1; s_pos[j] = 0; s_bd[j] = 1; d_s0[j] = 0; d_bd[j] = 1; d_ld[j] = 0; }
		int64_t cur_cell = -1, sbase = -1;
		int64_t addr[K];
		int sidx[K];
		#pragma unroll
		for (int k = 0; k < K; k++)
		{
			const int64_t e = wbase + k * 32 + lane;
			addr[k] = -2;
			sidx[k] = 0;
			if (e >= D.nstore) { continue; }
			while (e >= bend) { b++; bend = D.blk_off[b + 1]; }
			if (b != cur_b)
			{
				cur_b = b;
				boff = D.blk_off[b];
				int64_t cell = D.blk_grid[b];
				#pragma unroll
				for (int i = ND - 1; i >= 0; i--) {
					if (i < dn) {
						const int ns = D.nsec[i];
						const int sc = (int)(cell % ns); cell /= ns;
						const int s0 = D.secstart[i][sc];
						d_s0[i] = s0; d_bd[i] = D.secstart[i][sc + 1] - s0;
					}
				}
				cur_row = 0xFFFFFFFFu;
			}
			const uint32_t r = (uint32_t)(e - boff);
			const uint32_t bl = (uint32_t)pick<ND>(d_bd, last);
			const uint32_t row = r / bl;
			const int pos = (int)(r - row * bl);
			if (row != cur_row)
			{
				cur_row = row;
				uint32_t q = row;
				#pragma unroll
				for (int i = ND - 2; i >= 0; i--) {
					if (i < last) {
						const uint32_t bd = (uint32_t)d_bd[i];
						const uint32_t qq = q / bd;
						d_ld[i] = D.log_of[i][d_s0[i] + (int)(q - qq * bd)];
						q = qq;
					}
				}
				if (p.op == CTBD_REMAP_SLICE && p.i_ax != last) { ind_c = p.ind[pick<ND>(d_ld, p.i_ax)]; }
			}
			const int ll = D.log_of[last][pick<ND>(d_s0, last) + pos];
			#pragma unroll
			for (int i = 0; i < ND; i++) { if (i == last) { d_ld[i] = ll; } }
			if (p.op == CTBD_REMAP_SLICE && p.i_ax == last) { ind_c = p.ind[ll]; }
			/* source logical multi-index */
			int ls[ND];
			#pragma unroll
			for (int j = 0; j < ND; j++) { ls[j] = 0; }
			switch (p.op)
			{
				case CTBD_REMAP_TRANSPOSE:
					#pragma unroll
					for (int i = 0; i < ND; i++) { if (i < dn) {
						#pragma unroll
						for (int j = 0; j < ND; j++) { if (p.perm[i] == j) { ls[j] = d_ld[i]; } }
					} }
					break;
				case CTBD_REMAP_FLATTEN:
				{
					const int sd1 = (int)S.dim[p.i_ax + 1];
					#pragma unroll
					for (int i = 0; i < ND; i++) { if (i < dn) {
						if (i < p.i_ax) { ls[i] = d_ld[i]; }
						else if (i == p.i_ax) {
							const int hi = d_ld[i] / sd1;
							#pragma unroll
							for (int j = 0; j < ND - 1; j++) { if (j == i) { ls[j] = hi; ls[j + 1] = d_ld[i] - hi * sd1; } }
						}
						else { if (i + 1 < ND) { ls[i + 1] = d_ld[i]; } }
					} }
					break;
				}
				case CTBD_REMAP_SPLIT:
				{
					const int dd1 = (int)D.dim[p.i_ax + 1];
					#pragma unroll
					for (int i = 0; i < ND; i++) { if (i < dn) {
						if (i < p.i_ax) { ls[i] = d_ld[i]; }
						else if (i == p.i_ax) { ls[i] = d_ld[i] * dd1; }
						else if (i == p.i_ax + 1) { if (i >= 1) { ls[i - 1] += d_ld[i]; } }
						else { if (i >= 1) { ls[i - 1] = d_ld[i]; } }
					} }
					break;
				}
				case CTBD_REMAP_SLICE:
					#pragma unroll
					for (int i = 0; i < ND; i++) { if (i < dn) { ls[i] = (i == p.i_ax) ? (int)ind_c : d_ld[i]; } }
					break;
				default:
					#pragma unroll
					for (int i = 0; i < ND; i++) { if (i < dn) { ls[i] = d_ld[i]; } }
					break;
			}
			int64_t scell = 0;
			uint32_t soff = 0;
			#pragma unroll
			for (int j = 0; j < ND; j++) {
				if (j < sn) {
					if (ls[j] != s_ls[j]) {
						s_ls[j] = ls[j];
						const int sc = S.sec_of[j][ls[j]];
						s_pos[j] = S.pos_of[j][ls[j]];
						if (sc != s_sec[j]) { s_bd[j] = S.secstart[j][sc + 1] - S.secstart[j][sc]; }
						s_sec[j] = sc;
					}
					scell = scell * S.nsec[j] + s_sec[j];
					soff = soff * (uint32_t)s_bd[j] + (uint32_t)s_pos[j];
				}
			}
			if (scell != cur_cell) { cur_cell = scell; sbase = S.grid_off[scell]; }
			addr[k] = (sbase >= 0) ? sbase + (int64_t)soff : -1;
			if (p.scale_ax >= 0) { sidx[k] = pick<ND>(d_ld, p.scale_ax); }
		}
		T vals[K];
		double scl[K];
		#pragma unroll
		for (int k = 0; k < K; k++) {
			vals[k] = (addr[k] >= 0) ? src[addr[k]] : zero_of<T>();
			scl[k] = (p.scale_ax >= 0 && addr[k] != -2) ? p.scale[sidx[k]] : 1.0;
		}
		#pragma unroll
		for (int k = 0; k < K; k++) {
			if (addr[k] != -2) {
				T v = conj_if<T>(vals[k], p.conj);
				if (p.scale_ax >= 0) { v = scale_by(v, scl[k]); }
				dst[wbase + k * 32 + lane] = v;
			}
		}
	}
}

/* ---- run form of the gather (round 2, second step) -------------------------------------------------------------------------------
 * ncu of remap_fast_kernel above: 340 warp instructions per 32 entries, issue-bound at 0.7 - 0.9 TB/s.  Almost all entries of the
 * tensors of a sweep lie in long runs: consecutive destination entries of one destination row whose source entries are equally
 * spaced in ONE source block.  This kernel finds the runs instead of decoding every entry:
 *   - a warp owns 1024 consecutive destination entries = 16 sub-chunks of 64; lane l decodes ONE probe entry -- the first (l even) or
 *     last (l odd) entry of sub-chunk l / 2 -- with the complete decode (block search, sector tables, source block lookup);
 *   - a sub-chunk is a run iff both probes lie in the same destination row, their innermost logical indices differ by the entry
 *     distance n - 1, they map into the same source block, agree in every source position except on the one source axis the
 *     innermost destination axis drives (jstar), and differ there by n - 1.  Logical indices ascend along a destination row and
 *     positions ascend with the logical index inside a sector, so the entries in between are then equally spaced as well
 *     (a SLICE of the innermost axis needs an ascending index list for this; the host checks it);
 *   - a run is copied with two coalesced loads / stores per lane (address = first + i * stride); sub-chunks that are not runs are
 *     decoded entry by entry with the same decode function.
 * About 1 warp instruction per destination entry instead of 10. */
struct Probe
{
	int b;              /* destination block */
	uint32_t row;       /* destination row inside the block (outer positions) */
	int ll;             /* logical index on the innermost destination axis */
	int pos, bl;        /* position in the destination row, length of the row */
	int srem;           /* positions left in the source block along axis jstar, this one included */
	int sidx;           /* index into the scale vector */
	int64_t scell;      /* source grid cell */
	int64_t sbase;      /* element offset of the source block, < 0: not stored */
	uint32_t soff;      /* offset inside the source block */
	uint32_t srest;     /* soff without the contribution of source axis jstar */
	int spos;           /* position on source axis jstar */
	uint32_t sstr;      /* stride of source axis jstar inside the source block */
};

template <int ND>
__device__ __forceinline__ void remap_decode(const LayoutDev& D, const LayoutDev& S, const RemapParams& p, const int jstar, const int64_t e, int b, Probe& o)
{
	const int dn = D.ndim, sn = S.ndim;
	while (e >= D.blk_off[b + 1]) { b++; }
	o.b = b; o.row = 0; o.ll = 0; o.pos = 0; o.bl = 1;
	uint32_t r = (uint32_t)(e - D.blk_off[b]);
	int64_t cell = D.blk_grid[b];
	int ld[ND];
	uint32_t bd_last = 1;
	/* sectors of the block (innermost first), then the positions of the entry */
	int s0[ND], bd[ND];
	#pragma unroll
	for (int i = ND - 1; i >= 0; i--) {
		s0[i] = 0; bd[i] = 1;
		if (i < dn) {
			const int ns = D.nsec[i];
			const int sc = (int)(cell % ns); cell /= ns;
			s0[i] = D.secstart[i][sc];
			bd[i] = D.secstart[i][sc + 1] - s0[i];
		}
	}
	#pragma unroll
	for (int i = ND - 1; i >= 0; i--) {
		ld[i] = 0;
		if (i < dn) {
			const uint32_t q = r / (uint32_t)bd[i];
			ld[i] = D.log_of[i][s0[i] + (int)(r - q * (uint32_t)bd[i])];
			if (i == dn - 1) { bd_last = (uint32_t)bd[i]; o.row = q; o.ll = ld[i]; o.pos = (int)(r - q * (uint32_t)bd[i]); o.bl = bd[i]; }
			r = q;
		}
	}
	(void)bd_last;
	o.sidx = 0;
	if (p.scale_ax >= 0) {
		#pragma unroll
		for (int i = 0; i < ND; i++) { if (i == p.scale_ax) { o.sidx = ld[i]; } }
	}
	int ls[ND];
	#pragma unroll
	for (int j = 0; j < ND; j++) { ls[j] = 0; }
	if (p.op == CTBD_REMAP_TRANSPOSE) {
		#pragma unroll
		for (int i = 0; i < ND; i++) { if (i < dn) {
			#pragma unroll
			for (int j = 0; j < ND; j++) { if (p.perm[i] == j) { ls[j] = ld[i]; } }
		} }
	}
	else if (p.op == CTBD_REMAP_FLATTEN) {
		const int sd1 = (int)S.dim[p.i_ax + 1];
		#pragma unroll
		for (int i = 0; i < ND; i++) { if (i < dn) {
			if (i < p.i_ax) { ls[i] = ld[i]; }
			else if (i == p.i_ax) {
				const int hi = ld[i] / sd1;
				#pragma unroll
				for (int j = 0; j < ND - 1; j++) { if (j == i) { ls[j] = hi; ls[j + 1] = ld[i] - hi * sd1; } }
			}
			else { if (i + 1 < ND) { ls[i + 1] = ld[i]; } }
		} }
	}
	else if (p.op == CTBD_REMAP_SPLIT) {
		const int dd1 = (int)D.dim[p.i_ax + 1];
		#pragma unroll
		for (int i = 0; i < ND; i++) { if (i < dn) {
			if (i < p.i_ax) { ls[i] = ld[i]; }
			else if (i == p.i_ax) { ls[i] = ld[i] * dd1; }
			else if (i == p.i_ax + 1) { if (i >= 1) { ls[i - 1] += ld[i]; } }
			else { if (i >= 1) { ls[i - 1] = ld[i]; } }
		} }
	}
	else if (p.op == CTBD_REMAP_SLICE) {
		#pragma unroll
		for (int i = 0; i < ND; i++) { if (i < dn) { ls[i] = (i == p.i_ax) ? (int)p.ind[ld[i]] : ld[i]; } }
	}
	else {
		#pragma unroll
		for (int i = 0; i < ND; i++) { if (i < dn) { ls[i] = ld[i]; } }
	}
	int64_t scell = 0;
	uint32_t soff = 0, srest = 0, sstr = 1;
	int spos = 0, srem = 1;
	#pragma unroll
	for (int j = 0; j < ND; j++) {
		if (j < sn) {
			const int sc = S.sec_of[j][ls[j]];
			const int ps = S.pos_of[j][ls[j]];
			const uint32_t sb = (uint32_t)(S.secstart[j][sc + 1] - S.secstart[j][sc]);
			scell = scell * S.nsec[j] + sc;
			soff = soff * sb + (uint32_t)ps;
			srest = srest * sb + (j == jstar ? 0u : (uint32_t)ps);
			sstr = (j == jstar) ? 1u : sstr * sb;
			if (j == jstar) { spos = ps; srem = (int)sb - ps; }
		}
	}
	/* sstr so far = product of the block extents of the axes behind jstar (reset to 1 at jstar, multiplied afterwards) */
	o.scell = scell; o.soff = soff; o.srest = srest; o.spos = spos; o.sstr = sstr; o.srem = srem;
	o.sbase = S.grid_off[scell];
}

template <typename T, int ND>
__global__ void __launch_bounds__(256, 2) remap_run_kernel(const LayoutDev D, const LayoutDev S, const RemapParams p, const int jstar, const int lin_ok,
	T* __restrict__ dst, const T* __restrict__ src)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	constexpr int SUB = 64, NSUB = 16;
	constexpr int64_t WCH = SUB * NSUB;           /* entries per warp chunk */
	constexpr int64_t CH = 8 * WCH;
	const unsigned FULL = 0xFFFFFFFFu;
	for (int64_t cbase = (int64_t)blockIdx.x * CH; cbase < D.nstore; cbase += (int64_t)gridDim.x * CH)
	{
		const int64_t wbase = cbase + (int64_t)warp * WCH;
		if (wbase >= D.nstore) { continue; }
		/* probe of this lane */
		const int64_t sub0 = wbase + (int64_t)(lane >> 1) * SUB;
		int64_t pe = sub0 + ((lane & 1) ? (SUB - 1) : 0);
		if (pe > D.nstore - 1) { pe = D.nstore - 1; }
		const bool sub_valid = (sub0 < D.nstore);
		Probe pr;
		pr.b = 0; pr.row = 0; pr.ll = 0; pr.sidx = 0; pr.pos = 0; pr.bl = 1; pr.srem = 1; pr.scell = -1; pr.sbase = -1; pr.soff = 0; pr.srest = 0; pr.spos = 0; pr.sstr = 1;
		if (sub_valid)
		{
			int lo = 0, hi = D.nblk - 1;
			while (lo < hi) {
				const int mid = (lo + hi + 1) >> 1;
				if (D.blk_off[mid] <= pe) { lo = mid; } else { hi = mid - 1; }
			}
			remap_decode<ND>(D, S, p, jstar, pe, lo, pr);
		}
		/* exchange with the partner probe.  Every sub-chunk is described by two PIECES, piece 0 held by the even lane (entries
		 * [0, n0) of the sub-chunk) and piece 1 by the odd lane (entries [n0, n)): a sub-chunk that is one run has n0 = n; one that
		 * crosses into the NEXT row of the same block is cut at the row end and both pieces are judged on their own (second probe
		 * pass: last entry of the first row, first entry of the second); anything else is decoded entry by entry */
		const int     o_b     = __shfl_xor_sync(FULL, pr.b, 1);
		const uint32_t o_row  = __shfl_xor_sync(FULL, pr.row, 1);
		const int     o_ll    = __shfl_xor_sync(FULL, pr.ll, 1);
		const int64_t o_scell = __shfl_xor_sync(FULL, pr.scell, 1);
		const uint32_t o_rest = __shfl_xor_sync(FULL, pr.srest, 1);
		const int     o_spos  = __shfl_xor_sync(FULL, pr.spos, 1);
		const int64_t o_pe    = __shfl_xor_sync(FULL, pe, 1);
		const bool even = (lane & 1) == 0;
		const int nn = (int)(even ? o_pe - pe : pe - o_pe);      /* entries of the sub-chunk - 1 (both lanes) */
		int p_run = 0, p_n = 0, p_si = pr.sidx, p_b = pr.b;
		int64_t p_a0 = pr.sbase + (int64_t)pr.soff, p_st = (int64_t)pr.sstr;
		int cand = 0, n0 = 0;
		if (sub_valid && even) {
			p_n = nn + 1;
			p_run = (lin_ok != 0) && (o_b == pr.b) && (o_row == pr.row) && (o_ll - pr.ll == nn) && (o_scell == pr.scell) && (pr.sbase >= 0)
				&& (o_rest == pr.srest) && (o_spos - pr.spos == nn);
			if (nn == 0 && pr.sbase >= 0) { p_run = 1; }
			/* longest run the first probe can start: to the end of its destination row or of its source block along jstar */
			n0 = min(pr.bl - pr.pos, pr.srem);
			cand = (!p_run) && (lin_ok != 0) && (pr.sbase >= 0) && (n0 >= 1) && (n0 <= nn);
		}
		const int cand_pair = __shfl_sync(FULL, cand, lane & ~1);
		const int n0_pair = __shfl_sync(FULL, n0, lane & ~1);
		if (__any_sync(FULL, cand_pair))
		{
			if (cand_pair)
			{
				Probe pz;
				const int64_t e2 = sub0 + n0_pair - 1 + (even ? 0 : 1);
				remap_decode<ND>(D, S, p, jstar, e2, even ? pr.b : o_b, pz);
				if (even) {
					/* piece 0 = [first probe, end of its row] */
					const int c1 = n0_pair - 1;
					p_n = n0_pair;
					p_run = (pz.b == pr.b) && (pz.row == pr.row) && (pz.ll - pr.ll == c1) && (pz.scell == pr.scell) && (pr.sbase >= 0)
						&& (pz.srest == pr.srest) && (pz.spos - pr.spos == c1);
				}
				else {
					/* piece 1 = [start of the next row, last probe] */
					const int c1 = nn - n0_pair;
					p_n = nn + 1 - n0_pair;
					p_run = (pz.b == pr.b) && (pz.row == pr.row) && (pr.ll - pz.ll == c1) && (pz.scell == pr.scell) && (pz.sbase >= 0)
						&& (pz.srest == pr.srest) && (pr.spos - pz.spos == c1);
					p_a0 = pz.sbase + (int64_t)pz.soff; p_st = (int64_t)pz.sstr; p_si = pz.sidx; p_b = pz.b;
				}
			}
		}
		const int inner = (p.scale_ax == D.ndim - 1);
		/* four sub-chunks per pass: when all their pieces are runs the eight loads of a lane are in flight together */
		#pragma unroll 1
		for (int q4 = 0; q4 < NSUB; q4 += 4)
		{
			if (wbase + (int64_t)q4 * SUB >= D.nstore) { break; }
			int c0[4], c1[4];
			bool all_runs = true;
			#pragma unroll
			for (int u = 0; u < 4; u++) {
				c0[u] = __shfl_sync(FULL, p_n, 2 * (q4 + u));
				c1[u] = __shfl_sync(FULL, p_n, 2 * (q4 + u) + 1);
				const int r0 = __shfl_sync(FULL, p_run, 2 * (q4 + u)), r1 = __shfl_sync(FULL, p_run, 2 * (q4 + u) + 1);
				all_runs = all_runs && (c0[u] == 0 || r0) && (c1[u] == 0 || r1);
			}
			if (all_runs)
			{
				T v[4][2];
				double sc[4][2];
				#pragma unroll
				for (int u = 0; u < 4; u++) {
					#pragma unroll
					for (int t = 0; t < 2; t++) {
						const int i = lane + 32 * t;
						const int from = 2 * (q4 + u) + (i < c0[u] ? 0 : 1);      /* lane that holds the piece of entry i */
						const int64_t a0 = __shfl_sync(FULL, p_a0, from);
						const int64_t st = __shfl_sync(FULL, p_st, from);
						const int si = __shfl_sync(FULL, p_si, from);
						const int ip = (i < c0[u]) ? i : i - c0[u];
						v[u][t] = zero_of<T>(); sc[u][t] = 1.0;
						if (i < c0[u] + c1[u]) {
							v[u][t] = src[a0 + (int64_t)ip * st];
							if (p.scale_ax >= 0) { sc[u][t] = p.scale[si + (inner ? ip : 0)]; }
						}
					}
				}
				#pragma unroll
				for (int u = 0; u < 4; u++) {
					const int64_t first = wbase + (int64_t)(q4 + u) * SUB;
					#pragma unroll
					for (int t = 0; t < 2; t++) {
						const int i = lane + 32 * t;
						if (i < c0[u] + c1[u]) {
							T x = conj_if<T>(v[u][t], p.conj);
							if (p.scale_ax >= 0) { x = scale_by(x, sc[u][t]); }
							dst[first + i] = x;
						}
					}
				}
				continue;
			}
			#pragma unroll 1
			for (int u = 0; u < 4; u++)
			{
				const int q = q4 + u;
				const int64_t first = wbase + (int64_t)q * SUB;
				if (first >= D.nstore) { break; }
				#pragma unroll 1
				for (int pc = 0; pc < 2; pc++)
				{
					const int from = 2 * q + pc;
					const int cnt = __shfl_sync(FULL, p_n, from);
					if (cnt == 0) { continue; }
					const int off = (pc == 0) ? 0 : __shfl_sync(FULL, p_n, 2 * q);
					if (__shfl_sync(FULL, p_run, from))
					{
						const int64_t a0 = __shfl_sync(FULL, p_a0, from);
						const int64_t st = __shfl_sync(FULL, p_st, from);
						const int si = __shfl_sync(FULL, p_si, from);
						#pragma unroll
						for (int t = 0; t < 2; t++) {
							const int i = lane + 32 * t;
							if (i < cnt) {
								T x = conj_if<T>(src[a0 + (int64_t)i * st], p.conj);
								if (p.scale_ax >= 0) { x = scale_by(x, p.scale[si + (inner ? i : 0)]); }
								dst[first + off + i] = x;
							}
						}
					}
					else
					{
						/* entry by entry */
						const int bh = __shfl_sync(FULL, p_b, 2 * q);
						#pragma unroll 1
						for (int t = 0; t < 2; t++) {
							const int i = lane + 32 * t;
							if (i < cnt) {
								Probe z;
								remap_decode<ND>(D, S, p, jstar, first + off + i, bh, z);
								T x = (z.sbase >= 0) ? src[z.sbase + (int64_t)z.soff] : zero_of<T>();
								x = conj_if<T>(x, p.conj);
								if (p.scale_ax >= 0) { x = scale_by(x, p.scale[z.sidx]); }
								dst[first + off + i] = x;
							}
						}
					}
				}
			}
		}
	}
}

/* scatter form (CTBD_REMAP_UNSLICE): every SOURCE entry is written to the destination entry whose index on axis i_ax is
 * ind[source index]; used to merge the column slices computed by different GPUs back into the full tensor */
template <typename T>
__global__ void __launch_bounds__(256) unslice_kernel(const LayoutDev D, const LayoutDev S, const RemapParams p, T* __restrict__ dst, const T* __restrict__ src)
{
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < S.nstore; e += stride)
	{
		int lo = 0, hi = S.nblk - 1;
		while (lo < hi) {
			const int mid = (lo + hi + 1) >> 1;
			if (S.blk_off[mid] <= e) { lo = mid; } else { hi = mid - 1; }
		}
		int64_t cell = S.blk_grid[lo];
		int64_t r = e - S.blk_off[lo];
		int sec[CTBD_MAXDIM];
		#pragma unroll
		for (int i = CTBD_MAXDIM - 1; i >= 0; i--) {
			if (i < S.ndim) { sec[i] = (int)(cell % S.nsec[i]); cell /= S.nsec[i]; }
		}
		int64_t ld[CTBD_MAXDIM];
		#pragma unroll
		for (int i = CTBD_MAXDIM - 1; i >= 0; i--) {
			if (i < S.ndim) {
				const int s0 = S.secstart[i][sec[i]];
				const int bd = S.secstart[i][sec[i] + 1] - s0;
				const int pos = (int)(r % bd); r /= bd;
				const int64_t ls = S.log_of[i][s0 + pos];
				ld[i] = (i == p.i_ax) ? p.ind[ls] : ls;
			}
		}
		int64_t dcell = 0, doff = 0;
		#pragma unroll
		for (int i = 0; i < CTBD_MAXDIM; i++) {
			if (i < D.ndim) {
				const int s = D.sec_of[i][ld[i]];
				dcell = dcell * D.nsec[i] + s;
				doff = doff * (D.secstart[i][s + 1] - D.secstart[i][s]) + D.pos_of[i][ld[i]];
			}
		}
		const int64_t dbase = D.grid_off[dcell];
		if (dbase >= 0) { dst[dbase + doff] = src[e]; }
	}
}

/* ---- batched strided 2-d copy ---- */
struct CopyTile { int32_t desc, row0; };
struct CopyPlan { int dtype; int ndesc, ntiles; ctbd_copy2d* descs; CopyTile* tiles; };
static constexpr int COPY_ROWS = 16;      /* rows per tile: 8 warps x 2 rows */

template <typename T>
__global__ void __launch_bounds__(256) copy2d_kernel(int ntiles, const CopyTile* __restrict__ tiles, const ctbd_copy2d* __restrict__ descs,
	const T* __restrict__ src, T* __restrict__ dst)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int t = blockIdx.x; t < ntiles; t += gridDim.x)
	{
		const CopyTile tl = tiles[t];
		const ctbd_copy2d d = descs[tl.desc];
		const int r1 = min(tl.row0 + COPY_ROWS, d.rows);
		for (int i = tl.row0 + warp; i < r1; i += 8) {
			const T* __restrict__ s = src + d.src_off + (int64_t)i * d.src_ld;
			T* __restrict__ o = dst + d.dst_off + (int64_t)i * d.dst_ld;
			for (int j = lane; j < d.cols; j += 32) { o[j] = s[j]; }
		}
	}
}

struct CopySrcs { const void* p[8]; int64_t stride; int vec_ok; };

/* source rows may live in peer memory: 16-byte loads where the run allows it, every lane keeps several loads in flight */
template <typename T>
__global__ void __launch_bounds__(256) copy2d_multi_kernel(int ntiles, const CopyTile* __restrict__ tiles, const ctbd_copy2d* __restrict__ descs,
	const CopySrcs srcs, T* __restrict__ dst)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int t = blockIdx.x; t < ntiles; t += gridDim.x)
	{
		const CopyTile tl = tiles[t];
		const ctbd_copy2d d = descs[tl.desc];
		const int64_t q = d.src_off / srcs.stride, off = d.src_off - q * srcs.stride;
		const T* __restrict__ sbase = reinterpret_cast<const T*>(srcs.p[q]) + off;
		const int r1 = min(tl.row0 + COPY_ROWS, d.rows);
		const bool vec = (sizeof(T) == 8) && srcs.vec_ok && ((d.cols & 1) == 0) && ((d.src_ld & 1) == 0) && ((d.dst_ld & 1) == 0) && ((off & 1) == 0) && ((d.dst_off & 1) == 0);
		for (int i = tl.row0 + warp; i < r1; i += 8) {
			const T* __restrict__ s = sbase + (int64_t)i * d.src_ld;
			T* __restrict__ o = dst + d.dst_off + (int64_t)i * d.dst_ld;
			if (vec) {
				const double2* __restrict__ s2 = reinterpret_cast<const double2*>(s);
				double2* __restrict__ o2 = reinterpret_cast<double2*>(o);
				for (int j = lane; j < d.cols / 2; j += 32) { o2[j] = s2[j]; }
			}
			else { for (int j = lane; j < d.cols; j += 32) { o[j] = s[j]; } }
		}
	}
}

struct CopyDsts { void* p[8]; int n; int vec_ok; };

/* one local read, the same strided block written to several (peer-mapped) destinations: 16 bytes per lane, whole row runs per warp */
template <typename T>
__global__ void __launch_bounds__(256) copy2d_push_kernel(int ntiles, const CopyTile* __restrict__ tiles, const ctbd_copy2d* __restrict__ descs,
	const T* __restrict__ src, const CopyDsts dsts)
{
	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	for (int t = blockIdx.x; t < ntiles; t += gridDim.x)
	{
		const CopyTile tl = tiles[t];
		const ctbd_copy2d d = descs[tl.desc];
		const int r1 = min(tl.row0 + COPY_ROWS, d.rows);
		const bool vec = (sizeof(T) == 8) && dsts.vec_ok && ((d.cols & 1) == 0) && ((d.src_ld & 1) == 0) && ((d.dst_ld & 1) == 0) && ((d.src_off & 1) == 0) && ((d.dst_off & 1) == 0);
		for (int i = tl.row0 + warp; i < r1; i += 8) {
			const T* __restrict__ s = src + d.src_off + (int64_t)i * d.src_ld;
			const int64_t doff = d.dst_off + (int64_t)i * d.dst_ld;
			if (vec) {
				const double2* __restrict__ s2 = reinterpret_cast<const double2*>(s);
				for (int j = lane; j < d.cols / 2; j += 32) {
					const double2 v = s2[j];
					#pragma unroll
					for (int q = 0; q < 8; q++) { if (q < dsts.n) { reinterpret_cast<double2*>(reinterpret_cast<T*>(dsts.p[q]) + doff)[j] = v; } }
				}
			}
			else {
				for (int j = lane; j < d.cols; j += 32) {
					const T v = s[j];
					#pragma unroll
					for (int q = 0; q < 8; q++) { if (q < dsts.n) { (reinterpret_cast<T*>(dsts.p[q]) + doff)[j] = v; } }
				}
			}
		}
	}
}

} // namespace ctbd

using namespace ctbd;

extern "C" {

int ctbd_layout_create(const struct ctbd_layout_host* h, void** layout)
{
	CTBD_REQUIRE_INIT();
	Layout* L = new Layout();
	memset(&L->d, 0, sizeof(L->d));
	L->arena = nullptr;
	/* pack all tables into one host buffer (8-byte aligned pieces), upload once */
	size_t bytes = 0;
	auto reserve = [&](size_t n) { size_t off = bytes; bytes += (n + 7) & ~(size_t)7; return off; };
	size_t o_sec[CTBD_MAXDIM], o_pos[CTBD_MAXDIM], o_start[CTBD_MAXDIM], o_log[CTBD_MAXDIM];
	for (int i = 0; i < h->ndim; i++) {
		o_sec[i]   = reserve((size_t)h->dim[i] * 4);
		o_pos[i]   = reserve((size_t)h->dim[i] * 4);
		o_start[i] = reserve((size_t)(h->nsec[i] + 1) * 4);
		o_log[i]   = reserve((size_t)h->dim[i] * 4);
	}
	const size_t o_grid = reserve((size_t)h->ngrid * 8);
	const size_t o_bg   = reserve((size_t)(h->nblk > 0 ? h->nblk : 1) * 8);
	const size_t o_bo   = reserve((size_t)(h->nblk + 1) * 8);
	std::vector<unsigned char> buf(bytes > 0 ? bytes : 8, 0);
	for (int i = 0; i < h->ndim; i++) {
		memcpy(&buf[o_sec[i]],   h->sec_of[i],   (size_t)h->dim[i] * 4);
		memcpy(&buf[o_pos[i]],   h->pos_of[i],   (size_t)h->dim[i] * 4);
		memcpy(&buf[o_start[i]], h->secstart[i], (size_t)(h->nsec[i] + 1) * 4);
		memcpy(&buf[o_log[i]],   h->log_of[i],   (size_t)h->dim[i] * 4);
	}
	memcpy(&buf[o_grid], h->grid_off, (size_t)h->ngrid * 8);
	if (h->nblk > 0) { memcpy(&buf[o_bg], h->blk_grid, (size_t)h->nblk * 8); }
	memcpy(&buf[o_bo], h->blk_off, (size_t)(h->nblk + 1) * 8);
	if (upload(buf.data(), buf.size(), &L->arena) < 0) { delete L; return -1; }
	/* the staging vector dies at return: make sure the copy has been consumed */
	CTBD_CUDA(cudaStreamSynchronize(rt().stream));
	unsigned char* base = (unsigned char*)L->arena;
	LayoutDev& d = L->d;
	d.ndim = h->ndim; d.dtype = h->dtype; d.ngrid = h->ngrid; d.nblk = h->nblk; d.nstore = h->nstore;
	for (int i = 0; i < h->ndim; i++) {
		d.dim[i] = h->dim[i]; d.nsec[i] = h->nsec[i];
		d.sec_of[i]   = (const int32_t*)(base + o_sec[i]);
		d.pos_of[i]   = (const int32_t*)(base + o_pos[i]);
		d.secstart[i] = (const int32_t*)(base + o_start[i]);
		d.log_of[i]   = (const int32_t*)(base + o_log[i]);
	}
	d.grid_off = (const int64_t*)(base + o_grid);
	d.blk_grid = (const int64_t*)(base + o_bg);
	d.blk_off  = (const int64_t*)(base + o_bo);
	L->max_block = 0; L->max_dim = 0;
	for (int b = 0; b < h->nblk; b++) { L->max_block = std::max<int64_t>(L->max_block, h->blk_off[b + 1] - h->blk_off[b]); }
	for (int i = 0; i < h->ndim; i++) { L->max_dim = std::max<int64_t>(L->max_dim, h->dim[i]); }
	*layout = L;
	return 0;
}

int ctbd_layout_destroy(void* layout)
{
	Layout* L = (Layout*)layout;
	if (L == nullptr) { return 0; }
	ctbd_free(L->arena);
	delete L;
	return 0;
}

int ctbd_copy_plan_create(int dtype, int n, const struct ctbd_copy2d* descs_host, void** plan)
{
	CTBD_REQUIRE_INIT();
	if (dtype != CTBD_F64 && dtype != CTBD_C128) { return fail_msg("copy plan: unsupported dtype"); }
	CopyPlan* p = new CopyPlan();
	p->dtype = dtype; p->ndesc = n; p->descs = nullptr; p->tiles = nullptr;
	std::vector<CopyTile> tiles;
	for (int i = 0; i < n; i++) {
		if (descs_host[i].cols <= 0) { continue; }
		for (int r0 = 0; r0 < descs_host[i].rows; r0 += COPY_ROWS) { CopyTile t; t.desc = i; t.row0 = r0; tiles.push_back(t); }
	}
	p->ntiles = (int)tiles.size();
	int rc = upload(descs_host, (size_t)n * sizeof(ctbd_copy2d), (void**)&p->descs);
	rc |= upload(tiles.data(), tiles.size() * sizeof(CopyTile), (void**)&p->tiles);
	if (rc < 0) { ctbd_copy_plan_destroy(p); return -1; }
	*plan = p;
	return 0;
}

int ctbd_copy_plan_run(void* plan, const void* src, void* dst)
{
	CopyPlan* p = (CopyPlan*)plan;
	if (p == nullptr || p->ntiles == 0) { return 0; }
	const int grid = std::min(p->ntiles, rt().sm_count * 16);
	if (p->dtype == CTBD_F64) { copy2d_kernel<double><<<grid, 256, 0, rt().stream>>>(p->ntiles, p->tiles, p->descs, (const double*)src, (double*)dst); }
	else                      { copy2d_kernel<double2><<<grid, 256, 0, rt().stream>>>(p->ntiles, p->tiles, p->descs, (const double2*)src, (double2*)dst); }
	CTBD_LAUNCH_CHECK();
	return 0;
}

int ctbd_copy_plan_run_multi(void* plan, int nsrc, const void* const* srcs, int64_t src_stride, void* dst)
{
	CopyPlan* p = (CopyPlan*)plan;
	if (p == nullptr || p->ntiles == 0) { return 0; }
	if (nsrc < 1 || nsrc > 8 || src_stride <= 0) { return fail_msg("copy plan: between 1 and 8 sources"); }
	CopySrcs cs;
	for (int i = 0; i < 8; i++) { cs.p[i] = srcs[i < nsrc ? i : 0]; }
	cs.stride = src_stride;
	bool aligned = (((uintptr_t)dst) & 15) == 0;
	for (int i = 0; i < nsrc; i++) { aligned = aligned && ((((uintptr_t)srcs[i]) & 15) == 0); }
	cs.vec_ok = aligned ? 1 : 0;
	const int grid = std::min(p->ntiles, rt().sm_count * 8);
	if (p->dtype == CTBD_F64) { copy2d_multi_kernel<double><<<grid, 256, 0, rt().stream>>>(p->ntiles, p->tiles, p->descs, cs, (double*)dst); }
	else                      { copy2d_multi_kernel<double2><<<grid, 256, 0, rt().stream>>>(p->ntiles, p->tiles, p->descs, cs, (double2*)dst); }
	CTBD_LAUNCH_CHECK();
	return 0;
}

int ctbd_copy_plan_run_push(void* plan, const void* src, int ndst, void* const* dsts)
{
	CopyPlan* p = (CopyPlan*)plan;
	if (p == nullptr || p->ntiles == 0) { return 0; }
	if (ndst < 1 || ndst > 8) { return fail_msg("copy plan: between 1 and 8 destinations"); }
	CopyDsts cd;
	bool aligned = (((uintptr_t)src) & 15) == 0;
	for (int i = 0; i < 8; i++) { cd.p[i] = dsts[i < ndst ? i : 0]; }
	for (int i = 0; i < ndst; i++) { aligned = aligned && ((((uintptr_t)dsts[i]) & 15) == 0); }
	cd.n = ndst; cd.vec_ok = aligned ? 1 : 0;
	const int grid = std::min(p->ntiles, rt().sm_count * 8);
	if (p->dtype == CTBD_F64) { copy2d_push_kernel<double><<<grid, 256, 0, rt().stream>>>(p->ntiles, p->tiles, p->descs, (const double*)src, cd); }
	else                      { copy2d_push_kernel<double2><<<grid, 256, 0, rt().stream>>>(p->ntiles, p->tiles, p->descs, (const double2*)src, cd); }
	CTBD_LAUNCH_CHECK();
	return 0;
}

int ctbd_copy_plan_destroy(void* plan)
{
	CopyPlan* p = (CopyPlan*)plan;
	if (p == nullptr) { return 0; }
	ctbd_free(p->descs); ctbd_free(p->tiles);
	delete p;
	return 0;
}

int ctbd_remap(const struct ctbd_remap_args* a)
{
	CTBD_REQUIRE_INIT();
	const Layout* D = (const Layout*)a->dst_layout;
	const Layout* S = (const Layout*)a->src_layout;
	if (D->d.dtype != S->d.dtype) { return fail_msg("remap: dtype mismatch"); }
	if (D->d.nstore == 0) { return 0; }
	if (a->op == CTBD_REMAP_UNSLICE)
	{
		if (S->d.nstore == 0) { return 0; }
		RemapParams p;
		memset(&p, 0, sizeof(p));
		p.op = a->op; p.i_ax = a->i_ax; p.scale_ax = -1;
		void* ind_dev = nullptr;
		if (upload(a->ind, (size_t)S->d.dim[a->i_ax] * sizeof(int64_t), &ind_dev) < 0) { return -1; }
		p.ind = (const int64_t*)ind_dev;
		int64_t blocks = std::min<int64_t>(ceil_div(S->d.nstore, 256), (int64_t)rt().sm_count * 16);
		if (D->d.dtype == CTBD_F64) { unslice_kernel<double><<<(int)blocks, 256, 0, rt().stream>>>(D->d, S->d, p, (double*)a->dst, (const double*)a->src); }
		else if (D->d.dtype == CTBD_C128) { unslice_kernel<double2><<<(int)blocks, 256, 0, rt().stream>>>(D->d, S->d, p, (double2*)a->dst, (const double2*)a->src); }
		else { return fail_msg("remap: unsupported dtype"); }
		CTBD_LAUNCH_CHECK();
		ctbd_free(ind_dev);
		return 0;
	}
	RemapParams p;
	memset(&p, 0, sizeof(p));
	p.op = a->op; p.i_ax = a->i_ax; p.conj = a->conj; p.scale_ax = a->scale_ax; p.scale = a->scale;
	for (int i = 0; i < CTBD_MAXDIM; i++) { p.perm[i] = a->perm[i]; }
	void* ind_dev = nullptr;
	if (a->op == CTBD_REMAP_SLICE) {
		if (upload(a->ind, (size_t)D->d.dim[a->i_ax] * sizeof(int64_t), &ind_dev) < 0) { return -1; }
		p.ind = (const int64_t*)ind_dev;
	}
	const int threads = 256;
	static const int use_old = (getenv("CTB_REMAP_OLD") != nullptr) ? 1 : 0;
	if (D->d.dtype != CTBD_F64 && D->d.dtype != CTBD_C128) { return fail_msg("remap: unsupported dtype"); }
	if (!use_old && D->max_block < ((int64_t)1 << 31) && S->max_block < ((int64_t)1 << 31) && D->max_dim < ((int64_t)1 << 31) && S->max_dim < ((int64_t)1 << 31))
	{
		/* run form (probe decode + equally spaced copies); CTB_REMAP_FAST=1 selects the per-entry row-cached kernel */
		static const int use_fast = (getenv("CTB_REMAP_FAST") != nullptr) ? 1 : 0;
		const bool small = (D->d.ndim <= 4 && S->d.ndim <= 4);
		if (!use_fast)
		{
			const int dn = D->d.ndim, sn = S->d.ndim;
			int jstar = sn - 1;
			if (a->op == CTBD_REMAP_TRANSPOSE) { jstar = a->perm[dn - 1]; }
			int lin_ok = 1;
			if (a->op == CTBD_REMAP_SLICE && a->i_ax == dn - 1) {
				/* a run needs the index list of the innermost axis to ascend */
				for (int64_t j = 1; j < D->d.dim[dn - 1]; j++) { if (a->ind[j] <= a->ind[j - 1]) { lin_ok = 0; break; } }
			}
			int64_t blocks = ceil_div(D->d.nstore, (int64_t)8 * 1024);
			const int64_t maxb = (int64_t)rt().sm_count * 8;
			if (blocks > maxb) { blocks = maxb; }
			if (D->d.dtype == CTBD_F64) {
				if (small) { remap_run_kernel<double, 4><<<(int)blocks, threads, 0, rt().stream>>>(D->d, S->d, p, jstar, lin_ok, (double*)a->dst, (const double*)a->src); }
				else       { remap_run_kernel<double, 8><<<(int)blocks, threads, 0, rt().stream>>>(D->d, S->d, p, jstar, lin_ok, (double*)a->dst, (const double*)a->src); }
			}
			else {
				if (small) { remap_run_kernel<double2, 4><<<(int)blocks, threads, 0, rt().stream>>>(D->d, S->d, p, jstar, lin_ok, (double2*)a->dst, (const double2*)a->src); }
				else       { remap_run_kernel<double2, 8><<<(int)blocks, threads, 0, rt().stream>>>(D->d, S->d, p, jstar, lin_ok, (double2*)a->dst, (const double2*)a->src); }
			}
		}
		else
		{
		constexpr int K = 4;
		int64_t blocks = ceil_div(D->d.nstore, (int64_t)threads * K);
		const int64_t maxb = (int64_t)rt().sm_count * 8;
		if (blocks > maxb) { blocks = maxb; }
		if (D->d.dtype == CTBD_F64) {
			if (small) { remap_fast_kernel<double, 4, K><<<(int)blocks, threads, 0, rt().stream>>>(D->d, S->d, p, (double*)a->dst, (const double*)a->src); }
			else       { remap_fast_kernel<double, 8, K><<<(int)blocks, threads, 0, rt().stream>>>(D->d, S->d, p, (double*)a->dst, (const double*)a->src); }
		}
		else {
			if (small) { remap_fast_kernel<double2, 4, K><<<(int)blocks, threads, 0, rt().stream>>>(D->d, S->d, p, (double2*)a->dst, (const double2*)a->src); }
			else       { remap_fast_kernel<double2, 8, K><<<(int)blocks, threads, 0, rt().stream>>>(D->d, S->d, p, (double2*)a->dst, (const double2*)a->src); }
		}
		}
	}
	else
	{
		int64_t blocks = ceil_div(D->d.nstore, threads);
		const int64_t maxb = (int64_t)rt().sm_count * 16;
		if (blocks > maxb) { blocks = maxb; }
		if (D->d.dtype == CTBD_F64) {
			remap_kernel<double><<<(int)blocks, threads, 0, rt().stream>>>(D->d, S->d, p, (double*)a->dst, (const double*)a->src);
		}
		else {
			remap_kernel<double2><<<(int)blocks, threads, 0, rt().stream>>>(D->d, S->d, p, (double2*)a->dst, (const double2*)a->src);
		}
	}
	CTBD_LAUNCH_CHECK();
	if (ind_dev != nullptr) { ctbd_free(ind_dev); }
	return 0;
}

} // extern "C"
