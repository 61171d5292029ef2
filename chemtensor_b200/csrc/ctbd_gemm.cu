/*
 * ctbd_gemm.cu -- grouped block GEMM on the FP64 tensor pipe (DMMA) for sm_100a.
 *
 * Replaces the per-block cblas_dgemm/zgemm swarm of the reference's block_sparse_tensor_dot
 * (src/tensor/block_sparse_tensor.c:1935-1994 -> dense_tensor_dot_update, src/tensor/dense_tensor.c:1761-1828)
 * by ONE persistent launch over a device-resident work list:
 *   - an "output block" C (m x n) is the sum over its "segments" (contracted sector tuples, in the
 *     reference's row-major order) of op(A_seg) op(B_seg);
 *   - the tiles of all output blocks are packed at plan time into one queue per resident CTA
 *     (longest-processing-time-first bin packing on the known K extents), so the launch is one wave;
 *   - a CTA walks its queue with ONE software pipeline that runs across tile and segment boundaries:
 *     a multi-stage cp.async (LDGSTS) shared-memory ring is always STAGES-1 steps ahead in the flattened
 *     (tile, segment, k-chunk) sequence, so the many short K extents of small sectors never drain it and the
 *     epilogue of one tile overlaps the loads of the next;
 *   - the math is mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4; the larger PTX f64 shapes lower to the same
 *     instruction on sm_100a), complex128 as four real DMMAs on (re, im) fragments with optional
 *     conjugation of either operand fused into the fragment load;
 *   - the epilogue scatters through per-block row/column offset tables, which fuses the
 *     block_sparse_tensor_transpose (:785) that follows each contraction in chain_ops.c.
 * tcgen05/TMEM has no f64 kind, so the legacy warp-level tensor path is the FP64 tensor pipe on Blackwell.
 */
#include <vector>
#include <algorithm>
#include "ctbd_common.cuh"

namespace ctbd {

struct GemmTile { int32_t out, m0, n0, nsteps; };   /* nsteps = sum over segments of ceil(k / BK) */

struct GemmPlan
{
	int dtype = 0, a_kcontig = 0, b_ncontig = 0, conj_a = 0, conj_b = 0;
	int nouts = 0, nsegs = 0, ntab = 0;
	ctbd_gemm_out* outs = nullptr;     /* device */
	ctbd_gemm_seg* segs = nullptr;
	int32_t* tab = nullptr;
	int cfg = 0;                       /* tile class of this plan */
	int ntiles = 0, grid = 0;
	GemmTile* tiles = nullptr;         /* device: tiles grouped by CTA queue */
	int32_t* queue = nullptr;          /* device: [grid + 1] queue boundaries */
	void* a_packed = nullptr;          /* device: optional plan-owned A operand (gathered once at plan creation) */
	int64_t* b_rowtab = nullptr;       /* device: optional row table of the B operand */
	/* mixing form (see ctbd_mix_group in ctb_device.h) */
	int n_mix_groups = 0, n_mix_rows = 0, n_mix_tiles = 0;
	ctbd_mix_group* mix_groups = nullptr;
	ctbd_mix_row* mix_rows = nullptr;
	struct MixTile* mix_tiles = nullptr;
	int32_t* mix_nzk = nullptr;        /* [n_a_gather] compacted column indices of the non-zero packed entries, per row at row.a_off */
	void* mix_nzv = nullptr;           /* [n_a_gather] their values */
	int32_t* mix_cnt = nullptr;        /* [n_mix_rows] number of non-zeros of each row */
};

struct MixTile { int32_t group, col0; };

/* ---- PTX helpers ---- */

__device__ __forceinline__ void dmma884(double& c0, double& c1, const double a, const double b)
{
	asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
		: "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc, int src_bytes)
{
	const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(d), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes)
{
	const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
	asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(gsrc), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N)); }

/* ---- tile configuration ---- */

template <typename T, int BM_, int BN_, int WM_, int WN_, int STAGES_, int BK_ = 16>
struct TileCfg
{
	static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, STAGES = STAGES_;
	static constexpr int BK = BK_;
	static constexpr int NWARP = (BM / WM) * (BN / WN);
	static constexpr int NT = NWARP * 32;
	static constexpr bool CPLX = (sizeof(T) == 16);
	/* shared-memory strides in elements, chosen so that a DMMA fragment load is bank-conflict free:
	 *   K-inner tile S[x][k]: stride BK+4;  X-inner tile S[k][x]: stride BX+4 (real) / BX+2 (complex) */
	static constexpr int SK = BK + 4;
	static constexpr int PADX = CPLX ? 2 : 4;
	static constexpr int A_ELEMS_KIN = BM * SK, A_ELEMS_XIN = BK * (BM + PADX);
	static constexpr int B_ELEMS_KIN = BN * SK, B_ELEMS_XIN = BK * (BN + PADX);
	static constexpr int A_ELEMS = A_ELEMS_KIN > A_ELEMS_XIN ? A_ELEMS_KIN : A_ELEMS_XIN;
	static constexpr int B_ELEMS = B_ELEMS_KIN > B_ELEMS_XIN ? B_ELEMS_KIN : B_ELEMS_XIN;
	static constexpr size_t SMEM = (size_t)STAGES * (A_ELEMS + B_ELEMS) * sizeof(T);
};

/* copy one operand tile (BX x BK) of a segment into shared memory, zero-filling out-of-range elements.
 * KIN: global elem(x,k) = g[base + x*ld + k], shared S[x*SK + k];  else global g[base + k*ld + x], shared S[k*SX + x] */
template <typename T, bool KIN, int BX, int BK, int SK, int SX, int NT>
__device__ __forceinline__ void load_tile(T* __restrict__ S, const T* __restrict__ g, const int64_t base, const int ld,
	const int x0, const int X, const int k0, const int K, const bool vec2, const int64_t* __restrict__ rowtab = nullptr, const int odd = 0)
{
	const int tid = threadIdx.x;
	if constexpr (!KIN)
	{
		if (rowtab != nullptr)
		{
			/* row-table form: row kk of the operand starts at g[rowtab[base + kk]] (rows gathered from many small blocks) */
			if constexpr (sizeof(T) == 16) {
				#pragma unroll
				for (int c = tid; c < BX * BK; c += NT) {
					const int k = c / BX, x = c % BX;
					const bool ok = (x0 + x < X) && (k0 + k < K);
					const T* src = ok ? g + rowtab[base + k0 + k] + (x0 + x) : g;
					cp_async16(&S[k * SX + x], src, ok ? 16 : 0);
				}
			}
			else {
				constexpr int CPR = BX / 2;
				const T* g16 = reinterpret_cast<const T*>(reinterpret_cast<uintptr_t>(g) & ~(uintptr_t)15);   /* aligned dummy source of zero-fill copies */
				#pragma unroll
				for (int c = tid; c < BK * CPR; c += NT) {
					const int k = c / CPR, x = (c % CPR) * 2;
					if (k0 + k < K && x0 + x < X) {
						const int64_t ro = rowtab[base + k0 + k] + (x0 + x);
						const bool two = (x0 + x + 1 < X);
						if (((ro + odd) & 1) == 0) { cp_async16(&S[k * SX + x], g + ro, two ? 16 : 8); }
						else {
							cp_async8(&S[k * SX + x], g + ro, 8);
							cp_async8(&S[k * SX + x + 1], two ? g + ro + 1 : g, two ? 8 : 0);
						}
					}
					else { cp_async16(&S[k * SX + x], g16, 0); }
				}
			}
			return;
		}
	}
	if constexpr (sizeof(T) == 16)
	{
		/* complex128: one 16-byte element per copy */
		if constexpr (KIN) {
			#pragma unroll
			for (int c = tid; c < BX * BK; c += NT) {
				const int x = c / BK, k = c % BK;
				const bool ok = (x0 + x < X) && (k0 + k < K);
				const T* src = g + (ok ? base + (int64_t)(x0 + x) * ld + (k0 + k) : base);
				cp_async16(&S[x * SK + k], src, ok ? 16 : 0);
			}
		}
		else {
			#pragma unroll
			for (int c = tid; c < BX * BK; c += NT) {
				const int k = c / BX, x = c % BX;
				const bool ok = (x0 + x < X) && (k0 + k < K);
				const T* src = g + (ok ? base + (int64_t)(k0 + k) * ld + (x0 + x) : base);
				cp_async16(&S[k * SX + x], src, ok ? 16 : 0);
			}
		}
	}
	else
	{
		if (vec2)
		{
			/* two doubles per 16-byte copy (operand offsets and leading dimension are even) */
			if constexpr (KIN) {
				constexpr int CPR = BK / 2;
				#pragma unroll
				for (int c = tid; c < BX * CPR; c += NT) {
					const int x = c / CPR, k = (c % CPR) * 2;
					int nb = 0;
					if (x0 + x < X) { nb = (k0 + k + 1 < K) ? 16 : ((k0 + k < K) ? 8 : 0); }
					const T* src = g + (nb ? base + (int64_t)(x0 + x) * ld + (k0 + k) : base);
					cp_async16(&S[x * SK + k], src, nb);
				}
			}
			else {
				constexpr int CPR = BX / 2;
				#pragma unroll
				for (int c = tid; c < BK * CPR; c += NT) {
					const int k = c / CPR, x = (c % CPR) * 2;
					int nb = 0;
					if (k0 + k < K) { nb = (x0 + x + 1 < X) ? 16 : ((x0 + x < X) ? 8 : 0); }
					const T* src = g + (nb ? base + (int64_t)(k0 + k) * ld + (x0 + x) : base);
					cp_async16(&S[k * SX + x], src, nb);
				}
			}
		}
		else
		{
			if constexpr (KIN) {
				#pragma unroll
				for (int c = tid; c < BX * BK; c += NT) {
					const int x = c / BK, k = c % BK;
					const bool ok = (x0 + x < X) && (k0 + k < K);
					const T* src = g + (ok ? base + (int64_t)(x0 + x) * ld + (k0 + k) : base);
					cp_async8(&S[x * SK + k], src, ok ? 8 : 0);
				}
			}
			else {
				#pragma unroll
				for (int c = tid; c < BX * BK; c += NT) {
					const int k = c / BX, x = c % BX;
					const bool ok = (x0 + x < X) && (k0 + k < K);
					const T* src = g + (ok ? base + (int64_t)(k0 + k) * ld + (x0 + x) : base);
					cp_async8(&S[k * SX + x], src, ok ? 8 : 0);
				}
			}
		}
	}
}


/* Fast operand copy for one pipeline step: every thread owns a FIXED set of 16-byte chunks of the tile, so the per-step work is
 * one pointer increment, one size clamp and the LDGSTS per chunk (the generic load_tile above recomputes indices and bounds).
 * Requires 16-byte aligned chunks (complex128 always; real when offsets and leading dimension are even).
 *   KIN: chunk i of the thread = row xr + i*RP of the tile, columns [kc, kc + EPC): global p + i*RP*ld, valid rows by 'xmask',
 *        bytes limited by the remaining k extent 'krem';
 *   XIN: chunk i = k-row kr + i*RP, columns [xc, xc + EPC): global p + i*RP*ld, valid while kr + i*RP < krem, 'xbytes' fixed per tile. */
template <typename T, bool KIN, int BX, int BK, int SK, int SX, int NT>
struct FastCopy
{
	static constexpr int EPC = 16 / (int)sizeof(T);                 /* elements per 16-byte chunk */
	static constexpr int CPR = (KIN ? BK : BX) / EPC;               /* chunks per tile row */
	static constexpr int RP  = NT / CPR;                            /* tile rows covered per pass of the CTA */
	static constexpr int NC  = (KIN ? BX : BK) / RP;                /* chunks per thread */
	static_assert(NT % CPR == 0 && (KIN ? BX : BK) % RP == 0 && NC >= 1 && NC <= 32, "tile / thread-count mismatch");

	/* tile-constant part: called when the producer enters a segment */
	__device__ __forceinline__ static const T* base(const T* g, int64_t off, int ld, int x0, int k0)
	{
		const int r = threadIdx.x / CPR, c = (threadIdx.x % CPR) * EPC;
		return KIN ? g + off + (int64_t)(x0 + r) * ld + (k0 + c) : g + off + (int64_t)(k0 + r) * ld + (x0 + c);
	}
	/* KIN: bit i set when row xr + i*RP is inside the block;  XIN: bytes of this thread's column chunk (0, 8 or 16) */
	__device__ __forceinline__ static unsigned xinfo(int x0, int X)
	{
		if constexpr (KIN) {
			const int r = threadIdx.x / CPR;
			unsigned m = 0;
			#pragma unroll
			for (int i = 0; i < NC; i++) { m |= (x0 + r + i * RP < X) ? (1u << i) : 0u; }
			return m;
		}
		else {
			const int c = (threadIdx.x % CPR) * EPC;
			const int rem = X - (x0 + c);
			return (unsigned)(rem >= EPC ? 16 : (rem > 0 ? rem * (int)sizeof(T) : 0));
		}
	}
	__device__ __forceinline__ static void copy(T* __restrict__ S, const T* __restrict__ p, int ld, unsigned xi, int krem)
	{
		const int r = threadIdx.x / CPR, c = (threadIdx.x % CPR) * EPC;
		const unsigned sbase = (unsigned)__cvta_generic_to_shared(S) + (unsigned)((KIN ? r * SK + c : r * SX + c) * (int)sizeof(T));
		constexpr unsigned SSTEP = (unsigned)(RP * (KIN ? SK : SX) * (int)sizeof(T));
		const int64_t gstep = (int64_t)RP * ld;
		if constexpr (KIN) {
			const int kb = krem - c;                                    /* elements of this chunk still inside the k extent */
			const int nb = kb >= EPC ? 16 : (kb > 0 ? kb * (int)sizeof(T) : 0);
			#pragma unroll
			for (int i = 0; i < NC; i++) {
				const int bytes = ((xi >> i) & 1u) ? nb : 0;
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(sbase + i * SSTEP), "l"(p + i * gstep), "r"(bytes));
			}
		}
		else {
			#pragma unroll
			for (int i = 0; i < NC; i++) {
				const int bytes = (r + i * RP < krem) ? (int)xi : 0;
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(sbase + i * SSTEP), "l"(p + i * gstep), "r"(bytes));
			}
		}
	}
};

struct GemmArgs
{
	const GemmTile* tiles;
	const int32_t* queue;
	const ctbd_gemm_out* outs;
	const ctbd_gemm_seg* segs;
	const int32_t* tab;
	const void* A; const void* B; void* C;
	int conj_a, conj_b;
	int a_odd, b_odd;     /* operand base pointer is 8 (mod 16): shifts the parity test of the 16-byte copy path */
	int c_odd;            /* multi-destination runs: some destination base is 8 (mod 16), 16-byte stores are then off */
	const int64_t* b_rowtab;   /* optional: B rows gathered through a row table, seg.b_off indexes it (merged-row plans) */
	int ndst;                  /* >= 1: additional destinations of the epilogue (peer-mapped buffers of the other GPUs) follow */
	int mc;                    /* C is an NVSwitch multicast address: every element is stored once with multimem.st, the switch replicates it */
	void* Cx[7];
};

/* stores to a multicast address (multimem.st; SASS: STG.E.{64,128}.STRONG.SYS on the multicast mapping) */
__device__ __forceinline__ void mc_st8(double* p, double v)
{
	asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void mc_st16(void* p, double a, double b)
{
	asm volatile("{ .reg .b32 x0, x1, x2, x3;\n mov.b64 {x0, x1}, %1;\n mov.b64 {x2, x3}, %2;\n multimem.st.relaxed.sys.global.v4.f32 [%0], {x0, x1, x2, x3};\n }"
		:: "l"(p), "d"(a), "d"(b) : "memory");
}

template <typename T, typename Cfg, bool A_KC, bool B_NC>
__global__ void __launch_bounds__(Cfg::NT) grouped_gemm_kernel(const GemmArgs args)
{
	constexpr int BM = Cfg::BM, BN = Cfg::BN, BK = Cfg::BK, WM = Cfg::WM, WN = Cfg::WN, STAGES = Cfg::STAGES, NT = Cfg::NT;
	constexpr int SK = Cfg::SK, SXA = BM + Cfg::PADX, SXB = BN + Cfg::PADX;
	constexpr bool CPLX = Cfg::CPLX;
	constexpr int MI = WM / 8, NI = WN / 8;
	constexpr int NACC = CPLX ? 4 : 2;

	extern __shared__ __align__(16) unsigned char smem_raw[];
	T* As = reinterpret_cast<T*>(smem_raw);
	T* Bs = As + (size_t)STAGES * Cfg::A_ELEMS;

	const int qb = args.queue[blockIdx.x], qe = args.queue[blockIdx.x + 1];
	if (qb >= qe) { return; }
	const T* __restrict__ Ag = reinterpret_cast<const T*>(args.A);
	const T* __restrict__ Bg = reinterpret_cast<const T*>(args.B);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int wm0 = (warp / (BN / WN)) * WM, wn0 = (warp % (BN / WN)) * WN;
	const int lr = lane >> 2, lc = lane & 3;

	/* ---- producer: iterator over the flattened (tile, segment, k-chunk) sequence of this CTA's queue ---- */
	int pt = qb, ps = 0, pk0 = 0, p_m0 = 0, p_n0 = 0, p_M = 0, p_N = 0, p_seg_end = 0;
	ctbd_gemm_seg psg;
	psg.a_off = 0; psg.b_off = 0; psg.k = 0; psg.lda = 0; psg.ldb = 0; psg.pad_ = 0;
	auto producer_enter_tile = [&]() {
		while (pt < qe) {
			const GemmTile t = args.tiles[pt];
			if (t.nsteps > 0) {
				const ctbd_gemm_out o = args.outs[t.out];
				p_m0 = t.m0; p_n0 = t.n0; p_M = o.m; p_N = o.n; ps = o.seg_begin; p_seg_end = o.seg_end; pk0 = 0;
				psg = args.segs[ps];
				return;
			}
			pt++;
		}
	};
	producer_enter_tile();
	/* per-segment state of the fast copy path (16-byte aligned chunks owned by fixed threads) */
	typedef FastCopy<T, A_KC,  BM, BK, SK, SXA, NT> FcA;
	typedef FastCopy<T, !B_NC, BN, BK, SK, SXB, NT> FcB;
	const T* fa_p = nullptr; const T* fb_p = nullptr;
	unsigned fa_x = 0, fb_x = 0;
	bool fa_ok = false, fb_ok = false;
	auto seg_setup = [&]() {
		fa_ok = CPLX || (((psg.a_off + args.a_odd) | (int64_t)psg.lda) & 1) == 0;
		fb_ok = (args.b_rowtab == nullptr) && (CPLX || (((psg.b_off + args.b_odd) | (int64_t)psg.ldb) & 1) == 0);
		if (!A_KC) { fa_ok = fa_ok && (CPLX || ((p_m0 & 1) == 0)); }
		if (B_NC)  { fb_ok = fb_ok && (CPLX || ((p_n0 & 1) == 0)); }
		fa_p = FcA::base(Ag, psg.a_off, psg.lda, p_m0, 0); fa_x = FcA::xinfo(p_m0, p_M);
		fb_p = FcB::base(Bg, psg.b_off, psg.ldb, p_n0, 0); fb_x = FcB::xinfo(p_n0, p_N);
	};
	if (pt < qe) { seg_setup(); }
	auto issue = [&](int stage) {
		const int krem = psg.k - pk0;
		if (fa_ok) { FcA::copy(As + (size_t)stage * Cfg::A_ELEMS, fa_p, psg.lda, fa_x, krem); fa_p += A_KC ? (int64_t)BK : (int64_t)BK * psg.lda; }
		else { load_tile<T, A_KC, BM, BK, SK, SXA, NT>(As + (size_t)stage * Cfg::A_ELEMS, Ag, psg.a_off, psg.lda, p_m0, p_M, pk0, psg.k, false); }
		if (fb_ok) { FcB::copy(Bs + (size_t)stage * Cfg::B_ELEMS, fb_p, psg.ldb, fb_x, krem); fb_p += B_NC ? (int64_t)BK * psg.ldb : (int64_t)BK; }
		else {
			if constexpr (B_NC) { load_tile<T, false, BN, BK, SK, SXB, NT>(Bs + (size_t)stage * Cfg::B_ELEMS, Bg, psg.b_off, psg.ldb, p_n0, p_N, pk0, psg.k, false, args.b_rowtab, args.b_odd); }
			else                { load_tile<T, true,  BN, BK, SK, SXB, NT>(Bs + (size_t)stage * Cfg::B_ELEMS, Bg, psg.b_off, psg.ldb, p_n0, p_N, pk0, psg.k, false); }
		}
		pk0 += BK;
		if (pk0 >= psg.k) {
			pk0 = 0; ps++;
			if (ps < p_seg_end) { psg = args.segs[ps]; }
			else { pt++; producer_enter_tile(); }
			if (pt < qe) { seg_setup(); }
		}
	};

	#pragma unroll
	for (int st = 0; st < STAGES - 1; st++) {
		if (pt < qe) { issue(st); }
		cp_async_commit();
	}

	int gstep = 0;   /* consumer position in the flattened sequence */
	for (int ct = qb; ct < qe; ct++)
	{
		const GemmTile tile = args.tiles[ct];
		double acc[MI][NI][NACC];
		#pragma unroll
		for (int i = 0; i < MI; i++) {
			#pragma unroll
			for (int j = 0; j < NI; j++) {
				#pragma unroll
				for (int c = 0; c < NACC; c++) { acc[i][j][c] = 0.0; }
			}
		}

		for (int step = 0; step < tile.nsteps; step++, gstep++)
		{
			cp_async_wait<STAGES - 2>();
			__syncthreads();
			if (pt < qe) { issue((gstep + STAGES - 1) % STAGES); }
			cp_async_commit();

			const T* as = As + (size_t)(gstep % STAGES) * Cfg::A_ELEMS;
			const T* bs = Bs + (size_t)(gstep % STAGES) * Cfg::B_ELEMS;
			#pragma unroll
			for (int kk = 0; kk < BK; kk += 4)
			{
				T af[MI], bf[NI];
				#pragma unroll
				for (int i = 0; i < MI; i++) {
					const int row = wm0 + 8 * i + lr, k = kk + lc;
					af[i] = A_KC ? as[row * SK + k] : as[k * SXA + row];
				}
				#pragma unroll
				for (int j = 0; j < NI; j++) {
					const int col = wn0 + 8 * j + lr, k = kk + lc;
					bf[j] = B_NC ? bs[k * SXB + col] : bs[col * SK + k];
				}
				if constexpr (!CPLX)
				{
					#pragma unroll
					for (int i = 0; i < MI; i++) {
						#pragma unroll
						for (int j = 0; j < NI; j++) { dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]); }
					}
				}
				else
				{
					/* (ar + i ai)(br + i bi): re += ar br - ai bi, im += ar bi + ai br; conjugation flips the sign of ai / bi */
					double ar[MI], ai[MI], nai[MI];
					#pragma unroll
					for (int i = 0; i < MI; i++) {
						ar[i] = af[i].x;
						ai[i] = args.conj_a ? -af[i].y : af[i].y;
						nai[i] = -ai[i];
					}
					#pragma unroll
					for (int j = 0; j < NI; j++) {
						const double br = bf[j].x;
						const double bi = args.conj_b ? -bf[j].y : bf[j].y;
						#pragma unroll
						for (int i = 0; i < MI; i++) {
							dmma884(acc[i][j][0], acc[i][j][1], ar[i],  br);
							dmma884(acc[i][j][0], acc[i][j][1], nai[i], bi);
							dmma884(acc[i][j][2], acc[i][j][3], ar[i],  bi);
							dmma884(acc[i][j][2], acc[i][j][3], ai[i],  br);
						}
					}
				}
			}
		}

		/* epilogue (overlaps the in-flight loads of the next tile):
		 *   standard  C(i, j) -> c[c_off + rowtab[i] + coltab[j]]                      (output permutation fused)
		 *   merged    C(i, j) -> c[c_off + rowtab[i] + tab[rowcol[i] + j]], col_tab < 0 (rows of several output blocks) */
		const ctbd_gemm_out out = args.outs[tile.out];
		const int M = out.m, N = out.n, m0 = tile.m0, n0 = tile.n0;
		T* __restrict__ Cg = reinterpret_cast<T*>(args.C) + out.c_off;
		const int32_t* __restrict__ rowtab = args.tab + out.row_tab;
		const bool merged = (out.col_tab < 0);
		#pragma unroll
		for (int i = 0; i < MI; i++)
		{
			const int gr = m0 + wm0 + 8 * i + lr;
			if (gr >= M) { continue; }
			const int64_t ro = rowtab[gr];
			const int32_t* __restrict__ coltab = args.tab + (merged ? rowtab[M + gr] : out.col_tab);
			#pragma unroll
			for (int j = 0; j < NI; j++)
			{
				const int gc = n0 + wn0 + 8 * j + 2 * lc;
				T v0, v1;
				if constexpr (!CPLX) { v0 = acc[i][j][0]; v1 = acc[i][j][1]; }
				else { v0 = make_double2(acc[i][j][0], acc[i][j][2]); v1 = make_double2(acc[i][j][1], acc[i][j][3]); }
				const int64_t o0 = out.c_off + ro + (gc < N ? coltab[gc] : 0), o1 = out.c_off + ro + (gc + 1 < N ? coltab[gc + 1] : 0);
				if (args.mc) {
					/* fused all-gather over NVSwitch multicast: ONE store per element (16 bytes where two neighbouring columns are adjacent
					 * in the packed layout), replicated into the result buffers of all GPUs by the switch */
					if constexpr (!CPLX) {
						double* Cm = reinterpret_cast<double*>(args.C);
						const bool pair = (args.c_odd == 0) && (gc + 1 < N) && (o1 == o0 + 1) && ((o0 & 1) == 0);
						if (pair) { mc_st16(Cm + o0, v0, v1); }
						else {
							if (gc < N)     { mc_st8(Cm + o0, v0); }
							if (gc + 1 < N) { mc_st8(Cm + o1, v1); }
						}
					}
					else {
						double2* Cm = reinterpret_cast<double2*>(args.C);
						if (gc < N)     { mc_st16(Cm + o0, v0.x, v0.y); }
						if (gc + 1 < N) { mc_st16(Cm + o1, v1.x, v1.y); }
					}
				}
				else if (args.ndst > 1) {
					/* fused all-gather: the same element goes to the local result and to the peer-mapped result buffers of the other GPUs
					 * (posted NVLink stores).  Two neighbouring columns travel as ONE 16-byte store where the layout allows it: full
					 * sectors on the link instead of byte-masked halves. */
					bool pair = false;
					if constexpr (!CPLX) { pair = (args.c_odd == 0) && (gc + 1 < N) && (o1 == o0 + 1) && ((o0 & 1) == 0); }
					if (pair) {
						if constexpr (!CPLX) {
							const double2 vv = make_double2(v0, v1);
							*reinterpret_cast<double2*>(reinterpret_cast<T*>(args.C) + o0) = vv;
							for (int d = 0; d < args.ndst - 1; d++) { *reinterpret_cast<double2*>(reinterpret_cast<T*>(args.Cx[d]) + o0) = vv; }
						}
					}
					else {
						if (gc < N)     { Cg[o0 - out.c_off] = v0; }
						if (gc + 1 < N) { Cg[o1 - out.c_off] = v1; }
						for (int d = 0; d < args.ndst - 1; d++) {
							T* __restrict__ Cp = reinterpret_cast<T*>(args.Cx[d]);
							if (gc < N)     { Cp[o0] = v0; }
							if (gc + 1 < N) { Cp[o1] = v1; }
						}
					}
				}
				else {
					if (gc < N)     { Cg[o0 - out.c_off] = v0; }
					if (gc + 1 < N) { Cg[o1 - out.c_off] = v1; }
				}
			}
		}
	}
	cp_async_wait<0>();
}

/* dst[i] = idx[i] >= 0 ? src[idx[i]] : 0  (packs a small operand once per plan) */
template <typename T>
__global__ void gather_kernel(int64_t n, const int64_t* __restrict__ idx, const T* __restrict__ src, T* __restrict__ dst)
{
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
		const int64_t k = idx[i];
		T v; memset(&v, 0, sizeof(T));
		if (k >= 0) { v = src[k]; }
		dst[i] = v;
	}
}


/* ============================================================================================== */
/* mixing form: C rows = (small sparse constant matrix) x (gathered B rows), streamed at HBM speed   */
/* ============================================================================================== */

static constexpr int MIX_THREADS = 256;
static constexpr int MIX_U = 4;                       /* columns per thread */
static constexpr int MIX_COLS = MIX_THREADS * MIX_U;  /* columns per tile */

__device__ __forceinline__ void mix_fma(double& acc, const double v, const double b, int, int) { acc = fma(v, b, acc); }
__device__ __forceinline__ void mix_fma(double2& acc, const double2 v, const double2 b, int conj_a, int conj_b)
{
	const double vi = conj_a ? -v.y : v.y, bi = conj_b ? -b.y : b.y;
	acc.x = fma(v.x, b.x, acc.x); acc.x = fma(-vi, bi, acc.x);
	acc.y = fma(v.x, bi, acc.y);  acc.y = fma(vi, b.x, acc.y);
}
__device__ __forceinline__ bool mix_nonzero(const double v)  { return v != 0.0; }
__device__ __forceinline__ bool mix_nonzero(const double2 v) { return v.x != 0.0 || v.y != 0.0; }

/* one thread per row: compact the non-zero entries of the packed operand row (run once per plan) */
template <typename T>
__global__ void mix_compact_kernel(int nrows, const ctbd_mix_row* __restrict__ rows, const int32_t* __restrict__ row_kp,
	const T* __restrict__ apacked, int32_t* __restrict__ nzk, T* __restrict__ nzv, int32_t* __restrict__ cnt)
{
	const int r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= nrows) { return; }
	const int64_t a0 = rows[r].a_off;
	const int kp = row_kp[r];
	int c = 0;
	for (int k = 0; k < kp; k++) {
		const T v = apacked[a0 + k];
		if (mix_nonzero(v)) { nzk[a0 + c] = k; nzv[a0 + c] = v; c++; }
	}
	cnt[r] = c;
}

struct MixArgs
{
	const MixTile* tiles; int ntiles;
	const ctbd_mix_group* groups;
	const ctbd_mix_row* rows;
	const int64_t* brow;
	const int32_t* nzk; const void* nzv; const int32_t* cnt;
	const void* B; void* C;
	int conj_a, conj_b;
};

template <typename T>
__global__ void __launch_bounds__(MIX_THREADS) mix_kernel(const MixArgs args)
{
	const T* __restrict__ Bg = reinterpret_cast<const T*>(args.B);
	T* __restrict__ Cg = reinterpret_cast<T*>(args.C);
	const T* __restrict__ nzv = reinterpret_cast<const T*>(args.nzv);
	for (int tile = blockIdx.x; tile < args.ntiles; tile += gridDim.x)
	{
		const MixTile t = args.tiles[tile];
		const ctbd_mix_group g = args.groups[t.group];
		/* this thread's columns and their digits over the free axes of the right operand */
		int jc[MIX_U]; bool ok[MIX_U]; int dig[MIX_U][4];
		#pragma unroll
		for (int u = 0; u < MIX_U; u++) {
			const int j = t.col0 + (int)threadIdx.x + u * MIX_THREADS;
			ok[u] = (j < g.n);
			jc[u] = ok[u] ? j : 0;
			int rem = jc[u];
			#pragma unroll
			for (int a = 3; a >= 0; a--) {
				if (a < g.ndig) { dig[u][a] = rem % g.dig_dim[a]; rem /= g.dig_dim[a]; } else { dig[u][a] = 0; }
			}
		}
		const int64_t* __restrict__ brow = args.brow + g.brow_begin;
		for (int r = g.row_begin; r < g.row_end; r++)
		{
			const ctbd_mix_row rw = args.rows[r];
			const int cnt = args.cnt[r];
			T acc[MIX_U];
			#pragma unroll
			for (int u = 0; u < MIX_U; u++) { memset(&acc[u], 0, sizeof(T)); }
			for (int q = 0; q < cnt; q++)
			{
				const T v = nzv[rw.a_off + q];
				const T* __restrict__ b = Bg + brow[args.nzk[rw.a_off + q]];
				T bv[MIX_U];
				#pragma unroll
				for (int u = 0; u < MIX_U; u++) { bv[u] = b[jc[u]]; }
				#pragma unroll
				for (int u = 0; u < MIX_U; u++) { mix_fma(acc[u], v, bv[u], args.conj_a, args.conj_b); }
			}
			#pragma unroll
			for (int u = 0; u < MIX_U; u++) {
				if (ok[u]) {
					const int64_t off = rw.c_off + (int64_t)dig[u][0] * rw.cs[0] + (int64_t)dig[u][1] * rw.cs[1] + (int64_t)dig[u][2] * rw.cs[2] + (int64_t)dig[u][3] * rw.cs[3];
					Cg[off] = acc[u];
				}
			}
		}
	}
}

/* ---- tile classes: one class per plan, chosen for the least padded work ---- */

/* real:    0 = 64x64 (4 warps of 32x32), 1 = 32x32 (4 warps of 16x16), 2 = 128x128 (8 warps of 64x32),
 *          3 = 64x128 (8 warps of 32x32), 4 = 32x128 (4 warps of 32x32, BK 8) for skinny M,
 *          5 = 64x128 (4 warps of 32x64: tile and warp shape of cuBLAS' cutlass_80_tensorop_d884gemm_64x128_16x3),
 *          6 = 64x64 with 3 stages (4 CTAs per SM instead of 2)
 * complex: 0 = 64x32 (4 warps of 32x16), 1 = 32x32 (4 warps of 16x16), 2 = 32x64 (4 warps of 32x16, BK 8) */
typedef TileCfg<double, 64, 64, 32, 32, 4>      CfgD0;
typedef TileCfg<double, 32, 32, 16, 16, 4>      CfgD1;
typedef TileCfg<double, 128, 128, 64, 32, 3>    CfgD2;
typedef TileCfg<double, 64, 128, 32, 32, 3>     CfgD3;
typedef TileCfg<double, 32, 128, 32, 32, 4, 8>  CfgD4;
typedef TileCfg<double, 64, 128, 32, 64, 3>     CfgD5;
typedef TileCfg<double, 64, 64, 32, 32, 3>      CfgD6;
typedef TileCfg<double, 64, 64, 32, 32, 3, 32>  CfgD7;   /* 64x64 with 32-deep k chunks: half the barriers per flop, two CTAs per SM */
typedef TileCfg<double2, 64, 32, 32, 16, 3>     CfgZ0;
typedef TileCfg<double2, 32, 32, 16, 16, 3>     CfgZ1;
typedef TileCfg<double2, 32, 64, 32, 16, 4, 8>  CfgZ2;

/* eff: relative cost per padded multiply-add of the class */
struct ClassShape { int bm, bn, bk; double eff; };
/* eff from the measured large-block throughput of each class (tools/gemm_sweep.py, profiles/) */
static const ClassShape g_shapes_d[8] = { { 64, 64, 16, 1.03 }, { 32, 32, 16, 1.22 }, { 128, 128, 16, 1.09 }, { 64, 128, 16, 1.075 }, { 32, 128, 8, 1.08 }, { 64, 128, 16, 1.0 }, { 64, 64, 16, 1.04 }, { 64, 64, 32, 1.5 } };
static const ClassShape g_shapes_z[3] = { { 64, 32, 16, 1.0 }, { 32, 32, 16, 1.06 }, { 32, 64, 8, 1.03 } };

template <typename T, typename Cfg>
static void (*select_kernel(const GemmPlan* p))(const GemmArgs)
{
	if (p->a_kcontig) { return p->b_ncontig ? grouped_gemm_kernel<T, Cfg, true, true>  : grouped_gemm_kernel<T, Cfg, true, false>; }
	return p->b_ncontig ? grouped_gemm_kernel<T, Cfg, false, true> : grouped_gemm_kernel<T, Cfg, false, false>;
}

/* resident CTAs per SM of the plan's kernel (also opts the kernel into its dynamic shared memory size) */
template <typename T, typename Cfg>
static int occupancy_cfg(const GemmPlan* p, int* occ)
{
	auto kern = select_kernel<T, Cfg>(p);
	CTBD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
	CTBD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, Cfg::NT, Cfg::SMEM));
	return 0;
}

template <typename T, typename Cfg>
static int launch_cfg(const GemmPlan* p, const GemmArgs& args)
{
	auto kern = select_kernel<T, Cfg>(p);
	kern<<<p->grid, Cfg::NT, Cfg::SMEM, rt().stream>>>(args);
	CTBD_LAUNCH_CHECK();
	return 0;
}

#define CTBD_GEMM_DISPATCH(FN, ...) \
	(p->dtype == CTBD_F64 \
		? (p->cfg == 0 ? FN<double, CfgD0>(__VA_ARGS__) : p->cfg == 1 ? FN<double, CfgD1>(__VA_ARGS__) : p->cfg == 2 ? FN<double, CfgD2>(__VA_ARGS__) \
			: p->cfg == 3 ? FN<double, CfgD3>(__VA_ARGS__) : p->cfg == 4 ? FN<double, CfgD4>(__VA_ARGS__) : p->cfg == 5 ? FN<double, CfgD5>(__VA_ARGS__) \
			: p->cfg == 6 ? FN<double, CfgD6>(__VA_ARGS__) : FN<double, CfgD7>(__VA_ARGS__)) \
		: (p->cfg == 0 ? FN<double2, CfgZ0>(__VA_ARGS__) : p->cfg == 1 ? FN<double2, CfgZ1>(__VA_ARGS__) : FN<double2, CfgZ2>(__VA_ARGS__)))

} // namespace ctbd

using namespace ctbd;

extern "C" {

int ctbd_gemm_plan_create(const struct ctbd_gemm_plan_host* h, void** plan_out)
{
	CTBD_REQUIRE_INIT();
	if (h->dtype != CTBD_F64 && h->dtype != CTBD_C128) { return fail_msg("grouped GEMM: unsupported dtype"); }
	GemmPlan* p = new GemmPlan();
	p->dtype = h->dtype; p->a_kcontig = h->a_kcontig; p->b_ncontig = h->b_ncontig; p->conj_a = h->conj_a; p->conj_b = h->conj_b;
	p->nouts = h->nouts; p->nsegs = h->nsegs; p->ntab = h->ntab;

	const bool cplx = (h->dtype == CTBD_C128);
	if (h->mix_groups != nullptr && h->n_mix_groups > 0)
	{
		/* mixing form: upload the group / row descriptors, gather the packed operand, compact its non-zeros, list the column tiles */
		if (h->a_gather == nullptr || h->b_rowtab == nullptr) { delete p; return fail_msg("mixing plan needs a packed left operand and a B row table"); }
		const size_t esize = cplx ? 16 : 8;
		p->n_mix_groups = h->n_mix_groups; p->n_mix_rows = h->n_mix_rows;
		int rc = 0;
		rc |= upload(h->mix_groups, (size_t)h->n_mix_groups * sizeof(ctbd_mix_group), (void**)&p->mix_groups);
		rc |= upload(h->mix_rows, (size_t)h->n_mix_rows * sizeof(ctbd_mix_row), (void**)&p->mix_rows);
		rc |= upload(h->b_rowtab, (size_t)h->n_b_rowtab * sizeof(int64_t), (void**)&p->b_rowtab);
		std::vector<MixTile> tiles;
		std::vector<int32_t> row_kp((size_t)h->n_mix_rows, 0);
		std::vector<std::pair<double, int>> order((size_t)h->n_mix_groups);
		for (int g = 0; g < h->n_mix_groups; g++) {
			const ctbd_mix_group& mg = h->mix_groups[g];
			order[g] = std::make_pair((double)(mg.row_end - mg.row_begin) * mg.kp, g);
			for (int r = mg.row_begin; r < mg.row_end; r++) { row_kp[r] = mg.kp; }
		}
		std::stable_sort(order.begin(), order.end(), [](const std::pair<double, int>& a, const std::pair<double, int>& b) { return a.first > b.first; });
		for (const auto& og : order) {
			const ctbd_mix_group& mg = h->mix_groups[og.second];
			for (int c0 = 0; c0 < mg.n; c0 += MIX_COLS) { MixTile t; t.group = og.second; t.col0 = c0; tiles.push_back(t); }
		}
		p->n_mix_tiles = (int)tiles.size();
		p->ntiles = p->n_mix_tiles;
		rc |= upload(tiles.data(), tiles.size() * sizeof(MixTile), (void**)&p->mix_tiles);
		void* idx = nullptr; void* d_kp = nullptr;
		rc |= upload(h->a_gather, (size_t)h->n_a_gather * sizeof(int64_t), &idx);
		rc |= upload(row_kp.data(), row_kp.size() * sizeof(int32_t), &d_kp);
		rc |= ctbd_malloc(&p->a_packed, (size_t)h->n_a_gather * esize);
		rc |= ctbd_malloc((void**)&p->mix_nzk, (size_t)h->n_a_gather * sizeof(int32_t));
		rc |= ctbd_malloc(&p->mix_nzv, (size_t)h->n_a_gather * esize);
		rc |= ctbd_malloc((void**)&p->mix_cnt, (size_t)h->n_mix_rows * sizeof(int32_t));
		if (rc == 0 && h->n_a_gather > 0 && h->n_mix_rows > 0) {
			const int blocks = (int)std::min<int64_t>(ceil_div(h->n_a_gather, 256), 1024);
			const int rblocks = (int)ceil_div(h->n_mix_rows, 128);
			if (cplx) {
				gather_kernel<double2><<<blocks, 256, 0, rt().stream>>>(h->n_a_gather, (const int64_t*)idx, (const double2*)h->a_src, (double2*)p->a_packed);
				mix_compact_kernel<double2><<<rblocks, 128, 0, rt().stream>>>(h->n_mix_rows, p->mix_rows, (const int32_t*)d_kp, (const double2*)p->a_packed, p->mix_nzk, (double2*)p->mix_nzv, p->mix_cnt);
			}
			else {
				gather_kernel<double><<<blocks, 256, 0, rt().stream>>>(h->n_a_gather, (const int64_t*)idx, (const double*)h->a_src, (double*)p->a_packed);
				mix_compact_kernel<double><<<rblocks, 128, 0, rt().stream>>>(h->n_mix_rows, p->mix_rows, (const int32_t*)d_kp, (const double*)p->a_packed, p->mix_nzk, (double*)p->mix_nzv, p->mix_cnt);
			}
			rt().launches += 2;
			if (cudaGetLastError() != cudaSuccess) { rc = -1; }
		}
		ctbd_free(idx); ctbd_free(d_kp);
		if (rc < 0) { ctbd_gemm_plan_destroy(p); return -1; }
		*plan_out = p;
		return 0;
	}
	const ClassShape* shapes = cplx ? g_shapes_z : g_shapes_d;
	const int nshapes = cplx ? 3 : 8;

	/* tile class with the shortest estimated launch: for every class the tiles are packed longest-processing-time-first into
	 * the resident CTA slots (sm_count x occupancy) and the makespan is priced with the class' measured cost per multiply-add;
	 * this accounts for padding, the per-tile overhead AND wave quantisation (few huge tiles leave SMs idle) */
	int best = 0; double best_cost = 0;
	static int occ_cache[2][8] = { { 0 } };
	/* k-steps of every output block for the three chunk depths the classes use (one pass over the segments instead of one per class) */
	std::vector<double> steps_bk[3];      /* 8, 16, 32 */
	for (int q = 0; q < 3; q++) { steps_bk[q].assign((size_t)h->nouts, 0.0); }
	for (int b = 0; b < h->nouts; b++) {
		const ctbd_gemm_out& o = h->outs[b];
		double s8 = 0, s16 = 0, s32 = 0;
		for (int s = o.seg_begin; s < o.seg_end; s++) { const int k = h->segs[s].k; s8 += (double)((k + 7) >> 3); s16 += (double)((k + 15) >> 4); s32 += (double)((k + 31) >> 5); }
		steps_bk[0][b] = s8; steps_bk[1][b] = s16; steps_bk[2][b] = s32;
	}
	/* pass 1: per class the (weight, count) groups of its tiles and a lower bound of the launch cost (perfect split of the padded
	 * work, or the heaviest tile); pass 2: the exact longest-processing-time-first packing only for classes whose lower bound can
	 * still beat the best exact cost found so far -- the exact packing of all eight classes was more than half of the time spent
	 * in this function on the thousand small plans of a D = 1024 sweep */
	struct ClassEval { int c, grid, per_sm; int64_t ntl; double total, wmax, lb; std::vector<std::pair<double, int64_t>> grp; };
	std::vector<ClassEval> evals;
	evals.reserve((size_t)nshapes);
	for (int c = 0; c < nshapes; c++)
	{
		if (occ_cache[cplx][c] == 0) {
			int occ = 1;
			p->cfg = c;
			if (CTBD_GEMM_DISPATCH(occupancy_cfg, p, &occ) < 0) { occ = 1; }
			occ_cache[cplx][c] = occ < 1 ? 1 : occ;
		}
		const int slots = rt().sm_count * occ_cache[cplx][c];
		ClassEval ev;
		ev.c = c; ev.total = 0; ev.wmax = 0; ev.ntl = 0;
		/* all tiles of one output block weigh the same: keep (weight, count) groups instead of one entry per tile */
		ev.grp.reserve((size_t)h->nouts);
		const std::vector<double>& stp = steps_bk[shapes[c].bk == 8 ? 0 : (shapes[c].bk == 16 ? 1 : 2)];
		for (int b = 0; b < h->nouts; b++) {
			const ctbd_gemm_out& o = h->outs[b];
			if (o.m <= 0 || o.n <= 0) { continue; }
			const int64_t nt = ceil_div(o.m, shapes[c].bm) * ceil_div(o.n, shapes[c].bn);
			const double w = stp[b] + 32.0 / shapes[c].bk;
			ev.total += w * (double)nt; ev.wmax = std::max(ev.wmax, w); ev.ntl += nt;
			ev.grp.push_back(std::make_pair(w, nt));
		}
		if (ev.ntl == 0) { continue; }
		ev.grid = (int)std::min<int64_t>(ev.ntl, slots);
		ev.per_sm = (int)ceil_div(ev.grid, rt().sm_count);      /* CTAs sharing the tensor pipe of one SM */
		const double unit = (double)shapes[c].bk * shapes[c].bm * shapes[c].bn * shapes[c].eff * ev.per_sm;
		ev.lb = std::max(ev.total / ev.grid, ev.wmax) * unit;
		evals.push_back(std::move(ev));
	}
	std::stable_sort(evals.begin(), evals.end(), [](const ClassEval& x, const ClassEval& y) { return x.lb < y.lb; });
	bool have_best = false;
	for (ClassEval& ev : evals)
	{
		if (have_best && ev.lb >= best_cost) { break; }      /* sorted by lower bound: none of the remaining classes can win */
		const int c = ev.c;
		const int grid = ev.grid;
		double makespan;
		if (ev.ntl <= 8 * (int64_t)(rt().sm_count * occ_cache[cplx][c])) {
			/* few tiles per slot: the exact longest-processing-time-first packing (wave quantisation matters here) */
			std::sort(ev.grp.begin(), ev.grp.end(), [](const std::pair<double, int64_t>& a, const std::pair<double, int64_t>& b) { return a.first > b.first; });
			std::vector<double> heap(grid, 0.0);     /* min-heap of slot loads */
			auto cmpd = [](double a, double b) { return a > b; };
			for (const auto& g : ev.grp) {
				for (int64_t i = 0; i < g.second; i++) { std::pop_heap(heap.begin(), heap.end(), cmpd); heap.back() += g.first; std::push_heap(heap.begin(), heap.end(), cmpd); }
			}
			makespan = *std::max_element(heap.begin(), heap.end());
		}
		else {
			/* many tiles per slot: the packing ends within about half a (mean) tile of the perfect split */
			makespan = std::max(ev.total / grid + 0.5 * ev.total / (double)ev.ntl, ev.wmax);
		}
		const double cost = makespan * shapes[c].bk * shapes[c].bm * shapes[c].bn * shapes[c].eff * ev.per_sm;
		/* ties go to the lower class index, as with the former scan in class order */
		if (!have_best || cost < best_cost || (cost == best_cost && c < best)) { best = c; best_cost = cost; have_best = true; }
	}
	const char* force = getenv("CTB_GEMM_CLASS");   /* tuning knob: force one tile class */
	if (force != nullptr && atoi(force) >= 0 && atoi(force) < nshapes) { best = atoi(force); }
	p->cfg = best;
	const int bm = shapes[best].bm, bn = shapes[best].bn, bk = shapes[best].bk;

	/* tiles in longest-processing-time-first order: all tiles of an output block weigh the same, so the blocks are sorted (stable,
	 * heaviest first) and their tiles listed block after block */
	struct Item { GemmTile t; double w; };
	std::vector<Item> items;
	{
		std::vector<std::pair<int, int>> blk;      /* (k-steps, output block) */
		blk.reserve((size_t)h->nouts);
		int64_t ntot = 0;
		for (int b = 0; b < h->nouts; b++) {
			const ctbd_gemm_out& o = h->outs[b];
			if (o.m <= 0 || o.n <= 0) { continue; }
			const int nsteps = (int)steps_bk[bk == 8 ? 0 : (bk == 16 ? 1 : 2)][b];
			blk.push_back(std::make_pair(nsteps, b));
			ntot += ceil_div(o.m, bm) * ceil_div(o.n, bn);
		}
		std::stable_sort(blk.begin(), blk.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first > b.first; });
		items.reserve((size_t)ntot);
		for (const auto& pb : blk) {
			const ctbd_gemm_out& o = h->outs[pb.second];
			for (int m0 = 0; m0 < o.m; m0 += bm) {
				for (int n0 = 0; n0 < o.n; n0 += bn) {
					Item it; it.t.out = pb.second; it.t.m0 = m0; it.t.n0 = n0; it.t.nsteps = pb.first;
					it.w = (double)pb.first + 2.0;
					items.push_back(it);
				}
			}
		}
	}
	p->ntiles = (int)items.size();
	int rc = 0;
	rc |= upload(h->outs, (size_t)h->nouts * sizeof(ctbd_gemm_out), (void**)&p->outs);
	rc |= upload(h->segs, (size_t)h->nsegs * sizeof(ctbd_gemm_seg), (void**)&p->segs);
	rc |= upload(h->tab,  (size_t)h->ntab * sizeof(int32_t), (void**)&p->tab);
	if (h->b_rowtab != nullptr && h->n_b_rowtab > 0) {
		if (!h->b_ncontig) { ctbd_gemm_plan_destroy(p); return fail_msg("grouped GEMM: the B row table needs an n-contiguous B operand"); }
		rc |= upload(h->b_rowtab, (size_t)h->n_b_rowtab * sizeof(int64_t), (void**)&p->b_rowtab);
	}
	if (rc == 0 && p->ntiles > 0)
	{
		int occ = 1;
		rc = CTBD_GEMM_DISPATCH(occupancy_cfg, p, &occ);
		if (occ < 1) { occ = 1; }
		const int grid = std::min(p->ntiles, rt().sm_count * occ);
		p->grid = grid;
		/* longest-processing-time-first packing into one queue per resident CTA.  The packing unit is a run of up to 'strip' tiles
		 * that follow each other along n in the same row strip of a block (they share the A tile, so the run also helps L2 / shared
		 * memory locality); 'strip' keeps about 16 units per queue, which is plenty for the balance and cuts the packing cost */
		const int strip = (int)std::max<int64_t>(1, std::min<int64_t>(8, (int64_t)p->ntiles / (16 * (int64_t)grid)));
		struct Unit { int first, count; double w; };
		std::vector<Unit> units;
		units.reserve((size_t)p->ntiles / strip + 16);
		for (int i = 0; i < p->ntiles; ) {
			int j = i + 1;
			while (j < p->ntiles && j - i < strip && items[j].t.out == items[i].t.out && items[j].t.m0 == items[i].t.m0) { j++; }
			Unit u; u.first = i; u.count = j - i; u.w = items[i].w * (j - i);
			units.push_back(u);
			i = j;
		}
		std::stable_sort(units.begin(), units.end(), [](const Unit& a, const Unit& b) { return a.w > b.w; });
		std::vector<std::vector<int>> bins(grid);
		/* binary heap of (load, bin) */
		std::vector<std::pair<double, int>> heap(grid);
		for (int g = 0; g < grid; g++) { heap[g] = std::make_pair(0.0, g); }
		auto cmp = [](const std::pair<double, int>& a, const std::pair<double, int>& b) { return a.first > b.first || (a.first == b.first && a.second > b.second); };
		std::make_heap(heap.begin(), heap.end(), cmp);
		for (const Unit& u : units) {
			std::pop_heap(heap.begin(), heap.end(), cmp);
			std::pair<double, int>& top = heap.back();
			for (int i = 0; i < u.count; i++) { bins[top.second].push_back(u.first + i); }
			top.first += u.w;
			std::push_heap(heap.begin(), heap.end(), cmp);
		}
		std::vector<GemmTile> tl; tl.reserve(p->ntiles);
		std::vector<int32_t> queue(grid + 1, 0);
		for (int g = 0; g < grid; g++) {
			/* inside a queue: ascending output order for L2 locality of the operand blocks */
			std::sort(bins[g].begin(), bins[g].end(), [&](int a, int b) {
				const GemmTile& x = items[a].t; const GemmTile& y = items[b].t;
				if (x.out != y.out) { return x.out < y.out; }
				if (x.m0 != y.m0) { return x.m0 < y.m0; }
				return x.n0 < y.n0; });
			for (int i : bins[g]) { tl.push_back(items[i].t); }
			queue[g + 1] = (int32_t)tl.size();
		}
		rc |= upload(tl.data(), tl.size() * sizeof(GemmTile), (void**)&p->tiles);
		rc |= upload(queue.data(), queue.size() * sizeof(int32_t), (void**)&p->queue);
	}
	if (rc == 0 && h->a_gather != nullptr && h->n_a_gather > 0)
	{
		/* plan-owned packed A operand: gathered once from the (constant) source operand */
		const size_t esize = cplx ? 16 : 8;
		void* idx = nullptr;
		rc |= upload(h->a_gather, (size_t)h->n_a_gather * sizeof(int64_t), &idx);
		rc |= ctbd_malloc(&p->a_packed, (size_t)h->n_a_gather * esize);
		if (rc == 0) {
			const int blocks = (int)std::min<int64_t>(ceil_div(h->n_a_gather, 256), 1024);
			if (cplx) { gather_kernel<double2><<<blocks, 256, 0, rt().stream>>>(h->n_a_gather, (const int64_t*)idx, (const double2*)h->a_src, (double2*)p->a_packed); }
			else      { gather_kernel<double><<<blocks, 256, 0, rt().stream>>>(h->n_a_gather, (const int64_t*)idx, (const double*)h->a_src, (double*)p->a_packed); }
			rt().launches++;
			if (cudaGetLastError() != cudaSuccess) { rc = -1; }
		}
		ctbd_free(idx);
	}
	if (rc < 0) { ctbd_gemm_plan_destroy(p); return -1; }
	*plan_out = p;
	return 0;
}

int ctbd_gemm_plan_destroy(void* plan)
{
	GemmPlan* p = (GemmPlan*)plan;
	if (p == nullptr) { return 0; }
	ctbd_free(p->tiles); ctbd_free(p->queue); ctbd_free(p->a_packed); ctbd_free(p->b_rowtab);
	ctbd_free(p->mix_groups); ctbd_free(p->mix_rows); ctbd_free(p->mix_tiles); ctbd_free(p->mix_nzk); ctbd_free(p->mix_nzv); ctbd_free(p->mix_cnt);
	ctbd_free(p->outs); ctbd_free(p->segs); ctbd_free(p->tab);
	delete p;
	return 0;
}

int ctbd_gemm_plan_info(void* plan, int* ntiles, int* nlaunches)
{
	GemmPlan* p = (GemmPlan*)plan;
	if (ntiles != nullptr) { *ntiles = p->ntiles; }
	if (nlaunches != nullptr) { *nlaunches = p->ntiles > 0 ? 1 : 0; }
	return 0;
}

int ctbd_gemm_run(void* plan, const void* A, const void* B, void* C)
{
	GemmPlan* p = (GemmPlan*)plan;
	if (p->ntiles == 0) { return 0; }
	if (p->n_mix_tiles > 0)
	{
		MixArgs ma;
		ma.tiles = p->mix_tiles; ma.ntiles = p->n_mix_tiles; ma.groups = p->mix_groups; ma.rows = p->mix_rows; ma.brow = p->b_rowtab;
		ma.nzk = p->mix_nzk; ma.nzv = p->mix_nzv; ma.cnt = p->mix_cnt; ma.B = B; ma.C = C; ma.conj_a = p->conj_a; ma.conj_b = p->conj_b;
		const int grid = std::min(p->n_mix_tiles, rt().sm_count * 32);
		if (p->dtype == CTBD_C128) { mix_kernel<double2><<<grid, MIX_THREADS, 0, rt().stream>>>(ma); }
		else                       { mix_kernel<double><<<grid, MIX_THREADS, 0, rt().stream>>>(ma); }
		CTBD_LAUNCH_CHECK();
		return 0;
	}
	GemmArgs args;
	args.tiles = p->tiles; args.queue = p->queue;
	args.outs = p->outs; args.segs = p->segs; args.tab = p->tab;
	args.A = (p->a_packed != nullptr) ? p->a_packed : A; args.B = B; args.C = C;
	args.conj_a = p->conj_a; args.conj_b = p->conj_b;
	args.a_odd = (int)(((uintptr_t)args.A >> 3) & 1); args.b_odd = (int)(((uintptr_t)args.B >> 3) & 1);
	args.b_rowtab = p->b_rowtab;
	args.ndst = 1; args.mc = 0; args.c_odd = 0;
	return CTBD_GEMM_DISPATCH(launch_cfg, p, args);
}

int ctbd_gemm_run_multi(void* plan, const void* A, const void* B, int ndst, void* const* Cs)
{
	GemmPlan* p = (GemmPlan*)plan;
	if (ndst < 1 || ndst > 8) { return fail_msg("grouped GEMM: between 1 and 8 destinations"); }
	if (p->n_mix_tiles > 0) { return fail_msg("grouped GEMM: the mixing form has a single destination"); }
	if (p->ntiles == 0) { return 0; }
	GemmArgs args;
	args.tiles = p->tiles; args.queue = p->queue;
	args.outs = p->outs; args.segs = p->segs; args.tab = p->tab;
	args.A = (p->a_packed != nullptr) ? p->a_packed : A; args.B = B; args.C = Cs[0];
	args.conj_a = p->conj_a; args.conj_b = p->conj_b;
	args.a_odd = (int)(((uintptr_t)args.A >> 3) & 1); args.b_odd = (int)(((uintptr_t)args.B >> 3) & 1);
	args.b_rowtab = p->b_rowtab;
	args.ndst = ndst; args.mc = 0;
	args.c_odd = 0;
	for (int d = 0; d < ndst; d++) { if ((((uintptr_t)Cs[d]) >> 3) & 1) { args.c_odd = 1; } }
	for (int d = 1; d < ndst; d++) { args.Cx[d - 1] = Cs[d]; }
	return CTBD_GEMM_DISPATCH(launch_cfg, p, args);
}

int ctbd_gemm_run_mc(void* plan, const void* A, const void* B, void* C_mc)
{
	GemmPlan* p = (GemmPlan*)plan;
	if (p->n_mix_tiles > 0) { return fail_msg("grouped GEMM: the mixing form has no multicast epilogue"); }
	if (p->ntiles == 0) { return 0; }
	GemmArgs args;
	args.tiles = p->tiles; args.queue = p->queue;
	args.outs = p->outs; args.segs = p->segs; args.tab = p->tab;
	args.A = (p->a_packed != nullptr) ? p->a_packed : A; args.B = B; args.C = C_mc;
	args.conj_a = p->conj_a; args.conj_b = p->conj_b;
	args.a_odd = (int)(((uintptr_t)args.A >> 3) & 1); args.b_odd = (int)(((uintptr_t)args.B >> 3) & 1);
	args.b_rowtab = p->b_rowtab;
	args.ndst = 1; args.mc = 1;
	args.c_odd = (int)(((uintptr_t)C_mc >> 3) & 1);
	return CTBD_GEMM_DISPATCH(launch_cfg, p, args);
}

} // extern "C"
