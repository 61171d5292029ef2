/*
 * ctbd_gemm.cu -- grouped block GEMM on the FP64 tensor pipe (DMMA) for sm_100a.
 *
 * Replaces the per-block cblas_dgemm/zgemm swarm of the reference's block_sparse_tensor_dot
 * (src/tensor/block_sparse_tensor.c:1935-1994 -> dense_tensor_dot_update, src/tensor/dense_tensor.c:1761-1828)
 * by ONE launch per tile class over a device-resident work list:
 *   - an "output block" C (m x n) is the sum over its "segments" (contracted sector tuples, in the
 *     reference's row-major order) of op(A_seg) op(B_seg);
 *   - a CTA owns one tile of one output block and walks the concatenated K extent of all segments
 *     through a multi-stage cp.async (LDGSTS) shared-memory ring, so that many tiny contracted
 *     sectors still keep the pipeline full;
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

struct GemmTile { int32_t out, m0, n0, pad_; };

struct GemmClass
{
	int cfg = 0;
	int ntiles = 0;
	GemmTile* tiles = nullptr;    /* device */
};

struct GemmPlan
{
	int dtype = 0, a_kcontig = 0, b_ncontig = 0, conj_a = 0, conj_b = 0;
	int nouts = 0, nsegs = 0, ntab = 0;
	ctbd_gemm_out* outs = nullptr;     /* device */
	ctbd_gemm_seg* segs = nullptr;
	int32_t* tab = nullptr;
	std::vector<GemmClass> classes;
	int ntiles_total = 0;
};

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

template <typename T, int BM_, int BN_, int WM_, int WN_, int STAGES_>
struct TileCfg
{
	static constexpr int BM = BM_, BN = BN_, WM = WM_, WN = WN_, STAGES = STAGES_;
	static constexpr int BK = 16;
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
	const int x0, const int X, const int k0, const int K, const bool vec2)
{
	const int tid = threadIdx.x;
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

struct GemmArgs
{
	const GemmTile* tiles;
	const ctbd_gemm_out* outs;
	const ctbd_gemm_seg* segs;
	const int32_t* tab;
	const void* A; const void* B; void* C;
	int conj_a, conj_b;
};

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

	const GemmTile tile = args.tiles[blockIdx.x];
	const ctbd_gemm_out out = args.outs[tile.out];
	const int M = out.m, N = out.n, m0 = tile.m0, n0 = tile.n0;
	const T* __restrict__ Ag = reinterpret_cast<const T*>(args.A);
	const T* __restrict__ Bg = reinterpret_cast<const T*>(args.B);

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int wm0 = (warp / (BN / WN)) * WM, wn0 = (warp % (BN / WN)) * WN;
	const int lr = lane >> 2, lc = lane & 3;

	double acc[MI][NI][NACC];
	#pragma unroll
	for (int i = 0; i < MI; i++) {
		#pragma unroll
		for (int j = 0; j < NI; j++) {
			#pragma unroll
			for (int c = 0; c < NACC; c++) { acc[i][j][c] = 0.0; }
		}
	}

	/* total pipeline steps over the concatenated K extent of all segments */
	int total = 0;
	for (int s = out.seg_begin; s < out.seg_end; s++) { total += (args.segs[s].k + BK - 1) / BK; }

	/* producer iterator */
	int ps = out.seg_begin, pk0 = 0;
	auto issue = [&](int stage) {
		const ctbd_gemm_seg sg = args.segs[ps];
		const bool va = !CPLX && ((sg.a_off | (int64_t)sg.lda) & 1) == 0;
		const bool vb = !CPLX && ((sg.b_off | (int64_t)sg.ldb) & 1) == 0;
		load_tile<T, A_KC,  BM, BK, SK, SXA, NT>(As + (size_t)stage * Cfg::A_ELEMS, Ag, sg.a_off, sg.lda, m0, M, pk0, sg.k, va);
		load_tile<T, !B_NC, BN, BK, SK, SXB, NT>(Bs + (size_t)stage * Cfg::B_ELEMS, Bg, sg.b_off, sg.ldb, n0, N, pk0, sg.k, vb);
		pk0 += BK;
		if (pk0 >= sg.k) { ps++; pk0 = 0; }
	};

	#pragma unroll
	for (int st = 0; st < STAGES - 1; st++) {
		if (st < total) { issue(st); }
		cp_async_commit();
	}

	for (int step = 0; step < total; step++)
	{
		cp_async_wait<STAGES - 2>();
		__syncthreads();
		if (step + STAGES - 1 < total) { issue((step + STAGES - 1) % STAGES); }
		cp_async_commit();

		const T* as = As + (size_t)(step % STAGES) * Cfg::A_ELEMS;
		const T* bs = Bs + (size_t)(step % STAGES) * Cfg::B_ELEMS;
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
	cp_async_wait<0>();

	/* epilogue: C(i, j) -> c[c_off + rowtab[i] + coltab[j]] (output permutation fused) */
	T* __restrict__ Cg = reinterpret_cast<T*>(args.C) + out.c_off;
	const int32_t* __restrict__ rowtab = args.tab + out.row_tab;
	const int32_t* __restrict__ coltab = args.tab + out.col_tab;
	#pragma unroll
	for (int i = 0; i < MI; i++)
	{
		const int gr = m0 + wm0 + 8 * i + lr;
		if (gr >= M) { continue; }
		const int64_t ro = rowtab[gr];
		#pragma unroll
		for (int j = 0; j < NI; j++)
		{
			const int gc = n0 + wn0 + 8 * j + 2 * lc;
			if constexpr (!CPLX) {
				if (gc < N)     { Cg[ro + coltab[gc]]     = acc[i][j][0]; }
				if (gc + 1 < N) { Cg[ro + coltab[gc + 1]] = acc[i][j][1]; }
			}
			else {
				if (gc < N)     { Cg[ro + coltab[gc]]     = make_double2(acc[i][j][0], acc[i][j][2]); }
				if (gc + 1 < N) { Cg[ro + coltab[gc + 1]] = make_double2(acc[i][j][1], acc[i][j][3]); }
			}
		}
	}
}

/* ---- tile classes ---- */

/* real: 0 = 64x64 (4 warps of 32x32), 1 = 32x32 (4 warps of 16x16), 2 = 128x128 (8 warps of 64x32)
 * complex: 0 = 64x32 (4 warps of 32x16), 1 = 32x32 (4 warps of 16x16) */
typedef TileCfg<double, 64, 64, 32, 32, 4>   CfgD0;
typedef TileCfg<double, 32, 32, 16, 16, 4>   CfgD1;
typedef TileCfg<double, 128, 128, 64, 32, 3> CfgD2;
typedef TileCfg<double2, 64, 32, 32, 16, 3>  CfgZ0;
typedef TileCfg<double2, 32, 32, 16, 16, 3>  CfgZ1;

struct ClassShape { int bm, bn; double eff; };
static const ClassShape g_shapes_d[3] = { { 64, 64, 1.0 }, { 32, 32, 1.35 }, { 128, 128, 0.9 } };
static const ClassShape g_shapes_z[2] = { { 64, 32, 1.0 }, { 32, 32, 1.25 } };

template <typename T, typename Cfg>
static int launch_cfg(const GemmPlan* p, const GemmClass& cl, const GemmArgs& args)
{
	void (*kern)(const GemmArgs) = nullptr;
	if (p->a_kcontig) { kern = p->b_ncontig ? grouped_gemm_kernel<T, Cfg, true, true>  : grouped_gemm_kernel<T, Cfg, true, false>; }
	else              { kern = p->b_ncontig ? grouped_gemm_kernel<T, Cfg, false, true> : grouped_gemm_kernel<T, Cfg, false, false>; }
	static bool attr_set[4] = { false, false, false, false };
	const int v = (p->a_kcontig ? 2 : 0) + (p->b_ncontig ? 1 : 0);
	if (!attr_set[v]) {
		CTBD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
		attr_set[v] = true;
	}
	kern<<<cl.ntiles, Cfg::NT, Cfg::SMEM, rt().stream>>>(args);
	CTBD_LAUNCH_CHECK();
	return 0;
}

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
	const ClassShape* shapes = cplx ? g_shapes_z : g_shapes_d;
	const int nshapes = cplx ? 2 : 3;
	const char* force = getenv("CTB_GEMM_CLASS");   /* tuning/debug knob: force one tile class */
	const int forced = force != nullptr ? atoi(force) : -1;

	struct Item { GemmTile t; double w; };
	std::vector<std::vector<Item>> items(nshapes);
	for (int b = 0; b < h->nouts; b++)
	{
		const ctbd_gemm_out& o = h->outs[b];
		if (o.m <= 0 || o.n <= 0) { continue; }
		double ktot = 0;
		for (int s = o.seg_begin; s < o.seg_end; s++) { ktot += h->segs[s].k; }
		/* pick the class with the least padded work (scaled by its relative efficiency) */
		int best = 0; double best_cost = 0;
		for (int c = 0; c < nshapes; c++)
		{
			const double pm = (double)ceil_div(o.m, shapes[c].bm) * shapes[c].bm;
			const double pn = (double)ceil_div(o.n, shapes[c].bn) * shapes[c].bn;
			const double cost = pm * pn * shapes[c].eff;
			if (c == 0 || cost < best_cost) { best = c; best_cost = cost; }
		}
		if (forced >= 0 && forced < nshapes) { best = forced; }
		const int bm = shapes[best].bm, bn = shapes[best].bn;
		for (int m0 = 0; m0 < o.m; m0 += bm) {
			for (int n0 = 0; n0 < o.n; n0 += bn) {
				const double tm = std::min(bm, o.m - m0), tn = std::min(bn, o.n - n0);
				Item it; it.t.out = b; it.t.m0 = m0; it.t.n0 = n0; it.t.pad_ = 0;
				it.w = ktot * (double)(bm * bn) + tm * tn;
				items[best].push_back(it);
			}
		}
	}
	int rc = 0;
	rc |= upload(h->outs, (size_t)h->nouts * sizeof(ctbd_gemm_out), (void**)&p->outs);
	rc |= upload(h->segs, (size_t)h->nsegs * sizeof(ctbd_gemm_seg), (void**)&p->segs);
	rc |= upload(h->tab,  (size_t)h->ntab * sizeof(int32_t), (void**)&p->tab);
	for (int c = 0; c < nshapes && rc == 0; c++)
	{
		if (items[c].empty()) { continue; }
		/* heaviest first: the hardware block scheduler then acts as a longest-processing-time list scheduler */
		std::stable_sort(items[c].begin(), items[c].end(), [](const Item& a, const Item& b) { return a.w > b.w; });
		std::vector<GemmTile> tl(items[c].size());
		for (size_t i = 0; i < tl.size(); i++) { tl[i] = items[c][i].t; }
		GemmClass cl;
		cl.cfg = c; cl.ntiles = (int)tl.size();
		rc |= upload(tl.data(), tl.size() * sizeof(GemmTile), (void**)&cl.tiles);
		p->classes.push_back(cl);
		p->ntiles_total += cl.ntiles;
	}
	if (rc < 0) { ctbd_gemm_plan_destroy(p); return -1; }
	*plan_out = p;
	return 0;
}

int ctbd_gemm_plan_destroy(void* plan)
{
	GemmPlan* p = (GemmPlan*)plan;
	if (p == nullptr) { return 0; }
	for (auto& cl : p->classes) { ctbd_free(cl.tiles); }
	ctbd_free(p->outs); ctbd_free(p->segs); ctbd_free(p->tab);
	delete p;
	return 0;
}

int ctbd_gemm_plan_info(void* plan, int* ntiles, int* nlaunches)
{
	GemmPlan* p = (GemmPlan*)plan;
	if (ntiles != nullptr) { *ntiles = p->ntiles_total; }
	if (nlaunches != nullptr) { *nlaunches = (int)p->classes.size(); }
	return 0;
}

int ctbd_gemm_run(void* plan, const void* A, const void* B, void* C)
{
	GemmPlan* p = (GemmPlan*)plan;
	GemmArgs args;
	args.outs = p->outs; args.segs = p->segs; args.tab = p->tab;
	args.A = A; args.B = B; args.C = C;
	args.conj_a = p->conj_a; args.conj_b = p->conj_b;
	for (const GemmClass& cl : p->classes)
	{
		args.tiles = cl.tiles;
		int rc = 0;
		if (p->dtype == CTBD_F64) {
			switch (cl.cfg) {
				case 0: rc = launch_cfg<double, CfgD0>(p, cl, args); break;
				case 1: rc = launch_cfg<double, CfgD1>(p, cl, args); break;
				default: rc = launch_cfg<double, CfgD2>(p, cl, args); break;
			}
		}
		else {
			switch (cl.cfg) {
				case 0: rc = launch_cfg<double2, CfgZ0>(p, cl, args); break;
				default: rc = launch_cfg<double2, CfgZ1>(p, cl, args); break;
			}
		}
		if (rc < 0) { return rc; }
	}
	return 0;
}

} // extern "C"
