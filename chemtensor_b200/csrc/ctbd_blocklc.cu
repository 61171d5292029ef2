/*
 * ctbd_blocklc.cu -- batched strided block linear combinations (ctbd_lc_*, include/ctb_device.h).
 *
 * The data movement of the SU(2) layer (host/su2.c).  The reference performs these steps sector by sector on the host:
 * su2_tensor_fmove (src/tensor/su2_tensor.c:770-886: copy_dense_tensor / dense_tensor_scalar_multiply_add per pair of charge sectors),
 * su2_tensor_transpose (:647-739: dense_tensor_transpose per sector), the scalings of su2_tensor_reverse_axis_simple (:916-990) and
 * su2_tensor_swap_tree_axes (:467-502), and the (de)normalisation of su2_tensor_(de)serialize_renormalized_entries (:4725-4942).
 * Here ONE launch handles all sectors of a tensor: the work list (blocks, terms, 1024-element chunks) is built once per plan on the
 * host and stays on the device, so an F-move inside the Lanczos loop is a single streaming pass.
 *
 * Mapping: one warp per chunk of up to 1024 destination entries of one block (degeneracy tensors of an SU(2) MPS range from a
 * single entry to millions: a warp granule keeps the tiny ones cheap and the large ones coalesced); consecutive lanes take
 * consecutive destination entries, the multi-index decode runs over the (host-merged) axes of the block.  HBM-bound.
 */
#include "ctbd_common.cuh"

namespace ctbd {

static constexpr int LC_THREADS = 256;
static constexpr int LC_WARPS = LC_THREADS / 32;
static constexpr int64_t LC_CHUNK = 1024;

struct LcChunk { int32_t blk; int32_t pad; int64_t begin, end; };

struct LcPlan
{
	int dtype = 0, conj = 0, nblk = 0, nterm = 0;
	int64_t nchunk = 0;
	ctbd_lc_block* blocks = nullptr;
	ctbd_lc_term* terms = nullptr;
	LcChunk* chunks = nullptr;
};

template <bool CPLX>
__global__ void __launch_bounds__(LC_THREADS) lc_kernel(int64_t nchunk, const LcChunk* __restrict__ chunks, const ctbd_lc_block* __restrict__ blocks,
	const ctbd_lc_term* __restrict__ terms, const double* src, double* dst, int conj)
{
	__shared__ ctbd_lc_block sb[LC_WARPS];
	const int wl = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int64_t w = (int64_t)blockIdx.x * LC_WARPS + wl;
	if (w >= nchunk) { return; }
	const LcChunk c = chunks[w];
	{
		/* the block descriptor of this warp into shared memory, word by word */
		const int32_t* g = reinterpret_cast<const int32_t*>(blocks + c.blk);
		int32_t* s = reinterpret_cast<int32_t*>(&sb[wl]);
		for (int i = lane; i < (int)(sizeof(ctbd_lc_block) / 4); i += 32) { s[i] = g[i]; }
	}
	__syncwarp();
	const ctbd_lc_block& b = sb[wl];
	const int nd = b.ndim;
	const double sgn = (CPLX && conj) ? -1.0 : 1.0;
	/* fast path of the common case (F-moves, scalings, packing): source and destination contiguous.  16 bytes per lane and access,
	 * four independent accesses per term in flight (a streaming kernel needs ~40 KB in flight per SM to cover the HBM latency) */
	if (nd == 1 && b.dstride[0] == 1 && b.sstride[0] == 1)
	{
		bool even = (CPLX || (((b.dst_off + c.begin) & 1) == 0)) && (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0);
		for (int t = b.term_begin; t < b.term_end && even; t++) { even = CPLX || (((terms[t].src_off + c.begin) & 1) == 0); }
		if (even)
		{
			const int64_t n2 = CPLX ? (c.end - c.begin) : (c.end - c.begin) / 2;      /* 16-byte units */
			double2* d2 = reinterpret_cast<double2*>(dst + (CPLX ? 2 : 1) * (b.dst_off + c.begin));
			for (int64_t i0 = 0; i0 < n2; i0 += 128)
			{
				double2 acc[4];
				#pragma unroll
				for (int u = 0; u < 4; u++) { acc[u] = make_double2(0.0, 0.0); }
				for (int t = b.term_begin; t < b.term_end; t++)
				{
					const ctbd_lc_term tm = terms[t];
					const double2* s2 = reinterpret_cast<const double2*>(src + (CPLX ? 2 : 1) * (tm.src_off + c.begin));
					double2 v[4];
					#pragma unroll
					for (int u = 0; u < 4; u++) { const int64_t i = i0 + u * 32 + lane; v[u] = (i < n2) ? s2[i] : make_double2(0.0, 0.0); }
					#pragma unroll
					for (int u = 0; u < 4; u++) { acc[u].x += tm.coef * v[u].x; acc[u].y += tm.coef * v[u].y; }
				}
				#pragma unroll
				for (int u = 0; u < 4; u++) {
					const int64_t i = i0 + u * 32 + lane;
					if (i < n2) { if (CPLX) { acc[u].y *= sgn; } d2[i] = acc[u]; }
				}
			}
			if (!CPLX && ((c.end - c.begin) & 1) && lane == 0)
			{
				const int64_t e = c.end - 1;
				double acc1 = 0;
				for (int t = b.term_begin; t < b.term_end; t++) { acc1 += terms[t].coef * src[terms[t].src_off + e]; }
				dst[b.dst_off + e] = acc1;
			}
			return;
		}
	}
	for (int64_t e = c.begin + lane; e < c.end; e += 32)
	{
		int64_t r = e, doff = b.dst_off, soff = 0;
		for (int a = nd - 1; a > 0; a--)
		{
			const int64_t d = b.dim[a];
			const int64_t q = r / d;
			const int64_t i = r - q * d;
			doff += i * b.dstride[a];
			soff += i * b.sstride[a];
			r = q;
		}
		doff += r * b.dstride[0];
		soff += r * b.sstride[0];
		if (CPLX)
		{
			double re = 0, im = 0;
			for (int t = b.term_begin; t < b.term_end; t++)
			{
				const ctbd_lc_term tm = terms[t];
				const double2 v = *reinterpret_cast<const double2*>(src + 2 * (tm.src_off + soff));
				re += tm.coef * v.x;
				im += tm.coef * v.y;
			}
			*reinterpret_cast<double2*>(dst + 2 * doff) = make_double2(re, sgn * im);
		}
		else
		{
			double acc = 0;
			for (int t = b.term_begin; t < b.term_end; t++)
			{
				const ctbd_lc_term tm = terms[t];
				acc += tm.coef * src[tm.src_off + soff];
			}
			dst[doff] = acc;
		}
	}
}

} // namespace ctbd

using namespace ctbd;

extern "C" int ctbd_lc_plan_create(int dtype, int conj, int nblk, const struct ctbd_lc_block* blocks_host, int nterm, const struct ctbd_lc_term* terms_host, void** plan)
{
	CTBD_REQUIRE_INIT();
	if (dtype != CTBD_F64 && dtype != CTBD_C128) { return fail_msg("ctbd_lc_plan_create: dtype must be f64 or c128"); }
	LcPlan* p = new LcPlan();
	p->dtype = dtype; p->conj = conj; p->nblk = nblk; p->nterm = nterm;
	/* chunk list */
	int64_t nchunk = 0;
	for (int i = 0; i < nblk; i++)
	{
		const ctbd_lc_block& b = blocks_host[i];
		if (b.ndim < 1 || b.ndim > CTBD_LC_MAXDIM) { delete p; return fail_msg("ctbd_lc_plan_create: block rank out of range"); }
		int64_t n = 1;
		for (int a = 0; a < b.ndim; a++) { n *= b.dim[a]; }
		nchunk += ceil_div(n, LC_CHUNK);
	}
	LcChunk* ch = (LcChunk*)malloc((size_t)(nchunk > 0 ? nchunk : 1) * sizeof(LcChunk));
	int64_t k = 0;
	for (int i = 0; i < nblk; i++)
	{
		const ctbd_lc_block& b = blocks_host[i];
		int64_t n = 1;
		for (int a = 0; a < b.ndim; a++) { n *= b.dim[a]; }
		for (int64_t s = 0; s < n; s += LC_CHUNK) {
			ch[k].blk = i; ch[k].pad = 0; ch[k].begin = s; ch[k].end = (s + LC_CHUNK < n) ? s + LC_CHUNK : n;
			k++;
		}
	}
	p->nchunk = nchunk;
	/* ONE device allocation and ONE copy for the three tables (a local solve of the SU(2) layer creates ~14 of these plans;
	 * with three uploads each the plan creation was a fifth of a small-bond local solve) */
	const size_t nb = ((size_t)nblk * sizeof(ctbd_lc_block) + 15) & ~(size_t)15;
	const size_t nt = ((size_t)nterm * sizeof(ctbd_lc_term) + 15) & ~(size_t)15;
	const size_t nc = (size_t)nchunk * sizeof(LcChunk);
	const size_t total = nb + nt + nc;
	int rc = 0;
	if (total > 0)
	{
		char* host = (char*)malloc(total);
		if (nblk > 0)  { memcpy(host, blocks_host, (size_t)nblk * sizeof(ctbd_lc_block)); }
		if (nterm > 0) { memcpy(host + nb, terms_host, (size_t)nterm * sizeof(ctbd_lc_term)); }
		if (nchunk > 0) { memcpy(host + nb + nt, ch, nc); }
		void* dev = nullptr;
		rc = upload(host, total, &dev);
		free(host);
		if (rc == 0) {
			p->blocks = (ctbd_lc_block*)dev;
			p->terms = (ctbd_lc_term*)((char*)dev + nb);
			p->chunks = (LcChunk*)((char*)dev + nb + nt);
		}
	}
	free(ch);
	if (rc < 0) { ctbd_lc_plan_destroy(p); return rc; }
	*plan = p;
	return 0;
}

extern "C" int ctbd_lc_plan_run(void* plan, const void* src, void* dst)
{
	LcPlan* p = (LcPlan*)plan;
	if (p == nullptr) { return fail_msg("ctbd_lc_plan_run: null plan"); }
	if (p->nchunk == 0) { return 0; }
	const int64_t grid = ceil_div(p->nchunk, LC_WARPS);
	if (p->dtype == CTBD_C128) {
		lc_kernel<true><<<(unsigned)grid, LC_THREADS, 0, rt().stream>>>(p->nchunk, p->chunks, p->blocks, p->terms, (const double*)src, (double*)dst, p->conj);
	}
	else {
		lc_kernel<false><<<(unsigned)grid, LC_THREADS, 0, rt().stream>>>(p->nchunk, p->chunks, p->blocks, p->terms, (const double*)src, (double*)dst, 0);
	}
	CTBD_LAUNCH_CHECK();
	return 0;
}

extern "C" int ctbd_lc_plan_destroy(void* plan)
{
	LcPlan* p = (LcPlan*)plan;
	if (p == nullptr) { return 0; }
	if (p->blocks) { ctbd_free(p->blocks); }      /* terms and chunks live in the same allocation */
	delete p;
	return 0;
}
