/*
 * tensor_ops.c -- plan builders (host, integer-only) for the block-sparse primitives.
 *
 * Each primitive derives the sector structure of its result exactly as the
 * reference does and turns the per-block work into ONE device launch:
 *   ctb_dot_prepare   <- block_sparse_tensor_dot, src/tensor/block_sparse_tensor.c:1826-2001
 *                        (+ the block_sparse_tensor_transpose :785 that follows it in
 *                        chain_ops.c, fused as an output permutation)
 *   ctb_transpose     <- block_sparse_tensor_transpose          :785
 *   ctb_flatten_axes  <- block_sparse_tensor_flatten_axes        :950
 *   ctb_split_axis    <- block_sparse_tensor_split_axis          :1123
 *   ctb_slice         <- block_sparse_tensor_slice               :1446
 *   ctb_scale_axis    <- block_sparse_tensor_multiply_pointwise_vector :1654
 *   ctb_svd/qr/rq     <- block_sparse_tensor_svd :2686 / _qr :2402 / _rq :2544
 */
#include "ctb_internal.h"

/* ---------------------------------------------------------------------------------------------- */
/* grouped-GEMM contraction plans                                                                  */
/* ---------------------------------------------------------------------------------------------- */

/* growable int32 table */
struct i32vec { int32_t* v; size_t n, cap; };
static void i32vec_reserve(struct i32vec* t, size_t extra)
{
	while (t->n + extra > t->cap) { t->cap = t->cap ? 2 * t->cap : 4096; t->v = realloc(t->v, t->cap * sizeof(int32_t)); }
}

/* element offsets of the rows (free axes of s, 'first' = 0) or columns (free axes of t, 'first' = nfs) of a natural
 * result block inside the permuted result block */
/* optional embedding of the result into a larger tensor along one result axis (sharded effective Hamiltonian: the rank's column
 * slice is written straight into the packed layout of the full tensor); set by ctb_dot_prepare_embed for the duration of one call */
static _Thread_local const struct ctb_embed* g_embed = NULL;     /* per calling thread: set for the duration of one plan build */
static _Thread_local int g_map_nat = -1;              /* natural axis the position map applies to, -1: none */
static _Thread_local const int32_t* g_map_pos = NULL; /* position inside the piece sector -> position inside the full sector */

static void append_offset_table(struct i32vec* tab, int first, int count, const struct ctb_axis* const* nat, const int* nat_sec,
	const int* pos_of_nat, const ct_long* stride_r, ct_long base)
{
	ct_long total = 1;
	for (int a = 0; a < count; a++) { total *= nat[first + a]->secdim[nat_sec[first + a]]; }
	i32vec_reserve(tab, (size_t)total);
	if (count == 0) { CTB_REQUIRE(base < ((ct_long)1 << 31)); tab->v[tab->n++] = (int32_t)base; return; }
	/* the innermost axis is a plain strided run (or a lookup through the position map); the outer axes advance a digit counter */
	const int last = count - 1;
	const ct_long nlast = nat[first + last]->secdim[nat_sec[first + last]];
	const ct_long slast = stride_r[pos_of_nat[first + last]];
	const bool map_last = (first + last == g_map_nat);
	int dig[CTB_MAXDIM] = { 0 };
	int32_t* out = tab->v + tab->n;
	for (ct_long i = 0; i < total; i += nlast)
	{
		ct_long off = base;
		for (int a = 0; a < last; a++) { off += (first + a == g_map_nat ? (ct_long)g_map_pos[dig[a]] : (ct_long)dig[a]) * stride_r[pos_of_nat[first + a]]; }
		const ct_long hi = off + (map_last ? (nlast > 0 ? (ct_long)g_map_pos[nlast - 1] : 0) : nlast - 1) * slast;
		CTB_REQUIRE(off < ((ct_long)1 << 31) && hi < ((ct_long)1 << 31));
		if (map_last) {
			for (ct_long j = 0; j < nlast; j++) { CTB_REQUIRE(g_map_pos[j] * slast + off < ((ct_long)1 << 31)); out[i + j] = (int32_t)(off + (ct_long)g_map_pos[j] * slast); }
		}
		else {
			for (ct_long j = 0; j < nlast; j++) { out[i + j] = (int32_t)(off + j * slast); }
		}
		for (int a = last - 1; a >= 0; a--) {
			if (++dig[a] < nat[first + a]->secdim[nat_sec[first + a]]) { break; }
			dig[a] = 0;
		}
	}
	tab->n += (size_t)total;
}

/* CTB_TRACE_PLAN: [0] result tensor, [1] host lists incl. [0], [2] device plan, [3] plans, [4] output blocks, [5] table entries */
double ctb_plan_profile[16] = { 0 };
/* fine-grained timers only when tracing (they sit inside the per-output-block loop) */
static int plan_trace_on(void) { static int on = -1; if (on < 0) { on = (getenv("CTB_TRACE_PLAN") != NULL) ? 1 : 0; } return on; }
#define TP_NOW() (plan_trace_on() ? ctb_wall_ms() : 0.0)

/* Offset tables depend only on the extents and strides of their axes (and the base offset): blocks of one block row / column with
 * equal sector dimensions share them.  The cache maps that key to the table already emitted into 'tab' -- on the 5-leg intermediates
 * of a bond this removes most of the table volume (generation, upload and L2 footprint of the epilogue alike). */
struct tab_key { int32_t count; int32_t ext[4]; int64_t stride[4]; int64_t base; };
struct tab_cache { struct tab_key* key; int32_t* off; size_t cap, n; };
static uint64_t tab_key_hash(const struct tab_key* k)
{
	uint64_t h = 1469598103934665603ull ^ (uint64_t)k->count;
	for (int a = 0; a < 4; a++) { h = (h ^ (uint64_t)(uint32_t)k->ext[a]) * 1099511628211ull; h = (h ^ (uint64_t)k->stride[a]) * 1099511628211ull; }
	h = (h ^ (uint64_t)k->base) * 1099511628211ull;
	return h ^ (h >> 31);
}
static void tab_cache_init(struct tab_cache* c) { c->cap = 1 << 12; c->n = 0; c->key = calloc(c->cap, sizeof(*c->key)); c->off = malloc(c->cap * sizeof(int32_t)); for (size_t i = 0; i < c->cap; i++) { c->off[i] = -1; } }
static void tab_cache_free(struct tab_cache* c) { free(c->key); free(c->off); }
static void tab_cache_grow(struct tab_cache* c)
{
	struct tab_cache big; big.cap = c->cap * 2; big.n = c->n;
	big.key = calloc(big.cap, sizeof(*big.key)); big.off = malloc(big.cap * sizeof(int32_t));
	for (size_t i = 0; i < big.cap; i++) { big.off[i] = -1; }
	for (size_t i = 0; i < c->cap; i++) {
		if (c->off[i] < 0) { continue; }
		size_t q = (size_t)tab_key_hash(&c->key[i]) & (big.cap - 1);
		while (big.off[q] >= 0) { q = (q + 1) & (big.cap - 1); }
		big.key[q] = c->key[i]; big.off[q] = c->off[i];
	}
	tab_cache_free(c);
	*c = big;
}
/* start index in 'tab' of the offset table of the given axes: an existing identical table, or a freshly appended one */
static int32_t cached_offset_table(struct tab_cache* c, struct i32vec* tab, int first, int count, const struct ctb_axis* const* nat, const int* nat_sec,
	const int* pos_of_nat, const ct_long* stride_r, ct_long base)
{
	bool mapped = false;
	for (int a = 0; a < count; a++) { mapped = mapped || (first + a == g_map_nat); }
	if (mapped || count > 4) {
		const int32_t at = (int32_t)tab->n;
		append_offset_table(tab, first, count, nat, nat_sec, pos_of_nat, stride_r, base);
		return at;
	}
	struct tab_key k;
	memset(&k, 0, sizeof(k));
	k.count = count; k.base = base;
	for (int a = 0; a < count; a++) { k.ext[a] = nat[first + a]->secdim[nat_sec[first + a]]; k.stride[a] = stride_r[pos_of_nat[first + a]]; }
	size_t q = (size_t)tab_key_hash(&k) & (c->cap - 1);
	while (c->off[q] >= 0) {
		if (memcmp(&c->key[q], &k, sizeof(k)) == 0) { return c->off[q]; }
		q = (q + 1) & (c->cap - 1);
	}
	const int32_t at = (int32_t)tab->n;
	append_offset_table(tab, first, count, nat, nat_sec, pos_of_nat, stride_r, base);
	c->key[q] = k; c->off[q] = at;
	if (++c->n * 2 > c->cap) { tab_cache_grow(c); }
	return at;
}

struct merge_key { ct_long key; int blk; };
/* sector of axis 'ax' of x that makes the block idx[] (all other axes given) charge-conserving, or -1 (reference rule, sum dir q = 0) */
static int conserved_sector(const struct ctb_tensor* x, const int* idx, int ax)
{
	qnumber sum = 0;
	for (int i = 0; i < x->ndim; i++) { if (i != ax) { sum += x->ax[i].dir * x->ax[i].qsec[idx[i]]; } }
	return ctb_axis_find_sector(&x->ax[ax], -x->ax[ax].dir * sum);
}

static int cmp_merge_key(const void* a, const void* b)
{
	const struct merge_key* x = a; const struct merge_key* y = b;
	if (x->key != y->key) { return x->key < y->key ? -1 : 1; }
	return (x->blk > y->blk) - (x->blk < y->blk);
}

static uint64_t hash_i64(const int64_t* v, size_t n)
{
	uint64_t h = 1469598103934665603ull;
	for (size_t i = 0; i < n; i++) { h ^= (uint64_t)v[i]; h *= 1099511628211ull; h ^= h >> 29; }
	return h;
}

struct ctb_tensor* ctb_dot_prepare_ex(const struct ctb_tensor* s, int axrange_s, int conj_s,
	const struct ctb_tensor* t, int axrange_t, int conj_t, int ndim_mult, const int* perm,
	int alloc_result, int flags, struct ctb_dot_plan* plan)
{
	CTB_REQUIRE(s->dtype == t->dtype);
	CTB_REQUIRE(ndim_mult >= 1 && s->ndim >= ndim_mult && t->ndim >= ndim_mult);
	const double tp_begin = ctb_wall_ms();
	const int nfs = s->ndim - ndim_mult;   /* free axes of s */
	const int nft = t->ndim - ndim_mult;   /* free axes of t */
	const int ndimr = nfs + nft;
	CTB_REQUIRE(ndimr <= CTB_MAXDIM);
	const int shift_s  = (axrange_s == TENSOR_AXIS_RANGE_LEADING ? 0 : nfs);          /* first contracted axis */
	const int shift_t  = (axrange_t == TENSOR_AXIS_RANGE_LEADING ? 0 : nft);
	const int offset_s = (axrange_s == TENSOR_AXIS_RANGE_LEADING ? ndim_mult : 0);    /* first free axis */
	const int offset_t = (axrange_t == TENSOR_AXIS_RANGE_LEADING ? ndim_mult : 0);

	/* contracted legs must carry identical quantum numbers and opposite directions (reference :1836-1844) */
	for (int i = 0; i < ndim_mult; i++) {
		CTB_REQUIRE(s->ax[shift_s + i].dir == -t->ax[shift_t + i].dir);
		CTB_REQUIRE(ctb_axis_same_qnums(&s->ax[shift_s + i], &t->ax[shift_t + i]));
	}

	/* "natural" result axes: free axes of s, then free axes of t; result axis i = natural axis perm[i] */
	const struct ctb_axis* nat[CTB_MAXDIM];
	for (int i = 0; i < nfs; i++) { nat[i] = &s->ax[offset_s + i]; }
	for (int i = 0; i < nft; i++) { nat[nfs + i] = &t->ax[offset_t + i]; }
	int p[CTB_MAXDIM], pos_of_nat[CTB_MAXDIM];
	for (int i = 0; i < ndimr; i++) { p[i] = (perm != NULL ? perm[i] : i); }
	for (int i = 0; i < ndimr; i++) { pos_of_nat[p[i]] = i; }

	struct ctb_axis raxes[CTB_MAXDIM];
	for (int i = 0; i < ndimr; i++) { ctb_axis_copy(&raxes[i], nat[p[i]]); }
	struct ctb_tensor* r = ctb_tensor_from_axes(s->dtype, ndimr, raxes, alloc_result);
	ctb_plan_profile[0] += ctb_wall_ms() - tp_begin;

	const int a_kcontig = (axrange_s == TENSOR_AXIS_RANGE_TRAILING);
	const int b_ncontig = (axrange_t == TENSOR_AXIS_RANGE_LEADING);
	bool merge = (flags & CTB_DOT_MERGE_ROWS) != 0 && s->d != NULL && nft > 0 && r->nblk > 0;
	{
		/* the merged (non-mixing) form addresses the whole result with 32-bit offsets: beyond that the plain per-block form is used */
		ct_long kfull0 = 1;
		for (int i = 0; i < ndim_mult; i++) { kfull0 *= s->ax[shift_s + i].dim; }
		const bool mix_ok = b_ncontig && nft <= 4 && kfull0 <= CTB_MIX_KMAX;
		if (merge && !mix_ok && r->nstore >= ((ct_long)1 << 31)) { merge = false; }
	}

	/* growable host arrays */
	size_t cap_seg = 1024, nseg = 0;
	struct ctbd_gemm_seg* segs = malloc(cap_seg * sizeof(*segs));
	struct ctbd_gemm_out* outs = malloc((r->nblk > 0 ? r->nblk : 1) * sizeof(*outs));
	int nouts = 0;
	struct i32vec tab = { NULL, 0, 0 };
	struct tab_cache tcache;
	tab_cache_init(&tcache);
	double flops = 0;
	const double flop_factor = ctb_is_complex(s->dtype) ? 8.0 : 2.0;

	/* contracted sector grid: the leading ndim_mult - 1 contracted axes are enumerated (row-major, the order in which the reference
	 * accumulates, :1951-1953), the sector of the last one is fixed by charge conservation -- O(1) instead of a scan over its sectors */
	ct_long nlead = 1;
	for (int i = 0; i < ndim_mult - 1; i++) { nlead *= s->ax[shift_s + i].nsec; }

	/* merged form: packed copy of the (small) s operand, one dense M' x K' matrix per distinct group content */
	int64_t* gather = NULL; size_t ngather = 0, cap_gather = 0;
	/* merged form with an n-contiguous t: the rows of all contracted tuples of one output are addressed through a row table,
	 * so the output is ONE segment of extent K' (a pipeline step of the kernel then spans many of the 1-4 row blocks) */
	const bool use_rowtab = merge && b_ncontig;
	int64_t* browtab = NULL; size_t nbrow = 0, cap_brow = 0;
	/* mixing form (ctbd_mix_group): chosen when the whole contracted extent is small, i.e. s is an MPO tensor of a short-range
	 * Hamiltonian; the launch then streams t and the result once, with no per-element offset tables at all */
	ct_long kfull = 1;
	for (int i = 0; i < ndim_mult; i++) { kfull *= s->ax[shift_s + i].dim; }
	const bool mixmode = use_rowtab && nft <= 4 && kfull <= CTB_MIX_KMAX;
	struct ctbd_mix_group* mgroups = NULL; size_t nmg = 0, cap_mg = 0;
	struct ctbd_mix_row* mrows = NULL; size_t nmr = 0, cap_mr = 0;
	struct merge_key* mk = NULL;
	struct { uint64_t h; size_t base, len, sig, slen; }* packed = NULL; size_t npacked = 0, cap_packed = 0;
	int64_t* sig = NULL; size_t nsig = 0, cap_sig = 0;      /* signatures of the packed matrices */
	int32_t* pidx = NULL; size_t pidx_cap = 0;      /* hash index into packed[] */
	if (merge)
	{
		CTB_REQUIRE(mixmode || r->nstore < ((ct_long)1 << 31));
		mk = malloc((size_t)r->nblk * sizeof(*mk));
		for (int b = 0; b < r->nblk; b++) {
			int idx_r[CTB_MAXDIM];
			ctb_grid_unravel(r, r->blk_grid[b], idx_r);
			ct_long key = 0;
			for (int i = 0; i < nft; i++) { key = key * nat[nfs + i]->nsec + idx_r[pos_of_nat[nfs + i]]; }
			mk[b].key = key; mk[b].blk = b;
		}
		qsort(mk, (size_t)r->nblk, sizeof(*mk), cmp_merge_key);
	}

	for (int b0 = 0; b0 < r->nblk; )
	{
		/* members of this output: one result block, or all result blocks sharing the free sectors of t */
		int b1 = b0 + 1;
		if (merge) { while (b1 < r->nblk && mk[b1].key == mk[b0].key) { b1++; } }

		int idx_r[CTB_MAXDIM], nat_sec[CTB_MAXDIM];
		ctb_grid_unravel(r, r->blk_grid[merge ? mk[b0].blk : b0], idx_r);
		for (int i = 0; i < ndimr; i++) { nat_sec[p[i]] = idx_r[i]; }
		ct_long N = 1;
		for (int i = 0; i < nft; i++) { N *= nat[nfs + i]->secdim[nat_sec[nfs + i]]; }

		struct ctbd_gemm_out o_scratch;
		struct ctbd_gemm_out* o = mixmode ? &o_scratch : &outs[nouts++];
		o->n = (int32_t)N;
		o->seg_begin = (int32_t)nseg;

		if (!merge)
		{
			const int b = b0;
			ct_long M = 1;
			for (int i = 0; i < nfs; i++) { M *= nat[i]->secdim[nat_sec[i]]; }
			o->c_off = r->blk_off[b];
			o->m = (int32_t)M;
			int idx_full[CTB_MAXDIM];
			int32_t* posmap = NULL;
			if (g_embed != NULL)
			{
				/* block of the full tensor this piece block lands in, and where the piece columns sit inside it */
				const struct ctb_tensor* full = g_embed->full;
				const int ea = g_embed->axis;
				for (int i = 0; i < ndimr; i++) { idx_full[i] = idx_r[i]; }
				idx_full[ea] = ctb_axis_find_sector(&full->ax[ea], r->ax[ea].qsec[idx_r[ea]]);
				CTB_REQUIRE(idx_full[ea] >= 0);
				const ct_long off_full = ctb_grid_offset(full, ctb_grid_ravel(full, idx_full));
				CTB_REQUIRE(off_full >= 0);
				o->c_off = off_full;
				const int np = r->ax[ea].secdim[idx_r[ea]];
				posmap = malloc((size_t)np * sizeof(int32_t));
				for (int d = 0; d < np; d++) {
					const ct_long lg = g_embed->ind[r->ax[ea].log_of[r->ax[ea].secstart[idx_r[ea]] + d]];
					CTB_REQUIRE(full->ax[ea].sec_of[lg] == idx_full[ea]);
					posmap[d] = full->ax[ea].pos_of[lg];
				}
				g_map_nat = p[ea]; g_map_pos = posmap;
			}
			/* enumerate contracted sector tuples in row-major order (reference :1951-1953) */
			int idx_s[CTB_MAXDIM], idx_t[CTB_MAXDIM], kap[CTB_MAXDIM] = { 0 };
			for (int i = 0; i < nfs; i++) { idx_s[offset_s + i] = nat_sec[i]; }
			for (int i = 0; i < nft; i++) { idx_t[offset_t + i] = nat_sec[nfs + i]; }
			/* the last contracted sector follows from charge conservation in s: only the leading contracted sectors are enumerated */
			const double tq0 = TP_NOW();
			for (ct_long cl = 0; cl < nlead; cl++)
			{
				for (int i = 0; i < ndim_mult - 1; i++) { idx_s[shift_s + i] = kap[i]; }
				kap[ndim_mult - 1] = conserved_sector(s, idx_s, shift_s + ndim_mult - 1);
				const bool present = (kap[ndim_mult - 1] >= 0);
				if (present) { for (int i = 0; i < ndim_mult; i++) { idx_s[shift_s + i] = kap[i]; idx_t[shift_t + i] = kap[i]; } }
				const ct_long a_off = present ? ctb_grid_offset(s, ctb_grid_ravel(s, idx_s)) : -1;
				if (a_off >= 0)
				{
					const ct_long b_off = ctb_grid_offset(t, ctb_grid_ravel(t, idx_t));
					CTB_REQUIRE(b_off >= 0);   /* conservation in t follows from s and r (reference :1976-1984) */
					ct_long K = 1;
					for (int i = 0; i < ndim_mult; i++) { K *= s->ax[shift_s + i].secdim[kap[i]]; }
					if (nseg == cap_seg) { cap_seg *= 2; segs = realloc(segs, cap_seg * sizeof(*segs)); }
					struct ctbd_gemm_seg* g = &segs[nseg++];
					g->a_off = a_off; g->b_off = b_off; g->k = (int32_t)K;
					g->lda = (int32_t)(a_kcontig ? K : M);
					g->ldb = (int32_t)(b_ncontig ? N : K);
					g->pad_ = 0;
					flops += flop_factor * (double)M * (double)N * (double)K;
				}
				for (int i = ndim_mult - 2; i >= 0; i--) {
					if (++kap[i] < s->ax[shift_s + i].nsec) { break; }
					kap[i] = 0;
				}
			}
			o->seg_end = (int32_t)nseg;
			ctb_plan_profile[6] += TP_NOW() - tq0;
			const double tq5 = TP_NOW();
			/* row/column offset tables: where element (i, j) of the natural block lands in the permuted block */
			ct_long stride_r[CTB_MAXDIM];
			ct_long st = 1;
			for (int i = ndimr - 1; i >= 0; i--) {
				stride_r[i] = st;
				st *= (g_embed != NULL) ? g_embed->full->ax[i].secdim[idx_full[i]] : r->ax[i].secdim[idx_r[i]];
			}
			CTB_REQUIRE(st < ((ct_long)1 << 31));
			o->row_tab = cached_offset_table(&tcache, &tab, 0, nfs, nat, nat_sec, pos_of_nat, stride_r, 0);
			o->col_tab = cached_offset_table(&tcache, &tab, nfs, nft, nat, nat_sec, pos_of_nat, stride_r, 0);
			if (posmap != NULL) { free(posmap); g_map_nat = -1; g_map_pos = NULL; }
			ctb_plan_profile[11] += TP_NOW() - tq5;
		}
		else
		{
			/* contracted tuples present in t for these free sectors: the common K' of all members */
			int idx_t[CTB_MAXDIM], kap[CTB_MAXDIM] = { 0 };
			for (int i = 0; i < nft; i++) { idx_t[offset_t + i] = nat_sec[nfs + i]; }
			ct_long Ktot = 0;
			const size_t seg_first = nseg;
			const double tq1 = TP_NOW();
			for (ct_long cl = 0; cl < nlead; cl++)
			{
				for (int i = 0; i < ndim_mult - 1; i++) { idx_t[shift_t + i] = kap[i]; }
				kap[ndim_mult - 1] = conserved_sector(t, idx_t, shift_t + ndim_mult - 1);
				const bool present = (kap[ndim_mult - 1] >= 0);
				if (present) { idx_t[shift_t + ndim_mult - 1] = kap[ndim_mult - 1]; }
				const ct_long b_off = present ? ctb_grid_offset(t, ctb_grid_ravel(t, idx_t)) : -1;
				const ct_long c = cl * s->ax[shift_s + ndim_mult - 1].nsec + (present ? kap[ndim_mult - 1] : 0);      /* contracted grid cell */
				if (b_off >= 0)
				{
					ct_long K = 1;
					for (int i = 0; i < ndim_mult; i++) { K *= s->ax[shift_s + i].secdim[kap[i]]; }
					if (nseg == cap_seg) { cap_seg *= 2; segs = realloc(segs, cap_seg * sizeof(*segs)); }
					struct ctbd_gemm_seg* g = &segs[nseg++];
					g->a_off = Ktot;          /* column start inside the packed matrix; base added below */
					g->b_off = b_off; g->k = (int32_t)K;
					g->ldb = (int32_t)(b_ncontig ? N : K);
					g->pad_ = (int32_t)c;     /* contracted grid cell, consumed below */
					Ktot += K;
				}
				for (int i = ndim_mult - 2; i >= 0; i--) {
					if (++kap[i] < s->ax[shift_s + i].nsec) { break; }
					kap[i] = 0;
				}
			}
			o->seg_end = (int32_t)nseg;
			ctb_plan_profile[7] += TP_NOW() - tq1;
			const double tq2 = TP_NOW();
			double tq_pack = 0;
			/* stacked rows of all members */
			ct_long Mtot = 0;
			for (int q = b0; q < b1; q++) {
				int ir[CTB_MAXDIM];
				ctb_grid_unravel(r, r->blk_grid[mk[q].blk], ir);
				ct_long M = 1;
				for (int i = 0; i < nfs; i++) { M *= nat[i]->secdim[ir[pos_of_nat[i]]]; }
				Mtot += M;
			}
			o->c_off = 0;
			o->m = (int32_t)Mtot;
			flops += flop_factor * (double)Mtot * (double)N * (double)Ktot;
			o->row_tab = (int32_t)tab.n;
			o->col_tab = -1;
			if (!mixmode) {
				i32vec_reserve(&tab, (size_t)(2 * Mtot));
				tab.n += (size_t)(2 * Mtot);
			}
			const size_t rowtab0 = (size_t)o->row_tab, rowcol0 = rowtab0 + (size_t)Mtot;
			size_t mrow_first = nmr;
			if (mixmode) {
				while (nmr + (size_t)Mtot > cap_mr) { cap_mr = cap_mr ? 2 * cap_mr : 4096; mrows = realloc(mrows, cap_mr * sizeof(*mrows)); }
			}
			/* the packed Mtot x Ktot matrix (k contiguous): its gather list is generated further down, for a new matrix only */
			const size_t glen = (size_t)(Mtot * Ktot);
			const size_t sig_first = nsig;
			ct_long row0 = 0;
			for (int q = b0; q < b1; q++)
			{
				const int b = mk[q].blk;
				int ir[CTB_MAXDIM], ns[CTB_MAXDIM], idx_s[CTB_MAXDIM];
				ctb_grid_unravel(r, r->blk_grid[b], ir);
				for (int i = 0; i < ndimr; i++) { ns[p[i]] = ir[i]; }
				ct_long M = 1;
				for (int i = 0; i < nfs; i++) { M *= nat[i]->secdim[ns[i]]; idx_s[offset_s + i] = ns[i]; }
				ct_long stride_r[CTB_MAXDIM];
				ct_long st = 1;
				for (int i = ndimr - 1; i >= 0; i--) { stride_r[i] = st; st *= r->ax[i].secdim[ir[i]]; }
				if (mixmode)
				{
					/* one row descriptor per stacked row: absolute offset of its first column and the stride of every column digit */
					int dig[CTB_MAXDIM] = { 0 };
					for (ct_long i = 0; i < M; i++)
					{
						struct ctbd_mix_row* mr = &mrows[nmr++];
						ct_long off = r->blk_off[b];
						for (int a = 0; a < nfs; a++) { off += dig[a] * stride_r[pos_of_nat[a]]; }
						mr->c_off = off;
						mr->a_off = (row0 + i) * Ktot;      /* base of the packed matrix added below */
						for (int a = 0; a < 4; a++) { mr->cs[a] = 0; }
						for (int a = 0; a < nft; a++) {
							CTB_REQUIRE(stride_r[pos_of_nat[nfs + a]] < ((ct_long)1 << 31));
							mr->cs[a] = (int32_t)stride_r[pos_of_nat[nfs + a]];
						}
						for (int a = nfs - 1; a >= 0; a--) {
							if (++dig[a] < nat[a]->secdim[ns[a]]) { break; }
							dig[a] = 0;
						}
					}
				}
				else
				{
				/* absolute row offsets and the column table of this member */
				const size_t coltab_idx = tab.n + (size_t)M;
				{
					struct i32vec rows = { NULL, 0, 0 };
					append_offset_table(&rows, 0, nfs, nat, ns, pos_of_nat, stride_r, r->blk_off[b]);
					for (ct_long i = 0; i < M; i++) {
						tab.v[rowtab0 + (size_t)(row0 + i)] = rows.v[i];
						CTB_REQUIRE(coltab_idx < ((size_t)1 << 31));
					}
					free(rows.v);
				}
				/* member column table */
				const size_t ct0 = (size_t)cached_offset_table(&tcache, &tab, nfs, nft, nat, ns, pos_of_nat, stride_r, 0);
				for (ct_long i = 0; i < M; i++) { tab.v[rowcol0 + (size_t)(row0 + i)] = (int32_t)ct0; }
				}
				/* signature of this member's part of the packed matrix: (rows, offset of the s block of every contracted tuple) */
				const double tq3 = TP_NOW();
				while (nsig + 1 + (nseg - seg_first) > cap_sig) { cap_sig = cap_sig ? 2 * cap_sig : 4096; sig = realloc(sig, cap_sig * sizeof(int64_t)); }
				sig[nsig++] = M;
				for (size_t sg = seg_first; sg < nseg; sg++)
				{
					ct_long cell = segs[sg].pad_;
					for (int i = ndim_mult - 1; i >= 0; i--) { idx_s[shift_s + i] = (int)(cell % s->ax[shift_s + i].nsec); cell /= s->ax[shift_s + i].nsec; }
					const ct_long a_off = ctb_grid_offset(s, ctb_grid_ravel(s, idx_s));
					CTB_REQUIRE(a_off >= 0);   /* conservation in s follows from r and t */
					sig[nsig++] = a_off;
				}
				tq_pack += TP_NOW() - tq3;
				row0 += M;
			}
			ctb_plan_profile[8] += TP_NOW() - tq2 - tq_pack; ctb_plan_profile[9] += tq_pack;
			const double tq4 = TP_NOW();
			/* reuse an identical packed matrix if one exists.  The packed Mtot x Ktot matrix is a function of its signature -- the
			 * rows of every member, the extents of the contracted tuples and the offsets of the s blocks -- so identical matrices are
			 * recognised from a few numbers per (member, tuple) and the gather list is generated only for a new one (generating and
			 * hashing every list first was more than half of the plan-building time of a molecular bond) */
			while (nsig + (nseg - seg_first) > cap_sig) { cap_sig = cap_sig ? 2 * cap_sig : 4096; sig = realloc(sig, cap_sig * sizeof(int64_t)); }
			for (size_t sg = seg_first; sg < nseg; sg++) { sig[nsig++] = segs[sg].k; }
			const size_t slen = nsig - sig_first;
			const uint64_t hsh = hash_i64(sig + sig_first, slen) ^ (uint64_t)glen;
			size_t base = ngather;
			bool found = false;
			if (2 * (npacked + 1) > pidx_cap) {
				pidx_cap = pidx_cap ? 2 * pidx_cap : 1024;
				free(pidx); pidx = malloc(pidx_cap * sizeof(int32_t));
				for (size_t q = 0; q < pidx_cap; q++) { pidx[q] = -1; }
				for (size_t q = 0; q < npacked; q++) {
					size_t sl = (size_t)packed[q].h & (pidx_cap - 1);
					while (pidx[sl] >= 0) { sl = (sl + 1) & (pidx_cap - 1); }
					pidx[sl] = (int32_t)q;
				}
			}
			size_t slot = (size_t)hsh & (pidx_cap - 1);
			while (pidx[slot] >= 0) {
				const size_t q = (size_t)pidx[slot];
				if (packed[q].h == hsh && packed[q].len == glen && packed[q].slen == slen && memcmp(sig + packed[q].sig, sig + sig_first, slen * sizeof(int64_t)) == 0) { base = packed[q].base; found = true; break; }
				slot = (slot + 1) & (pidx_cap - 1);
			}
			if (!found) {
				if (npacked == cap_packed) { cap_packed = cap_packed ? 2 * cap_packed : 256; packed = realloc(packed, cap_packed * sizeof(*packed)); }
				packed[npacked].h = hsh; packed[npacked].base = base; packed[npacked].len = glen; packed[npacked].sig = sig_first; packed[npacked].slen = slen;
				pidx[slot] = (int32_t)npacked;
				npacked++;
				/* gather list of the new packed matrix (k contiguous) */
				while (ngather + glen > cap_gather) { cap_gather = cap_gather ? 2 * cap_gather : 65536; gather = realloc(gather, cap_gather * sizeof(int64_t)); }
				int64_t* gl = gather + ngather;
				const int64_t* sp = sig + sig_first;
				ct_long r0 = 0;
				for (int q = b0; q < b1; q++) {
					const ct_long M = *sp++;
					for (size_t sg = seg_first; sg < nseg; sg++) {
						const ct_long a_off = *sp++;
						const ct_long K = segs[sg].k, kcol = segs[sg].a_off;
						for (ct_long i = 0; i < M; i++) {
							int64_t* row = gl + (r0 + i) * Ktot + kcol;
							if (a_kcontig) { const ct_long a0 = a_off + i * K; for (ct_long kk = 0; kk < K; kk++) { row[kk] = a0 + kk; } }
							else { for (ct_long kk = 0; kk < K; kk++) { row[kk] = a_off + kk * M + i; } }
						}
					}
					r0 += M;
				}
				ngather += glen;
			}
			else { nsig = sig_first; }      /* the signature of a reused matrix is not kept */
			ctb_plan_profile[10] += TP_NOW() - tq4;
			for (size_t sg = seg_first; sg < nseg; sg++) {
				segs[sg].a_off += (int64_t)base;
				segs[sg].lda = (int32_t)Ktot;
				segs[sg].pad_ = 0;
			}
			if (use_rowtab && nseg > seg_first)
			{
				while (nbrow + (size_t)Ktot > cap_brow) { cap_brow = cap_brow ? 2 * cap_brow : 16384; browtab = realloc(browtab, cap_brow * sizeof(int64_t)); }
				const size_t row0_idx = nbrow;
				for (size_t sg = seg_first; sg < nseg; sg++) {
					for (ct_long kk = 0; kk < segs[sg].k; kk++) { browtab[nbrow++] = segs[sg].b_off + kk * N; }
				}
				CTB_REQUIRE(nbrow - row0_idx == (size_t)Ktot);
				struct ctbd_gemm_seg* g = &segs[seg_first];
				g->a_off = (int64_t)base; g->b_off = (int64_t)row0_idx; g->k = (int32_t)Ktot; g->lda = (int32_t)Ktot; g->ldb = (int32_t)N; g->pad_ = 0;
				nseg = seg_first + 1;
				o->seg_end = (int32_t)nseg;
				if (mixmode)
				{
					if (nmg == cap_mg) { cap_mg = cap_mg ? 2 * cap_mg : 1024; mgroups = realloc(mgroups, cap_mg * sizeof(*mgroups)); }
					struct ctbd_mix_group* mg = &mgroups[nmg++];
					memset(mg, 0, sizeof(*mg));
					mg->brow_begin = (int64_t)row0_idx;
					mg->n = (int32_t)N; mg->kp = (int32_t)Ktot;
					mg->row_begin = (int32_t)mrow_first; mg->row_end = (int32_t)nmr;
					mg->ndig = nft;
					for (int a = 0; a < 4; a++) { mg->dig_dim[a] = 1; }
					for (int a = 0; a < nft; a++) { mg->dig_dim[a] = nat[nfs + a]->secdim[nat_sec[nfs + a]]; }
					for (size_t q = mrow_first; q < nmr; q++) { mrows[q].a_off += (int64_t)base; }
					nseg = seg_first;      /* the mixing form carries no segment list */
				}
			}
			else if (mixmode) { nmr = mrow_first; nseg = seg_first; }
		}
		CTB_REQUIRE(tab.n < ((size_t)1 << 31));
		b0 = b1;
	}

	struct ctbd_gemm_plan_host h;
	memset(&h, 0, sizeof(h));
	h.dtype = s->dtype;
	h.a_kcontig = merge ? 1 : a_kcontig; h.b_ncontig = b_ncontig;
	h.conj_a = conj_s; h.conj_b = conj_t;
	h.nouts = nouts; h.nsegs = (int32_t)nseg; h.ntab = (int32_t)tab.n;
	h.outs = outs; h.segs = segs; h.tab = tab.v;
	h.flops = flops;
	if (merge) { h.a_gather = gather; h.n_a_gather = (int64_t)ngather; h.a_src = s->d; }
	if (use_rowtab && nbrow > 0) { h.b_rowtab = browtab; h.n_b_rowtab = (int64_t)nbrow; }
	if (mixmode && nmg > 0) { h.mix_groups = mgroups; h.n_mix_groups = (int32_t)nmg; h.mix_rows = mrows; h.n_mix_rows = (int32_t)nmr; }
	if (getenv("CTB_DUMP_PLAN") != NULL) {
		/* tuning aid: (m, n, total k) of every output block, appended as text */
		FILE* f = fopen(getenv("CTB_DUMP_PLAN"), "a");
		if (f != NULL) {
			fprintf(f, "plan %d\n", nouts);
			for (int b = 0; b < nouts; b++) { long long kt = 0; for (int q = outs[b].seg_begin; q < outs[b].seg_end; q++) { kt += segs[q].k; } fprintf(f, "%d %d %lld %d\n", outs[b].m, outs[b].n, kt, outs[b].seg_end - outs[b].seg_begin); }
			fclose(f);
		}
	}
	plan->dev = NULL;
	const double tpc = ctb_wall_ms();
	CTB_CHECK_ABORT(ctbd_gemm_plan_create(&h, &plan->dev));
	ctb_plan_profile[1] += tpc - tp_begin; ctb_plan_profile[2] += ctb_wall_ms() - tpc; ctb_plan_profile[3] += 1; ctb_plan_profile[4] += nouts; ctb_plan_profile[5] += (double)tab.n;
	ctb_global_stats.plan_ms += ctb_wall_ms() - tp_begin;
	if (getenv("CTB_TRACE") != NULL) { fprintf(stderr, "  dot plan: host lists %.2f ms, device plan %.2f ms (%d blocks, %d segments, %d table entries)\n", tpc - tp_begin, ctb_wall_ms() - tpc, nouts, (int)nseg, (int)tab.n); }
	plan->flops = flops;
	plan->nouts = nouts; plan->nsegs = (int)nseg;
	plan->ntiles = 0;
	CTB_CHECK_ABORT(ctbd_gemm_plan_info(plan->dev, &plan->ntiles, NULL));

	tab_cache_free(&tcache);
	free(tab.v); free(outs); free(segs); free(gather); free(mk); free(packed); free(pidx); free(sig); free(browtab); free(mgroups); free(mrows);
	return r;
}

struct ctb_tensor* ctb_dot_prepare_embed(const struct ctb_tensor* s, int axrange_s, int conj_s,
	const struct ctb_tensor* t, int axrange_t, int conj_t, int ndim_mult, const struct ctb_embed* emb, struct ctb_dot_plan* plan)
{
	CTB_REQUIRE(emb != NULL && emb->full != NULL && emb->ind != NULL);
	g_embed = emb;
	struct ctb_tensor* r = ctb_dot_prepare_ex(s, axrange_s, conj_s, t, axrange_t, conj_t, ndim_mult, NULL, 0, 0, plan);
	g_embed = NULL;
	return r;
}

struct ctb_tensor* ctb_dot_prepare(const struct ctb_tensor* s, int axrange_s, int conj_s,
	const struct ctb_tensor* t, int axrange_t, int conj_t, int ndim_mult, const int* perm,
	int alloc_result, struct ctb_dot_plan* plan)
{
	return ctb_dot_prepare_ex(s, axrange_s, conj_s, t, axrange_t, conj_t, ndim_mult, perm, alloc_result, 0, plan);
}

int ctb_dot_exec(const struct ctb_dot_plan* plan, const void* s_data, const void* t_data, void* r_data)
{
	if (plan->ntiles == 0) { return 0; }
	return ctbd_gemm_run(plan->dev, s_data, t_data, r_data);
}

int ctb_dot_exec_multi(const struct ctb_dot_plan* plan, const void* s_data, const void* t_data, int ndst, void* const* r_datas)
{
	if (plan->ntiles == 0) { return 0; }
	return ctbd_gemm_run_multi(plan->dev, s_data, t_data, ndst, r_datas);
}

int ctb_dot_exec_mc(const struct ctb_dot_plan* plan, const void* s_data, const void* t_data, void* r_mc)
{
	if (plan->ntiles == 0) { return 0; }
	return ctbd_gemm_run_mc(plan->dev, s_data, t_data, r_mc);
}

void ctb_dot_plan_free(struct ctb_dot_plan* plan)
{
	if (plan->dev != NULL) { ctbd_gemm_plan_destroy(plan->dev); plan->dev = NULL; }
}

struct ctb_tensor* ctb_dot(const struct ctb_tensor* s, int axrange_s, int conj_s,
	const struct ctb_tensor* t, int axrange_t, int conj_t, int ndim_mult, const int* perm)
{
	struct ctb_dot_plan plan;
	struct ctb_tensor* r = ctb_dot_prepare(s, axrange_s, conj_s, t, axrange_t, conj_t, ndim_mult, perm, 1, &plan);
	CTB_CHECK_ABORT(ctb_dot_exec(&plan, s->d, t->d, r->d));
	ctb_dot_plan_free(&plan);
	return r;
}

/* ---------------------------------------------------------------------------------------------- */
/* logical-index remaps                                                                            */
/* ---------------------------------------------------------------------------------------------- */

static void run_remap(struct ctbd_remap_args* args, struct ctb_tensor* dst, struct ctb_tensor* src)
{
	if (dst->nstore == 0) { return; }
	const double t0 = ctb_wall_ms();
	args->dst_layout = ctb_tensor_layout(dst); args->dst = dst->d;
	args->src_layout = ctb_tensor_layout(src); args->src = src->d;
	CTB_CHECK_ABORT(ctbd_remap(args));
	ctb_global_stats.remap_ms += ctb_wall_ms() - t0;
}

struct ctb_tensor* ctb_transpose(struct ctb_tensor* t, const int* perm, int conj)
{
	struct ctb_axis axes[CTB_MAXDIM];
	for (int i = 0; i < t->ndim; i++) { ctb_axis_copy(&axes[i], &t->ax[perm[i]]); }
	struct ctb_tensor* r = ctb_tensor_from_axes(t->dtype, t->ndim, axes, 1);
	struct ctbd_remap_args args;
	memset(&args, 0, sizeof(args));
	args.op = CTBD_REMAP_TRANSPOSE;
	for (int i = 0; i < t->ndim; i++) { args.perm[i] = perm[i]; }
	args.conj = conj;
	args.scale_ax = -1;
	run_remap(&args, r, t);
	return r;
}

struct ctb_tensor* ctb_flatten_axes(struct ctb_tensor* t, int i_ax, int new_dir)
{
	CTB_REQUIRE(0 <= i_ax && i_ax + 1 < t->ndim);
	const struct ctb_axis* a0 = &t->ax[i_ax];
	const struct ctb_axis* a1 = &t->ax[i_ax + 1];
	const ct_long dflat = a0->dim * a1->dim;
	qnumber* qflat = ctb_malloc(dflat * sizeof(qnumber));
	/* reference :972-979 */
	for (ct_long j = 0; j < a0->dim; j++) {
		for (ct_long k = 0; k < a1->dim; k++) {
			qflat[j * a1->dim + k] = new_dir * (a0->dir * a0->qlog[j] + a1->dir * a1->qlog[k]);
		}
	}
	struct ctb_axis axes[CTB_MAXDIM];
	for (int i = 0; i < i_ax; i++) { ctb_axis_copy(&axes[i], &t->ax[i]); }
	ctb_axis_init(&axes[i_ax], dflat, new_dir, qflat);
	for (int i = i_ax + 2; i < t->ndim; i++) { ctb_axis_copy(&axes[i - 1], &t->ax[i]); }
	ctb_free(qflat);
	struct ctb_tensor* r = ctb_tensor_from_axes(t->dtype, t->ndim - 1, axes, 1);
	struct ctbd_remap_args args;
	memset(&args, 0, sizeof(args));
	args.op = CTBD_REMAP_FLATTEN;
	args.i_ax = i_ax;
	args.scale_ax = -1;
	run_remap(&args, r, t);
	return r;
}

struct ctb_tensor* ctb_split_axis(struct ctb_tensor* t, int i_ax, const ct_long new_dim[2], const int new_dir[2], const qnumber* const new_qnums[2])
{
	CTB_REQUIRE(0 <= i_ax && i_ax < t->ndim && t->ndim + 1 <= CTB_MAXDIM);
	CTB_REQUIRE(new_dim[0] * new_dim[1] == t->ax[i_ax].dim);
	/* consistency of the provided quantum numbers (reference asserts :1127-1133) */
	for (ct_long j = 0; j < new_dim[0]; j++) {
		for (ct_long k = 0; k < new_dim[1]; k++) {
			CTB_REQUIRE(t->ax[i_ax].dir * t->ax[i_ax].qlog[j * new_dim[1] + k] == new_dir[0] * new_qnums[0][j] + new_dir[1] * new_qnums[1][k]);
		}
	}
	struct ctb_axis axes[CTB_MAXDIM];
	for (int i = 0; i < i_ax; i++) { ctb_axis_copy(&axes[i], &t->ax[i]); }
	ctb_axis_init(&axes[i_ax],     new_dim[0], new_dir[0], new_qnums[0]);
	ctb_axis_init(&axes[i_ax + 1], new_dim[1], new_dir[1], new_qnums[1]);
	for (int i = i_ax + 1; i < t->ndim; i++) { ctb_axis_copy(&axes[i + 1], &t->ax[i]); }
	struct ctb_tensor* r = ctb_tensor_from_axes(t->dtype, t->ndim + 1, axes, 1);
	struct ctbd_remap_args args;
	memset(&args, 0, sizeof(args));
	args.op = CTBD_REMAP_SPLIT;
	args.i_ax = i_ax;
	args.scale_ax = -1;
	run_remap(&args, r, t);
	return r;
}

struct ctb_tensor* ctb_slice(struct ctb_tensor* t, int i_ax, const ct_long* ind, ct_long nind)
{
	CTB_REQUIRE(0 <= i_ax && i_ax < t->ndim && nind > 0);
	qnumber* q = ctb_malloc(nind * sizeof(qnumber));
	for (ct_long j = 0; j < nind; j++) {
		CTB_REQUIRE(0 <= ind[j] && ind[j] < t->ax[i_ax].dim);
		q[j] = t->ax[i_ax].qlog[ind[j]];
	}
	struct ctb_axis axes[CTB_MAXDIM];
	for (int i = 0; i < t->ndim; i++) {
		if (i == i_ax) { ctb_axis_init(&axes[i], nind, t->ax[i].dir, q); }
		else { ctb_axis_copy(&axes[i], &t->ax[i]); }
	}
	ctb_free(q);
	struct ctb_tensor* r = ctb_tensor_from_axes(t->dtype, t->ndim, axes, 1);
	struct ctbd_remap_args args;
	memset(&args, 0, sizeof(args));
	args.op = CTBD_REMAP_SLICE;
	args.i_ax = i_ax;
	args.ind = (const int64_t*)ind;
	args.scale_ax = -1;
	run_remap(&args, r, t);
	return r;
}

struct ctb_tensor* ctb_scale_axis(struct ctb_tensor* t, int i_ax, const double* scale_dev)
{
	struct ctb_tensor* r = ctb_tensor_like(t, 1);
	struct ctbd_remap_args args;
	memset(&args, 0, sizeof(args));
	args.op = CTBD_REMAP_IDENTITY;
	args.scale_ax = i_ax;
	args.scale = scale_dev;
	run_remap(&args, r, t);
	return r;
}

struct ctb_tensor* ctb_drop_dummy_axes(const struct ctb_tensor* t, int ntrace)
{
	CTB_REQUIRE(t->ndim >= 2 * ntrace);
	struct ctb_axis axes[CTB_MAXDIM];
	for (int i = 0; i < ntrace; i++) {
		/* tracing a pair of dimension-1 legs is a copy; the pair must match (reference dense_tensor.c:269) */
		CTB_REQUIRE(t->ax[i].dim == 1 && t->ax[t->ndim - ntrace + i].dim == 1);
		CTB_REQUIRE(t->ax[i].qlog[0] == t->ax[t->ndim - ntrace + i].qlog[0]);
		CTB_REQUIRE(t->ax[i].dir == -t->ax[t->ndim - ntrace + i].dir);
	}
	for (int i = ntrace; i < t->ndim - ntrace; i++) { ctb_axis_copy(&axes[i - ntrace], &t->ax[i]); }
	struct ctb_tensor* r = ctb_tensor_from_axes(t->dtype, t->ndim - 2 * ntrace, axes, t->d != NULL);
	/* block order and block contents are unchanged by removing unit axes */
	CTB_REQUIRE(r->nstore == t->nstore);
	if (t->nstore > 0 && t->d != NULL) {
		CTB_CHECK_ABORT(ctbd_d2d(r->d, t->d, (size_t)t->nstore * ctb_sizeof_dtype(t->dtype)));
	}
	return r;
}

/* General cyclic partial trace (reference block_sparse_tensor.c:1560-1610): r[free] = sum_i t[i, free, i] over the 'ntrace' leading
 * and trailing axes.  Composed of the device primitives: transpose to [free, lead, trail], then ONE grouped-GEMM contraction over
 * (lead, trail) with the identity tensor delta[i, i'] of the traced legs (stored blocks: equal sectors only).  The pairs of unit
 * legs of the hot path never get here (ctb_drop_dummy_axes is a plain copy). */
struct ctb_tensor* ctb_cyclic_partial_trace(struct ctb_tensor* t, int ntrace)
{
	CTB_REQUIRE(ntrace >= 0 && t->ndim >= 2 * ntrace);
	bool unit = true;
	for (int i = 0; i < ntrace; i++) {
		const struct ctb_axis* a = &t->ax[i]; const struct ctb_axis* b = &t->ax[t->ndim - ntrace + i];
		CTB_REQUIRE(a->dim == b->dim && a->dir == -b->dir && ctb_axis_same_qnums(a, b));
		if (a->dim != 1) { unit = false; }
	}
	if (ntrace == 0) { return ctb_tensor_clone(t); }
	if (unit) { return ctb_drop_dummy_axes(t, ntrace); }
	const int nfree = t->ndim - 2 * ntrace;
	CTB_REQUIRE(2 * ntrace <= CTB_MAXDIM);
	/* [free..., lead..., trail...] */
	int perm[CTB_MAXDIM];
	for (int i = 0; i < nfree; i++) { perm[i] = ntrace + i; }
	for (int i = 0; i < ntrace; i++) { perm[nfree + i] = i; perm[nfree + ntrace + i] = t->ndim - ntrace + i; }
	struct ctb_tensor* tp = ctb_transpose(t, perm, 0);
	/* delta over the traced legs, directions opposite to those of tp's trailing legs */
	struct ctb_axis axes[CTB_MAXDIM];
	for (int i = 0; i < 2 * ntrace; i++) {
		ctb_axis_copy(&axes[i], &tp->ax[nfree + i]);
		axes[i].dir = -axes[i].dir;
	}
	struct ctb_tensor* delta = ctb_tensor_from_axes(t->dtype, 2 * ntrace, axes, 1);
	{
		/* host image of the packed entries: 1 where the leading multi-index equals the trailing one */
		const size_t esize = ctb_sizeof_dtype(t->dtype);
		void* ent = ctb_calloc((size_t)(delta->nstore > 0 ? delta->nstore : 1), esize);
		ct_long total = 1;
		for (int i = 0; i < ntrace; i++) { total *= delta->ax[i].dim; }
		ct_long idx[CTB_MAXDIM] = { 0 };
		for (ct_long k = 0; k < total; k++)
		{
			int sec[CTB_MAXDIM];
			for (int i = 0; i < ntrace; i++) { idx[ntrace + i] = idx[i]; }
			for (int i = 0; i < 2 * ntrace; i++) { sec[i] = delta->ax[i].sec_of[idx[i]]; }
			const ct_long base = ctb_grid_offset(delta, ctb_grid_ravel(delta, sec));
			CTB_REQUIRE(base >= 0);      /* equal quantum numbers with opposite directions always conserve */
			ct_long off = 0;
			for (int i = 0; i < 2 * ntrace; i++) { off = off * delta->ax[i].secdim[sec[i]] + delta->ax[i].pos_of[idx[i]]; }
			((double*)ent)[(size_t)(base + off) * (esize / sizeof(double))] = 1.0;
			for (int i = ntrace - 1; i >= 0; i--) {
				if (++idx[i] < delta->ax[i].dim) { break; }
				idx[i] = 0;
			}
		}
		CTB_CHECK_ABORT(ctb_upload_entries(delta, ent));
		ctb_free(ent);
	}
	struct ctb_tensor* r = ctb_dot(tp, TENSOR_AXIS_RANGE_TRAILING, 0, delta, TENSOR_AXIS_RANGE_LEADING, 0, 2 * ntrace, NULL);
	ctb_tensor_free(delta);
	ctb_tensor_free(tp);
	return r;
}

/* ---------------------------------------------------------------------------------------------- */
/* block-wise SVD / QR / RQ                                                                        */
/* ---------------------------------------------------------------------------------------------- */

/* Connecting-axis quantum numbers (reference :2694-2724 for SVD/QR with column sectors outer and ascending,
 * :2552-2581 for RQ with row sectors outer).  Returns the number of entries; 'q' has room for max(dim0, dim1). */
static ct_long interm_qnums(const struct ctb_tensor* a, int by_rows, qnumber* q)
{
	ct_long n = 0;
	const struct ctb_axis* outer = by_rows ? &a->ax[0] : &a->ax[1];
	const struct ctb_axis* inner = by_rows ? &a->ax[1] : &a->ax[0];
	for (int jo = 0; jo < outer->nsec; jo++) {
		for (int ji = 0; ji < inner->nsec; ji++) {
			if (outer->dir * outer->qsec[jo] + inner->dir * inner->qsec[ji] != 0) { continue; }
			const ct_long k = outer->secdim[jo] < inner->secdim[ji] ? outer->secdim[jo] : inner->secdim[ji];
			for (ct_long l = 0; l < k; l++) { q[n + l] = outer->qsec[jo]; }
			n += k;
		}
	}
	return n;
}

/* descriptors of all stored blocks of the matrix 'a' with their targets in (o0, o1); o0 = [rows, interm], o1 = [interm, cols] */
static int build_mat_descs(const struct ctb_tensor* a, const struct ctb_tensor* o0, const struct ctb_tensor* o1, int by_rows, struct ctbd_mat_desc* descs)
{
	int n = 0;
	for (int b = 0; b < a->nblk; b++)
	{
		int idx[2];
		ctb_grid_unravel(a, a->blk_grid[b], idx);
		const qnumber qc = by_rows ? a->ax[0].qsec[idx[0]] : a->ax[1].qsec[idx[1]];
		const int kk = ctb_axis_find_sector(&o0->ax[1], qc);
		CTB_REQUIRE(kk >= 0 && o1->ax[0].qsec[kk] == qc);
		struct ctbd_mat_desc* d = &descs[n++];
		d->a_off = a->blk_off[b];
		d->m = a->ax[0].secdim[idx[0]];
		d->n = a->ax[1].secdim[idx[1]];
		const int i0[2] = { idx[0], kk };
		const int i1[2] = { kk, idx[1] };
		d->o0_off = ctb_grid_offset(o0, ctb_grid_ravel(o0, i0));
		d->o1_off = ctb_grid_offset(o1, ctb_grid_ravel(o1, i1));
		CTB_REQUIRE(d->o0_off >= 0 && d->o1_off >= 0);
		const int kmin = d->m < d->n ? d->m : d->n;
		CTB_REQUIRE(o0->ax[1].secdim[kk] == kmin);
		/* logical indices of one quantum number on the connecting axis are contiguous (reference :2824-2834) */
		d->s_off = o0->ax[1].log_of[o0->ax[1].secstart[kk]];
	}
	return n;
}

/* One process per GPU ("SVD/QR: sectors are independent -> distribute sector blocks by cost m n min(m, n)"): the
 * sector blocks are dealt to the ranks longest-processing-time first, every rank factorises its own blocks, packs U, Vh and the
 * singular values of those blocks into one buffer, and ONE all-gather hands every rank the factors of all blocks (each block is
 * computed by exactly one rank, so all ranks end up with bit-identical tensors, as the replicated form gave them). */
struct svd_owner { double cost; int blk; };
static int cmp_svd_owner(const void* x, const void* y)
{
	const struct svd_owner* a = x; const struct svd_owner* b = y;
	if (a->cost != b->cost) { return a->cost > b->cost ? -1 : 1; }
	return (a->blk > b->blk) - (a->blk < b->blk);
}
static int svd_sharded(struct ctb_tensor* a, struct ctb_tensor* u, struct ctb_tensor* vh, double* s_dev, int nmat, const struct ctbd_mat_desc* descs)
{
	const int W = ctb_dist_world, me = ctb_dist_rank;
	const size_t es = ctb_sizeof_dtype(a->dtype);
	struct svd_owner* ord = malloc((size_t)nmat * sizeof(*ord));
	for (int b = 0; b < nmat; b++) {
		const double r = descs[b].m < descs[b].n ? descs[b].m : descs[b].n, c = descs[b].m < descs[b].n ? descs[b].n : descs[b].m;
		ord[b].cost = r * r * c; ord[b].blk = b;
	}
	qsort(ord, (size_t)nmat, sizeof(*ord), cmp_svd_owner);
	int* owner = malloc((size_t)nmat * sizeof(int));
	double* load = calloc((size_t)W, sizeof(double));
	int64_t* bytes = calloc((size_t)W, sizeof(int64_t));      /* packed size of every rank's results */
	int64_t* pos = malloc((size_t)nmat * sizeof(int64_t));    /* byte offset of a block inside its owner's packed buffer */
	for (int q = 0; q < nmat; q++) {
		int best = 0;
		for (int p = 1; p < W; p++) { if (load[p] < load[best]) { best = p; } }
		const int b = ord[q].blk;
		owner[b] = best; load[best] += ord[q].cost;
	}
	for (int b = 0; b < nmat; b++) {
		const int64_t k = descs[b].m < descs[b].n ? descs[b].m : descs[b].n;
		pos[b] = bytes[owner[b]];
		int64_t sz = ((int64_t)descs[b].m * k + k * (int64_t)descs[b].n) * (int64_t)es + k * (int64_t)sizeof(double);
		bytes[owner[b]] += (sz + 15) / 16 * 16;
	}
	int64_t slot = 16;
	for (int p = 0; p < W; p++) { if (bytes[p] > slot) { slot = bytes[p]; } }
	slot = (slot + 255) / 256 * 256;

	/* own blocks */
	struct ctbd_mat_desc* mine = malloc((size_t)nmat * sizeof(*mine));
	int nmine = 0;
	for (int b = 0; b < nmat; b++) { if (owner[b] == me) { mine[nmine++] = descs[b]; } }
	int rc = (nmine > 0) ? ctbd_svd_batched(a->dtype, nmine, mine, a->d, u->d, vh->d, s_dev) : 0;
	free(mine);

	void* pack = NULL;
	if (rc == 0) { rc = ctbd_malloc_noinit(&pack, (size_t)slot * (size_t)W); }
	if (rc == 0)
	{
		char* my = (char*)pack + (size_t)me * (size_t)slot;
		for (int b = 0; b < nmat && rc == 0; b++) {
			if (owner[b] != me) { continue; }
			const int64_t k = descs[b].m < descs[b].n ? descs[b].m : descs[b].n;
			const size_t nu = (size_t)descs[b].m * (size_t)k * es, nv = (size_t)k * (size_t)descs[b].n * es;
			rc = ctbd_d2d(my + pos[b], (const char*)u->d + (size_t)descs[b].o0_off * es, nu);
			if (rc == 0) { rc = ctbd_d2d(my + pos[b] + nu, (const char*)vh->d + (size_t)descs[b].o1_off * es, nv); }
			if (rc == 0) { rc = ctbd_d2d(my + pos[b] + nu + nv, (const char*)s_dev + (size_t)descs[b].s_off * sizeof(double), (size_t)k * sizeof(double)); }
		}
		if (rc == 0) { rc = ctbd_allgather(my, pack, (size_t)slot); }
		for (int b = 0; b < nmat && rc == 0; b++) {
			if (owner[b] == me) { continue; }
			const char* src = (const char*)pack + (size_t)owner[b] * (size_t)slot + pos[b];
			const int64_t k = descs[b].m < descs[b].n ? descs[b].m : descs[b].n;
			const size_t nu = (size_t)descs[b].m * (size_t)k * es, nv = (size_t)k * (size_t)descs[b].n * es;
			rc = ctbd_d2d((char*)u->d + (size_t)descs[b].o0_off * es, src, nu);
			if (rc == 0) { rc = ctbd_d2d((char*)vh->d + (size_t)descs[b].o1_off * es, src + nu, nv); }
			if (rc == 0) { rc = ctbd_d2d((char*)s_dev + (size_t)descs[b].s_off * sizeof(double), src + nu + nv, (size_t)k * sizeof(double)); }
		}
	}
	if (pack != NULL) { ctbd_free(pack); }
	free(pos); free(bytes); free(load); free(owner); free(ord);
	return rc;
}

static int svd_direct(struct ctb_tensor* a, struct ctb_tensor** u, double** s_dev, ct_long* ns, struct ctb_tensor** vh)
{
	CTB_REQUIRE(a->ndim == 2);
	const ct_long maxn = a->ax[0].dim > a->ax[1].dim ? a->ax[0].dim : a->ax[1].dim;
	qnumber* q = ctb_calloc(maxn, sizeof(qnumber));
	ct_long dim_interm = interm_qnums(a, 0, q);
	bool dummy = false;
	if (dim_interm == 0)
	{
		/* no stored block: dummy bond of dimension 1 (reference :2726-2776) */
		dummy = true;
		dim_interm = 1;
		q[0] = -a->ax[0].dir * a->ax[1].dir * a->ax[0].qsec[0];
	}
	{
		const ct_long dim_u[2] = { a->ax[0].dim, dim_interm };
		const int dir_u[2] = { a->ax[0].dir, a->ax[1].dir };
		const qnumber* qn_u[2] = { a->ax[0].qlog, q };
		*u = ctb_tensor_create(a->dtype, 2, dim_u, dir_u, qn_u, 1);
		const ct_long dim_v[2] = { dim_interm, a->ax[1].dim };
		const int dir_v[2] = { -a->ax[1].dir, a->ax[1].dir };
		const qnumber* qn_v[2] = { q, a->ax[1].qlog };
		*vh = ctb_tensor_create(a->dtype, 2, dim_v, dir_v, qn_v, 1);
	}
	ctb_free(q);
	*ns = dim_interm;
	CTB_CHECK(ctbd_malloc((void**)s_dev, (size_t)dim_interm * sizeof(double)));
	if (dummy)
	{
		const int iu[2] = { 0, ctb_axis_find_sector(&(*u)->ax[1], (*u)->ax[1].qlog[0]) };
		const ct_long off = ctb_grid_offset((*u), ctb_grid_ravel(*u, iu));
		CTB_REQUIRE(off >= 0);
		CTB_CHECK(ctb_set_entry(*u, off, 1.0, 0.0));
		return 0;
	}
	struct ctbd_mat_desc* descs = malloc((size_t)a->nblk * sizeof(*descs));
	const int nmat = build_mat_descs(a, *u, *vh, 0, descs);
	int rc;
	if (ctb_dist_world > 1 && nmat > 1 && getenv("CTB_NO_SHARDED_SVD") == NULL) { rc = svd_sharded(a, *u, *vh, *s_dev, nmat, descs); }
	else { rc = ctbd_svd_batched(a->dtype, nmat, descs, a->d, (*u)->d, (*vh)->d, *s_dev); }
	free(descs);
	if (rc < 0) { fprintf(stderr, "chemtensor_b200: batched SVD failed: %s\n", ctbd_last_error()); return -1; }
	return 0;
}

/* QR-preconditioned SVD (Drmac-Veselic flavour), built from the device primitives alone: a = q r per sector block (Householder),
 * one-sided Jacobi on the ROWS of the triangular factor r, u = q u_r by the grouped GEMM.  On the graded spectra of two-site
 * tensors the Jacobi iteration on r needs 3-4x fewer rotations than on a itself (tools/proto_block_jacobi.py and DESIGN.md
 * section 6: 7-9 sweeps instead of 23-31 over 8-16 decades), and for tall blocks its rows are shorter.  Sector structure, bond
 * quantum numbers and singular values are those of the direct path.  Opt-in (CTB_SVD_PRECONDITION=1) until it has been timed
 * on the GPU against the Householder kernel's cost on large blocks. */
static int svd_preconditioned_tall(struct ctb_tensor* a, struct ctb_tensor** u, double** s_dev, ct_long* ns, struct ctb_tensor** vh)
{
	struct ctb_tensor *q = NULL, *r = NULL, *u_r = NULL;
	int rc = ctb_qr(a, &q, &r);
	if (rc < 0) { return rc; }
	rc = svd_direct(r, &u_r, s_dev, ns, vh);
	ctb_tensor_free(r);
	if (rc < 0) { ctb_tensor_free(q); return rc; }
	*u = ctb_dot(q, TENSOR_AXIS_RANGE_TRAILING, 0, u_r, TENSOR_AXIS_RANGE_LEADING, 0, 1, NULL);
	ctb_tensor_free(q);
	ctb_tensor_free(u_r);
	return 0;
}

int ctb_svd(struct ctb_tensor* a, struct ctb_tensor** u, double** s_dev, ct_long* ns, struct ctb_tensor** vh)
{
	CTB_REQUIRE(a->ndim == 2);
	const char* env = getenv("CTB_SVD_PRECONDITION");
	if (env == NULL || atoi(env) == 0 || a->nblk == 0) { return svd_direct(a, u, s_dev, ns, vh); }
	/* orientation: the Householder step pays off on blocks with at least as many rows as columns */
	double tall = 0, wide = 0;
	for (int b = 0; b < a->nblk; b++) {
		int idx[2];
		ctb_grid_unravel(a, a->blk_grid[b], idx);
		const double m = a->ax[0].secdim[idx[0]], n = a->ax[1].secdim[idx[1]];
		if (m >= n) { tall += m * n * n; } else { wide += m * m * n; }
	}
	if (wide <= tall || a->ax[0].dir != -a->ax[1].dir) { return svd_preconditioned_tall(a, u, s_dev, ns, vh); }
	/* mostly wide blocks: factorise the conjugate transpose and swap the factors back.  With opposite leg directions the bond of
	 * the transposed problem carries the same quantum numbers in the same order (q_row = q_col for every conserving block). */
	const int perm[2] = { 1, 0 };
	const int cj = ctb_is_complex(a->dtype);
	struct ctb_tensor* at = ctb_transpose(a, perm, cj);
	struct ctb_tensor *ut = NULL, *vht = NULL;
	int rc = svd_preconditioned_tall(at, &ut, s_dev, ns, &vht);
	ctb_tensor_free(at);
	if (rc < 0) { return rc; }
	*u = ctb_transpose(vht, perm, cj);
	*vh = ctb_transpose(ut, perm, cj);
	ctb_tensor_free(ut); ctb_tensor_free(vht);
	return 0;
}

static int qr_common(struct ctb_tensor* a, int rq, struct ctb_tensor** o0, struct ctb_tensor** o1)
{
	CTB_REQUIRE(a->ndim == 2);
	const ct_long maxn = a->ax[0].dim > a->ax[1].dim ? a->ax[0].dim : a->ax[1].dim;
	qnumber* q = ctb_calloc(maxn, sizeof(qnumber));
	ct_long dim_interm = interm_qnums(a, rq, q);
	bool dummy = false;
	if (dim_interm == 0)
	{
		dummy = true;
		dim_interm = 1;
		if (!rq) { q[0] = -a->ax[0].dir * a->ax[1].dir * a->ax[0].qsec[0]; }   /* reference :2443-2452 */
		else     { q[0] = -a->ax[0].dir * a->ax[1].dir * a->ax[1].qsec[0]; }   /* reference :2585-2594 */
	}
	{
		/* QR: q keeps a's directions, r = (-dir1, dir1);  RQ: r = (dir0, -dir0), q keeps a's directions */
		const ct_long dim0[2] = { a->ax[0].dim, dim_interm };
		const int dir0[2] = { a->ax[0].dir, rq ? -a->ax[0].dir : a->ax[1].dir };
		const qnumber* qn0[2] = { a->ax[0].qlog, q };
		*o0 = ctb_tensor_create(a->dtype, 2, dim0, dir0, qn0, 1);
		const ct_long dim1[2] = { dim_interm, a->ax[1].dim };
		const int dir1[2] = { rq ? a->ax[0].dir : -a->ax[1].dir, a->ax[1].dir };
		const qnumber* qn1[2] = { q, a->ax[1].qlog };
		*o1 = ctb_tensor_create(a->dtype, 2, dim1, dir1, qn1, 1);
	}
	const qnumber q0 = q[0];
	ctb_free(q);
	if (dummy)
	{
		/* the isometry gets a single entry 1, the triangular factor stays zero */
		if (!rq) {
			const int iq[2] = { 0, ctb_axis_find_sector(&(*o0)->ax[1], q0) };
			const ct_long off = ctb_grid_offset((*o0), ctb_grid_ravel(*o0, iq));
			CTB_REQUIRE(off >= 0);
			CTB_CHECK(ctb_set_entry(*o0, off, 1.0, 0.0));
		}
		else {
			const int iq[2] = { ctb_axis_find_sector(&(*o1)->ax[0], q0), 0 };
			const ct_long off = ctb_grid_offset((*o1), ctb_grid_ravel(*o1, iq));
			CTB_REQUIRE(off >= 0);
			CTB_CHECK(ctb_set_entry(*o1, off, 1.0, 0.0));
		}
		return 0;
	}
	struct ctbd_mat_desc* descs = malloc((size_t)a->nblk * sizeof(*descs));
	const int nmat = build_mat_descs(a, *o0, *o1, rq, descs);
	int rc = ctbd_qr_batched(a->dtype, rq, nmat, descs, a->d, (*o0)->d, (*o1)->d);
	free(descs);
	if (rc < 0) { fprintf(stderr, "chemtensor_b200: batched QR failed: %s\n", ctbd_last_error()); return -1; }
	return 0;
}

int ctb_qr(struct ctb_tensor* a, struct ctb_tensor** q, struct ctb_tensor** r) { return qr_common(a, 0, q, r); }
int ctb_rq(struct ctb_tensor* a, struct ctb_tensor** r, struct ctb_tensor** q) { return qr_common(a, 1, r, q); }
