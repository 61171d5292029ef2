/*
 * ctbd_emu.c -- TEST DOUBLE for the thin CUDA layer (include/ctb_device.h).
 *
 * NOT PART OF THE PRODUCT.  This file implements the ctbd_* C-ABI with plain host loops so that
 * the host-side plan builders (sector bookkeeping, work lists, offset tables, DMRG driver logic)
 * can be exercised by `pytest -m "not gpu"` on a machine without a GPU.  It is compiled only into
 * tests/emu/libctb_hostlogic_emu.so by the test-suite; libchemtensor_b200.so never contains it and
 * never falls back to it (ctbd_backend() reports 2 here, 1 for the CUDA layer).
 *
 * The algorithms mirror the CUDA kernels' formulation (grouped GEMM over segments with offset-table
 * epilogue, logical-index remap, one-sided Jacobi SVD on rows, Householder QR with the RQ index
 * transform) so that the mathematics of the kernels is validated independently of the GPU.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <complex.h>
#include <fcntl.h>
#include <unistd.h>
#include <sys/mman.h>
#include "ctb_device.h"

static char g_err[512] = "";
static long long g_launches = 0;
static long long g_bytes = 0;

int ctbd_init(int device) { (void)device; return 0; }
int ctbd_shutdown(void) { return 0; }
int ctbd_backend(void) { return 2; }
const char* ctbd_last_error(void) { return g_err; }
long long ctbd_launch_count(void) { return g_launches; }
int ctbd_sm_count(void) { return 1; }
void* ctbd_stream(void) { return NULL; }
int ctbd_event_create(void** ev) { *ev = malloc(8); return 0; }
int ctbd_event_record(void* ev) { (void)ev; return 0; }
int ctbd_event_elapsed_ms(void* a, void* b, float* ms) { (void)a; (void)b; *ms = 0.f; return 0; }
int ctbd_event_destroy(void* ev) { free(ev); return 0; }

struct hdr { size_t bytes; size_t pad; };
int ctbd_malloc(void** dptr, size_t bytes)
{
	struct hdr* h = calloc(1, sizeof(struct hdr) + (bytes ? bytes : 1));
	if (!h) { snprintf(g_err, sizeof g_err, "emu: out of memory"); return -1; }
	h->bytes = bytes; g_bytes += (long long)bytes;
	*dptr = h + 1;
	return 0;
}
/* the test double poisons un-initialised memory (all-ones bytes = NaN doubles) so that the CPU suite catches reads of it */
int ctbd_malloc_noinit(void** dptr, size_t bytes)
{
	int rc = ctbd_malloc(dptr, bytes);
	if (rc == 0) { memset(*dptr, 0xFF, bytes); }
	return rc;
}
int ctbd_free(void* dptr) { if (dptr) { struct hdr* h = (struct hdr*)dptr - 1; g_bytes -= (long long)h->bytes; free(h); } return 0; }
int ctbd_memset_zero(void* dptr, size_t bytes) { memset(dptr, 0, bytes); return 0; }
int ctbd_h2d(void* d, const void* h, size_t bytes) { memcpy(d, h, bytes); return 0; }
int ctbd_d2h(void* h, const void* d, size_t bytes) { memcpy(h, d, bytes); return 0; }
int ctbd_h2d_blocks(void* dptr, int nblk, const void* const* hptrs, const int64_t* dst_off, const int64_t* nbytes)
{
	for (int b = 0; b < nblk; b++) { memcpy((char*)dptr + dst_off[b], hptrs[b], (size_t)nbytes[b]); }
	return 0;
}
int ctbd_d2h_blocks(const void* dptr, int nblk, void* const* hptrs, const int64_t* src_off, const int64_t* nbytes)
{
	for (int b = 0; b < nblk; b++) { memcpy(hptrs[b], (const char*)dptr + src_off[b], (size_t)nbytes[b]); }
	return 0;
}
/* exchange step: the test double only knows the host-callback form (gloo in the tests) */
static ctbd_allgather_fn g_ag_fn = NULL; static void* g_ag_ctx = NULL; static int g_world = 1, g_rank_emu = 0;
int ctbd_dist_unique_id(void* id_out) { memset(id_out, 0, CTBD_UNIQUE_ID_BYTES); return 0; }
int ctbd_dist_init(int rank, int world, const void* unique_id) { (void)unique_id; g_world = world; g_rank_emu = rank; return 0; }
int ctbd_dist_set_allgather(ctbd_allgather_fn fn, void* ctx) { g_ag_fn = fn; g_ag_ctx = ctx; return 0; }
int ctbd_dist_finalize(void) { g_world = 1; g_ag_fn = NULL; g_ag_ctx = NULL; return 0; }
int ctbd_allgather(const void* sendbuf, void* recvbuf, size_t bytes_per_rank)
{
	if (g_world == 1) { if (sendbuf != recvbuf) { memmove(recvbuf, sendbuf, bytes_per_rank); } return 0; }
	if (g_ag_fn == NULL) { snprintf(g_err, sizeof(g_err), "dist: no all-gather callback registered"); return -1; }
	return g_ag_fn(g_ag_ctx, sendbuf, recvbuf, bytes_per_rank, NULL);
}
int ctbd_barrier(void)
{
	if (g_world == 1) { return 0; }
	static long long scratch[64];
	return ctbd_allgather(scratch, scratch + 1, 8);
}
/* peer-mapped buffers of the test double: POSIX shared memory between the rank processes of one test run */
struct emu_peer { int world; size_t bytes; void** ptrs; char (*names)[96]; };
static int g_peer_counter = 0;
int ctbd_peer_buffer_create(size_t bytes, void** handle)
{
	*handle = NULL;
	if (g_world == 1 || g_ag_fn == NULL || getenv("CTB_NO_PEER") != NULL) { snprintf(g_err, sizeof(g_err), "peer buffers need a multi-rank run"); return -1; }
	struct emu_peer* pb = calloc(1, sizeof(*pb));
	pb->world = g_world; pb->bytes = bytes ? bytes : 64;
	pb->ptrs = calloc((size_t)g_world, sizeof(void*));
	pb->names = calloc((size_t)g_world, sizeof(*pb->names));
	const char* port = getenv("MASTER_PORT");
	const int id = g_peer_counter++;
	for (int p = 0; p < g_world; p++) { snprintf(pb->names[p], sizeof(pb->names[p]), "/ctb_emu_%s_%d_%d", port ? port : "0", id, p); }
	int fd = shm_open(pb->names[g_rank_emu], O_CREAT | O_RDWR, 0600);
	if (fd < 0 || ftruncate(fd, (off_t)pb->bytes) != 0) { snprintf(g_err, sizeof(g_err), "shm_open failed"); return -1; }
	pb->ptrs[g_rank_emu] = mmap(NULL, pb->bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
	close(fd);
	if (ctbd_barrier() < 0) { return -1; }
	for (int p = 0; p < g_world; p++) {
		if (p == g_rank_emu) { continue; }
		fd = shm_open(pb->names[p], O_RDWR, 0600);
		if (fd < 0) { snprintf(g_err, sizeof(g_err), "shm_open of a peer failed"); return -1; }
		pb->ptrs[p] = mmap(NULL, pb->bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
		close(fd);
	}
	if (ctbd_barrier() < 0) { return -1; }
	shm_unlink(pb->names[g_rank_emu]);      /* mappings stay valid; the name disappears */
	*handle = pb;
	return 0;
}
int ctbd_peer_buffer_ptrs(void* handle, void** ptrs) { struct emu_peer* pb = handle; for (int p = 0; p < pb->world; p++) { ptrs[p] = pb->ptrs[p]; } return 0; }
int ctbd_peer_buffer_destroy(void* handle)
{
	struct emu_peer* pb = handle;
	if (pb == NULL) { return 0; }
	ctbd_barrier();
	for (int p = 0; p < pb->world; p++) { if (pb->ptrs[p] != NULL) { munmap(pb->ptrs[p], pb->bytes); } }
	free(pb->ptrs); free(pb->names); free(pb);
	return 0;
}
/* "multicast" buffers of the test double: the same shared-memory team; the multicast address is a token (the handle itself), and a
 * store to it (ctbd_gemm_run_mc) is carried out as stores into the mappings of all ranks -- what the NVSwitch does in hardware */
int ctbd_mc_buffer_create(size_t bytes, void** handle, void** local_ptr, void** mc_ptr)
{
	*local_ptr = NULL; *mc_ptr = NULL;
	if (getenv("CTB_NO_MULTICAST") != NULL) { *handle = NULL; snprintf(g_err, sizeof(g_err), "multicast disabled"); return -1; }
	if (ctbd_peer_buffer_create(bytes, handle) < 0) { return -1; }
	struct emu_peer* pb = *handle;
	*local_ptr = pb->ptrs[g_rank_emu];
	*mc_ptr = pb;
	return 0;
}
int ctbd_mc_buffer_destroy(void* handle) { return ctbd_peer_buffer_destroy(handle); }
int ctbd_gemm_run_multi(void* plan, const void* A, const void* B, int ndst, void* const* Cs);
int ctbd_gemm_run_mc(void* plan, const void* A, const void* B, void* C_mc)
{
	struct emu_peer* pb = C_mc;
	return ctbd_gemm_run_multi(plan, A, B, pb->world, (void* const*)pb->ptrs);
}
int ctbd_gemm_run_multi(void* plan, const void* A, const void* B, int ndst, void* const* Cs)
{
	for (int d = 0; d < ndst; d++) { if (Cs[d] != NULL) { int rc = ctbd_gemm_run(plan, A, B, Cs[d]); if (rc < 0) { return rc; } } }
	return 0;
}
int ctbd_d2d(void* dst, const void* src, size_t bytes) { memmove(dst, src, bytes); return 0; }
int ctbd_host_prefault(int nblk, void* const* hptrs, const int64_t* nbytes) { (void)nblk; (void)hptrs; (void)nbytes; return 0; }
int ctbd_sync(void) { return 0; }
int ctbd_host_alloc(void** hptr, size_t bytes) { *hptr = malloc(bytes ? bytes : 1); return *hptr ? 0 : -1; }
int ctbd_host_free(void* hptr) { free(hptr); return 0; }
long long ctbd_bytes_in_use(void) { return g_bytes; }

/* ---- grouped GEMM ---- */
struct emu_plan { struct ctbd_gemm_plan_host h; struct ctbd_gemm_out* outs; struct ctbd_gemm_seg* segs; int32_t* tab; void* a_packed; int64_t* b_rowtab;
	struct ctbd_mix_group* mix_groups; struct ctbd_mix_row* mix_rows; };

static void* dup_mem(const void* p, size_t n) { void* q = malloc(n ? n : 1); if (n) { memcpy(q, p, n); } return q; }

int ctbd_gemm_plan_create(const struct ctbd_gemm_plan_host* h, void** plan)
{
	struct emu_plan* p = calloc(1, sizeof(*p));
	p->h = *h;
	p->outs  = dup_mem(h->outs,  (size_t)h->nouts  * sizeof(*h->outs));
	p->segs  = dup_mem(h->segs,  (size_t)h->nsegs  * sizeof(*h->segs));
	p->tab   = dup_mem(h->tab,   (size_t)h->ntab   * sizeof(int32_t));
	p->b_rowtab = NULL;
	if (h->b_rowtab != NULL && h->n_b_rowtab > 0) { p->b_rowtab = dup_mem(h->b_rowtab, (size_t)h->n_b_rowtab * sizeof(int64_t)); }
	p->mix_groups = NULL; p->mix_rows = NULL;
	if (h->mix_groups != NULL && h->n_mix_groups > 0) {
		p->mix_groups = dup_mem(h->mix_groups, (size_t)h->n_mix_groups * sizeof(*h->mix_groups));
		p->mix_rows = dup_mem(h->mix_rows, (size_t)h->n_mix_rows * sizeof(*h->mix_rows));
	}
	p->a_packed = NULL;
	if (h->a_gather != NULL && h->n_a_gather > 0) {
		const size_t es = (h->dtype == CTBD_C128) ? 16 : 8;
		p->a_packed = calloc((size_t)h->n_a_gather, es);
		for (int64_t i = 0; i < h->n_a_gather; i++) {
			if (h->a_gather[i] >= 0) { memcpy((char*)p->a_packed + (size_t)i * es, (const char*)h->a_src + (size_t)h->a_gather[i] * es, es); }
		}
	}
	/* analysis aid: CTB_EMU_PLAN_DUMP=<file> appends the block shapes of every plan (one line per output block: m n k1 k2 ...) */
	const char* dump = getenv("CTB_EMU_PLAN_DUMP");
	if (dump != NULL && h->n_mix_groups == 0) {
		FILE* f = fopen(dump, "a");
		if (f != NULL) {
			fprintf(f, "plan %d %d %d %d\n", h->dtype, h->a_kcontig, h->b_ncontig, h->nouts);
			for (int b = 0; b < h->nouts; b++) {
				fprintf(f, "%d %d", h->outs[b].m, h->outs[b].n);
				for (int sg = h->outs[b].seg_begin; sg < h->outs[b].seg_end; sg++) { fprintf(f, " %d", h->segs[sg].k); }
				fprintf(f, "\n");
			}
			fclose(f);
		}
	}
	*plan = p;
	return 0;
}
int ctbd_gemm_plan_destroy(void* plan)
{
	struct emu_plan* p = plan;
	free(p->outs); free(p->segs); free(p->tab); free(p->a_packed); free(p->b_rowtab); free(p->mix_groups); free(p->mix_rows); free(p);
	return 0;
}

int ctbd_gemm_plan_info(void* plan, int* ntiles, int* nlaunches)
{
	struct emu_plan* p = plan;
	if (ntiles) { *ntiles = p->h.nouts + p->h.n_mix_groups; }
	if (nlaunches) { *nlaunches = 1; }
	return 0;
}

int ctbd_gemm_run(void* plan, const void* A, const void* B, void* C)
{
	struct emu_plan* p = plan;
	g_launches++;
	const int cplx = (p->h.dtype == CTBD_C128);
	if (p->a_packed != NULL) { A = p->a_packed; }
	for (int gi = 0; gi < (p->mix_groups != NULL ? p->h.n_mix_groups : 0); gi++)
	{
		/* mixing form, straight from its definition in ctb_device.h */
		const struct ctbd_mix_group* g = &p->mix_groups[gi];
		for (int r = g->row_begin; r < g->row_end; r++) {
			const struct ctbd_mix_row* rw = &p->mix_rows[r];
			for (int j = 0; j < g->n; j++) {
				double complex acc = 0;
				for (int k = 0; k < g->kp; k++) {
					const int64_t ib = p->b_rowtab[g->brow_begin + k] + j;
					if (cplx) {
						double complex a = ((const double complex*)A)[rw->a_off + k], b = ((const double complex*)B)[ib];
						if (p->h.conj_a) { a = conj(a); }
						if (p->h.conj_b) { b = conj(b); }
						acc += a * b;
					}
					else { acc += ((const double*)A)[rw->a_off + k] * ((const double*)B)[ib]; }
				}
				int64_t off = rw->c_off; int rem = j;
				for (int a = g->ndig - 1; a >= 0; a--) { off += (int64_t)(rem % g->dig_dim[a]) * rw->cs[a]; rem /= g->dig_dim[a]; }
				if (cplx) { ((double complex*)C)[off] = acc; } else { ((double*)C)[off] = creal(acc); }
			}
		}
	}
	for (int t = 0; t < p->h.nouts; t++)
	{
		const struct ctbd_gemm_out* o = &p->outs[t];
		for (int i = 0; i < o->m; i++) {
			for (int j = 0; j < o->n; j++)
			{
				double complex acc = 0;
				for (int s = o->seg_begin; s < o->seg_end; s++)
				{
					const struct ctbd_gemm_seg* g = &p->segs[s];
					for (int kk = 0; kk < g->k; kk++)
					{
						const int64_t ia = p->h.a_kcontig ? g->a_off + (int64_t)i * g->lda + kk : g->a_off + (int64_t)kk * g->lda + i;
						const int64_t ib = (p->b_rowtab != NULL) ? p->b_rowtab[g->b_off + kk] + j
							: (p->h.b_ncontig ? g->b_off + (int64_t)kk * g->ldb + j : g->b_off + (int64_t)j * g->ldb + kk);
						if (cplx) {
							double complex a = ((const double complex*)A)[ia], b = ((const double complex*)B)[ib];
							if (p->h.conj_a) { a = conj(a); }
							if (p->h.conj_b) { b = conj(b); }
							acc += a * b;
						}
						else {
							acc += ((const double*)A)[ia] * ((const double*)B)[ib];
						}
					}
				}
				const int64_t ic = (o->col_tab >= 0) ? o->c_off + p->tab[o->row_tab + i] + p->tab[o->col_tab + j]
					: o->c_off + p->tab[o->row_tab + i] + p->tab[p->tab[o->row_tab + o->m + i] + j];
				if (cplx) { ((double complex*)C)[ic] = acc; } else { ((double*)C)[ic] = creal(acc); }
			}
		}
	}
	return 0;
}

/* ---- layouts and remaps ---- */
struct emu_layout
{
	int ndim, dtype; int64_t dim[CTBD_MAXDIM]; int nsec[CTBD_MAXDIM];
	int32_t *sec_of[CTBD_MAXDIM], *pos_of[CTBD_MAXDIM], *secstart[CTBD_MAXDIM], *log_of[CTBD_MAXDIM];
	int64_t ngrid; int64_t* grid_off; int nblk; int64_t* blk_grid; int64_t* blk_off; int64_t nstore;
};

int ctbd_layout_create(const struct ctbd_layout_host* h, void** layout)
{
	struct emu_layout* L = calloc(1, sizeof(*L));
	L->ndim = h->ndim; L->dtype = h->dtype; L->ngrid = h->ngrid; L->nblk = h->nblk; L->nstore = h->nstore;
	for (int i = 0; i < h->ndim; i++) {
		L->dim[i] = h->dim[i]; L->nsec[i] = h->nsec[i];
		L->sec_of[i]   = dup_mem(h->sec_of[i],   (size_t)h->dim[i] * 4);
		L->pos_of[i]   = dup_mem(h->pos_of[i],   (size_t)h->dim[i] * 4);
		L->secstart[i] = dup_mem(h->secstart[i], (size_t)(h->nsec[i] + 1) * 4);
		L->log_of[i]   = dup_mem(h->log_of[i],   (size_t)h->dim[i] * 4);
	}
	L->grid_off = dup_mem(h->grid_off, (size_t)h->ngrid * 8);
	L->blk_grid = dup_mem(h->blk_grid, (size_t)h->nblk * 8);
	L->blk_off  = dup_mem(h->blk_off,  (size_t)(h->nblk + 1) * 8);
	*layout = L;
	return 0;
}
int ctbd_layout_destroy(void* layout)
{
	struct emu_layout* L = layout;
	for (int i = 0; i < L->ndim; i++) { free(L->sec_of[i]); free(L->pos_of[i]); free(L->secstart[i]); free(L->log_of[i]); }
	free(L->grid_off); free(L->blk_grid); free(L->blk_off); free(L);
	return 0;
}

struct emu_copy_plan { int dtype, n; struct ctbd_copy2d* descs; };
int ctbd_copy_plan_create(int dtype, int n, const struct ctbd_copy2d* descs_host, void** plan)
{
	struct emu_copy_plan* p = calloc(1, sizeof(*p));
	p->dtype = dtype; p->n = n; p->descs = dup_mem(descs_host, (size_t)n * sizeof(*descs_host));
	*plan = p;
	return 0;
}
int ctbd_copy_plan_run(void* plan, const void* src, void* dst)
{
	const struct emu_copy_plan* p = plan;
	g_launches++;
	const size_t es = (p->dtype == CTBD_C128) ? 16 : 8;
	for (int k = 0; k < p->n; k++) {
		const struct ctbd_copy2d* d = &p->descs[k];
		for (int i = 0; i < d->rows; i++) {
			memcpy((char*)dst + (size_t)(d->dst_off + (int64_t)i * d->dst_ld) * es, (const char*)src + (size_t)(d->src_off + (int64_t)i * d->src_ld) * es, (size_t)d->cols * es);
		}
	}
	return 0;
}
int ctbd_copy_plan_run_multi(void* plan, int nsrc, const void* const* srcs, int64_t src_stride, void* dst)
{
	const struct emu_copy_plan* p = plan;
	g_launches++;
	const size_t es = (p->dtype == CTBD_C128) ? 16 : 8;
	for (int k = 0; k < p->n; k++) {
		const struct ctbd_copy2d* d = &p->descs[k];
		const int64_t q = d->src_off / src_stride, off = d->src_off % src_stride;
		if (q < 0 || q >= nsrc) { return -1; }
		for (int i = 0; i < d->rows; i++) {
			memcpy((char*)dst + (size_t)(d->dst_off + (int64_t)i * d->dst_ld) * es, (const char*)srcs[q] + (size_t)(off + (int64_t)i * d->src_ld) * es, (size_t)d->cols * es);
		}
	}
	return 0;
}
int ctbd_copy_plan_run_push(void* plan, const void* src, int ndst, void* const* dsts)
{
	for (int q = 0; q < ndst; q++) { if (ctbd_copy_plan_run(plan, src, dsts[q]) < 0) { return -1; } }
	return 0;
}
int ctbd_copy_plan_destroy(void* plan) { struct emu_copy_plan* p = plan; if (p) { free(p->descs); free(p); } return 0; }

int ctbd_remap(const struct ctbd_remap_args* a)
{
	g_launches++;
	const struct emu_layout* D = a->dst_layout; const struct emu_layout* S = a->src_layout;
	const int cplx = (D->dtype == CTBD_C128);
	if (a->op == CTBD_REMAP_UNSLICE)
	{
		/* scatter: every source entry goes to the destination entry with index ind[.] on axis i_ax */
		for (int b = 0; b < S->nblk; b++)
		{
			int sec[CTBD_MAXDIM]; int64_t bdim[CTBD_MAXDIM];
			int64_t cell = S->blk_grid[b];
			for (int i = S->ndim - 1; i >= 0; i--) { sec[i] = (int)(cell % S->nsec[i]); cell /= S->nsec[i]; }
			int64_t numel = 1;
			for (int i = 0; i < S->ndim; i++) { bdim[i] = S->secstart[i][sec[i] + 1] - S->secstart[i][sec[i]]; numel *= bdim[i]; }
			for (int64_t e = 0; e < numel; e++)
			{
				int64_t ld[CTBD_MAXDIM];
				int64_t r = e;
				for (int i = S->ndim - 1; i >= 0; i--) {
					const int64_t pos = r % bdim[i]; r /= bdim[i];
					const int64_t ls = S->log_of[i][S->secstart[i][sec[i]] + pos];
					ld[i] = (i == a->i_ax) ? a->ind[ls] : ls;
				}
				int64_t dcell = 0, doff = 0;
				for (int i = 0; i < D->ndim; i++) {
					const int sc = D->sec_of[i][ld[i]];
					dcell = dcell * D->nsec[i] + sc;
					doff = doff * (D->secstart[i][sc + 1] - D->secstart[i][sc]) + D->pos_of[i][ld[i]];
				}
				const int64_t dbase = D->grid_off[dcell];
				if (dbase < 0) { continue; }
				if (cplx) { ((double complex*)a->dst)[dbase + doff] = ((const double complex*)a->src)[S->blk_off[b] + e]; }
				else { ((double*)a->dst)[dbase + doff] = ((const double*)a->src)[S->blk_off[b] + e]; }
			}
		}
		return 0;
	}
	for (int b = 0; b < D->nblk; b++)
	{
		int sec[CTBD_MAXDIM]; int64_t bdim[CTBD_MAXDIM];
		int64_t cell = D->blk_grid[b];
		for (int i = D->ndim - 1; i >= 0; i--) { sec[i] = (int)(cell % D->nsec[i]); cell /= D->nsec[i]; }
		int64_t numel = 1;
		for (int i = 0; i < D->ndim; i++) { bdim[i] = D->secstart[i][sec[i] + 1] - D->secstart[i][sec[i]]; numel *= bdim[i]; }
		for (int64_t e = 0; e < numel; e++)
		{
			int64_t ld[CTBD_MAXDIM], ls[CTBD_MAXDIM];
			int64_t r = e;
			for (int i = D->ndim - 1; i >= 0; i--) { const int64_t pos = r % bdim[i]; r /= bdim[i]; ld[i] = D->log_of[i][D->secstart[i][sec[i]] + pos]; }
			switch (a->op) {
				case CTBD_REMAP_TRANSPOSE: for (int i = 0; i < D->ndim; i++) { ls[a->perm[i]] = ld[i]; } break;
				case CTBD_REMAP_FLATTEN:
					for (int i = 0; i < a->i_ax; i++) { ls[i] = ld[i]; }
					ls[a->i_ax] = ld[a->i_ax] / S->dim[a->i_ax + 1]; ls[a->i_ax + 1] = ld[a->i_ax] % S->dim[a->i_ax + 1];
					for (int i = a->i_ax + 1; i < D->ndim; i++) { ls[i + 1] = ld[i]; }
					break;
				case CTBD_REMAP_SPLIT:
					for (int i = 0; i < a->i_ax; i++) { ls[i] = ld[i]; }
					ls[a->i_ax] = ld[a->i_ax] * D->dim[a->i_ax + 1] + ld[a->i_ax + 1];
					for (int i = a->i_ax + 2; i < D->ndim; i++) { ls[i - 1] = ld[i]; }
					break;
				case CTBD_REMAP_SLICE: for (int i = 0; i < D->ndim; i++) { ls[i] = (i == a->i_ax) ? a->ind[ld[i]] : ld[i]; } break;
				default: for (int i = 0; i < D->ndim; i++) { ls[i] = ld[i]; } break;
			}
			int64_t scell = 0, soff = 0;
			int64_t sbd[CTBD_MAXDIM], spos[CTBD_MAXDIM];
			for (int i = 0; i < S->ndim; i++) {
				const int s = S->sec_of[i][ls[i]];
				scell = scell * S->nsec[i] + s;
				sbd[i] = S->secstart[i][s + 1] - S->secstart[i][s];
				spos[i] = S->pos_of[i][ls[i]];
			}
			for (int i = 0; i < S->ndim; i++) { soff = soff * sbd[i] + spos[i]; }
			const int64_t sbase = S->grid_off[scell];
			double complex v = 0;
			if (sbase >= 0) { v = cplx ? ((const double complex*)a->src)[sbase + soff] : ((const double*)a->src)[sbase + soff]; }
			if (a->conj) { v = conj(v); }
			if (a->scale_ax >= 0) { v *= a->scale[ld[a->scale_ax]]; }
			if (cplx) { ((double complex*)a->dst)[D->blk_off[b] + e] = v; } else { ((double*)a->dst)[D->blk_off[b] + e] = creal(v); }
		}
	}
	return 0;
}

/* ---- level 1 ---- */
#define GETC(p, i) (cplx ? ((const double complex*)(p))[i] : (double complex)((const double*)(p))[i])
#define PUTC(p, i, v) do { if (cplx) { ((double complex*)(p))[i] = (v); } else { ((double*)(p))[i] = creal(v); } } while (0)

int ctbd_dotc(int dtype, int64_t n, const void* x, const void* y, double* out)
{
	g_launches++;
	const int cplx = (dtype == CTBD_C128);
	double complex s = 0;
	for (int64_t i = 0; i < n; i++) { s += conj(GETC(x, i)) * GETC(y, i); }
	out[0] = creal(s); out[1] = cimag(s);
	return 0;
}
int ctbd_nrm2(int dtype, int64_t n, const void* x, double* out)
{
	g_launches++;
	const int cplx = (dtype == CTBD_C128);
	double s = 0;
	for (int64_t i = 0; i < n; i++) { const double complex v = GETC(x, i); s += creal(v) * creal(v) + cimag(v) * cimag(v); }
	out[0] = sqrt(s);
	return 0;
}
int ctbd_rscale(int dtype, int64_t n, const void* x, const double* s, int divide, void* y)
{
	g_launches++;
	const int cplx = (dtype == CTBD_C128);
	const double f = divide ? 1.0 / s[0] : s[0];
	for (int64_t i = 0; i < n; i++) { PUTC(y, i, f * GETC(x, i)); }
	return 0;
}
int ctbd_lanczos_update(int dtype, int64_t n, void* w, const void* vj, const void* vjm1, const double* alpha, const double* beta_prev, double* out)
{
	g_launches++;
	const int cplx = (dtype == CTBD_C128);
	double s = 0;
	for (int64_t i = 0; i < n; i++) {
		double complex v = GETC(w, i) - alpha[0] * GETC(vj, i);
		if (vjm1 != NULL) { v -= beta_prev[0] * GETC(vjm1, i); }
		PUTC(w, i, v);
		s += creal(v) * creal(v) + cimag(v) * cimag(v);
	}
	out[0] = sqrt(s);
	return 0;
}
int ctbd_lincomb(int dtype, int64_t n, const void* V, int64_t ldv, int m, const double* coef, void* out)
{
	g_launches++;
	const int cplx = (dtype == CTBD_C128);
	for (int64_t i = 0; i < n; i++) {
		double complex s = 0;
		for (int j = 0; j < m; j++) { s += coef[j] * GETC(V, (int64_t)j * ldv + i); }
		PUTC(out, i, s);
	}
	return 0;
}
int ctbd_zscale_host(int dtype, int64_t n, void* x, double re, double im)
{
	g_launches++;
	const int cplx = (dtype == CTBD_C128);
	const double complex f = cplx ? re + im * I : re;
	for (int64_t i = 0; i < n; i++) { PUTC(x, i, f * GETC(x, i)); }
	return 0;
}

int ctbd_scale_host(int dtype, int64_t n, void* x, double alpha)
{
	g_launches++;
	const int cplx = (dtype == CTBD_C128);
	for (int64_t i = 0; i < n; i++) { PUTC(x, i, alpha * GETC(x, i)); }
	return 0;
}

/* ---- batched SVD: one-sided Jacobi on the rows of the wide orientation ---- */
static int jacobi_rows(int R, int C, double complex* G /* R x (C+R): [G | W] */)
{
	const int ld = C + R;
	const double tol = 1e-15;
	for (int sweep = 0; sweep < 60; sweep++)
	{
		int rotated = 0;
		for (int p = 0; p < R - 1; p++) {
			for (int q = p + 1; q < R; q++)
			{
				double alpha = 0, beta = 0; double complex gamma = 0;
				for (int k = 0; k < C; k++) {
					const double complex x = G[p * ld + k], y = G[q * ld + k];
					alpha += creal(x) * creal(x) + cimag(x) * cimag(x);
					beta  += creal(y) * creal(y) + cimag(y) * cimag(y);
					gamma += x * conj(y);
				}
				const double ag = cabs(gamma);
				if (ag == 0 || ag <= tol * sqrt(alpha * beta)) { continue; }
				rotated++;
				const double complex ph = gamma / ag;          /* e^{i phi} */
				const double zeta = (beta - alpha) / (2 * ag);
				const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1 + zeta * zeta));
				const double c = 1 / sqrt(1 + t * t), s = c * t;
				for (int k = 0; k < ld; k++) {
					const double complex x = G[p * ld + k], y = ph * G[q * ld + k];
					G[p * ld + k] = c * x - s * y;
					G[q * ld + k] = s * x + c * y;
				}
			}
		}
		if (rotated == 0) { return 0; }
	}
	return 0;   /* best effort, as LAPACK's jacobi drivers */
}

/* work matrix [G | W] of one block in complex arithmetic (wide: G = A, tall: G = A^H; W = identity) */
static double complex* svd_work_matrix(int cplx, const struct ctbd_mat_desc* d, const void* A)
{
	const int m = d->m, n = d->n;
	const int wide = (m <= n);
	const int R = wide ? m : n, C = wide ? n : m;
	const int ld = C + R;
	double complex* G = calloc((size_t)R * ld, sizeof(double complex));
	for (int i = 0; i < R; i++) {
		for (int k = 0; k < C; k++) {
			G[i * ld + k] = wide ? GETC(A, d->a_off + (int64_t)i * n + k) : conj(GETC(A, d->a_off + (int64_t)k * n + i));
		}
		G[i * ld + C + i] = 1;
	}
	return G;
}

/* singular values = row norms of G (descending), vectors from the rows of [G | W] */
static void svd_write_out(int cplx, const struct ctbd_mat_desc* d, const double complex* G, double unscale, void* U, void* Vh, double* S)
{
	const int m = d->m, n = d->n;
	const int wide = (m <= n);
	const int R = wide ? m : n, C = wide ? n : m;
	const int ld = C + R;
	double* sig = malloc((size_t)R * sizeof(double)); double* wn = malloc((size_t)R * sizeof(double)); int* ord = malloc((size_t)R * sizeof(int));
	for (int i = 0; i < R; i++) {
		double s = 0, w = 0;
		for (int k = 0; k < C; k++) { const double complex x = G[i * ld + k]; s += creal(x) * creal(x) + cimag(x) * cimag(x); }
		for (int k = 0; k < R; k++) { const double complex x = G[i * ld + C + k]; w += creal(x) * creal(x) + cimag(x) * cimag(x); }
		sig[i] = sqrt(s); wn[i] = (w > 0) ? 1.0 / sqrt(w) : 1.0; ord[i] = i;
	}
	for (int i = 0; i < R; i++) { for (int j = i + 1; j < R; j++) { if (sig[ord[j]] > sig[ord[i]]) { int t = ord[i]; ord[i] = ord[j]; ord[j] = t; } } }
	for (int r = 0; r < R; r++)
	{
		const int i = ord[r];
		S[d->s_off + r] = unscale * sig[i];
		const double inv = sig[i] > 0 ? 1.0 / sig[i] : 0.0;
		if (wide) {
			/* A = W^H G: Vh[r,:] = G[i,:]/sigma, U[:,r] = conj(W[i,:]) */
			for (int k = 0; k < n; k++) { PUTC(Vh, d->o1_off + (int64_t)r * n + k, inv * G[i * ld + k]); }
			for (int k = 0; k < m; k++) { PUTC(U, d->o0_off + (int64_t)k * R + r, wn[i] * conj(G[i * ld + C + k])); }
		}
		else {
			/* A^H = W^H G  =>  A = G^H W: U[:,r] = conj(G[i,:])/sigma, Vh[r,:] = W[i,:] */
			for (int k = 0; k < m; k++) { PUTC(U, d->o0_off + (int64_t)k * R + r, inv * conj(G[i * ld + k])); }
			for (int k = 0; k < n; k++) { PUTC(Vh, d->o1_off + (int64_t)r * n + k, wn[i] * G[i * ld + C + k]); }
		}
	}
	free(sig); free(wn); free(ord);
}

int ctbd_svd_batched(int dtype, int nmat, const struct ctbd_mat_desc* d, const void* A, void* U, void* Vh, double* S)
{
	g_launches++;
	const int cplx = (dtype == CTBD_C128);
	for (int b = 0; b < nmat; b++)
	{
		const int R = d[b].m <= d[b].n ? d[b].m : d[b].n, C = d[b].m <= d[b].n ? d[b].n : d[b].m;
		double complex* G = svd_work_matrix(cplx, &d[b], A);
		jacobi_rows(R, C, G);
		svd_write_out(cplx, &d[b], G, 1.0, U, Vh, S);
		free(G);
	}
	return 0;
}

/* ---- the big-block SVD in pieces (see ctb_device.h): work matrices in the element type of the tensors, packed in one buffer ---- */
struct emu_svdws { int dtype, nmat; struct ctbd_mat_desc* d; int64_t* g_off; int64_t g_total; void* G; };

int ctbd_svdws_create(int dtype, int nmat, const struct ctbd_mat_desc* descs, const void* A, void** ws, void** Gout, int64_t* g_total)
{
	g_launches++;
	const int cplx = (dtype == CTBD_C128);
	struct emu_svdws* w = calloc(1, sizeof(*w));
	w->dtype = dtype; w->nmat = nmat;
	w->d = dup_mem(descs, (size_t)nmat * sizeof(*descs));
	w->g_off = calloc((size_t)nmat, sizeof(int64_t));
	for (int b = 0; b < nmat; b++) {
		const int64_t R = descs[b].m <= descs[b].n ? descs[b].m : descs[b].n, C = descs[b].m <= descs[b].n ? descs[b].n : descs[b].m;
		w->g_off[b] = w->g_total; w->g_total += R * (C + R);
	}
	w->G = calloc((size_t)w->g_total, cplx ? 16 : 8);
	for (int b = 0; b < nmat; b++) {
		const int64_t R = descs[b].m <= descs[b].n ? descs[b].m : descs[b].n, C = descs[b].m <= descs[b].n ? descs[b].n : descs[b].m;
		double complex* G = svd_work_matrix(cplx, &descs[b], A);
		for (int64_t e = 0; e < R * (C + R); e++) { PUTC(w->G, w->g_off[b] + e, G[e]); }
		free(G);
	}
	*ws = w; *Gout = w->G; *g_total = w->g_total;
	return 0;
}

int ctbd_svdws_finish(void* ws, const void* G_cur, int polish, void* U, void* Vh, double* S)
{
	struct emu_svdws* w = ws;
	if (w == NULL) { return 0; }
	g_launches++;
	const int cplx = (w->dtype == CTBD_C128);
	const void* src = (G_cur != NULL) ? G_cur : w->G;
	for (int b = 0; b < w->nmat; b++) {
		const int R = w->d[b].m <= w->d[b].n ? w->d[b].m : w->d[b].n, C = w->d[b].m <= w->d[b].n ? w->d[b].n : w->d[b].m;
		double complex* G = malloc((size_t)R * (C + R) * sizeof(double complex));
		for (int64_t e = 0; e < (int64_t)R * (C + R); e++) { G[e] = GETC(src, w->g_off[b] + e); }
		if (polish) { jacobi_rows(R, C, G); }
		svd_write_out(cplx, &w->d[b], G, 1.0, U, Vh, S);
		free(G);
	}
	free(w->d); free(w->g_off); free(w->G); free(w);
	return 0;
}

int ctbd_gram_offdiag(int dtype, int ngram, const int64_t* off, const int32_t* dim, const void* G, double floor_rel, double* out)
{
	g_launches++;
	const int cplx = (dtype == CTBD_C128);
	for (int k = 0; k < ngram; k++) {
		for (int i = 0; i < dim[k]; i++) { const double v = cabs(GETC(G, off[k] + (int64_t)i * dim[k] + i)); if (v > out[1]) { out[1] = v; } }
	}
	const double fl2 = (floor_rel * out[1]) * (floor_rel * out[1]);
	for (int k = 0; k < ngram; k++) {
		const int n = dim[k];
		for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) {
			if (i == j) { continue; }
			const double dii = cabs(GETC(G, off[k] + (int64_t)i * n + i)), djj = cabs(GETC(G, off[k] + (int64_t)j * n + j));
			const double den = dii * djj > fl2 ? dii * djj : fl2;
			const double complex g = GETC(G, off[k] + (int64_t)i * n + j);
			const double v = creal(g) * creal(g) + cimag(g) * cimag(g);
			if (den > 0 && v / den > out[0]) { out[0] = v / den; }
		} }
	}
	return 0;
}

/* ---- batched QR / RQ: Householder on a work matrix X (rows x cols), with the RQ index transform ---- */
static void householder_qr(int rows, int cols, double complex* X, double complex* Q /* rows x k */, double complex* Rm /* k x cols */)
{
	const int k = rows < cols ? rows : cols;
	double complex* tau = calloc((size_t)k, sizeof(double complex));
	for (int j = 0; j < k; j++)
	{
		double nrm = 0;
		for (int i = j; i < rows; i++) { const double complex x = X[i * cols + j]; nrm += creal(x) * creal(x) + cimag(x) * cimag(x); }
		nrm = sqrt(nrm);
		const double complex x0 = X[j * cols + j];
		double xn = 0;
		for (int i = j + 1; i < rows; i++) { const double complex x = X[i * cols + j]; xn += creal(x) * creal(x) + cimag(x) * cimag(x); }
		if (xn == 0 && cimag(x0) == 0) { tau[j] = 0; continue; }
		const double beta = -(creal(x0) >= 0 ? 1.0 : -1.0) * nrm;
		tau[j] = (beta - x0) / beta;
		const double complex scal = 1.0 / (x0 - beta);
		for (int i = j + 1; i < rows; i++) { X[i * cols + j] *= scal; }
		X[j * cols + j] = beta;
		/* apply H^H = I - conj(tau) v v^H to the trailing columns */
		for (int c = j + 1; c < cols; c++)
		{
			double complex s = X[j * cols + c];
			for (int i = j + 1; i < rows; i++) { s += conj(X[i * cols + j]) * X[i * cols + c]; }
			s *= conj(tau[j]);
			X[j * cols + c] -= s;
			for (int i = j + 1; i < rows; i++) { X[i * cols + c] -= X[i * cols + j] * s; }
		}
	}
	for (int i = 0; i < k; i++) { for (int c = 0; c < cols; c++) { Rm[i * cols + c] = (c >= i) ? X[i * cols + c] : 0; } }
	/* Q = H_0 H_1 ... H_{k-1} [I; 0] */
	for (int i = 0; i < rows; i++) { for (int c = 0; c < k; c++) { Q[i * k + c] = (i == c) ? 1 : 0; } }
	for (int j = k - 1; j >= 0; j--)
	{
		for (int c = j; c < k; c++)
		{
			double complex s = Q[j * k + c];
			for (int i = j + 1; i < rows; i++) { s += conj(X[i * cols + j]) * Q[i * k + c]; }
			s *= tau[j];
			Q[j * k + c] -= s;
			for (int i = j + 1; i < rows; i++) { Q[i * k + c] -= X[i * cols + j] * s; }
		}
	}
	free(tau);
}

int ctbd_qr_batched(int dtype, int rq, int nmat, const struct ctbd_mat_desc* d, const void* A, void* O0, void* O1)
{
	g_launches++;
	const int cplx = (dtype == CTBD_C128);
	for (int b = 0; b < nmat; b++)
	{
		const int m = d[b].m, n = d[b].n;
		const int k = m < n ? m : n;
		const int rows = rq ? n : m, cols = rq ? m : n;
		double complex* X = malloc((size_t)rows * cols * sizeof(double complex));
		double complex* Q = malloc((size_t)rows * k * sizeof(double complex));
		double complex* Rm = malloc((size_t)k * cols * sizeof(double complex));
		for (int i = 0; i < rows; i++) { for (int j = 0; j < cols; j++) {
			/* RQ: X = (E A E)^H, i.e. X[i][j] = conj(A[m-1-j][n-1-i]) */
			X[i * cols + j] = rq ? conj(GETC(A, d[b].a_off + (int64_t)(m - 1 - j) * n + (n - 1 - i))) : GETC(A, d[b].a_off + (int64_t)i * n + j);
		} }
		householder_qr(rows, cols, X, Q, Rm);
		if (!rq) {
			for (int i = 0; i < m; i++) { for (int c = 0; c < k; c++) { PUTC(O0, d[b].o0_off + (int64_t)i * k + c, Q[i * k + c]); } }
			for (int i = 0; i < k; i++) { for (int c = 0; c < n; c++) { PUTC(O1, d[b].o1_off + (int64_t)i * n + c, Rm[i * n + c]); } }
		}
		else {
			/* R[i][j] = conj(Rt[k-1-j][m-1-i]) (m x k), Q[i][j] = conj(Qt[n-1-j][k-1-i]) (k x n) */
			for (int i = 0; i < m; i++) { for (int j = 0; j < k; j++) { PUTC(O0, d[b].o0_off + (int64_t)i * k + j, conj(Rm[(k - 1 - j) * cols + (m - 1 - i)])); } }
			for (int i = 0; i < k; i++) { for (int j = 0; j < n; j++) { PUTC(O1, d[b].o1_off + (int64_t)i * n + j, conj(Q[(n - 1 - j) * k + (k - 1 - i)])); } }
		}
		free(X); free(Q); free(Rm);
	}
	return 0;
}

/* ---- singular-value selection (see ctb_device.h): plain loops with the reference's rule, equal values ordered by index ---- */
struct emu_sv { double v; int64_t i; };
static int emu_cmp_sv(const void* a, const void* b)
{
	const struct emu_sv* x = a; const struct emu_sv* y = b;
	if (x->v < y->v) { return -1; }
	if (x->v > y->v) { return 1; }
	return (x->i > y->i) - (x->i < y->i);
}
int ctbd_truncate_select(int64_t n, const double* S, double tol, int relative, int64_t max_vdim, int renormalize,
	int64_t* nret, int64_t* ind_host, double* info3, double* s_ret)
{
	*nret = 0; info3[0] = 0; info3[1] = 0; info3[2] = tol;
	if (n <= 0) { return 0; }
	struct emu_sv* srt = malloc((size_t)n * sizeof(*srt));
	for (int64_t i = 0; i < n; i++) { srt[i].v = S[i]; srt[i].i = i; }
	qsort(srt, (size_t)n, sizeof(*srt), emu_cmp_sv);
	double sqsum = 0;
	for (int64_t i = 0; i < n; i++) { srt[i].v = srt[i].v * srt[i].v; sqsum += srt[i].v; }
	if (sqsum == 0) { free(srt); s_ret[0] = 0; return 0; }
	if (relative) { for (int64_t i = 0; i < n; i++) { srt[i].v /= sqsum; } }
	for (int64_t i = 1; i < n; i++) { srt[i].v += srt[i - 1].v; }
	if (max_vdim < n) {
		info3[2] = fmax(tol, srt[n - max_vdim - 1].v);
		for (int64_t i = 0; i < n - max_vdim; i++) { srt[i].v = 0; }
	}
	double* accum = malloc((size_t)n * sizeof(double));
	for (int64_t i = 0; i < n; i++) { accum[srt[i].i] = srt[i].v; }
	free(srt);
	int64_t num = 0;
	double scale = 0, ssq = 1, nrm_all = 0;
	for (int64_t i = 0; i < n; i++) {
		nrm_all += S[i] * S[i];
		if (!(accum[i] > tol)) { continue; }
		ind_host[num++] = i;
		const double a = fabs(S[i]);
		if (a > 0) {
			if (scale < a) { ssq = 1 + ssq * (scale / a) * (scale / a); scale = a; }
			else { ssq += (a / scale) * (a / scale); }
		}
	}
	free(accum);
	if (num == 0) { s_ret[0] = 0; return 0; }
	const double norm_sigma = scale * sqrt(ssq);
	const double resc = renormalize ? sqrt(nrm_all) / norm_sigma : 1.0;
	double entropy = 0;
	for (int64_t k = 0; k < num; k++) {
		const double p = S[ind_host[k]] / norm_sigma;
		if (p > 0) { const double sq = p * p; entropy -= sq * log(sq); }
		s_ret[k] = renormalize ? S[ind_host[k]] * resc : S[ind_host[k]];
	}
	*nret = num; info3[0] = norm_sigma; info3[1] = entropy;
	return 0;
}

/* ---- batched strided block linear combinations (twin of csrc/ctbd_blocklc.cu) ---- */
struct emu_lc_plan { int dtype, conj, nblk, nterm; struct ctbd_lc_block* blocks; struct ctbd_lc_term* terms; };
int ctbd_lc_plan_create(int dtype, int conj, int nblk, const struct ctbd_lc_block* blocks, int nterm, const struct ctbd_lc_term* terms, void** plan)
{
	if (dtype != CTBD_F64 && dtype != CTBD_C128) { snprintf(g_err, sizeof g_err, "emu: lc plan dtype"); return -1; }
	struct emu_lc_plan* p = calloc(1, sizeof *p);
	p->dtype = dtype; p->conj = conj; p->nblk = nblk; p->nterm = nterm;
	p->blocks = dup_mem(blocks, (size_t)nblk * sizeof *blocks);
	p->terms = dup_mem(terms, (size_t)nterm * sizeof *terms);
	*plan = p;
	return 0;
}
int ctbd_lc_plan_run(void* plan, const void* src, void* dst)
{
	const struct emu_lc_plan* p = plan;
	const int cplx = (p->dtype == CTBD_C128);
	g_launches++;
	for (int ib = 0; ib < p->nblk; ib++)
	{
		const struct ctbd_lc_block* b = &p->blocks[ib];
		int64_t n = 1;
		for (int a = 0; a < b->ndim; a++) { n *= b->dim[a]; }
		for (int64_t e = 0; e < n; e++)
		{
			int64_t r = e, doff = b->dst_off, soff = 0;
			for (int a = b->ndim - 1; a >= 0; a--) {
				const int64_t i = r % b->dim[a]; r /= b->dim[a];
				doff += i * b->dstride[a]; soff += i * b->sstride[a];
			}
			if (cplx) {
				double complex acc = 0;
				for (int t = b->term_begin; t < b->term_end; t++) { acc += p->terms[t].coef * ((const double complex*)src)[p->terms[t].src_off + soff]; }
				((double complex*)dst)[doff] = p->conj ? conj(acc) : acc;
			}
			else {
				double acc = 0;
				for (int t = b->term_begin; t < b->term_end; t++) { acc += p->terms[t].coef * ((const double*)src)[p->terms[t].src_off + soff]; }
				((double*)dst)[doff] = acc;
			}
		}
	}
	return 0;
}
int ctbd_lc_plan_destroy(void* plan) { struct emu_lc_plan* p = plan; if (p) { free(p->blocks); free(p->terms); free(p); } return 0; }

/* CUDA graphs: not available on the test double (the caller replays launch by launch) */
int ctbd_graph_capture_begin(void) { return 1; }
int ctbd_graph_capture_end(void** graph) { *graph = NULL; return 1; }
int ctbd_graph_launch(void* graph) { (void)graph; snprintf(g_err, sizeof g_err, "emu: no graphs"); return -1; }
int ctbd_graph_destroy(void* graph) { (void)graph; return 0; }
