/*
 * tensor_meta.c -- sector structure of block-sparse tensors and the host <-> device boundary.
 *
 * Structural rules follow the reference bit-exactly:
 *   - sectors of an axis = distinct logical quantum numbers, ascending
 *     (reference allocate_block_sparse_tensor, src/tensor/block_sparse_tensor.c:79-117)
 *   - position inside a sector = order of appearance along the logical axis (:536-550, :922-939)
 *   - a grid cell holds a block iff sum_i dir_i * q_i == 0 (:126-133)
 *   - packed order = stored blocks in row-major grid order, each row-major (:3131-3150)
 */
#include <time.h>
#include "ctb_internal.h"

double ctb_wall_ms(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}

static int cmp_qnumber(const void* a, const void* b)
{
	const qnumber x = *(const qnumber*)a, y = *(const qnumber*)b;
	return (x > y) - (x < y);
}

void ctb_axis_init(struct ctb_axis* ax, ct_long dim, int dir, const qnumber* qlog)
{
	CTB_REQUIRE(dim > 0);
	ax->dim = dim;
	ax->dir = dir;
	ax->qlog = ctb_malloc(dim * sizeof(qnumber));
	memcpy(ax->qlog, qlog, dim * sizeof(qnumber));

	/* distinct quantum numbers, ascending */
	qnumber* sorted = ctb_malloc(dim * sizeof(qnumber));
	memcpy(sorted, qlog, dim * sizeof(qnumber));
	qsort(sorted, dim, sizeof(qnumber), cmp_qnumber);
	int nsec = 0;
	for (ct_long j = 0; j < dim; j++) {
		if (j == 0 || sorted[j] != sorted[j - 1]) { sorted[nsec++] = sorted[j]; }
	}
	ax->nsec = nsec;
	ax->qsec = ctb_malloc(nsec * sizeof(qnumber));
	memcpy(ax->qsec, sorted, nsec * sizeof(qnumber));
	ctb_free(sorted);

	ax->secdim   = ctb_calloc(nsec, sizeof(int32_t));
	ax->sec_of   = ctb_malloc(dim * sizeof(int32_t));
	ax->pos_of   = ctb_malloc(dim * sizeof(int32_t));
	ax->secstart = ctb_calloc(nsec + 1, sizeof(int32_t));
	ax->log_of   = ctb_malloc(dim * sizeof(int32_t));
	for (ct_long j = 0; j < dim; j++)
	{
		const int s = ctb_axis_find_sector(ax, qlog[j]);
		ax->sec_of[j] = s;
		ax->pos_of[j] = ax->secdim[s]++;
	}
	for (int s = 0; s < nsec; s++) {
		ax->secstart[s + 1] = ax->secstart[s] + ax->secdim[s];
	}
	for (ct_long j = 0; j < dim; j++) {
		ax->log_of[ax->secstart[ax->sec_of[j]] + ax->pos_of[j]] = (int32_t)j;
	}
}

void ctb_axis_copy(struct ctb_axis* dst, const struct ctb_axis* src)
{
	const ct_long dim = src->dim;
	const int nsec = src->nsec;
	*dst = *src;
	dst->qlog     = ctb_malloc(dim * sizeof(qnumber));        memcpy(dst->qlog,     src->qlog,     dim * sizeof(qnumber));
	dst->qsec     = ctb_malloc(nsec * sizeof(qnumber));       memcpy(dst->qsec,     src->qsec,     nsec * sizeof(qnumber));
	dst->secdim   = ctb_malloc(nsec * sizeof(int32_t));       memcpy(dst->secdim,   src->secdim,   nsec * sizeof(int32_t));
	dst->sec_of   = ctb_malloc(dim * sizeof(int32_t));        memcpy(dst->sec_of,   src->sec_of,   dim * sizeof(int32_t));
	dst->pos_of   = ctb_malloc(dim * sizeof(int32_t));        memcpy(dst->pos_of,   src->pos_of,   dim * sizeof(int32_t));
	dst->secstart = ctb_malloc((nsec + 1) * sizeof(int32_t)); memcpy(dst->secstart, src->secstart, (nsec + 1) * sizeof(int32_t));
	dst->log_of   = ctb_malloc(dim * sizeof(int32_t));        memcpy(dst->log_of,   src->log_of,   dim * sizeof(int32_t));
}

void ctb_axis_free(struct ctb_axis* ax)
{
	ctb_free(ax->qlog); ctb_free(ax->qsec); ctb_free(ax->secdim); ctb_free(ax->sec_of);
	ctb_free(ax->pos_of); ctb_free(ax->secstart); ctb_free(ax->log_of);
	memset(ax, 0, sizeof(*ax));
}

static int cmp_ct_long(const void* a, const void* b) { const ct_long x = *(const ct_long*)a, y = *(const ct_long*)b; return (x > y) - (x < y); }

/* binary search in the sorted sector list; -1 if absent */
int ctb_axis_find_sector(const struct ctb_axis* ax, qnumber q)
{
	int lo = 0, hi = ax->nsec - 1;
	while (lo <= hi)
	{
		const int mid = (lo + hi) / 2;
		if (ax->qsec[mid] == q) { return mid; }
		if (ax->qsec[mid] < q) { lo = mid + 1; } else { hi = mid - 1; }
	}
	return -1;
}

bool ctb_axis_same_qnums(const struct ctb_axis* a, const struct ctb_axis* b)
{
	return a->dim == b->dim && memcmp(a->qlog, b->qlog, a->dim * sizeof(qnumber)) == 0;
}

void ctb_grid_unravel(const struct ctb_tensor* t, ct_long cell, int* idx)
{
	for (int i = t->ndim - 1; i >= 0; i--) {
		idx[i] = (int)(cell % t->ax[i].nsec);
		cell /= t->ax[i].nsec;
	}
}

ct_long ctb_grid_ravel(const struct ctb_tensor* t, const int* idx)
{
	ct_long cell = 0;
	for (int i = 0; i < t->ndim; i++) {
		cell = cell * t->ax[i].nsec + idx[i];
	}
	return cell;
}

struct ctb_tensor* ctb_tensor_from_axes(int dtype, int ndim, struct ctb_axis* axes, int alloc)
{
	CTB_REQUIRE(ndim >= 0 && ndim <= CTB_MAXDIM);
	CTB_REQUIRE(dtype == CT_DOUBLE_REAL || dtype == CT_DOUBLE_COMPLEX);
	struct ctb_tensor* t = ctb_calloc(1, sizeof(struct ctb_tensor));
	t->dtype = dtype;
	t->ndim = ndim;
	for (int i = 0; i < ndim; i++) {
		t->ax[i] = axes[i];   /* move */
		memset(&axes[i], 0, sizeof(struct ctb_axis));
	}
	t->ngrid = 1;
	for (int i = 0; i < ndim; i++) { t->ngrid *= t->ax[i].nsec; }
	/* stored blocks = the charge-conserving cells of the sector grid (reference block_sparse_tensor.c:126-133), in row-major grid order.
	 * The axis with the MOST sectors is not enumerated: its sector follows from sum dir q = 0 (binary search in its sorted sector
	 * list), so the cost is ngrid / max_i nsec_i lookups instead of a scan of the whole grid -- the 5- and 6-leg intermediates of a
	 * bond have 10^5 - 10^7 cells (with a dummy last leg the former "solve the last axis" rule still walked all of them).  The cells
	 * come out of order when the solved axis is not the last one and are sorted afterwards. */
	int solved = ndim - 1;
	for (int i = 0; i < ndim; i++) { if (t->ax[i].nsec > t->ax[solved].nsec) { solved = i; } }
	ct_long cell_stride[CTB_MAXDIM];
	{
		ct_long st = 1;
		for (int i = ndim - 1; i >= 0; i--) { cell_stride[i] = st; st *= t->ax[i].nsec; }
	}
	const ct_long nlead = (ndim > 0 && t->ax[solved].nsec > 0) ? t->ngrid / t->ax[solved].nsec : (ndim == 0 ? 1 : 0);
	int idx[CTB_MAXDIM] = { 0 };
	size_t cap = 256;
	int nblk = 0;
	t->blk_grid = ctb_malloc(cap * sizeof(ct_long));
	for (ct_long cl = 0; cl < nlead; cl++)
	{
		qnumber qsum = 0;
		ct_long cell = 0;
		for (int i = 0; i < ndim; i++) { if (i != solved) { qsum += t->ax[i].dir * t->ax[i].qsec[idx[i]]; cell += idx[i] * cell_stride[i]; } }
		if (ndim == 0) { cell = 0; }
		else {
			const int sl = ctb_axis_find_sector(&t->ax[solved], -t->ax[solved].dir * qsum);
			cell = (sl >= 0) ? cell + sl * cell_stride[solved] : -1;
		}
		if (cell >= 0) {
			if ((size_t)nblk == cap) { cap *= 2; t->blk_grid = realloc(t->blk_grid, cap * sizeof(ct_long)); }      /* glibc: realloc keeps the 16-byte alignment */
			t->blk_grid[nblk++] = cell;
		}
		/* next combination of the enumerated axes, row-major (the last enumerated axis runs fastest) */
		for (int i = ndim - 1; i >= 0; i--) {
			if (i == solved) { continue; }
			if (++idx[i] < t->ax[i].nsec) { break; }
			idx[i] = 0;
		}
	}
	if (solved != ndim - 1 && nblk > 1) { qsort(t->blk_grid, (size_t)nblk, sizeof(ct_long), cmp_ct_long); }
	/* cell -> offset: a dense table for ordinary grids, a hash index over the stored blocks for large ones */
	const char* dense_env = getenv("CTB_GRID_DENSE_MAX");      /* test knob: force the hash index on small grids */
	const ct_long dense_max = (dense_env != NULL) ? (ct_long)atoll(dense_env) : CTB_GRID_DENSE_MAX;
	t->grid_off = NULL; t->gh_key = NULL; t->gh_val = NULL; t->gh_cap = 0;
	if (t->ngrid <= dense_max) {
		t->grid_off = ctb_malloc((t->ngrid > 0 ? t->ngrid : 1) * sizeof(ct_long));
		memset(t->grid_off, 0xFF, (size_t)t->ngrid * sizeof(ct_long));      /* -1 everywhere */
	}
	else {
		ct_long hc = 16;
		while (hc < 2 * (ct_long)nblk) { hc *= 2; }
		t->gh_cap = hc;
		t->gh_key = ctb_malloc(hc * sizeof(ct_long));
		t->gh_val = ctb_malloc(hc * sizeof(ct_long));
		memset(t->gh_key, 0xFF, (size_t)hc * sizeof(ct_long));
	}
	t->nblk = nblk;
	t->blk_off  = ctb_malloc((nblk + 1) * sizeof(ct_long));
	ct_long off = 0, nelem = 0;
	for (int b = 0; b < nblk; b++)
	{
		ct_long rem = t->blk_grid[b], numel = 1;
		for (int i = ndim - 1; i >= 0; i--) { numel *= t->ax[i].secdim[rem % t->ax[i].nsec]; rem /= t->ax[i].nsec; }
		if (t->grid_off != NULL) { t->grid_off[t->blk_grid[b]] = off; }
		else {
			const ct_long cell = t->blk_grid[b];
			ct_long q = (ct_long)(((uint64_t)cell * 0x9E3779B97F4A7C15ull) >> 20) & (t->gh_cap - 1);
			while (t->gh_key[q] >= 0) { q = (q + 1) & (t->gh_cap - 1); }
			t->gh_key[q] = cell; t->gh_val[q] = off;
		}
		t->blk_off[b] = off;
		nelem += numel;
		off += numel;
		off = (off + CTB_BLOCK_ALIGN - 1) / CTB_BLOCK_ALIGN * CTB_BLOCK_ALIGN;
	}
	t->blk_off[nblk] = off;
	t->nelem = nelem;
	t->nstore = off;
	t->d = NULL;
	t->layout = NULL;
	if (alloc) {
		CTB_CHECK_ABORT(ctbd_malloc(&t->d, (size_t)(t->nstore > 0 ? t->nstore : 1) * ctb_sizeof_dtype(dtype)));
	}
	return t;
}

struct ctb_tensor* ctb_tensor_create(int dtype, int ndim, const ct_long* dim, const int* dirs, const qnumber* const* qnums, int alloc)
{
	struct ctb_axis axes[CTB_MAXDIM];
	for (int i = 0; i < ndim; i++) {
		ctb_axis_init(&axes[i], dim[i], dirs[i], qnums[i]);
	}
	return ctb_tensor_from_axes(dtype, ndim, axes, alloc);
}

struct ctb_tensor* ctb_tensor_like(const struct ctb_tensor* s, int alloc)
{
	struct ctb_axis axes[CTB_MAXDIM];
	for (int i = 0; i < s->ndim; i++) {
		ctb_axis_copy(&axes[i], &s->ax[i]);
	}
	return ctb_tensor_from_axes(s->dtype, s->ndim, axes, alloc);
}

struct ctb_tensor* ctb_tensor_clone(const struct ctb_tensor* s)
{
	struct ctb_tensor* t = ctb_tensor_like(s, 1);
	if (s->nstore > 0) {
		CTB_CHECK_ABORT(ctbd_d2d(t->d, s->d, (size_t)s->nstore * ctb_sizeof_dtype(s->dtype)));
	}
	return t;
}

void ctb_tensor_free(struct ctb_tensor* t)
{
	if (t == NULL) { return; }
	for (int i = 0; i < t->ndim; i++) { ctb_axis_free(&t->ax[i]); }
	ctb_free(t->grid_off); ctb_free(t->gh_key); ctb_free(t->gh_val);
	ctb_free(t->blk_grid);
	ctb_free(t->blk_off);
	if (t->layout != NULL) { ctbd_layout_destroy(t->layout); }
	if (t->d != NULL && !t->borrowed) { ctbd_free(t->d); }
	ctb_free(t);
}

bool ctb_tensor_same_structure(const struct ctb_tensor* a, const struct ctb_tensor* b)
{
	if (a->dtype != b->dtype || a->ndim != b->ndim) { return false; }
	for (int i = 0; i < a->ndim; i++) {
		if (a->ax[i].dir != b->ax[i].dir || !ctb_axis_same_qnums(&a->ax[i], &b->ax[i])) { return false; }
	}
	return true;
}

void* ctb_tensor_layout(struct ctb_tensor* t)
{
	if (t->layout == NULL)
	{
		struct ctbd_layout_host h;
		memset(&h, 0, sizeof(h));
		h.ndim = t->ndim;
		h.dtype = t->dtype;
		for (int i = 0; i < t->ndim; i++)
		{
			h.dim[i]      = t->ax[i].dim;
			h.nsec[i]     = t->ax[i].nsec;
			h.sec_of[i]   = t->ax[i].sec_of;
			h.pos_of[i]   = t->ax[i].pos_of;
			h.secstart[i] = t->ax[i].secstart;
			h.log_of[i]   = t->ax[i].log_of;
		}
		h.ngrid = t->ngrid;
		ct_long* dense_tmp = NULL;
		if (t->grid_off == NULL) {
			/* a re-blocking kernel wants the dense table of a large grid (rare: the large grids belong to plan intermediates) */
			dense_tmp = ctb_malloc((t->ngrid > 0 ? t->ngrid : 1) * sizeof(ct_long));
			memset(dense_tmp, 0xFF, (size_t)t->ngrid * sizeof(ct_long));
			for (int b = 0; b < t->nblk; b++) { dense_tmp[t->blk_grid[b]] = t->blk_off[b]; }
		}
		h.grid_off = (t->grid_off != NULL) ? t->grid_off : dense_tmp;
		h.nblk = t->nblk;
		h.blk_grid = t->blk_grid;
		h.blk_off = t->blk_off;
		h.nstore = t->nstore;
		CTB_CHECK_ABORT(ctbd_layout_create(&h, &t->layout));
		ctb_free(dense_tmp);
	}
	return t->layout;
}

/* ---- host struct <-> device ---- */

/* allocate a reference-ABI host tensor (all payload with 16-byte aligned malloc; released by delete_block_sparse_tensor) */
static void host_allocate_bst(int dtype, int ndim, const ct_long* dim, const enum tensor_axis_direction* axis_dir, const qnumber* const* qnums, struct block_sparse_tensor* t, int zero);

void ctb_host_allocate_bst(int dtype, int ndim, const ct_long* dim, const enum tensor_axis_direction* axis_dir, const qnumber* const* qnums, struct block_sparse_tensor* t)
{
	host_allocate_bst(dtype, ndim, dim, axis_dir, qnums, t, 1);
}

/* zero == 0: block payloads are left uninitialised (the caller overwrites every entry) */
static void host_allocate_bst(int dtype, int ndim, const ct_long* dim, const enum tensor_axis_direction* axis_dir, const qnumber* const* qnums, struct block_sparse_tensor* t, int zero)
{
	t->dtype = (enum numeric_type)dtype;
	t->ndim = ndim;
	const size_t esize = ctb_sizeof_dtype(dtype);
	if (ndim == 0)
	{
		/* reference special case (block_sparse_tensor.c:44-60): a single scalar block, no axis metadata */
		t->dim_logical = NULL; t->dim_blocks = NULL; t->axis_dir = NULL; t->qnums_logical = NULL; t->qnums_blocks = NULL;
		t->blocks = ctb_calloc(1, sizeof(struct dense_tensor*));
		t->blocks[0] = ctb_calloc(1, sizeof(struct dense_tensor));
		t->blocks[0]->dtype = (enum numeric_type)dtype;
		t->blocks[0]->ndim = 0;
		t->blocks[0]->dim = NULL;
		t->blocks[0]->data = ctb_calloc(1, esize);
		return;
	}
	t->dim_logical = ctb_malloc(ndim * sizeof(ct_long));
	t->dim_blocks  = ctb_calloc(ndim, sizeof(ct_long));
	t->axis_dir    = ctb_malloc(ndim * sizeof(enum tensor_axis_direction));
	t->qnums_logical = ctb_calloc(ndim, sizeof(qnumber*));
	t->qnums_blocks  = ctb_calloc(ndim, sizeof(qnumber*));
	struct ctb_axis axes[CTB_MAXDIM];
	ct_long ngrid = 1;
	for (int i = 0; i < ndim; i++)
	{
		ctb_axis_init(&axes[i], dim[i], axis_dir[i], qnums[i]);
		t->dim_logical[i] = dim[i];
		t->axis_dir[i] = axis_dir[i];
		t->qnums_logical[i] = ctb_malloc(dim[i] * sizeof(qnumber));
		memcpy(t->qnums_logical[i], qnums[i], dim[i] * sizeof(qnumber));
		t->dim_blocks[i] = axes[i].nsec;
		t->qnums_blocks[i] = ctb_malloc(axes[i].nsec * sizeof(qnumber));
		memcpy(t->qnums_blocks[i], axes[i].qsec, axes[i].nsec * sizeof(qnumber));
		ngrid *= axes[i].nsec;
	}
	t->blocks = ctb_calloc(ngrid, sizeof(struct dense_tensor*));
	int idx[CTB_MAXDIM] = { 0 };
	for (ct_long c = 0; c < ngrid; c++)
	{
		qnumber qsum = 0;
		for (int i = 0; i < ndim; i++) { qsum += axis_dir[i] * axes[i].qsec[idx[i]]; }
		if (qsum == 0)
		{
			struct dense_tensor* b = ctb_calloc(1, sizeof(struct dense_tensor));
			b->dtype = (enum numeric_type)dtype;
			b->ndim = ndim;
			b->dim = ctb_malloc(ndim * sizeof(ct_long));
			ct_long numel = 1;
			for (int i = 0; i < ndim; i++) { b->dim[i] = axes[i].secdim[idx[i]]; numel *= b->dim[i]; }
			b->data = zero ? ctb_calloc(numel, esize) : ctb_malloc((size_t)numel * esize);
			t->blocks[c] = b;
		}
		for (int i = ndim - 1; i >= 0; i--) {
			if (++idx[i] < axes[i].nsec) { break; }
			idx[i] = 0;
		}
	}
	for (int i = 0; i < ndim; i++) { ctb_axis_free(&axes[i]); }
}

/* device tensor with the structure of the host tensor, buffer allocated, payload not yet copied */
struct ctb_tensor* ctb_upload_begin(const struct block_sparse_tensor* h)
{
	int dirs[CTB_MAXDIM];
	for (int i = 0; i < h->ndim; i++) { dirs[i] = (int)h->axis_dir[i]; }
	return ctb_tensor_create(h->dtype, h->ndim, h->dim_logical, dirs, (const qnumber* const*)h->qnums_logical, 1);
}

/* host block pointers, byte offsets in the packed device buffer and byte lengths of the stored blocks */
static void block_lists(const struct ctb_tensor* t, const struct block_sparse_tensor* h, void*** hptrs, int64_t** offs, int64_t** lens)
{
	const size_t esize = ctb_sizeof_dtype(t->dtype);
	*hptrs = malloc((size_t)t->nblk * sizeof(void*));
	*offs = malloc((size_t)t->nblk * sizeof(int64_t));
	*lens = malloc((size_t)t->nblk * sizeof(int64_t));
	for (int b = 0; b < t->nblk; b++)
	{
		const struct dense_tensor* hb = h->blocks[t->blk_grid[b]];
		CTB_REQUIRE(hb != NULL);
		ct_long numel = 1;
		for (int i = 0; i < hb->ndim; i++) { numel *= hb->dim[i]; }
		(*hptrs)[b] = hb->data; (*offs)[b] = (int64_t)t->blk_off[b] * (int64_t)esize; (*lens)[b] = (int64_t)numel * (int64_t)esize;
	}
}

/* payload copy of ctb_upload: the separately allocated host blocks stream through the pinned staging ring of the device layer.
 * One process per GPU (ctb_dist_world > 1): every rank holds the same host tensor, so each uploads only ITS 1/world byte range of the
 * packed layout over its own PCIe link and the ranges are all-gathered over NVLink -- the host-to-device traffic of a call shrinks
 * with the number of ranks instead of being repeated by every one of them (collective: all ranks upload the same tensors in the
 * same order, which the one-process-per-GPU contract already demands). */
#define CTB_SHARDED_UPLOAD_MIN ((int64_t)8 << 20)
/* > 0 only inside the entry points every rank calls together (apply_local_hamiltonian, dmrg_*): the sharded upload is a collective,
 * and a query a host makes on one rank only (ctb_heff_plan_info, block_sparse_tensor_dot, ...) must never wait for the others */
int ctb_collective_upload = 0;
int ctb_upload_data(struct ctb_tensor* t, const struct block_sparse_tensor* h)
{
	if (t->nstore == 0) { return 0; }
	void** hptrs; int64_t* offs; int64_t* lens;
	block_lists(t, h, &hptrs, &offs, &lens);
	const int64_t total = (int64_t)t->nstore * (int64_t)ctb_sizeof_dtype(t->dtype);
	int rc = 0;
	const char* env_min = getenv("CTB_SHARDED_UPLOAD_MIN");      /* bytes; the tests lower it to exercise the path on small tensors */
	const int64_t min_bytes = (env_min != NULL) ? (int64_t)atoll(env_min) : CTB_SHARDED_UPLOAD_MIN;
	if (ctb_dist_world > 1 && ctb_collective_upload > 0 && total >= min_bytes && getenv("CTB_NO_SHARDED_UPLOAD") == NULL)
	{
		const int W = ctb_dist_world, me = ctb_dist_rank;
		int64_t chunk = (total + W - 1) / W;
		chunk = (chunk + 255) / 256 * 256;
		const int64_t lo = (int64_t)me * chunk, hi = (lo + chunk < total) ? lo + chunk : total;
		void* tmp = NULL;
		rc = ctbd_malloc(&tmp, (size_t)chunk * (size_t)W);      /* zero-filled: the alignment padding between blocks stays zero */
		if (rc == 0)
		{
			/* the pieces of the host blocks that fall into [lo, hi) */
			int nb = 0;
			for (int b = 0; b < t->nblk; b++)
			{
				const int64_t b0 = offs[b], b1 = offs[b] + lens[b];
				const int64_t c0 = b0 > lo ? b0 : lo, c1 = b1 < hi ? b1 : hi;
				if (c1 <= c0) { continue; }
				hptrs[nb] = (char*)hptrs[b] + (c0 - b0); offs[nb] = c0; lens[nb] = c1 - c0; nb++;
			}
			rc = ctbd_h2d_blocks(tmp, nb, (const void* const*)hptrs, offs, lens);
			if (rc == 0) { rc = ctbd_allgather((char*)tmp + lo, tmp, (size_t)chunk); }      /* in place: the own range already sits in its slot */
			if (rc == 0) { rc = ctbd_d2d(t->d, tmp, (size_t)total); }
			ctbd_free(tmp);
		}
	}
	else {
		rc = ctbd_h2d_blocks(t->d, t->nblk, (const void* const*)hptrs, offs, lens);
	}
	free(hptrs); free(offs); free(lens);
	return rc;
}

struct ctb_tensor* ctb_upload(const struct block_sparse_tensor* h)
{
	struct ctb_tensor* t = ctb_upload_begin(h);
	CTB_CHECK_ABORT(ctb_upload_data(t, h));
	return t;
}

/* allocates the host payload of the result (uninitialised); with 'prefault' its pages are faulted in by the copy threads right away */
void ctb_download_begin(const struct ctb_tensor* t, struct block_sparse_tensor* h, int prefault)
{
	ct_long dim[CTB_MAXDIM];
	enum tensor_axis_direction dirs[CTB_MAXDIM];
	const qnumber* qn[CTB_MAXDIM];
	for (int i = 0; i < t->ndim; i++) {
		dim[i] = t->ax[i].dim;
		dirs[i] = (enum tensor_axis_direction)t->ax[i].dir;
		qn[i] = t->ax[i].qlog;
	}
	host_allocate_bst(t->dtype, t->ndim, dim, dirs, qn, h, 0);
	if (prefault && t->nstore > 0) {
		void** hptrs; int64_t* offs; int64_t* lens;
		block_lists(t, h, &hptrs, &offs, &lens);
		CTB_CHECK_ABORT(ctbd_host_prefault(t->nblk, hptrs, lens));
		free(hptrs); free(offs); free(lens);
	}
}

int ctb_download_data(const struct ctb_tensor* t, struct block_sparse_tensor* h)
{
	if (t->nstore == 0) { return 0; }
	void** hptrs; int64_t* offs; int64_t* lens;
	block_lists(t, h, &hptrs, &offs, &lens);
	const int rc = ctbd_d2h_blocks(t->d, t->nblk, hptrs, offs, lens);
	free(hptrs); free(offs); free(lens);
	return rc;
}

int ctb_download(const struct ctb_tensor* t, struct block_sparse_tensor* h)
{
	ctb_download_begin(t, h, 0);
	return ctb_download_data(t, h);
}

int ctb_upload_entries(struct ctb_tensor* t, const void* entries)
{
	const size_t esize = ctb_sizeof_dtype(t->dtype);
#if CTB_BLOCK_ALIGN == 1
	if (t->nelem > 0) { CTB_CHECK(ctbd_h2d(t->d, entries, (size_t)t->nelem * esize)); }
#else
	ct_long pos = 0;
	for (int b = 0; b < t->nblk; b++) {
		ct_long numel = (b + 1 < t->nblk ? t->blk_off[b + 1] : t->nstore) - t->blk_off[b];
		/* with alignment the true block size must be recomputed */
		int idx[CTB_MAXDIM]; ctb_grid_unravel(t, t->blk_grid[b], idx);
		numel = 1; for (int i = 0; i < t->ndim; i++) { numel *= t->ax[i].secdim[idx[i]]; }
		CTB_CHECK(ctbd_h2d((char*)t->d + (size_t)t->blk_off[b] * esize, (const char*)entries + (size_t)pos * esize, (size_t)numel * esize));
		pos += numel;
	}
#endif
	return 0;
}

int ctb_download_entries(const struct ctb_tensor* t, void* entries)
{
	const size_t esize = ctb_sizeof_dtype(t->dtype);
#if CTB_BLOCK_ALIGN == 1
	if (t->nelem > 0) { CTB_CHECK(ctbd_d2h(entries, t->d, (size_t)t->nelem * esize)); }
#else
	ct_long pos = 0;
	for (int b = 0; b < t->nblk; b++) {
		int idx[CTB_MAXDIM]; ctb_grid_unravel(t, t->blk_grid[b], idx);
		ct_long numel = 1; for (int i = 0; i < t->ndim; i++) { numel *= t->ax[i].secdim[idx[i]]; }
		CTB_CHECK(ctbd_d2h((char*)entries + (size_t)pos * esize, (const char*)t->d + (size_t)t->blk_off[b] * esize, (size_t)numel * esize));
		pos += numel;
	}
#endif
	return 0;
}

int ctb_set_entry(struct ctb_tensor* t, ct_long offset, double re, double im)
{
	const size_t esize = ctb_sizeof_dtype(t->dtype);
	double v[2] = { re, im };
	CTB_CHECK(ctbd_h2d((char*)t->d + (size_t)offset * esize, v, esize));
	return 0;
}
