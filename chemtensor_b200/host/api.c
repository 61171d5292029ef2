/*
 * api.c -- the reference-named entry points (include/chemtensor_b200.h) on host structs.
 *
 * Pattern of every compute wrapper: upload the host operands into packed device tensors,
 * run the device primitive(s), download the result into a freshly allocated host struct
 * ("memory will be allocated for r", reference ownership convention).  The multi-step
 * drivers (dmrg_*, compute_right_operator_blocks, mps_orthonormalize_qr) stay on the device
 * between steps.  Functions returning void abort with a message on a device failure
 * (the reference's void functions cannot fail); int functions return <0.
 */
#include <pthread.h>
#include "ctb_internal.h"
#include "chemtensor_b200.h"

static void ensure_init(void) { CTB_CHECK_ABORT(ctbd_init(-1)); }

int ctb_init(int device) { return ctbd_init(device); }

/* ---- one process per GPU ---- */
int ctb_dist_unique_id(void* id_out) { return ctbd_dist_unique_id(id_out); }
int ctb_dist_init(int rank, int world, const void* unique_id)
{
	CTB_CHECK(ctbd_init(-1));
	CTB_CHECK(ctbd_dist_init(rank, world, unique_id));
	ctb_dist_rank = rank; ctb_dist_world = world;
	return 0;
}
int ctb_dist_set_allgather(ctbd_allgather_fn fn, void* ctx) { return ctbd_dist_set_allgather(fn, ctx); }
/* out[0] = rank, out[1] = world, out[2] = exchanges done by the fused peer-store path, out[3] = by all-gather + scatter */
int ctb_dist_info(long long* out) { out[0] = ctb_dist_rank; out[1] = ctb_dist_world; ctb_dist_counters(&out[2], &out[3]); return 0; }
/* exchanges done by the pull path (peer-mapped send buffers read over NVLink) */
long long ctb_dist_pull_exchanges(void) { return ctb_dist_pull_count(); }
long long ctb_dist_push_exchanges(void) { return ctb_dist_push_count(); }
long long ctb_dist_multicast_exchanges(void) { return ctb_dist_multicast_count(); }
int ctb_dist_finalize(void) { ctb_dist_release_buffers(); ctb_dist_rank = 0; ctb_dist_world = 1; return ctbd_dist_finalize(); }
int ctb_backend(void) { return ctbd_backend(); }
long long ctb_launch_count(void) { return ctbd_launch_count(); }

/* ---- host containers ---- */

static ct_long product(const ct_long* x, int n) { ct_long p = 1; for (int i = 0; i < n; i++) { p *= x[i]; } return p; }

void allocate_dense_tensor(const enum numeric_type dtype, const int ndim, const ct_long* dim, struct dense_tensor* t)
{
	t->dtype = dtype;
	t->ndim = ndim;
	if (ndim > 0) {
		t->dim = ctb_malloc((size_t)ndim * sizeof(ct_long));
		memcpy(t->dim, dim, (size_t)ndim * sizeof(ct_long));
	}
	else { t->dim = NULL; }
	t->data = ctb_malloc((size_t)product(dim, ndim) * ctb_sizeof_dtype(dtype));
}

void allocate_zero_dense_tensor(const enum numeric_type dtype, const int ndim, const ct_long* dim, struct dense_tensor* t)
{
	allocate_dense_tensor(dtype, ndim, dim, t);
	memset(t->data, 0, (size_t)product(dim, ndim) * ctb_sizeof_dtype(dtype));
}

void delete_dense_tensor(struct dense_tensor* t)
{
	ctb_free(t->data); t->data = NULL;
	if (t->ndim > 0) { ctb_free(t->dim); }
	t->dim = NULL; t->ndim = 0;
}

void allocate_block_sparse_tensor(const enum numeric_type dtype, const int ndim, const ct_long* dim, const enum tensor_axis_direction* axis_dir, const qnumber** qnums, struct block_sparse_tensor* t)
{
	ctb_host_allocate_bst(dtype, ndim, dim, axis_dir, (const qnumber* const*)qnums, t);
}

void allocate_block_sparse_tensor_like(const struct block_sparse_tensor* s, struct block_sparse_tensor* t)
{
	ctb_host_allocate_bst(s->dtype, s->ndim, s->dim_logical, s->axis_dir, (const qnumber* const*)s->qnums_logical, t);
}

static ct_long host_grid_size(const struct block_sparse_tensor* t) { return product(t->dim_blocks, t->ndim); }

void delete_block_sparse_tensor(struct block_sparse_tensor* t)
{
	const ct_long ngrid = (t->ndim == 0 ? 1 : host_grid_size(t));
	for (ct_long k = 0; k < ngrid; k++) {
		if (t->blocks[k] != NULL) {
			delete_dense_tensor(t->blocks[k]);
			ctb_free(t->blocks[k]);
		}
	}
	ctb_free(t->blocks); t->blocks = NULL;
	if (t->ndim == 0) { return; }
	for (int i = 0; i < t->ndim; i++) { ctb_free(t->qnums_blocks[i]); ctb_free(t->qnums_logical[i]); }
	ctb_free(t->qnums_blocks); ctb_free(t->qnums_logical);
	ctb_free(t->axis_dir); ctb_free(t->dim_blocks); ctb_free(t->dim_logical);
	t->qnums_blocks = NULL; t->qnums_logical = NULL; t->axis_dir = NULL; t->dim_blocks = NULL; t->dim_logical = NULL;
}

static ct_long dense_numel(const struct dense_tensor* b) { return product(b->dim, b->ndim); }

void copy_block_sparse_tensor(const struct block_sparse_tensor* src, struct block_sparse_tensor* dst)
{
	allocate_block_sparse_tensor_like(src, dst);
	const ct_long ngrid = (src->ndim == 0 ? 1 : host_grid_size(src));
	for (ct_long k = 0; k < ngrid; k++) {
		if (src->blocks[k] != NULL) {
			memcpy(dst->blocks[k]->data, src->blocks[k]->data, (size_t)dense_numel(src->blocks[k]) * ctb_sizeof_dtype(src->dtype));
		}
	}
}

static void allocate_chain(const enum numeric_type dtype, const int nsites, const ct_long d, const qnumber* qsite, const ct_long* dim_bonds, const qnumber** qbonds,
	int nphys, struct block_sparse_tensor** a_out, qnumber** qsite_out)
{
	*qsite_out = ctb_malloc((size_t)d * sizeof(qnumber));
	memcpy(*qsite_out, qsite, (size_t)d * sizeof(qnumber));
	*a_out = ctb_calloc((size_t)nsites, sizeof(struct block_sparse_tensor));
	for (int i = 0; i < nsites; i++)
	{
		if (nphys == 1) {
			const ct_long dim[3] = { dim_bonds[i], d, dim_bonds[i + 1] };
			const enum tensor_axis_direction dirs[3] = { TENSOR_AXIS_OUT, TENSOR_AXIS_OUT, TENSOR_AXIS_IN };   /* reference mps.c:41 */
			const qnumber* qn[3] = { qbonds[i], qsite, qbonds[i + 1] };
			allocate_block_sparse_tensor(dtype, 3, dim, dirs, qn, &(*a_out)[i]);
		}
		else {
			const ct_long dim[4] = { dim_bonds[i], d, d, dim_bonds[i + 1] };
			const enum tensor_axis_direction dirs[4] = { TENSOR_AXIS_OUT, TENSOR_AXIS_OUT, TENSOR_AXIS_IN, TENSOR_AXIS_IN };   /* reference mpo.c:48 */
			const qnumber* qn[4] = { qbonds[i], qsite, qsite, qbonds[i + 1] };
			allocate_block_sparse_tensor(dtype, 4, dim, dirs, qn, &(*a_out)[i]);
		}
	}
}

void allocate_mps(const enum numeric_type dtype, const int nsites, const ct_long d, const qnumber* qsite, const ct_long* dim_bonds, const qnumber** qbonds, struct mps* mps)
{
	mps->nsites = nsites; mps->d = d;
	allocate_chain(dtype, nsites, d, qsite, dim_bonds, qbonds, 1, &mps->a, &mps->qsite);
}

void delete_mps(struct mps* mps)
{
	for (int i = 0; i < mps->nsites; i++) { delete_block_sparse_tensor(&mps->a[i]); }
	ctb_free(mps->a); mps->a = NULL;
	ctb_free(mps->qsite); mps->qsite = NULL;
	mps->nsites = 0;
}

void allocate_mpo(const enum numeric_type dtype, const int nsites, const ct_long d, const qnumber* qsite, const ct_long* dim_bonds, const qnumber** qbonds, struct mpo* mpo)
{
	mpo->nsites = nsites; mpo->d = d;
	allocate_chain(dtype, nsites, d, qsite, dim_bonds, qbonds, 2, &mpo->a, &mpo->qsite);
}

void delete_mpo(struct mpo* mpo)
{
	for (int i = 0; i < mpo->nsites; i++) { delete_block_sparse_tensor(&mpo->a[i]); }
	ctb_free(mpo->a); mpo->a = NULL;
	ctb_free(mpo->qsite); mpo->qsite = NULL;
	mpo->nsites = 0;
}

void delete_index_list(struct index_list* list)
{
	ctb_free(list->ind); list->ind = NULL; list->num = 0;
}

/* ---- (de)serialisation: defines the Lanczos vector layout ---- */

ct_long block_sparse_tensor_num_elements_blocks(const struct block_sparse_tensor* t)
{
	const ct_long ngrid = (t->ndim == 0 ? 1 : host_grid_size(t));
	ct_long n = 0;
	for (ct_long k = 0; k < ngrid; k++) {
		if (t->blocks[k] != NULL) { n += dense_numel(t->blocks[k]); }
	}
	return n;
}

void block_sparse_tensor_serialize_entries(const struct block_sparse_tensor* t, void* entries)
{
	const size_t esize = ctb_sizeof_dtype(t->dtype);
	const ct_long ngrid = (t->ndim == 0 ? 1 : host_grid_size(t));
	char* p = entries;
	for (ct_long k = 0; k < ngrid; k++) {
		if (t->blocks[k] != NULL) {
			const size_t nb = (size_t)dense_numel(t->blocks[k]) * esize;
			memcpy(p, t->blocks[k]->data, nb);
			p += nb;
		}
	}
}

void block_sparse_tensor_deserialize_entries(struct block_sparse_tensor* t, const void* entries)
{
	const size_t esize = ctb_sizeof_dtype(t->dtype);
	const ct_long ngrid = (t->ndim == 0 ? 1 : host_grid_size(t));
	const char* p = entries;
	for (ct_long k = 0; k < ngrid; k++) {
		if (t->blocks[k] != NULL) {
			const size_t nb = (size_t)dense_numel(t->blocks[k]) * esize;
			memcpy(t->blocks[k]->data, p, nb);
			p += nb;
		}
	}
}

/* ---- tensor primitives ---- */

static void finish(struct ctb_tensor* dev, struct block_sparse_tensor* out)
{
	CTB_CHECK_ABORT(ctb_download(dev, out));
	ctb_tensor_free(dev);
}

void block_sparse_tensor_transpose(const int* perm, const struct block_sparse_tensor* t, struct block_sparse_tensor* r)
{
	ensure_init();
	struct ctb_tensor* td = ctb_upload(t);
	struct ctb_tensor* rd = ctb_transpose(td, perm, 0);
	ctb_tensor_free(td);
	finish(rd, r);
}

void block_sparse_tensor_conjugate_transpose(const int* perm, const struct block_sparse_tensor* t, struct block_sparse_tensor* r)
{
	ensure_init();
	struct ctb_tensor* td = ctb_upload(t);
	/* transpose followed by entry-wise conjugation, axis directions unchanged (reference block_sparse_tensor.c:880-884) */
	struct ctb_tensor* rd = ctb_transpose(td, perm, ctb_is_complex(t->dtype));
	ctb_tensor_free(td);
	finish(rd, r);
}

void block_sparse_tensor_flatten_axes(const struct block_sparse_tensor* t, const int i_ax, const enum tensor_axis_direction new_axis_dir, struct block_sparse_tensor* r)
{
	ensure_init();
	struct ctb_tensor* td = ctb_upload(t);
	struct ctb_tensor* rd = ctb_flatten_axes(td, i_ax, (int)new_axis_dir);
	ctb_tensor_free(td);
	finish(rd, r);
}

void block_sparse_tensor_split_axis(const struct block_sparse_tensor* t, const int i_ax, const ct_long new_dim_logical[2], const enum tensor_axis_direction new_axis_dir[2], const qnumber* new_qnums_logical[2], struct block_sparse_tensor* r)
{
	ensure_init();
	struct ctb_tensor* td = ctb_upload(t);
	const int dirs[2] = { (int)new_axis_dir[0], (int)new_axis_dir[1] };
	struct ctb_tensor* rd = ctb_split_axis(td, i_ax, new_dim_logical, dirs, (const qnumber* const*)new_qnums_logical);
	ctb_tensor_free(td);
	finish(rd, r);
}

void block_sparse_tensor_slice(const struct block_sparse_tensor* t, const int i_ax, const ct_long* ind, const ct_long nind, struct block_sparse_tensor* r)
{
	ensure_init();
	struct ctb_tensor* td = ctb_upload(t);
	struct ctb_tensor* rd = ctb_slice(td, i_ax, ind, nind);
	ctb_tensor_free(td);
	finish(rd, r);
}

void block_sparse_tensor_cyclic_partial_trace(const struct block_sparse_tensor* t, const int ndim_trace, struct block_sparse_tensor* r)
{
	ensure_init();
	struct ctb_tensor* td = ctb_upload(t);
	struct ctb_tensor* rd = ctb_cyclic_partial_trace(td, ndim_trace);
	ctb_tensor_free(td);
	finish(rd, r);
}

void block_sparse_tensor_multiply_pointwise_vector(const struct block_sparse_tensor* s, const struct dense_tensor* t, const enum tensor_axis_range axrange, struct block_sparse_tensor* r)
{
	ensure_init();
	CTB_REQUIRE(t->ndim == 1 && t->dtype == CT_DOUBLE_REAL);
	const int i_ax = (axrange == TENSOR_AXIS_RANGE_LEADING ? 0 : s->ndim - 1);
	CTB_REQUIRE(t->dim[0] == s->dim_logical[i_ax]);
	struct ctb_tensor* sd = ctb_upload(s);
	double* vec = NULL;
	CTB_CHECK_ABORT(ctbd_malloc((void**)&vec, (size_t)t->dim[0] * sizeof(double)));
	CTB_CHECK_ABORT(ctbd_h2d(vec, t->data, (size_t)t->dim[0] * sizeof(double)));
	struct ctb_tensor* rd = ctb_scale_axis(sd, i_ax, vec);
	CTB_CHECK_ABORT(ctbd_free(vec));
	ctb_tensor_free(sd);
	finish(rd, r);
}

void block_sparse_tensor_dot(const struct block_sparse_tensor* s, const enum tensor_axis_range axrange_s, const struct block_sparse_tensor* t, const enum tensor_axis_range axrange_t, const int ndim_mult, struct block_sparse_tensor* r)
{
	ensure_init();
	struct ctb_tensor* sd = ctb_upload(s);
	struct ctb_tensor* td = ctb_upload(t);
	struct ctb_tensor* rd = ctb_dot(sd, (int)axrange_s, 0, td, (int)axrange_t, 0, ndim_mult, NULL);
	ctb_tensor_free(sd);
	ctb_tensor_free(td);
	finish(rd, r);
}

int block_sparse_tensor_qr(const struct block_sparse_tensor* a, const enum qr_mode mode, struct block_sparse_tensor* q, struct block_sparse_tensor* r)
{
	if (mode != QR_REDUCED) { fprintf(stderr, "chemtensor_b200: block_sparse_tensor_qr supports QR_REDUCED only\n"); return -1; }
	CTB_CHECK(ctbd_init(-1));
	struct ctb_tensor* ad = ctb_upload(a);
	struct ctb_tensor *qd = NULL, *rd = NULL;
	int rc = ctb_qr(ad, &qd, &rd);
	ctb_tensor_free(ad);
	if (rc < 0) { return rc; }
	finish(qd, q);
	finish(rd, r);
	return 0;
}

int block_sparse_tensor_rq(const struct block_sparse_tensor* a, const enum qr_mode mode, struct block_sparse_tensor* r, struct block_sparse_tensor* q)
{
	if (mode != QR_REDUCED) { fprintf(stderr, "chemtensor_b200: block_sparse_tensor_rq supports QR_REDUCED only\n"); return -1; }
	CTB_CHECK(ctbd_init(-1));
	struct ctb_tensor* ad = ctb_upload(a);
	struct ctb_tensor *qd = NULL, *rd = NULL;
	int rc = ctb_rq(ad, &rd, &qd);
	ctb_tensor_free(ad);
	if (rc < 0) { return rc; }
	finish(rd, r);
	finish(qd, q);
	return 0;
}

int block_sparse_tensor_svd(const struct block_sparse_tensor* a, struct block_sparse_tensor* u, struct dense_tensor* s, struct block_sparse_tensor* vh)
{
	CTB_CHECK(ctbd_init(-1));
	struct ctb_tensor* ad = ctb_upload(a);
	struct ctb_tensor *ud = NULL, *vd = NULL;
	double* s_dev = NULL;
	ct_long ns = 0;
	int rc = ctb_svd(ad, &ud, &s_dev, &ns, &vd);
	ctb_tensor_free(ad);
	if (rc < 0) { return rc; }
	const ct_long sdim[1] = { ns };
	allocate_zero_dense_tensor(CT_DOUBLE_REAL, 1, sdim, s);
	CTB_CHECK(ctbd_d2h(s->data, s_dev, (size_t)ns * sizeof(double)));
	CTB_CHECK(ctbd_free(s_dev));
	finish(ud, u);
	finish(vd, vh);
	return 0;
}

/* ---- truncation / bond operations ---- */

double von_neumann_entropy(const double* sigma, const ct_long n) { return ctb_von_neumann_entropy(sigma, n); }

/* Extension: the same selection through the DEVICE kernels the split uses (ctbd_truncate_select): sigma goes up, the index list and the
 * three scalars come back.  Exists so that the device rule can be tested against the reference's retained_bond_indices directly. */
int ctb_retained_bond_indices_device(const double* sigma, const ct_long n, const double tol, const bool relative_thresh, const ct_long max_vdim, struct index_list* list, struct trunc_info* info)
{
	CTB_CHECK(ctbd_init(-1));
	list->ind = NULL; list->num = 0;
	info->norm_sigma = 0; info->entropy = 0; info->tol_eff = tol;
	if (n <= 0) { return 0; }
	double *d_s = NULL, *d_ret = NULL;
	CTB_CHECK(ctbd_malloc((void**)&d_s, (size_t)n * sizeof(double)));
	CTB_CHECK(ctbd_malloc((void**)&d_ret, (size_t)n * sizeof(double)));
	CTB_CHECK(ctbd_h2d(d_s, sigma, (size_t)n * sizeof(double)));
	ct_long* ind = ctb_malloc((size_t)n * sizeof(ct_long));
	int64_t nret = 0;
	double info3[3];
	int rc = ctbd_truncate_select((int64_t)n, d_s, tol, relative_thresh ? 1 : 0, (int64_t)max_vdim, 0, &nret, (int64_t*)ind, info3, d_ret);
	ctbd_free(d_ret); ctbd_free(d_s);
	if (rc < 0) { ctb_free(ind); return rc; }
	info->norm_sigma = info3[0]; info->entropy = info3[1]; info->tol_eff = info3[2];
	if (nret == 0) { ctb_free(ind); return 0; }
	list->ind = ind; list->num = (ct_long)nret;
	return 0;
}

void retained_bond_indices(const double* sigma, const ct_long n, const double tol, const bool relative_thresh, const ct_long max_vdim, struct index_list* list, struct trunc_info* info)
{
	ctb_retained_bond_indices(sigma, n, tol, relative_thresh, max_vdim, list, info);
}

int split_block_sparse_matrix_svd(const struct block_sparse_tensor* a, const double tol, const bool relative_thresh, const ct_long max_vdim, const bool renormalize, const enum singular_value_distr svd_distr, struct block_sparse_tensor* a0, struct block_sparse_tensor* a1, struct trunc_info* info)
{
	CTB_CHECK(ctbd_init(-1));
	struct ctb_tensor* ad = ctb_upload(a);
	struct ctb_tensor *d0 = NULL, *d1 = NULL;
	int rc = ctb_split_matrix_svd(ad, tol, relative_thresh, max_vdim, renormalize, (int)svd_distr, &d0, &d1, info);
	ctb_tensor_free(ad);
	if (rc < 0) { return rc; }
	finish(d0, a0);
	finish(d1, a1);
	return 0;
}

/* ---- MPS / MPO pieces ---- */

void mps_local_orthonormalize_qr(struct block_sparse_tensor* a, struct block_sparse_tensor* a_next)
{
	ensure_init();
	struct ctb_tensor* ad = ctb_upload(a);
	struct ctb_tensor* nd = ctb_upload(a_next);
	CTB_CHECK_ABORT(ctb_mps_local_qr(&ad, &nd));
	delete_block_sparse_tensor(a);
	delete_block_sparse_tensor(a_next);
	finish(ad, a);
	finish(nd, a_next);
}

void mps_local_orthonormalize_rq(struct block_sparse_tensor* a, struct block_sparse_tensor* a_prev)
{
	ensure_init();
	struct ctb_tensor* ad = ctb_upload(a);
	struct ctb_tensor* pd = ctb_upload(a_prev);
	CTB_CHECK_ABORT(ctb_mps_local_rq(&ad, &pd));
	delete_block_sparse_tensor(a);
	delete_block_sparse_tensor(a_prev);
	finish(ad, a);
	finish(pd, a_prev);
}

double mps_orthonormalize_qr(struct mps* mps, const enum mps_orthonormalization_mode mode)
{
	ensure_init();
	const int L = mps->nsites;
	struct ctb_tensor** A = calloc((size_t)L, sizeof(struct ctb_tensor*));
	for (int i = 0; i < L; i++) { A[i] = ctb_upload(&mps->a[i]); }
	const int dtype = A[0]->dtype;
	const qnumber qzero[1] = { 0 };
	const ct_long dim1[3] = { 1, 1, 1 };
	const int dirs[3] = { TENSOR_AXIS_OUT, TENSOR_AXIS_OUT, TENSOR_AXIS_IN };
	struct ctb_tensor* cap = NULL;
	int edge;
	if (mode == MPS_ORTHONORMAL_LEFT)
	{
		for (int i = 0; i < L - 1; i++) { CTB_CHECK_ABORT(ctb_mps_local_qr(&A[i], &A[i + 1])); }
		edge = L - 1;
		CTB_REQUIRE(A[edge]->ax[2].dim == 1);
		const qnumber* qn[3] = { A[edge]->ax[2].qlog, qzero, A[edge]->ax[2].qlog };
		cap = ctb_tensor_create(dtype, 3, dim1, dirs, qn, 1);
		CTB_CHECK_ABORT(ctb_set_entry(cap, 0, 1.0, 0.0));
		CTB_CHECK_ABORT(ctb_mps_local_qr(&A[edge], &cap));
	}
	else
	{
		for (int i = L - 1; i > 0; i--) { CTB_CHECK_ABORT(ctb_mps_local_rq(&A[i], &A[i - 1])); }
		edge = 0;
		CTB_REQUIRE(A[0]->ax[0].dim == 1);
		const qnumber* qn[3] = { A[0]->ax[0].qlog, qzero, A[0]->ax[0].qlog };
		cap = ctb_tensor_create(dtype, 3, dim1, dirs, qn, 1);
		CTB_CHECK_ABORT(ctb_set_entry(cap, 0, 1.0, 0.0));
		CTB_CHECK_ABORT(ctb_mps_local_rq(&A[0], &cap));
	}
	double norm = 0;
	if (ctb_grid_offset(cap, 0) >= 0)
	{
		double v[2] = { 0, 0 };
		CTB_CHECK_ABORT(ctbd_d2h(v, (char*)cap->d + (size_t)ctb_grid_offset(cap, 0) * ctb_sizeof_dtype(dtype), ctb_sizeof_dtype(dtype)));
		norm = v[0];
		if (norm < 0) {
			CTB_CHECK_ABORT(ctbd_scale_host(dtype, A[edge]->nstore, A[edge]->d, -1.0));
			norm = -norm;
		}
	}
	ctb_tensor_free(cap);
	for (int i = 0; i < L; i++) {
		delete_block_sparse_tensor(&mps->a[i]);
		finish(A[i], &mps->a[i]);
	}
	free(A);
	return norm;
}

int mps_split_tensor_svd(const struct block_sparse_tensor* a, const ct_long d[2], const qnumber* new_qsite[2], const double tol, const ct_long max_vdim, const bool renormalize, const enum singular_value_distr svd_distr, struct block_sparse_tensor* a0, struct block_sparse_tensor* a1, struct trunc_info* info)
{
	CTB_CHECK(ctbd_init(-1));
	struct ctb_tensor* ad = ctb_upload(a);
	struct ctb_tensor *d0 = NULL, *d1 = NULL;
	int rc = ctb_mps_split_svd(ad, d, (const qnumber* const*)new_qsite, tol, max_vdim, renormalize, (int)svd_distr, &d0, &d1, info);
	ctb_tensor_free(ad);
	if (rc < 0) { return rc; }
	finish(d0, a0);
	finish(d1, a1);
	return 0;
}

void mps_merge_tensor_pair(const struct block_sparse_tensor* a0, const struct block_sparse_tensor* a1, struct block_sparse_tensor* a)
{
	ensure_init();
	struct ctb_tensor* d0 = ctb_upload(a0);
	struct ctb_tensor* d1 = ctb_upload(a1);
	struct ctb_tensor* ad = ctb_mps_merge_pair(d0, d1);
	ctb_tensor_free(d0); ctb_tensor_free(d1);
	finish(ad, a);
}

void mpo_merge_tensor_pair(const struct block_sparse_tensor* a0, const struct block_sparse_tensor* a1, struct block_sparse_tensor* a)
{
	ensure_init();
	struct ctb_tensor* d0 = ctb_upload(a0);
	struct ctb_tensor* d1 = ctb_upload(a1);
	struct ctb_tensor* ad = ctb_mpo_merge_pair(d0, d1);
	ctb_tensor_free(d0); ctb_tensor_free(d1);
	finish(ad, a);
}

/* ---- chain operations ---- */

void create_dummy_operator_block_right(const struct block_sparse_tensor* a, const struct block_sparse_tensor* b, const struct block_sparse_tensor* w, struct block_sparse_tensor* r)
{
	ensure_init();
	struct ctb_tensor* ad = ctb_upload(a); struct ctb_tensor* bd = ctb_upload(b); struct ctb_tensor* wd = ctb_upload(w);
	struct ctb_tensor* rd = ctb_dummy_block_right(ad, bd, wd);
	ctb_tensor_free(ad); ctb_tensor_free(bd); ctb_tensor_free(wd);
	finish(rd, r);
}

void create_dummy_operator_block_left(const struct block_sparse_tensor* a, const struct block_sparse_tensor* b, const struct block_sparse_tensor* w, struct block_sparse_tensor* l)
{
	ensure_init();
	struct ctb_tensor* ad = ctb_upload(a); struct ctb_tensor* bd = ctb_upload(b); struct ctb_tensor* wd = ctb_upload(w);
	struct ctb_tensor* ld = ctb_dummy_block_left(ad, bd, wd);
	ctb_tensor_free(ad); ctb_tensor_free(bd); ctb_tensor_free(wd);
	finish(ld, l);
}

void contraction_operator_step_right(const struct block_sparse_tensor* a, const struct block_sparse_tensor* b, const struct block_sparse_tensor* w, const struct block_sparse_tensor* r, struct block_sparse_tensor* r_next)
{
	ensure_init();
	struct ctb_tensor* ad = ctb_upload(a); struct ctb_tensor* bd = ctb_upload(b); struct ctb_tensor* wd = ctb_upload(w); struct ctb_tensor* rd = ctb_upload(r);
	struct ctb_tensor* nd = ctb_env_step_right(ad, bd, wd, rd);
	ctb_tensor_free(ad); ctb_tensor_free(bd); ctb_tensor_free(wd); ctb_tensor_free(rd);
	finish(nd, r_next);
}

void contraction_operator_step_left(const struct block_sparse_tensor* a, const struct block_sparse_tensor* b, const struct block_sparse_tensor* w, const struct block_sparse_tensor* l, struct block_sparse_tensor* l_next)
{
	ensure_init();
	struct ctb_tensor* ad = ctb_upload(a); struct ctb_tensor* bd = ctb_upload(b); struct ctb_tensor* wd = ctb_upload(w); struct ctb_tensor* ld = ctb_upload(l);
	struct ctb_tensor* nd = ctb_env_step_left(ad, bd, wd, ld);
	ctb_tensor_free(ad); ctb_tensor_free(bd); ctb_tensor_free(wd); ctb_tensor_free(ld);
	finish(nd, l_next);
}

void compute_right_operator_blocks(const struct mps* psi, const struct mps* chi, const struct mpo* op, struct block_sparse_tensor* r_list)
{
	ensure_init();
	const int L = op->nsites;
	CTB_REQUIRE(psi->nsites == L && chi->nsites == L && L >= 1);
	struct ctb_tensor* r = NULL;
	{
		struct ctb_tensor* ad = ctb_upload(&psi->a[L - 1]); struct ctb_tensor* bd = ctb_upload(&chi->a[L - 1]); struct ctb_tensor* wd = ctb_upload(&op->a[L - 1]);
		r = ctb_dummy_block_right(ad, bd, wd);
		ctb_tensor_free(ad); ctb_tensor_free(bd); ctb_tensor_free(wd);
	}
	CTB_CHECK_ABORT(ctb_download(r, &r_list[L - 1]));
	for (int i = L - 1; i > 0; i--)
	{
		struct ctb_tensor* ad = ctb_upload(&psi->a[i]); struct ctb_tensor* bd = ctb_upload(&chi->a[i]); struct ctb_tensor* wd = ctb_upload(&op->a[i]);
		struct ctb_tensor* rn = ctb_env_step_right(ad, bd, wd, r);
		ctb_tensor_free(ad); ctb_tensor_free(bd); ctb_tensor_free(wd);
		ctb_tensor_free(r);
		r = rn;
		CTB_CHECK_ABORT(ctb_download(r, &r_list[i - 1]));
	}
	ctb_tensor_free(r);
}

/* Host-struct entry point of the hot path.  The pieces that do not depend on each other run side by side:
 *   - a helper thread streams the payloads of l, r and a through the pinned staging ring (host packing threads + copy engine)
 *     while this thread builds the three contraction plans, which need sector structure only (and the small payload of w);
 *   - the host blocks of the result are allocated and their pages faulted in while the device runs the three launches;
 * so the call costs about max(upload, plans) + matvec + download instead of their sum. */
struct heff_upload_job
{
	struct ctb_tensor* t[3];
	const struct block_sparse_tensor* h[3];
	int rc;
	int first_done;                 /* payload of t[0] (= l) is enqueued on the stream */
	pthread_mutex_t mtx;
	pthread_cond_t cv;
};

static void* heff_upload_main(void* arg)
{
	struct heff_upload_job* job = arg;
	for (int i = 0; i < 3; i++) {
		const int rc = ctb_upload_data(job->t[i], job->h[i]);
		if (rc < 0) { job->rc = rc; }
		if (i == 0) {
			pthread_mutex_lock(&job->mtx);
			job->first_done = 1;
			pthread_cond_signal(&job->cv);
			pthread_mutex_unlock(&job->mtx);
		}
	}
	return NULL;
}

static void heff_wait_first(void* arg)
{
	struct heff_upload_job* job = arg;
	pthread_mutex_lock(&job->mtx);
	while (!job->first_done) { pthread_cond_wait(&job->cv, &job->mtx); }
	pthread_mutex_unlock(&job->mtx);
}

void apply_local_hamiltonian(const struct block_sparse_tensor* a, const struct block_sparse_tensor* w, const struct block_sparse_tensor* l, const struct block_sparse_tensor* r, struct block_sparse_tensor* b)
{
	ensure_init();
	const bool trace = getenv("CTB_TRACE") != NULL;
	const double t0 = ctb_wall_ms();
	struct ctb_tensor* ad = ctb_upload_begin(a); struct ctb_tensor* ld = ctb_upload_begin(l); struct ctb_tensor* rd = ctb_upload_begin(r);
	struct ctb_tensor* wd = ctb_upload(w);       /* small, and read by the plan of the MPO-mixing step */
	struct heff_upload_job job;
	job.t[0] = ld; job.t[1] = rd; job.t[2] = ad;
	job.h[0] = l;  job.h[1] = r;  job.h[2] = a;
	job.rc = 0; job.first_done = 0;
	pthread_mutex_init(&job.mtx, NULL);
	pthread_cond_init(&job.cv, NULL);
	pthread_t th;
	/* sharded: the plan builder cuts the column slice of r on the device right away, so all payloads go first */
	const bool threaded = (ctb_dist_world == 1) && (getenv("CTB_NO_OVERLAP") == NULL) && (pthread_create(&th, NULL, heff_upload_main, &job) == 0);
	if (!threaded) { ctb_collective_upload++; heff_upload_main(&job); ctb_collective_upload--; }
	const double t1 = ctb_wall_ms();
	struct ctb_heff h;
	CTB_CHECK_ABORT(ctb_heff_prepare_ex(ad, wd, ld, rd, &h, heff_wait_first, &job));
	const double t2 = ctb_wall_ms();
	if (threaded) { pthread_join(th, NULL); }
	pthread_mutex_destroy(&job.mtx);
	pthread_cond_destroy(&job.cv);
	CTB_CHECK_ABORT(job.rc);
	const double t3 = ctb_wall_ms();
	struct ctb_tensor* bd = ctb_tensor_like(h.b, 1);
	CTB_CHECK_ABORT(ctb_heff_apply(&h, ad->d, bd->d));
	/* the device is busy with the three launches: allocate the host result and take its page faults now */
	ctb_download_begin(bd, b, 1);
	if (trace) { ctbd_sync(); }
	const double t4 = ctb_wall_ms();
	ctb_heff_free(&h);
	ctb_tensor_free(ad); ctb_tensor_free(wd); ctb_tensor_free(ld); ctb_tensor_free(rd);
	CTB_CHECK_ABORT(ctb_download_data(bd, b));
	ctb_tensor_free(bd);
	if (trace) {
		fprintf(stderr, "apply_local_hamiltonian: setup %.2f ms, plans (upload alongside) %.2f ms, upload tail %.2f ms, matvec + result alloc %.2f ms, download %.2f ms\n",
			t1 - t0, t2 - t1, t3 - t2, t4 - t3, ctb_wall_ms() - t4);
	}
}

/* Extension: the effective Hamiltonian with the two single-site MPO tensors applied one after the other (no merged pair tensor,
 * see ctb_heff_prepare_pair).  `a` and the result are the reference's two-site tensors [Dl, d0*d1, Dr]; equals
 * apply_local_hamiltonian(a, mpo_merge_tensor_pair(w0, w1), l, r). */
int ctb_apply_local_hamiltonian_pair(const struct block_sparse_tensor* a, const struct block_sparse_tensor* w0, const struct block_sparse_tensor* w1,
	const struct block_sparse_tensor* l, const struct block_sparse_tensor* r, struct block_sparse_tensor* b)
{
	CTB_CHECK(ctbd_init(-1));
	struct ctb_tensor* ad = ctb_upload(a); struct ctb_tensor* w0d = ctb_upload(w0); struct ctb_tensor* w1d = ctb_upload(w1);
	struct ctb_tensor* ld = ctb_upload(l); struct ctb_tensor* rd = ctb_upload(r);
	const ct_long dims[2] = { w0d->ax[1].dim, w1d->ax[1].dim };
	const int dirs[2] = { TENSOR_AXIS_OUT, TENSOR_AXIS_OUT };
	const qnumber* qn[2] = { w0d->ax[1].qlog, w1d->ax[1].qlog };
	struct ctb_tensor* a4 = ctb_split_axis(ad, 1, dims, dirs, qn);
	struct ctb_heff h;
	int rc = ctb_heff_prepare_pair(a4, w0d, w1d, ld, rd, &h);
	if (rc == 0) {
		struct ctb_tensor* b4 = ctb_tensor_like(a4, 1);
		rc = ctb_heff_apply(&h, a4->d, b4->d);
		if (rc == 0) {
			struct ctb_tensor* b3 = ctb_flatten_axes(b4, 1, TENSOR_AXIS_OUT);
			rc = ctb_download(b3, b);
			ctb_tensor_free(b3);
		}
		ctb_tensor_free(b4);
		ctb_heff_free(&h);
	}
	ctb_tensor_free(a4);
	ctb_tensor_free(ad); ctb_tensor_free(w0d); ctb_tensor_free(w1d); ctb_tensor_free(ld); ctb_tensor_free(rd);
	return rc;
}

/* ---- measurement ---- */

int ctb_heff_benchmark(const struct block_sparse_tensor* a, const struct block_sparse_tensor* w, const struct block_sparse_tensor* l, const struct block_sparse_tensor* r,
	int warmup, int reps, int flush_l2, double* ms_per_matvec, double* flops_per_matvec, double* per_step_ms, double* per_step_flops)
{
	CTB_CHECK(ctbd_init(-1));
	struct ctb_tensor* ad = ctb_upload(a); struct ctb_tensor* wd = ctb_upload(w); struct ctb_tensor* ld = ctb_upload(l); struct ctb_tensor* rd = ctb_upload(r);
	struct ctb_heff h;
	CTB_CHECK(ctb_heff_prepare(ad, wd, ld, rd, &h));
	struct ctb_tensor* bd = ctb_tensor_like(h.b, 1);
	/* optional L2 flush buffer (larger than the 126 MB L2) rewritten between timed matvecs */
	void* flush = NULL;
	const size_t flush_bytes = (size_t)192 << 20;
	if (flush_l2) { CTB_CHECK(ctbd_malloc(&flush, flush_bytes)); }
	for (int i = 0; i < warmup; i++) { CTB_CHECK(ctb_heff_apply(&h, ad->d, bd->d)); }
	void *e0 = NULL, *e1 = NULL, *e2 = NULL, *e3 = NULL, *e4 = NULL;
	CTB_CHECK(ctbd_event_create(&e0)); CTB_CHECK(ctbd_event_create(&e1)); CTB_CHECK(ctbd_event_create(&e2)); CTB_CHECK(ctbd_event_create(&e3)); CTB_CHECK(ctbd_event_create(&e4));
	double tot = 0, t1 = 0, t2 = 0, t3 = 0;
	for (int i = 0; i < reps; i++)
	{
		if (flush_l2) { CTB_CHECK(ctbd_memset_zero(flush, flush_bytes)); }
		CTB_CHECK(ctbd_event_record(e0));
		CTB_CHECK(ctb_dot_exec(&h.p1, ad->d, h.r->d, h.t1->d));
		CTB_CHECK(ctbd_event_record(e1));
		CTB_CHECK(ctb_dot_exec(&h.p2, wd->d, h.t1->d, h.t2->d));
		CTB_CHECK(ctbd_event_record(e2));
		void* bl = ctb_heff_result_buffer(&h);      /* fused exchange: consume in place, as the Lanczos loop does */
		CTB_CHECK(ctb_heff_step3(&h, bl != NULL ? bl : bd->d));
		CTB_CHECK(ctbd_event_record(e3));
		CTB_CHECK(ctb_heff_exchange(&h, bl != NULL ? bl : bd->d));
		CTB_CHECK(ctbd_event_record(e4));
		float ms;
		CTB_CHECK(ctbd_event_elapsed_ms(e0, e4, &ms)); tot += ms;
		CTB_CHECK(ctbd_event_elapsed_ms(e0, e1, &ms)); t1 += ms;
		CTB_CHECK(ctbd_event_elapsed_ms(e1, e2, &ms)); t2 += ms;
		CTB_CHECK(ctbd_event_elapsed_ms(e2, e3, &ms)); t3 += ms;
	}
	*ms_per_matvec = tot / reps;
	*flops_per_matvec = h.flops;
	if (per_step_ms != NULL) { per_step_ms[0] = t1 / reps; per_step_ms[1] = t2 / reps; per_step_ms[2] = t3 / reps; }
	if (per_step_flops != NULL) { per_step_flops[0] = h.p1.flops; per_step_flops[1] = h.p2.flops; per_step_flops[2] = h.p3.flops; }
	ctbd_event_destroy(e0); ctbd_event_destroy(e1); ctbd_event_destroy(e2); ctbd_event_destroy(e3); ctbd_event_destroy(e4);
	if (flush != NULL) { CTB_CHECK(ctbd_free(flush)); }
	ctb_heff_free(&h);
	ctb_tensor_free(bd);
	ctb_tensor_free(ad); ctb_tensor_free(wd); ctb_tensor_free(ld); ctb_tensor_free(rd);
	return 0;
}

/* one contraction r = dot(s, t), device-resident, plan built once: mean device time per execution (CUDA events) */
int ctb_dot_benchmark(const struct block_sparse_tensor* s, const int axrange_s, const struct block_sparse_tensor* t, const int axrange_t, const int ndim_mult,
	int warmup, int reps, int flush_l2, double* ms_per_run, double* flops_per_run)
{
	CTB_CHECK(ctbd_init(-1));
	struct ctb_tensor* sd = ctb_upload(s); struct ctb_tensor* td = ctb_upload(t);
	struct ctb_dot_plan pl;
	struct ctb_tensor* rd = ctb_dot_prepare(sd, axrange_s, 0, td, axrange_t, 0, ndim_mult, NULL, 1, &pl);
	void* flush = NULL;
	const size_t flush_bytes = (size_t)192 << 20;
	if (flush_l2) { CTB_CHECK(ctbd_malloc(&flush, flush_bytes)); }
	for (int i = 0; i < warmup; i++) { CTB_CHECK(ctb_dot_exec(&pl, sd->d, td->d, rd->d)); }
	void *e0 = NULL, *e1 = NULL;
	CTB_CHECK(ctbd_event_create(&e0)); CTB_CHECK(ctbd_event_create(&e1));
	double tot = 0;
	for (int i = 0; i < reps; i++)
	{
		if (flush_l2) { CTB_CHECK(ctbd_memset_zero(flush, flush_bytes)); }
		CTB_CHECK(ctbd_event_record(e0));
		CTB_CHECK(ctb_dot_exec(&pl, sd->d, td->d, rd->d));
		CTB_CHECK(ctbd_event_record(e1));
		float ms;
		CTB_CHECK(ctbd_event_elapsed_ms(e0, e1, &ms)); tot += ms;
	}
	*ms_per_run = tot / (reps > 0 ? reps : 1);
	*flops_per_run = pl.flops;
	ctbd_event_destroy(e0); ctbd_event_destroy(e1);
	if (flush != NULL) { CTB_CHECK(ctbd_free(flush)); }
	ctb_dot_plan_free(&pl);
	ctb_tensor_free(rd); ctb_tensor_free(sd); ctb_tensor_free(td);
	return 0;
}

/* plan-only query of the (sharded) effective Hamiltonian: out[0] = algorithmic flops of rank 'rank' of 'world', out[1..3] = tiles of the
 * three launches, out[4] = stored entries of this rank's result slice, out[5] = entries of one all-gather slot, out[6] = entries of t1, out[7] = of t2 */
int ctb_heff_plan_info(const struct block_sparse_tensor* a, const struct block_sparse_tensor* w, const struct block_sparse_tensor* l, const struct block_sparse_tensor* r,
	int rank, int world, double* out)
{
	CTB_CHECK(ctbd_init(-1));
	const int rank0 = ctb_dist_rank, world0 = ctb_dist_world;
	struct ctb_tensor* ad = ctb_upload(a); struct ctb_tensor* wd = ctb_upload(w); struct ctb_tensor* ld = ctb_upload(l); struct ctb_tensor* rd = ctb_upload(r);
	ctb_dist_rank = rank; ctb_dist_world = world;
	struct ctb_heff h;
	int rc = ctb_heff_prepare(ad, wd, ld, rd, &h);
	ctb_dist_rank = rank0; ctb_dist_world = world0;
	if (rc == 0) {
		out[0] = h.flops; out[1] = h.p1.ntiles; out[2] = h.p2.ntiles; out[3] = h.p3.ntiles;
		out[4] = (h.world > 1) ? (double)h.piece[h.rank]->nstore : (double)h.b->nstore;
		out[5] = (double)h.piece_cap; out[6] = (double)h.t1->nstore; out[7] = (double)h.t2->nstore;
		ctb_heff_free(&h);
	}
	ctb_tensor_free(ad); ctb_tensor_free(wd); ctb_tensor_free(ld); ctb_tensor_free(rd);
	return rc;
}

/* device time (CUDA events around the kernel launch alone, 192 MiB L2 flush between repetitions) of the re-blocking kernel on a
 * device-resident tensor: out[2c], out[2c+1] = ms and algorithmic bytes (2 x stored entries x element size) of
 *   c = 0: transposing the axes to reverse order, c = 1: fusing axes (0, 1), c = 2: fusing the last two axes
 * -- the HBM-bound kernels of merge / split (reference block_sparse_tensor.c:785, :950).  The destination tensor, its sector tables
 * and their upload are prepared outside the timed region (round 1 timed them with the kernel). */
int ctb_remap_benchmark(const struct block_sparse_tensor* t, double* out)
{
	CTB_CHECK(ctbd_init(-1));
	struct ctb_tensor* td = ctb_upload(t);
	const size_t flush_bytes = (size_t)192 << 20;
	void* flush = NULL;
	CTB_CHECK(ctbd_malloc(&flush, flush_bytes));
	void *e0 = NULL, *e1 = NULL;
	CTB_CHECK(ctbd_event_create(&e0)); CTB_CHECK(ctbd_event_create(&e1));
	const double bytes = 2.0 * (double)td->nelem * (double)ctb_sizeof_dtype(td->dtype);
	int perm[CTB_MAXDIM];
	for (int i = 0; i < td->ndim; i++) { perm[i] = td->ndim - 1 - i; }
	for (int which = 0; which < 3; which++)
	{
		struct ctbd_remap_args args;
		memset(&args, 0, sizeof(args));
		args.scale_ax = -1;
		struct ctb_tensor* r = NULL;
		if (which == 0) {
			r = ctb_transpose(td, perm, 0);
			args.op = CTBD_REMAP_TRANSPOSE;
			for (int i = 0; i < td->ndim; i++) { args.perm[i] = perm[i]; }
		}
		else {
			const int i_ax = (which == 1) ? 0 : td->ndim - 2;
			r = ctb_flatten_axes(td, i_ax, td->ax[i_ax].dir);
			args.op = CTBD_REMAP_FLATTEN;
			args.i_ax = i_ax;
		}
		args.dst_layout = ctb_tensor_layout(r);  args.dst = r->d;
		args.src_layout = ctb_tensor_layout(td); args.src = td->d;
		double best = 1e30;
		for (int rep = 0; rep < 5; rep++)
		{
			CTB_CHECK(ctbd_memset_zero(flush, flush_bytes));
			CTB_CHECK(ctbd_event_record(e0));
			CTB_CHECK(ctbd_remap(&args));
			CTB_CHECK(ctbd_event_record(e1));
			float ms = 0;
			CTB_CHECK(ctbd_event_elapsed_ms(e0, e1, &ms));
			if (rep > 0 && ms < best) { best = ms; }
		}
		ctb_tensor_free(r);
		out[2 * which] = best; out[2 * which + 1] = bytes;
	}
	ctbd_event_destroy(e0); ctbd_event_destroy(e1);
	CTB_CHECK(ctbd_free(flush));
	ctb_tensor_free(td);
	return 0;
}

int ctb_get_stats(double* out, int n)
{
	for (int i = 11; i < n && i < 19; i++) { out[i] = ctb_global_stats.sweep_ms[i - 11]; }
	const double v[11] = {
		ctb_global_stats.heff_flops, (double)ctb_global_stats.heff_calls, ctb_global_stats.env_flops,
		ctb_global_stats.lanczos_ms, ctb_global_stats.svd_ms, ctb_global_stats.env_ms, ctb_global_stats.total_ms,
		(double)ctb_global_stats.max_vector_len, (double)ctb_global_stats.max_bond_dim, ctb_global_stats.plan_ms, ctb_global_stats.remap_ms };
	for (int i = 0; i < n && i < 11; i++) { out[i] = v[i]; }
	return 0;
}
