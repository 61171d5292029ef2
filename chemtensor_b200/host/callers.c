/*
 * callers.c -- the other callers of the hot-path primitives, behind the reference's names.
 *
 * Everything here is a composition of the device-resident primitives of tensor_ops.c / chain.c (grouped-GEMM contraction
 * with fused output permutation and conjugation, re-blocking kernels, batched SVD): operands are uploaded once, all
 * intermediates stay in HBM, only results cross back.
 *
 *   mps_vdot, mps_norm                        <- src/state/mps.c:279-425
 *   mpo_inner_product                         <- src/algorithm/chain_ops.c:274-321
 *   apply_mpo                                 <- src/algorithm/chain_ops.c:485-528
 *   compute_local_hamiltonian_environment     <- src/algorithm/chain_ops.c:424-478   (building block of
 *                                                operator_average_coefficient_gradient, src/algorithm/gradient.c:39-222)
 *   mps_local_orthonormalize_left/right_svd   <- src/state/mps.c:764-860
 *   mps_compress, mps_compress_rescale        <- src/state/mps.c:868-1112
 *   split_block_sparse_matrix_svd_isometry    <- src/algorithm/bond_ops.c:146-199
 */
#include "ctb_internal.h"
#include "chemtensor_b200.h"

static void ensure_init(void) { CTB_CHECK_ABORT(ctbd_init(-1)); }

static void finish(struct ctb_tensor* dev, struct block_sparse_tensor* out)
{
	CTB_CHECK_ABORT(ctb_download(dev, out));
	ctb_tensor_free(dev);
}

/* the single stored entry of a tensor whose logical dimensions are all 1 (0 if its one block does not conserve) */
static void read_scalar(const struct ctb_tensor* t, void* ret)
{
	const size_t esize = ctb_sizeof_dtype(t->dtype);
	memset(ret, 0, esize);
	CTB_REQUIRE(t->nelem <= 1);
	if (t->nstore >= 1) { CTB_CHECK_ABORT(ctbd_d2h(ret, t->d, esize)); }
}

/* ---- <chi | psi> ---- */

/* one transfer-matrix step from the right (mps.c:279-308): r[Dr_a, Dr_b, x] -> r'[Dl_a, Dl_b, x];
 * both transposes of the reference are fused into the output permutations of the two launches, conj(b) into the operand load */
static struct ctb_tensor* mps_step_right(const struct ctb_tensor* a, const struct ctb_tensor* b, const struct ctb_tensor* r)
{
	const int perm0[4] = { 3, 0, 1, 2 };
	struct ctb_tensor* s = ctb_dot(a, TENSOR_AXIS_RANGE_TRAILING, 0, r, TENSOR_AXIS_RANGE_LEADING, 0, 1, perm0);     /* [x, Dl_a, d, Dr_b] */
	struct ctb_tensor* br = ctb_view_reversed_dirs(b);
	const int perm1[3] = { 1, 2, 0 };
	struct ctb_tensor* rn = ctb_dot(s, TENSOR_AXIS_RANGE_TRAILING, 0, br, TENSOR_AXIS_RANGE_TRAILING, ctb_is_complex(b->dtype), 2, perm1);   /* [Dl_a, Dl_b, x] */
	ctb_tensor_free(br);
	ctb_tensor_free(s);
	return rn;
}

static void mps_vdot_device(struct ctb_tensor* const* chi, struct ctb_tensor* const* psi, int nsites, void* ret)
{
	const int dtype = psi[0]->dtype;
	CTB_REQUIRE(chi[0]->ax[0].dim == 1 && psi[0]->ax[0].dim == 1 && chi[nsites - 1]->ax[2].dim == 1 && psi[nsites - 1]->ax[2].dim == 1);
	struct ctb_tensor* r = NULL;
	{
		const ct_long dim[4] = { 1, 1, 1, 1 };
		const int dirs[4] = { TENSOR_AXIS_OUT, TENSOR_AXIS_IN, TENSOR_AXIS_IN, TENSOR_AXIS_OUT };
		const qnumber* qn[4] = { psi[nsites - 1]->ax[2].qlog, chi[nsites - 1]->ax[2].qlog, psi[nsites - 1]->ax[2].qlog, chi[nsites - 1]->ax[2].qlog };
		struct ctb_tensor* t = ctb_tensor_create(dtype, 4, dim, dirs, qn, 1);
		CTB_REQUIRE(t->nblk == 1);
		CTB_CHECK_ABORT(ctb_set_entry(t, 0, 1.0, 0.0));
		r = ctb_flatten_axes(t, 2, TENSOR_AXIS_IN);
		ctb_tensor_free(t);
	}
	for (int i = nsites - 1; i >= 0; i--) {
		struct ctb_tensor* rn = mps_step_right(psi[i], chi[i], r);
		ctb_tensor_free(r);
		r = rn;
	}
	struct ctb_tensor* flat = ctb_flatten_axes(r, 0, TENSOR_AXIS_OUT);
	ctb_tensor_free(r);
	CTB_REQUIRE(flat->ndim == 2 && flat->ax[0].dim == 1 && flat->ax[1].dim == 1);
	read_scalar(flat, ret);
	ctb_tensor_free(flat);
}

static struct ctb_tensor** upload_sites(const struct block_sparse_tensor* a, int nsites)
{
	struct ctb_tensor** A = calloc((size_t)nsites, sizeof(struct ctb_tensor*));
	for (int i = 0; i < nsites; i++) { A[i] = ctb_upload(&a[i]); }
	return A;
}

static void free_sites(struct ctb_tensor** A, int nsites)
{
	for (int i = 0; i < nsites; i++) { ctb_tensor_free(A[i]); }
	free(A);
}

void mps_vdot(const struct mps* chi, const struct mps* psi, void* ret)
{
	ensure_init();
	CTB_REQUIRE(psi->nsites == chi->nsites && psi->nsites >= 1);
	struct ctb_tensor** P = upload_sites(psi->a, psi->nsites);
	struct ctb_tensor** X = (chi == psi) ? P : upload_sites(chi->a, chi->nsites);
	mps_vdot_device(X, P, psi->nsites, ret);
	if (X != P) { free_sites(X, chi->nsites); }
	free_sites(P, psi->nsites);
}

double mps_norm(const struct mps* psi)
{
	if (psi->nsites == 0) { return 0; }
	double v[2] = { 0, 0 };
	mps_vdot(psi, psi, v);
	return sqrt(v[0] > 0 ? v[0] : 0);
}

/* ---- <chi | op | psi> ---- */

void mpo_inner_product(const struct mps* chi, const struct mpo* op, const struct mps* psi, void* ret)
{
	ensure_init();
	const int L = op->nsites;
	CTB_REQUIRE(chi->nsites == L && psi->nsites == L && L >= 1);
	CTB_REQUIRE(chi->a[0].dim_logical[0] == 1 && op->a[0].dim_logical[0] == 1 && psi->a[0].dim_logical[0] == 1);
	CTB_REQUIRE(chi->a[L - 1].dim_logical[2] == 1 && op->a[L - 1].dim_logical[3] == 1 && psi->a[L - 1].dim_logical[2] == 1);
	struct ctb_tensor* r = NULL;
	for (int i = L - 1; i >= 0; i--)
	{
		struct ctb_tensor* ad = ctb_upload(&psi->a[i]);
		struct ctb_tensor* bd = (chi == psi) ? ad : ctb_upload(&chi->a[i]);
		struct ctb_tensor* wd = ctb_upload(&op->a[i]);
		if (r == NULL) { r = ctb_dummy_block_right(ad, bd, wd); }
		struct ctb_tensor* rn = ctb_env_step_right(ad, bd, wd, r);
		ctb_tensor_free(r);
		r = rn;
		if (bd != ad) { ctb_tensor_free(bd); }
		ctb_tensor_free(ad); ctb_tensor_free(wd);
	}
	/* flatten the left virtual bonds (chain_ops.c:305-311): a 1 x 1 tensor remains */
	struct ctb_tensor* t = ctb_flatten_axes(r, 0, TENSOR_AXIS_OUT);
	ctb_tensor_free(r);
	r = ctb_flatten_axes(t, 0, TENSOR_AXIS_OUT);
	ctb_tensor_free(t);
	CTB_REQUIRE(r->ndim == 2 && r->ax[0].dim == 1 && r->ax[1].dim == 1);
	read_scalar(r, ret);
	ctb_tensor_free(r);
}

/* ---- op |psi> ---- */

void apply_mpo(const struct mpo* op, const struct mps* psi, struct mps* op_psi)
{
	ensure_init();
	CTB_REQUIRE(psi->d == op->d && psi->nsites == op->nsites && psi->nsites >= 1);
	for (ct_long j = 0; j < psi->d; j++) { CTB_REQUIRE(psi->qsite[j] == op->qsite[j]); }
	/* allocate_empty_mps (mps.c:18-29) */
	op_psi->nsites = psi->nsites;
	op_psi->d = psi->d;
	op_psi->qsite = ctb_malloc((size_t)psi->d * sizeof(qnumber));
	memcpy(op_psi->qsite, psi->qsite, (size_t)psi->d * sizeof(qnumber));
	op_psi->a = ctb_calloc((size_t)psi->nsites, sizeof(struct block_sparse_tensor));
	for (int i = 0; i < psi->nsites; i++)
	{
		struct ctb_tensor* wd = ctb_upload(&op->a[i]);
		struct ctb_tensor* ad = ctb_upload(&psi->a[i]);
		const int perm_op[4] = { 0, 1, 3, 2 };
		struct ctb_tensor* r = ctb_transpose(wd, perm_op, 0);          /* [Dw, d_out, Dw', d_in] */
		const int perm_psi[3] = { 1, 0, 2 };
		struct ctb_tensor* s = ctb_transpose(ad, perm_psi, 0);         /* [d_in, Dl, Dr] */
		ctb_tensor_free(wd); ctb_tensor_free(ad);
		/* contraction with the re-ordering [0,3,1,2,4] of chain_ops.c:514 fused into the epilogue: [Dw, Dl, d_out, Dw', Dr] */
		const int perm_ax[5] = { 0, 3, 1, 2, 4 };
		struct ctb_tensor* t = ctb_dot(r, TENSOR_AXIS_RANGE_TRAILING, 0, s, TENSOR_AXIS_RANGE_LEADING, 0, 1, perm_ax);
		ctb_tensor_free(r); ctb_tensor_free(s);
		struct ctb_tensor* f0 = ctb_flatten_axes(t, 0, TENSOR_AXIS_OUT);
		ctb_tensor_free(t);
		struct ctb_tensor* f1 = ctb_flatten_axes(f0, 2, TENSOR_AXIS_IN);
		ctb_tensor_free(f0);
		finish(f1, &op_psi->a[i]);
	}
}

/* ---- d<b| l w a r |..>/dw: the environment of the MPO tensor ---- */

void compute_local_hamiltonian_environment(const struct block_sparse_tensor* a, const struct block_sparse_tensor* b,
	const struct block_sparse_tensor* l, const struct block_sparse_tensor* r, struct block_sparse_tensor* dw)
{
	ensure_init();
	CTB_REQUIRE(a->ndim == 3 && b->ndim == 3 && l->ndim == 4 && r->ndim == 4);
	struct ctb_tensor* ad = ctb_upload(a); struct ctb_tensor* bd = ctb_upload(b); struct ctb_tensor* ld = ctb_upload(l); struct ctb_tensor* rd = ctb_upload(r);
	/* a . r with the last two legs swapped: [Dl, d, Dw', x, Dr'] */
	const int perm0[5] = { 0, 1, 2, 4, 3 };
	struct ctb_tensor* s = ctb_dot(ad, TENSOR_AXIS_RANGE_TRAILING, 0, rd, TENSOR_AXIS_RANGE_LEADING, 0, 1, perm0);
	/* conj(b) . s over Dr', MPS virtual bonds first: [Dl, Dl_b, d_b, d, Dw', x] */
	struct ctb_tensor* br = ctb_view_reversed_dirs(bd);
	const int perm2[6] = { 2, 0, 1, 3, 4, 5 };
	struct ctb_tensor* t = ctb_dot(br, TENSOR_AXIS_RANGE_TRAILING, ctb_is_complex(bd->dtype), s, TENSOR_AXIS_RANGE_TRAILING, 0, 1, perm2);
	ctb_tensor_free(br); ctb_tensor_free(s);
	/* l with its second and third leg swapped, contracted over (Dl, Dl_b): [x, Dw, d_b, d, Dw', x'] */
	const int perm1[4] = { 0, 2, 1, 3 };
	struct ctb_tensor* k = ctb_transpose(ld, perm1, 0);
	struct ctb_tensor* u = ctb_dot(k, TENSOR_AXIS_RANGE_TRAILING, 0, t, TENSOR_AXIS_RANGE_LEADING, 0, 2, NULL);
	ctb_tensor_free(k); ctb_tensor_free(t);
	struct ctb_tensor* res = ctb_drop_dummy_axes(u, 1);
	ctb_tensor_free(u);
	ctb_tensor_free(ad); ctb_tensor_free(bd); ctb_tensor_free(ld); ctb_tensor_free(rd);
	finish(res, dw);
}

/* ---- left isometry of a truncated SVD ---- */

int split_block_sparse_matrix_svd_isometry(const struct block_sparse_tensor* a, const double tol, const bool relative_thresh, const ct_long max_vdim,
	struct block_sparse_tensor* u, struct trunc_info* info)
{
	CTB_CHECK(ctbd_init(-1));
	CTB_REQUIRE(a->ndim == 2);
	struct ctb_tensor* ad = ctb_upload(a);
	struct ctb_tensor *w = NULL, *vh = NULL;
	double* s_dev = NULL;
	ct_long ns = 0;
	int rc = ctb_svd(ad, &w, &s_dev, &ns, &vh);
	ctb_tensor_free(ad);
	if (rc < 0) { return rc; }
	ctb_tensor_free(vh);
	double* sigma = ctb_malloc((size_t)(ns > 0 ? ns : 1) * sizeof(double));
	CTB_CHECK(ctbd_d2h(sigma, s_dev, (size_t)ns * sizeof(double)));
	CTB_CHECK(ctbd_free(s_dev));
	struct index_list retained;
	ctb_retained_bond_indices(sigma, ns, tol, relative_thresh, max_vdim, &retained, info);
	const ct_long ind0[1] = { 0 };
	struct ctb_tensor* us = (retained.num == 0) ? ctb_slice(w, 1, ind0, 1) : ctb_slice(w, 1, retained.ind, retained.num);
	ctb_tensor_free(w);
	ctb_free(retained.ind);
	ctb_free(sigma);
	finish(us, u);
	return 0;
}

/* ---- site-local SVD orthonormalisation with truncation ---- */

static int local_left_svd(double tol, ct_long max_vdim, bool renormalize, struct ctb_tensor** a, struct ctb_tensor** a_next, struct trunc_info* info)
{
	struct ctb_tensor* t = *a;
	CTB_REQUIRE(t->ndim == 3 && (*a_next)->ndim == 3);
	CTB_REQUIRE(t->ax[0].dir == TENSOR_AXIS_OUT && t->ax[1].dir == TENSOR_AXIS_OUT);
	struct ctb_tensor* a_mat = ctb_flatten_axes(t, 0, TENSOR_AXIS_OUT);
	struct ctb_tensor *m0 = NULL, *m1 = NULL;
	int rc = ctb_split_matrix_svd(a_mat, tol, true, max_vdim, renormalize, SVD_DISTR_RIGHT, &m0, &m1, info);
	ctb_tensor_free(a_mat);
	if (rc < 0) { return rc; }
	const ct_long dl[2] = { t->ax[0].dim, t->ax[1].dim };
	const int dirl[2] = { TENSOR_AXIS_OUT, TENSOR_AXIS_OUT };
	const qnumber* ql[2] = { t->ax[0].qlog, t->ax[1].qlog };
	struct ctb_tensor* a_new = ctb_split_axis(m0, 0, dl, dirl, ql);
	ctb_tensor_free(m0);
	ctb_tensor_free(t);
	*a = a_new;
	struct ctb_tensor* upd = ctb_dot(m1, TENSOR_AXIS_RANGE_TRAILING, 0, *a_next, TENSOR_AXIS_RANGE_LEADING, 0, 1, NULL);
	ctb_tensor_free(m1);
	ctb_tensor_free(*a_next);
	*a_next = upd;
	return 0;
}

static int local_right_svd(double tol, ct_long max_vdim, bool renormalize, struct ctb_tensor** a, struct ctb_tensor** a_prev, struct trunc_info* info)
{
	struct ctb_tensor* t = *a;
	CTB_REQUIRE(t->ndim == 3 && (*a_prev)->ndim == 3);
	CTB_REQUIRE(t->ax[1].dir == TENSOR_AXIS_OUT && t->ax[2].dir == TENSOR_AXIS_IN);
	struct ctb_tensor* a_mat = ctb_flatten_axes(t, 1, TENSOR_AXIS_IN);
	struct ctb_tensor *m0 = NULL, *m1 = NULL;
	int rc = ctb_split_matrix_svd(a_mat, tol, true, max_vdim, renormalize, SVD_DISTR_LEFT, &m0, &m1, info);
	ctb_tensor_free(a_mat);
	if (rc < 0) { return rc; }
	const ct_long dr[2] = { t->ax[1].dim, t->ax[2].dim };
	const int dirr[2] = { TENSOR_AXIS_OUT, TENSOR_AXIS_IN };
	const qnumber* qr[2] = { t->ax[1].qlog, t->ax[2].qlog };
	struct ctb_tensor* a_new = ctb_split_axis(m1, 1, dr, dirr, qr);
	ctb_tensor_free(m1);
	ctb_tensor_free(t);
	*a = a_new;
	struct ctb_tensor* upd = ctb_dot(*a_prev, TENSOR_AXIS_RANGE_TRAILING, 0, m0, TENSOR_AXIS_RANGE_LEADING, 0, 1, NULL);
	ctb_tensor_free(m0);
	ctb_tensor_free(*a_prev);
	*a_prev = upd;
	return 0;
}

int mps_local_orthonormalize_left_svd(const double tol, const ct_long max_vdim, const bool renormalize, struct block_sparse_tensor* a, struct block_sparse_tensor* a_next, struct trunc_info* info)
{
	CTB_CHECK(ctbd_init(-1));
	struct ctb_tensor* ad = ctb_upload(a);
	struct ctb_tensor* nd = ctb_upload(a_next);
	int rc = local_left_svd(tol, max_vdim, renormalize, &ad, &nd, info);
	if (rc < 0) { ctb_tensor_free(ad); ctb_tensor_free(nd); return rc; }
	delete_block_sparse_tensor(a);
	delete_block_sparse_tensor(a_next);
	finish(ad, a);
	finish(nd, a_next);
	return 0;
}

int mps_local_orthonormalize_right_svd(const double tol, const ct_long max_vdim, const bool renormalize, struct block_sparse_tensor* a, struct block_sparse_tensor* a_prev, struct trunc_info* info)
{
	CTB_CHECK(ctbd_init(-1));
	struct ctb_tensor* ad = ctb_upload(a);
	struct ctb_tensor* pd = ctb_upload(a_prev);
	int rc = local_right_svd(tol, max_vdim, renormalize, &ad, &pd, info);
	if (rc < 0) { ctb_tensor_free(ad); ctb_tensor_free(pd); return rc; }
	delete_block_sparse_tensor(a);
	delete_block_sparse_tensor(a_prev);
	finish(ad, a);
	finish(pd, a_prev);
	return 0;
}

/* ---- compression ---- */

/* 1 x 1 x 1 tensor holding a single 1 with the quantum number of the given boundary bond (the "tail"/"head" of mps.c:889-899, :989-999) */
static struct ctb_tensor* boundary_cap(int dtype, const struct ctb_axis* bond)
{
	CTB_REQUIRE(bond->dim == 1);
	const qnumber qzero[1] = { 0 };
	const ct_long dim1[3] = { 1, 1, 1 };
	const int dirs[3] = { TENSOR_AXIS_OUT, TENSOR_AXIS_OUT, TENSOR_AXIS_IN };
	const qnumber* qn[3] = { bond->qlog, qzero, bond->qlog };
	struct ctb_tensor* cap = ctb_tensor_create(dtype, 3, dim1, dirs, qn, 1);
	CTB_CHECK_ABORT(ctb_set_entry(cap, 0, 1.0, 0.0));
	return cap;
}

/* QR sweep of mps_orthonormalize_qr (mps.c:609-757) on device-resident site tensors; returns the norm */
static double orthonormalize_qr_device(struct ctb_tensor** A, int L, int mode)
{
	const int dtype = A[0]->dtype;
	struct ctb_tensor* cap = NULL;
	int edge;
	if (mode == MPS_ORTHONORMAL_LEFT) {
		for (int i = 0; i < L - 1; i++) { CTB_CHECK_ABORT(ctb_mps_local_qr(&A[i], &A[i + 1])); }
		edge = L - 1;
		cap = boundary_cap(dtype, &A[edge]->ax[2]);
		CTB_CHECK_ABORT(ctb_mps_local_qr(&A[edge], &cap));
	}
	else {
		for (int i = L - 1; i > 0; i--) { CTB_CHECK_ABORT(ctb_mps_local_rq(&A[i], &A[i - 1])); }
		edge = 0;
		cap = boundary_cap(dtype, &A[0]->ax[0]);
		CTB_CHECK_ABORT(ctb_mps_local_rq(&A[0], &cap));
	}
	double norm = 0;
	if (cap->nstore >= 1) {
		double v[2] = { 0, 0 };
		CTB_CHECK_ABORT(ctbd_d2h(v, cap->d, ctb_sizeof_dtype(dtype)));
		norm = v[0];
		if (norm < 0) {
			CTB_CHECK_ABORT(ctbd_scale_host(dtype, A[edge]->nstore, A[edge]->d, -1.0));
			norm = -norm;
		}
	}
	ctb_tensor_free(cap);
	return norm;
}

int mps_compress(const double tol, const ct_long max_vdim, const enum mps_orthonormalization_mode mode,
	struct mps* mps, double* norm, double* trunc_scale, struct trunc_info* info)
{
	CTB_CHECK(ctbd_init(-1));
	const bool renormalize = false;
	const int L = mps->nsites;
	CTB_REQUIRE(L >= 1);
	struct ctb_tensor** A = upload_sites(mps->a, L);
	const int dtype = A[0]->dtype;
	int rc = 0, edge;
	struct ctb_tensor* cap = NULL;
	if (mode == MPS_ORTHONORMAL_LEFT)
	{
		/* right-canonical form first, then SVD sweep to the right */
		*norm = orthonormalize_qr_device(A, L, MPS_ORTHONORMAL_RIGHT);
		for (int i = 0; i < L - 1 && rc == 0; i++) { rc = local_left_svd(tol, max_vdim, renormalize, &A[i], &A[i + 1], &info[i]); }
		edge = L - 1;
		if (rc == 0) {
			cap = boundary_cap(dtype, &A[edge]->ax[2]);
			rc = local_left_svd(tol, max_vdim, renormalize, &A[edge], &cap, &info[edge]);
		}
	}
	else
	{
		CTB_REQUIRE(mode == MPS_ORTHONORMAL_RIGHT);
		*norm = orthonormalize_qr_device(A, L, MPS_ORTHONORMAL_LEFT);
		for (int i = L - 1; i > 0 && rc == 0; i--) { rc = local_right_svd(tol, max_vdim, renormalize, &A[i], &A[i - 1], &info[i]); }
		edge = 0;
		if (rc == 0) {
			cap = boundary_cap(dtype, &A[0]->ax[0]);
			rc = local_right_svd(tol, max_vdim, renormalize, &A[0], &cap, &info[0]);
		}
	}
	if (rc < 0) { ctb_tensor_free(cap); free_sites(A, L); return rc; }
	/* the scalar left in the cap: its modulus is the truncation scale, its phase goes into the boundary site tensor (mps.c:923-979) */
	{
		double v[2] = { 0, 0 };
		CTB_REQUIRE(cap->nelem == 1);
		if (cap->nstore >= 1) { CTB_CHECK(ctbd_d2h(v, cap->d, ctb_sizeof_dtype(dtype))); }
		if (!ctb_is_complex(dtype)) {
			if (v[0] < 0) { CTB_CHECK(ctbd_scale_host(dtype, A[edge]->nstore, A[edge]->d, -1.0)); }
			*trunc_scale = fabs(v[0]);
		}
		else {
			const double abs_d = hypot(v[0], v[1]);
			if (abs_d != 0) { CTB_CHECK(ctbd_zscale_host(dtype, A[edge]->nstore, A[edge]->d, v[0] / abs_d, v[1] / abs_d)); }
			*trunc_scale = abs_d;
		}
	}
	ctb_tensor_free(cap);
	for (int i = 0; i < L; i++) {
		delete_block_sparse_tensor(&mps->a[i]);
		finish(A[i], &mps->a[i]);
	}
	free(A);
	return 0;
}

int mps_compress_rescale(const double tol, const ct_long max_vdim, const enum mps_orthonormalization_mode mode,
	struct mps* mps, double* trunc_scale, struct trunc_info* info)
{
	double norm = 0;
	int rc = mps_compress(tol, max_vdim, mode, mps, &norm, trunc_scale, info);
	if (rc < 0) { return rc; }
	/* rescale the boundary tensor by the original norm (mps.c:1082-1109); small, done on the host blocks */
	struct block_sparse_tensor* t = &mps->a[mode == MPS_ORTHONORMAL_LEFT ? mps->nsites - 1 : 0];
	ct_long ngrid = 1;
	for (int i = 0; i < t->ndim; i++) { ngrid *= t->dim_blocks[i]; }
	const int nreal = ctb_is_complex(t->dtype) ? 2 : 1;
	for (ct_long k = 0; k < ngrid; k++) {
		struct dense_tensor* blk = t->blocks[k];
		if (blk == NULL) { continue; }
		ct_long numel = 1;
		for (int i = 0; i < blk->ndim; i++) { numel *= blk->dim[i]; }
		double* x = blk->data;
		for (ct_long j = 0; j < numel * nreal; j++) { x[j] *= norm; }
	}
	return 0;
}
