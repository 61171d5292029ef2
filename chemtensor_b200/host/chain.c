/*
 * chain.c -- MPS/MPO chain operations composed from the device primitives.
 *
 * Every function mirrors one reference routine (same contraction order, same
 * sector conventions) but runs on device-resident tensors, with the transposes
 * that separate the reference's contractions folded into the GEMM epilogues:
 *   ctb_heff_*            <- apply_local_hamiltonian            src/algorithm/chain_ops.c:353-390
 *   ctb_env_step_right    <- contraction_operator_step_right    src/algorithm/chain_ops.c:116-165
 *   ctb_env_step_left     <- contraction_operator_step_left     src/algorithm/chain_ops.c:196-245
 *   ctb_dummy_block_*     <- create_dummy_operator_block_*      src/algorithm/chain_ops.c:14-87
 *   ctb_split_matrix_svd  <- split_block_sparse_matrix_svd      src/algorithm/bond_ops.c:15-138
 *   ctb_mps_merge_pair    <- mps_merge_tensor_pair              src/state/mps.c:1166-1178
 *   ctb_mps_split_svd     <- mps_split_tensor_svd               src/state/mps.c:1119-1159
 *   ctb_mps_local_qr/rq   <- mps_local_orthonormalize_qr/rq     src/state/mps.c:513-602
 *   ctb_mpo_merge_pair    <- mpo_merge_tensor_pair              src/operator/mpo.c:255-277
 */
#include "ctb_internal.h"

struct ctb_stats ctb_global_stats;
int ctb_dist_rank = 0, ctb_dist_world = 1;

/* metadata copy that shares the device buffer, with all axis directions reversed */
struct ctb_tensor* ctb_view_reversed_dirs(const struct ctb_tensor* t)
{
	struct ctb_axis axes[CTB_MAXDIM];
	for (int i = 0; i < t->ndim; i++) {
		ctb_axis_copy(&axes[i], &t->ax[i]);
		axes[i].dir = -axes[i].dir;
	}
	struct ctb_tensor* v = ctb_tensor_from_axes(t->dtype, t->ndim, axes, 0);
	CTB_REQUIRE(v->nstore == t->nstore);
	v->d = t->d;
	v->borrowed = 1;
	return v;
}

/* ---------------------------------------------------------------------------------------------- */

int ctb_split_matrix_svd(struct ctb_tensor* a, double tol, bool relative_thresh, ct_long max_vdim, bool renormalize,
	int svd_distr, struct ctb_tensor** a0, struct ctb_tensor** a1, struct trunc_info* info)
{
	CTB_REQUIRE(a->ndim == 2);
	struct ctb_tensor *u = NULL, *vh = NULL;
	double* s_dev = NULL;
	ct_long ns = 0;
	int rc = ctb_svd(a, &u, &s_dev, &ns, &vh);
	if (rc < 0) { return rc; }

	/* the selection rule runs on the device (ctbd_truncate_select); the host receives what its metadata needs -- the ascending list of
	 * retained indices and three scalars -- and the retained (possibly renormalised) values stay on the device for the scaling */
	ct_long* ind_buf = ctb_malloc((size_t)(ns > 0 ? ns : 1) * sizeof(ct_long));
	double* s_ret_dev = NULL;
	CTB_CHECK(ctbd_malloc((void**)&s_ret_dev, (size_t)(ns > 0 ? ns : 1) * sizeof(double)));
	int64_t nret64 = 0;
	double info3[3] = { 0, 0, tol };
	CTB_CHECK(ctbd_truncate_select((int64_t)ns, s_dev, tol, relative_thresh ? 1 : 0, (int64_t)max_vdim, renormalize ? 1 : 0, &nret64, (int64_t*)ind_buf, info3, s_ret_dev));
	info->norm_sigma = info3[0]; info->entropy = info3[1]; info->tol_eff = info3[2];

	ct_long nret = (ct_long)nret64;
	const ct_long ind0[1] = { 0 };
	const ct_long* ind = ind_buf;
	if (nret == 0)
	{
		/* dummy bond of dimension 1 carrying a zero singular value (reference bond_ops.c:50-84); s_ret_dev[0] = 0 */
		nret = 1;
		ind = ind0;
	}
	struct ctb_tensor* us  = ctb_slice(u, 1, ind, nret);
	struct ctb_tensor* vhs = ctb_slice(vh, 0, ind, nret);
	ctb_tensor_free(u);
	ctb_tensor_free(vh);
	CTB_CHECK(ctbd_free(s_dev));

	if (svd_distr == SVD_DISTR_LEFT) {
		*a0 = ctb_scale_axis(us, 1, s_ret_dev);
		ctb_tensor_free(us);
		*a1 = vhs;
	}
	else {
		*a1 = ctb_scale_axis(vhs, 0, s_ret_dev);
		ctb_tensor_free(vhs);
		*a0 = us;
	}
	CTB_CHECK(ctbd_free(s_ret_dev));
	ctb_free(ind_buf);
	return 0;
}

struct ctb_tensor* ctb_mps_merge_pair(const struct ctb_tensor* a0, const struct ctb_tensor* a1)
{
	CTB_REQUIRE(a0->ndim == 3 && a1->ndim == 3);
	struct ctb_tensor* t = ctb_dot(a0, TENSOR_AXIS_RANGE_TRAILING, 0, a1, TENSOR_AXIS_RANGE_LEADING, 0, 1, NULL);
	struct ctb_tensor* a = ctb_flatten_axes(t, 1, TENSOR_AXIS_OUT);
	ctb_tensor_free(t);
	return a;
}

int ctb_mps_split_svd(struct ctb_tensor* a, const ct_long d[2], const qnumber* const new_qsite[2], double tol, ct_long max_vdim,
	bool renormalize, int svd_distr, struct ctb_tensor** a0, struct ctb_tensor** a1, struct trunc_info* info)
{
	CTB_REQUIRE(a->ndim == 3 && d[0] * d[1] == a->ax[1].dim && a->ax[1].dir == TENSOR_AXIS_OUT);
	const int dirs_out[2] = { TENSOR_AXIS_OUT, TENSOR_AXIS_OUT };
	struct ctb_tensor* a_twosite = ctb_split_axis(a, 1, d, dirs_out, new_qsite);
	struct ctb_tensor* tmp = ctb_flatten_axes(a_twosite, 0, TENSOR_AXIS_OUT);
	struct ctb_tensor* a_mat = ctb_flatten_axes(tmp, 1, TENSOR_AXIS_IN);
	ctb_tensor_free(tmp);

	struct ctb_tensor *m0 = NULL, *m1 = NULL;
	int rc = ctb_split_matrix_svd(a_mat, tol, true, max_vdim, renormalize, svd_distr, &m0, &m1, info);
	ctb_tensor_free(a_mat);
	if (rc < 0) { ctb_tensor_free(a_twosite); return rc; }

	{
		const ct_long dl[2] = { a_twosite->ax[0].dim, a_twosite->ax[1].dim };
		const int dirl[2] = { a_twosite->ax[0].dir, a_twosite->ax[1].dir };
		const qnumber* ql[2] = { a_twosite->ax[0].qlog, a_twosite->ax[1].qlog };
		*a0 = ctb_split_axis(m0, 0, dl, dirl, ql);
		const ct_long dr[2] = { a_twosite->ax[2].dim, a_twosite->ax[3].dim };
		const int dirr[2] = { a_twosite->ax[2].dir, a_twosite->ax[3].dir };
		const qnumber* qr[2] = { a_twosite->ax[2].qlog, a_twosite->ax[3].qlog };
		*a1 = ctb_split_axis(m1, 1, dr, dirr, qr);
	}
	ctb_tensor_free(m0);
	ctb_tensor_free(m1);
	ctb_tensor_free(a_twosite);
	return 0;
}

int ctb_mps_local_qr(struct ctb_tensor** a, struct ctb_tensor** a_next)
{
	struct ctb_tensor* t = *a;
	CTB_REQUIRE(t->ndim == 3 && (*a_next)->ndim == 3);
	CTB_REQUIRE(t->ax[0].dir == TENSOR_AXIS_OUT && t->ax[1].dir == TENSOR_AXIS_OUT);
	struct ctb_tensor* a_mat = ctb_flatten_axes(t, 0, TENSOR_AXIS_OUT);
	struct ctb_tensor *q = NULL, *r = NULL;
	int rc = ctb_qr(a_mat, &q, &r);
	ctb_tensor_free(a_mat);
	if (rc < 0) { return rc; }
	const ct_long dl[2] = { t->ax[0].dim, t->ax[1].dim };
	const int dirl[2] = { TENSOR_AXIS_OUT, TENSOR_AXIS_OUT };
	const qnumber* ql[2] = { t->ax[0].qlog, t->ax[1].qlog };
	struct ctb_tensor* a_new = ctb_split_axis(q, 0, dl, dirl, ql);
	ctb_tensor_free(q);
	ctb_tensor_free(t);
	*a = a_new;
	struct ctb_tensor* upd = ctb_dot(r, TENSOR_AXIS_RANGE_TRAILING, 0, *a_next, TENSOR_AXIS_RANGE_LEADING, 0, 1, NULL);
	ctb_tensor_free(r);
	ctb_tensor_free(*a_next);
	*a_next = upd;
	return 0;
}

int ctb_mps_local_rq(struct ctb_tensor** a, struct ctb_tensor** a_prev)
{
	struct ctb_tensor* t = *a;
	CTB_REQUIRE(t->ndim == 3 && (*a_prev)->ndim == 3);
	CTB_REQUIRE(t->ax[1].dir == TENSOR_AXIS_OUT && t->ax[2].dir == TENSOR_AXIS_IN);
	struct ctb_tensor* a_mat = ctb_flatten_axes(t, 1, TENSOR_AXIS_IN);
	struct ctb_tensor *q = NULL, *r = NULL;
	int rc = ctb_rq(a_mat, &r, &q);
	ctb_tensor_free(a_mat);
	if (rc < 0) { return rc; }
	const ct_long dr[2] = { t->ax[1].dim, t->ax[2].dim };
	const int dirr[2] = { TENSOR_AXIS_OUT, TENSOR_AXIS_IN };
	const qnumber* qr[2] = { t->ax[1].qlog, t->ax[2].qlog };
	struct ctb_tensor* a_new = ctb_split_axis(q, 1, dr, dirr, qr);
	ctb_tensor_free(q);
	ctb_tensor_free(t);
	*a = a_new;
	struct ctb_tensor* upd = ctb_dot(*a_prev, TENSOR_AXIS_RANGE_TRAILING, 0, r, TENSOR_AXIS_RANGE_LEADING, 0, 1, NULL);
	ctb_tensor_free(r);
	ctb_tensor_free(*a_prev);
	*a_prev = upd;
	return 0;
}

struct ctb_tensor* ctb_mpo_merge_pair(const struct ctb_tensor* w0, const struct ctb_tensor* w1)
{
	CTB_REQUIRE(w0->ndim == 4 && w1->ndim == 4);
	/* dot + transpose [0,1,3,2,4,5] in one launch, then fuse the physical legs */
	const int perm[6] = { 0, 1, 3, 2, 4, 5 };
	struct ctb_tensor* t = ctb_dot(w0, TENSOR_AXIS_RANGE_TRAILING, 0, w1, TENSOR_AXIS_RANGE_LEADING, 0, 1, perm);
	struct ctb_tensor* tmp = ctb_flatten_axes(t, 1, TENSOR_AXIS_OUT);
	ctb_tensor_free(t);
	struct ctb_tensor* w = ctb_flatten_axes(tmp, 2, TENSOR_AXIS_IN);
	ctb_tensor_free(tmp);
	return w;
}

/* 1x1x1x1 identity environments with the boundary quantum numbers of the 6-leg construction */
static struct ctb_tensor* dummy_block_typed(int dtype, const qnumber* qa, const qnumber* qw, const qnumber* qb, int left)
{
	const ct_long dim[6] = { 1, 1, 1, 1, 1, 1 };
	const int dirs[6] = { TENSOR_AXIS_OUT, TENSOR_AXIS_OUT, TENSOR_AXIS_IN, TENSOR_AXIS_IN, TENSOR_AXIS_IN, TENSOR_AXIS_OUT };
	const qnumber* qn[6] = { qa, qw, qb, qa, qw, qb };
	struct ctb_tensor* s = ctb_tensor_create(dtype, 6, dim, dirs, qn, 1);
	CTB_REQUIRE(s->nblk == 1);
	CTB_CHECK_ABORT(ctb_set_entry(s, 0, 1.0, 0.0));
	struct ctb_tensor *t, *r;
	if (left) {
		t = ctb_flatten_axes(s, 0, TENSOR_AXIS_OUT);
		r = ctb_flatten_axes(t, 0, TENSOR_AXIS_OUT);
	}
	else {
		t = ctb_flatten_axes(s, 4, TENSOR_AXIS_IN);
		r = ctb_flatten_axes(t, 3, TENSOR_AXIS_IN);
	}
	ctb_tensor_free(s);
	ctb_tensor_free(t);
	return r;
}

struct ctb_tensor* ctb_dummy_block_right(const struct ctb_tensor* a, const struct ctb_tensor* b, const struct ctb_tensor* w)
{
	CTB_REQUIRE(a->ndim == 3 && b->ndim == 3 && w->ndim == 4);
	CTB_REQUIRE(a->ax[2].dim == 1 && b->ax[2].dim == 1 && w->ax[3].dim == 1);
	return dummy_block_typed(a->dtype, a->ax[2].qlog, w->ax[3].qlog, b->ax[2].qlog, 0);
}

struct ctb_tensor* ctb_dummy_block_left(const struct ctb_tensor* a, const struct ctb_tensor* b, const struct ctb_tensor* w)
{
	CTB_REQUIRE(a->ndim == 3 && b->ndim == 3 && w->ndim == 4);
	CTB_REQUIRE(a->ax[0].dim == 1 && b->ax[0].dim == 1 && w->ax[0].dim == 1);
	return dummy_block_typed(a->dtype, a->ax[0].qlog, w->ax[0].qlog, b->ax[0].qlog, 1);
}

struct ctb_tensor* ctb_env_step_right(const struct ctb_tensor* a, const struct ctb_tensor* b, const struct ctb_tensor* w, const struct ctb_tensor* r)
{
	CTB_REQUIRE(a->ndim == 3 && b->ndim == 3 && w->ndim == 4 && r->ndim == 4);
	/* a . r, stored as [d, Dw', Dl, Dr', x] */
	const int perm0[5] = { 1, 2, 0, 3, 4 };
	struct ctb_dot_plan pl;
	struct ctb_tensor* s = ctb_dot_prepare(a, TENSOR_AXIS_RANGE_TRAILING, 0, r, TENSOR_AXIS_RANGE_LEADING, 0, 1, perm0, 1, &pl);
	CTB_CHECK_ABORT(ctb_dot_exec(&pl, a->d, r->d, s->d));
	ctb_global_stats.env_flops += pl.flops;
	ctb_dot_plan_free(&pl);
	/* w . (a r), stored as [x, Dl, Dw, d_out, Dr'] */
	const int perm1[5] = { 4, 2, 0, 1, 3 };
	struct ctb_tensor* t = ctb_dot_prepare_ex(w, TENSOR_AXIS_RANGE_TRAILING, 0, s, TENSOR_AXIS_RANGE_LEADING, 0, 2, perm1, 1, CTB_DOT_MERGE_ROWS, &pl);
	CTB_CHECK_ABORT(ctb_dot_exec(&pl, w->d, s->d, t->d));
	ctb_global_stats.env_flops += pl.flops;
	ctb_dot_plan_free(&pl);
	ctb_tensor_free(s);
	/* contract with conj(b) over (d_out, Dr'); conjugation fused into the operand load; result [Dl, Dw, Dl', x] */
	struct ctb_tensor* br = ctb_view_reversed_dirs(b);
	const int perm2[4] = { 1, 2, 3, 0 };
	struct ctb_tensor* r_next = ctb_dot_prepare(t, TENSOR_AXIS_RANGE_TRAILING, 0, br, TENSOR_AXIS_RANGE_TRAILING, ctb_is_complex(b->dtype), 2, perm2, 1, &pl);
	CTB_CHECK_ABORT(ctb_dot_exec(&pl, t->d, br->d, r_next->d));
	ctb_global_stats.env_flops += pl.flops;
	ctb_dot_plan_free(&pl);
	ctb_tensor_free(br);
	ctb_tensor_free(t);
	return r_next;
}

struct ctb_tensor* ctb_env_step_left(const struct ctb_tensor* a, const struct ctb_tensor* b, const struct ctb_tensor* w, const struct ctb_tensor* l)
{
	CTB_REQUIRE(a->ndim == 3 && b->ndim == 3 && w->ndim == 4 && l->ndim == 4);
	struct ctb_dot_plan pl;
	/* l . conj(b), stored as [x, Dl, Dr', Dw, d'] */
	struct ctb_tensor* br = ctb_view_reversed_dirs(b);
	const int perm0[5] = { 0, 1, 4, 2, 3 };
	struct ctb_tensor* s = ctb_dot_prepare(l, TENSOR_AXIS_RANGE_TRAILING, 0, br, TENSOR_AXIS_RANGE_LEADING, ctb_is_complex(b->dtype), 1, perm0, 1, &pl);
	CTB_CHECK_ABORT(ctb_dot_exec(&pl, l->d, br->d, s->d));
	ctb_global_stats.env_flops += pl.flops;
	ctb_dot_plan_free(&pl);
	ctb_tensor_free(br);
	/* (l b*) . w over (Dw, d'), stored as [Dl, d_in, Dw', Dr', x] */
	const int perm1[5] = { 1, 3, 4, 2, 0 };
	struct ctb_tensor* t = ctb_dot_prepare(s, TENSOR_AXIS_RANGE_TRAILING, 0, w, TENSOR_AXIS_RANGE_LEADING, 0, 2, perm1, 1, &pl);
	CTB_CHECK_ABORT(ctb_dot_exec(&pl, s->d, w->d, t->d));
	ctb_global_stats.env_flops += pl.flops;
	ctb_dot_plan_free(&pl);
	ctb_tensor_free(s);
	/* a . (...) over (Dl, d_in), stored as [x, Dr, Dw', Dr'] */
	const int perm2[4] = { 3, 0, 1, 2 };
	struct ctb_tensor* l_next = ctb_dot_prepare(a, TENSOR_AXIS_RANGE_LEADING, 0, t, TENSOR_AXIS_RANGE_LEADING, 0, 2, perm2, 1, &pl);
	CTB_CHECK_ABORT(ctb_dot_exec(&pl, a->d, t->d, l_next->d));
	ctb_global_stats.env_flops += pl.flops;
	ctb_dot_plan_free(&pl);
	ctb_tensor_free(t);
	return l_next;
}

/* ---------------------------------------------------------------------------------------------- */
/* effective Hamiltonian: plans built once per bond, three launches per matvec                     */
/* ---------------------------------------------------------------------------------------------- */

/* ---- peer-mapped landing buffers of the fused exchange (two, used alternately; see ctb_heff_finish) ---- */
static void* g_land[2] = { NULL, NULL };
static void** g_land_ptrs[2] = { NULL, NULL };
static void* g_land_mc[2] = { NULL, NULL };      /* != NULL: the landing buffers are bound to an NVSwitch multicast object; the address all ranks store to */
static size_t g_land_bytes = 0;
static int g_land_flip = 0, g_land_failed = 0, g_mc_failed = 0;
static long long g_fused_count = 0, g_allgather_count = 0, g_mc_count = 0;

static void landing_release(void)
{
	for (int i = 0; i < 2; i++) {
		if (g_land[i] != NULL) { if (g_land_mc[i] != NULL) { ctbd_mc_buffer_destroy(g_land[i]); } else { ctbd_peer_buffer_destroy(g_land[i]); } g_land[i] = NULL; }
		g_land_mc[i] = NULL;
		free(g_land_ptrs[i]); g_land_ptrs[i] = NULL;
	}
	g_land_bytes = 0; g_land_flip = 0;
}

/* collective: every rank calls it with the same size; returns 1 when both landing buffers are available.  With multicast != 0 the
 * buffers are first tried as NVSwitch multicast buffers (one store per element reaches all ranks), else (or when the box has no
 * multicast) as CUDA-IPC peer-mapped buffers every rank stores into separately. */
static int landing_ensure(size_t bytes, int world, int multicast)
{
	if (g_land_failed || getenv("CTB_NO_FUSED_EXCHANGE") != NULL) { return 0; }
	const int want_mc = multicast && !g_mc_failed && getenv("CTB_NO_MULTICAST") == NULL;
	if (g_land[0] != NULL && bytes <= g_land_bytes && (g_land_mc[0] != NULL) == (want_mc != 0)) { return 1; }
	landing_release();
	const size_t want = bytes + bytes / 4 + 4096;
	if (want_mc)
	{
		int ok = 1;
		for (int i = 0; i < 2 && ok; i++) {
			void* local = NULL;
			if (ctbd_mc_buffer_create(want, &g_land[i], &local, &g_land_mc[i]) < 0) { ok = 0; break; }
			g_land_ptrs[i] = calloc((size_t)world, sizeof(void*));
			g_land_ptrs[i][ctb_dist_rank] = local;
		}
		if (ok) { g_land_bytes = want; return 1; }
		landing_release();
		g_mc_failed = 1;      /* the same decision on every rank: the creation is collective and agrees on failure */
	}
	for (int i = 0; i < 2; i++) {
		if (ctbd_peer_buffer_create(want, &g_land[i]) < 0) { g_land_failed = 1; landing_release(); return 0; }
		g_land_ptrs[i] = calloc((size_t)world, sizeof(void*));
		ctbd_peer_buffer_ptrs(g_land[i], g_land_ptrs[i]);
	}
	g_land_bytes = want;
	return 1;
}

/* ---- peer-mapped send buffers of the pull exchange (two, used alternately): every rank writes its result slice locally, and after
 * the barrier one kernel per rank reads the slices of all ranks over NVLink straight into the packed result ---- */
static void* g_send[2] = { NULL, NULL };
static void** g_send_ptrs[2] = { NULL, NULL };
static size_t g_send_bytes = 0;
static int g_send_flip = 0, g_send_failed = 0;
static long long g_pull_count = 0, g_push_count = 0;

static void send_release(void)
{
	for (int i = 0; i < 2; i++) {
		if (g_send[i] != NULL) { ctbd_peer_buffer_destroy(g_send[i]); g_send[i] = NULL; }
		free(g_send_ptrs[i]); g_send_ptrs[i] = NULL;
	}
	g_send_bytes = 0; g_send_flip = 0;
}

/* collective: every rank calls it with the same size; returns 1 when both send buffers are available */
static int send_ensure(size_t bytes, int world)
{
	if (g_send_failed) { return 0; }
	if (g_send[0] != NULL && bytes <= g_send_bytes) { return 1; }
	send_release();
	const size_t want = bytes + bytes / 4 + 4096;
	for (int i = 0; i < 2; i++) {
		if (ctbd_peer_buffer_create(want, &g_send[i]) < 0) { g_send_failed = 1; send_release(); return 0; }
		g_send_ptrs[i] = calloc((size_t)world, sizeof(void*));
		ctbd_peer_buffer_ptrs(g_send[i], g_send_ptrs[i]);
	}
	g_send_bytes = want;
	return 1;
}

/* exchange form of the sharded effective Hamiltonian: CTB_EXCHANGE = fused | pull | allgather (default: see exchange_mode) */
enum { EXCH_ALLGATHER = 0, EXCH_FUSED = 1, EXCH_PULL = 2, EXCH_PUSH = 3 };
static int exchange_mode(int world)
{
	const char* env = getenv("CTB_EXCHANGE");
	if (env != NULL) {
		if (strcmp(env, "fused") == 0) { return EXCH_FUSED; }
		if (strcmp(env, "pull") == 0) { return EXCH_PULL; }
		if (strcmp(env, "push") == 0) { return EXCH_PUSH; }
		if (strcmp(env, "allgather") == 0) { return EXCH_ALLGATHER; }
	}
	if (getenv("CTB_NO_FUSED_EXCHANGE") != NULL) { return EXCH_ALLGATHER; }
	(void)world;
	return EXCH_FUSED;
}

void ctb_dist_release_buffers(void) { landing_release(); g_land_failed = 0; g_mc_failed = 0; send_release(); g_send_failed = 0; }
long long ctb_dist_multicast_count(void) { return g_mc_count; }
long long ctb_dist_pull_count(void) { return g_pull_count; }
long long ctb_dist_push_count(void) { return g_push_count; }
void ctb_dist_counters(long long* fused, long long* allgather) { *fused = g_fused_count; *allgather = g_allgather_count; }

/* Split of the bra bond of r among 'world' ranks.  Every sector is cut into ceil(m / grain) nearly equal contiguous chunks
 * (grain = one GEMM tile width, so a rank's blocks keep full tiles), each chunk is weighted by the step-1 flops of its
 * columns (computed from the block structures of a and r), and the chunks are handed out longest-processing-time first.
 * Deterministic, identical on every rank.  Returns 0 if some rank would stay empty (bond too small to shard). */
struct shard_chunk { int sec; ct_long pos0, len; double cost; };
static int cmp_chunk_desc(const void* x, const void* y)
{
	const struct shard_chunk* a = x; const struct shard_chunk* b = y;
	if (a->cost != b->cost) { return a->cost > b->cost ? -1 : 1; }
	if (a->sec != b->sec) { return a->sec < b->sec ? -1 : 1; }
	return (a->pos0 > b->pos0) - (a->pos0 < b->pos0);
}

static int split_bond(const struct ctb_tensor* a, const struct ctb_tensor* l, const struct ctb_tensor* r, int world, ct_long** ind, ct_long* nind)
{
	const struct ctb_axis* ax = &r->ax[2];
	ct_long grain = (world >= 8) ? 64 : 128;      /* one or two 64-column GEMM tiles */
	const char* env = getenv("CTB_SHARD_GRAIN");
	if (env != NULL && atol(env) > 0) { grain = atol(env); }
	/* Padded GEMM work of a chunk of 'len' columns of bra sector s, in the 64 x 64 x 16 tiles of the kernel:
	 *   step 1, every block r[Dr, w', s]:   sum over blocks a[l, sigma, Dr] of pad64(m_l m_sigma) pad16(m_Dr)   times pad64(m_w' len)
	 *   step 3, every block b[l', sigma', s] (structure of a):  pad64(m_l') sum over blocks l[1, l, w, l'] of pad16(m_l m_w)   times pad64(m_sigma' len)
	 * so per sector the weights are collected by the multiplier (m_w' or m_sigma') of the column count. */
	enum { MULT_MAX = 64 };
#define PAD_UP(x, g) ((((x) + (g) - 1) / (g)) * (g))
	const int kb = a->ndim - 1;      /* ket bond of a: [Dl, d, Dr] or, with the physical legs kept apart, [Dl, d1, d2, Dr] */
	double* rows_a = ctb_calloc((size_t)a->ax[kb].nsec, sizeof(double));
	for (int b = 0; b < a->nblk; b++) {
		int idx[CTB_MAXDIM];
		ctb_grid_unravel(a, a->blk_grid[b], idx);
		ct_long rows = 1;
		for (int i = 0; i < kb; i++) { rows *= a->ax[i].secdim[idx[i]]; }
		rows_a[idx[kb]] += (double)PAD_UP(rows, 64) * (double)PAD_UP((ct_long)a->ax[kb].secdim[idx[kb]], 16);
	}
	double* wmul = ctb_calloc((size_t)ax->nsec * MULT_MAX, sizeof(double));     /* [sector][multiplier] */
	double* wcol = ctb_calloc((size_t)ax->nsec, sizeof(double));                /* unpadded work per column (small-bond case) */
	for (int b = 0; b < r->nblk; b++) {
		int idx[CTB_MAXDIM];
		ctb_grid_unravel(r, r->blk_grid[b], idx);
		const int sa = ctb_axis_find_sector(&a->ax[kb], r->ax[0].qsec[idx[0]]);
		if (sa < 0) { continue; }
		int mult = r->ax[1].secdim[idx[1]];
		wcol[idx[2]] += rows_a[sa] * mult;
		if (mult >= MULT_MAX) { mult = MULT_MAX - 1; }
		wmul[(size_t)idx[2] * MULT_MAX + mult] += rows_a[sa];
	}
	{
		double* kl = ctb_calloc((size_t)l->ax[3].nsec, sizeof(double));
		for (int b = 0; b < l->nblk; b++) {
			int idx[CTB_MAXDIM];
			ctb_grid_unravel(l, l->blk_grid[b], idx);
			kl[idx[3]] += (double)PAD_UP((ct_long)l->ax[1].secdim[idx[1]] * l->ax[2].secdim[idx[2]], 16);
		}
		for (int b = 0; b < a->nblk; b++) {
			int idx[CTB_MAXDIM];
			ctb_grid_unravel(a, a->blk_grid[b], idx);
			/* the result has the structure of a; its first leg carries the quantum numbers of the bra leg of l */
			const int sl = ctb_axis_find_sector(&l->ax[3], a->ax[0].qsec[idx[0]]);
			if (sl < 0) { continue; }
			const int sr = ctb_axis_find_sector(ax, a->ax[kb].qsec[idx[kb]]);
			if (sr < 0) { continue; }
			int mult = 1;
			for (int i = 1; i < kb; i++) { mult *= a->ax[i].secdim[idx[i]]; }
			const double wgt = (double)PAD_UP((ct_long)a->ax[0].secdim[idx[0]], 64) * kl[sl];
			wcol[sr] += wgt * mult;
			if (mult >= MULT_MAX) { mult = MULT_MAX - 1; }
			wmul[(size_t)sr * MULT_MAX + mult] += wgt;
		}
		ctb_free(kl);
	}
	/* the balanced chunks */
	size_t nch = 0, cap = 64;
	struct shard_chunk* ch = malloc(cap * sizeof(*ch));
	for (int s = 0; s < ax->nsec; s++) {
		const ct_long m = ax->secdim[s];
		ct_long parts = (m + grain - 1) / grain;
		/* never fewer chunks in total than ranks: small bonds are cut finer */
		const bool fine = (ax->dim < grain * world);
		if (fine) { parts = (m * world + ax->dim - 1) / ax->dim; if (parts > m) { parts = m; } if (parts < 1) { parts = 1; } }
		for (ct_long c = 0; c < parts; c++) {
			/* chunk boundaries on multiples of the grain: whatever set of chunks a rank receives, its blocks are whole tiles plus at
			 * most one remainder per sector */
			const ct_long p0 = fine ? (m * c) / parts : c * grain;
			const ct_long p1 = fine ? (m * (c + 1)) / parts : ((c + 1) * grain < m ? (c + 1) * grain : m);
			if (p1 <= p0) { continue; }
			if (nch == cap) { cap *= 2; ch = realloc(ch, cap * sizeof(*ch)); }
			ch[nch].sec = s; ch[nch].pos0 = p0; ch[nch].len = p1 - p0;
			if (fine) { ch[nch].cost = (double)(p1 - p0) * (wcol[s] > 0 ? wcol[s] : 1.0); }
			else {
				double cst = 0;
				for (int mu = 1; mu < MULT_MAX; mu++) { cst += wmul[(size_t)s * MULT_MAX + mu] * (double)PAD_UP((ct_long)mu * (p1 - p0), 64); }
				ch[nch].cost = cst > 0 ? cst : (double)(p1 - p0);
			}
			nch++;
		}
	}
	qsort(ch, nch, sizeof(*ch), cmp_chunk_desc);
	double* load = ctb_calloc((size_t)world, sizeof(double));
	unsigned char* owner = ctb_malloc((size_t)(ax->dim > 0 ? ax->dim : 1));
	for (size_t c = 0; c < nch; c++) {
		int best = 0;
		for (int p = 1; p < world; p++) { if (load[p] < load[best]) { best = p; } }
		load[best] += ch[c].cost;
		for (ct_long j = 0; j < ch[c].len; j++) { owner[ax->log_of[ax->secstart[ch[c].sec] + ch[c].pos0 + j]] = (unsigned char)best; }
	}
	for (int p = 0; p < world; p++) { ind[p] = ctb_malloc((size_t)(ax->dim > 0 ? ax->dim : 1) * sizeof(ct_long)); nind[p] = 0; }
	for (ct_long i = 0; i < ax->dim; i++) { const int p = owner[i]; ind[p][nind[p]++] = i; }
	ctb_free(owner); ctb_free(load); ctb_free(wcol); ctb_free(wmul); ctb_free(rows_a); free(ch);
#undef PAD_UP
	for (int p = 0; p < world; p++) { if (nind[p] == 0) { return 0; } }
	return 1;
}

int ctb_heff_prepare(const struct ctb_tensor* a, const struct ctb_tensor* w, struct ctb_tensor* l, const struct ctb_tensor* r, struct ctb_heff* h)
{
	return ctb_heff_prepare_ex(a, w, l, r, h, NULL, NULL);
}

static int heff_prepare_any(const struct ctb_tensor* a, const struct ctb_tensor* w, const struct ctb_tensor* w_second, struct ctb_tensor* l, const struct ctb_tensor* r,
	struct ctb_heff* h, void (*l_ready)(void*), void* ctx);

int ctb_heff_prepare_ex(const struct ctb_tensor* a, const struct ctb_tensor* w, struct ctb_tensor* l, const struct ctb_tensor* r, struct ctb_heff* h,
	void (*l_ready)(void*), void* ctx)
{
	CTB_REQUIRE(a->ndim == 3);
	return heff_prepare_any(a, w, NULL, l, r, h, l_ready, ctx);
}

/* The two single-site MPO tensors applied one after the other instead of the merged pair tensor: `a` is the
 * two-site tensor with its physical legs kept apart, [Dl, d1, d2, Dr], and so is the result.  The merged tensor of the reference
 * (mpo_merge_tensor_pair, src/operator/mpo.c:255) grows with Dw^2 d^4 and is what makes large-bond MPOs infeasible there; the
 * sequential form needs the two site tensors only and costs Dw d^2 (Dw' d) per (Dl, Dr') column instead of (Dw d^2)(d^2 Dw''). */
int ctb_heff_prepare_pair(const struct ctb_tensor* a4, const struct ctb_tensor* w0, const struct ctb_tensor* w1, struct ctb_tensor* l, const struct ctb_tensor* r, struct ctb_heff* h)
{
	CTB_REQUIRE(a4->ndim == 4 && w1 != NULL);
	return heff_prepare_any(a4, w0, w1, l, r, h, NULL, NULL);
}

static int heff_prepare_any(const struct ctb_tensor* a, const struct ctb_tensor* w, const struct ctb_tensor* w_second, struct ctb_tensor* l, const struct ctb_tensor* r,
	struct ctb_heff* h, void (*l_ready)(void*), void* ctx)
{
	const bool pair = (w_second != NULL);
	const int kb = a->ndim - 1;      /* ket bond of a */
	CTB_REQUIRE(a->ndim == (pair ? 4 : 3) && w->ndim == 4 && l->ndim == 4 && r->ndim == 4);
	memset(h, 0, sizeof(*h));
	h->w = w; h->w_second = w_second; h->r = r;
	h->world = 1; h->rank = 0;
	if (ctb_dist_world > 1)
	{
		/* the column slice of r is cut on the device right away: no overlap with in-flight payloads in the sharded case */
		if (l_ready != NULL) { l_ready(ctx); l_ready = NULL; }
		const int W = ctb_dist_world;
		CTB_REQUIRE(W <= 8);
		h->ind = ctb_calloc((size_t)W, sizeof(ct_long*));
		h->nind = ctb_calloc((size_t)W, sizeof(ct_long));
		if (split_bond(a, l, r, W, h->ind, h->nind))
		{
			h->world = W; h->rank = ctb_dist_rank;
			h->r_own = ctb_slice((struct ctb_tensor*)r, 2, h->ind[h->rank], h->nind[h->rank]);
			h->r = h->r_own;
			h->piece = ctb_calloc((size_t)W, sizeof(struct ctb_tensor*));
			for (int p = 0; p < W; p++) {
				qnumber* q = ctb_malloc((size_t)h->nind[p] * sizeof(qnumber));
				for (ct_long j = 0; j < h->nind[p]; j++) { q[j] = a->ax[kb].qlog[h->ind[p][j]]; }
				struct ctb_axis axes[4];
				for (int i = 0; i < kb; i++) { ctb_axis_copy(&axes[i], &a->ax[i]); }
				ctb_axis_init(&axes[kb], h->nind[p], a->ax[kb].dir, q);
				ctb_free(q);
				h->piece[p] = ctb_tensor_from_axes(a->dtype, a->ndim, axes, 0);
				if (h->piece[p]->nstore > h->piece_cap) { h->piece_cap = h->piece[p]->nstore; }
			}
			if (h->piece_cap == 0) { h->piece_cap = 1; }
			const size_t esize = ctb_sizeof_dtype(a->dtype);
			CTB_CHECK(ctbd_malloc(&h->send, (size_t)h->piece_cap * esize));
			CTB_CHECK(ctbd_malloc(&h->recv, (size_t)h->piece_cap * esize * (size_t)W));
			/* scatter plan: every stored block of every piece is a set of strided 2-d copies into the matching block of b, one per
			 * maximal run of consecutive sector positions owned by that rank */
			{
				size_t cap = 1024, nd = 0;
				struct ctbd_copy2d* descs = malloc(cap * sizeof(*descs));
				size_t nd_own0 = 0, nd_own1 = 0;      /* descriptors of this rank's own piece */
				for (int p = 0; p < W; p++)
				{
					const struct ctb_tensor* pc = h->piece[p];
					if (p == h->rank) { nd_own0 = nd; }
					for (int blk = 0; blk < pc->nblk; blk++)
					{
						int idx[CTB_MAXDIM];
						ctb_grid_unravel(pc, pc->blk_grid[blk], idx);
						const int sfull = ctb_axis_find_sector(&a->ax[kb], pc->ax[kb].qsec[idx[kb]]);
						CTB_REQUIRE(sfull >= 0);
						int ifull[CTB_MAXDIM];
						int rows = 1;
						for (int i = 0; i < kb; i++) { ifull[i] = idx[i]; rows *= pc->ax[i].secdim[idx[i]]; }
						ifull[kb] = sfull;
						const ct_long dst_blk = ctb_grid_offset(a, ctb_grid_ravel(a, ifull));
						CTB_REQUIRE(dst_blk >= 0);
						const int np = pc->ax[kb].secdim[idx[kb]], nfull = a->ax[kb].secdim[sfull];
						int j = 0;
						while (j < np)
						{
							/* position inside the full sector of the j-th piece column of this sector */
							const ct_long lg0 = h->ind[p][pc->ax[kb].log_of[pc->ax[kb].secstart[idx[kb]] + j]];
							const int pos0 = a->ax[kb].pos_of[lg0];
							int len = 1;
							while (j + len < np) {
								const ct_long lg = h->ind[p][pc->ax[kb].log_of[pc->ax[kb].secstart[idx[kb]] + j + len]];
								if (a->ax[kb].pos_of[lg] != pos0 + len) { break; }
								len++;
							}
							if (nd == cap) { cap *= 2; descs = realloc(descs, cap * sizeof(*descs)); }
							struct ctbd_copy2d* d = &descs[nd++];
							d->src_off = (int64_t)p * h->piece_cap + pc->blk_off[blk] + j;
							d->dst_off = dst_blk + pos0;
							d->rows = rows; d->cols = len; d->src_ld = np; d->dst_ld = nfull;
							j += len;
						}
					}
					if (p == h->rank) { nd_own1 = nd; }
				}
				CTB_CHECK(ctbd_copy_plan_create(a->dtype, (int)nd, descs, &h->scatter));
				/* push form of the exchange: the own piece, read from its local buffer (offsets relative to the piece) */
				for (size_t q = nd_own0; q < nd_own1; q++) { descs[q].src_off -= (int64_t)h->rank * h->piece_cap; }
				CTB_CHECK(ctbd_copy_plan_create(a->dtype, (int)(nd_own1 - nd_own0), descs + nd_own0, &h->push_plan));
				free(descs);
			}
		}
		else
		{
			/* a bond too small to split (chain ends): every rank computes the whole matvec, no exchange */
			for (int p = 0; p < W; p++) { ctb_free(h->ind[p]); }
			ctb_free(h->ind); ctb_free(h->nind); h->ind = NULL; h->nind = NULL;
		}
	}
	const struct ctb_tensor* ru = h->r;
	const int trace = getenv("CTB_TRACE") != NULL;
	const double tp0 = ctb_wall_ms();
	if (!pair)
	{
		/* step 1: a . r  -> t1 [dd, Dw', Dl, Dr', x'] */
		const int perm0[5] = { 1, 2, 0, 3, 4 };
		h->t1 = ctb_dot_prepare(a, TENSOR_AXIS_RANGE_TRAILING, 0, ru, TENSOR_AXIS_RANGE_LEADING, 0, 1, perm0, 1, &h->p1);
	}
	else
	{
		/* step 1: a . r  -> t1 [d2, Dw'', Dl, d1, Dr', x'] */
		const int perm0[6] = { 2, 3, 0, 1, 4, 5 };
		h->t1 = ctb_dot_prepare(a, TENSOR_AXIS_RANGE_TRAILING, 0, ru, TENSOR_AXIS_RANGE_LEADING, 0, 1, perm0, 1, &h->p1);
	}
	const double tp1 = ctb_wall_ms();
	if (!pair)
	{
		/* step 2: w . t1 over (dd_in, Dw') -> t2 [Dl, Dw, dd_out, Dr', x'] */
		const int perm1[5] = { 2, 0, 1, 3, 4 };
		h->t2 = ctb_dot_prepare_ex(w, TENSOR_AXIS_RANGE_TRAILING, 0, h->t1, TENSOR_AXIS_RANGE_LEADING, 0, 2, perm1, 1, CTB_DOT_MERGE_ROWS, &h->p2);
	}
	else
	{
		/* step 2a: w(site i+1) . t1 over (d2_in, Dw'') -> tm [d1, Dw', Dl, d2_out, Dr', x'] */
		const int perm1a[6] = { 3, 0, 2, 1, 4, 5 };
		h->tm = ctb_dot_prepare_ex(w_second, TENSOR_AXIS_RANGE_TRAILING, 0, h->t1, TENSOR_AXIS_RANGE_LEADING, 0, 2, perm1a, 1, CTB_DOT_MERGE_ROWS, &h->p2a);
		/* step 2b: w(site i) . tm over (d1_in, Dw') -> t2 [Dl, Dw, d1_out, d2_out, Dr', x'] */
		const int perm1b[6] = { 2, 0, 1, 3, 4, 5 };
		h->t2 = ctb_dot_prepare_ex(w, TENSOR_AXIS_RANGE_TRAILING, 0, h->tm, TENSOR_AXIS_RANGE_LEADING, 0, 2, perm1b, 1, CTB_DOT_MERGE_ROWS, &h->p2);
	}
	const double tp2 = ctb_wall_ms();
	/* step 3: k . t2 over (Dl, Dw), k = transpose(l, [0,3,1,2]) once per bond (the reference redoes it every matvec) */
	const int perm2[4] = { 0, 3, 1, 2 };
	if (l_ready != NULL) { l_ready(ctx); }
	h->k = ctb_transpose(l, perm2, 0);
	const double tp3 = ctb_wall_ms();
	struct ctb_tensor* s = NULL;
	const int exch = (h->world > 1) ? exchange_mode(h->world) : EXCH_ALLGATHER;
	if (h->world > 1 && exch == EXCH_PULL && send_ensure((size_t)h->piece_cap * ctb_sizeof_dtype(a->dtype), h->world)) { h->pull = 1; }
	/* push: step 3 writes the slice locally, one copy kernel then stores it into the landing buffers of all ranks (wide NVLink stores) */
	if (h->world > 1 && exch == EXCH_PUSH && landing_ensure((size_t)a->nstore * ctb_sizeof_dtype(a->dtype), h->world, 0)) { h->push = 1; }
	if (h->world > 1 && exch == EXCH_FUSED && landing_ensure((size_t)a->nstore * ctb_sizeof_dtype(a->dtype), h->world, 1))
	{
		/* fused exchange: the step-3 GEMM stores its column slice straight into the packed layout of the FULL result, in the
		 * peer-mapped landing buffer of every rank (NVLink stores from the epilogue); no all-gather, no scatter */
		struct ctb_axis axes[6];
		const int nphys = a->ndim - 2;      /* one fused or two separate physical legs */
		ctb_axis_copy(&axes[0], &h->k->ax[0]);
		ctb_axis_copy(&axes[1], &h->k->ax[1]);
		for (int i = 0; i < nphys; i++) { ctb_axis_copy(&axes[2 + i], &h->t2->ax[2 + i]); }
		ctb_axis_copy(&axes[2 + nphys], &r->ax[2]);
		ctb_axis_copy(&axes[3 + nphys], &h->t2->ax[3 + nphys]);
		h->bfull5 = ctb_tensor_from_axes(a->dtype, 4 + nphys, axes, 0);
		CTB_REQUIRE(h->bfull5->nstore == a->nstore && h->bfull5->nblk == a->nblk);
		const struct ctb_embed emb = { 2 + nphys, h->bfull5, h->ind[h->rank] };
		s = ctb_dot_prepare_embed(h->k, TENSOR_AXIS_RANGE_TRAILING, 0, h->t2, TENSOR_AXIS_RANGE_LEADING, 0, 2, &emb, &h->p3);
		h->fused = 1;
	}
	else {
		s = ctb_dot_prepare(h->k, TENSOR_AXIS_RANGE_TRAILING, 0, h->t2, TENSOR_AXIS_RANGE_LEADING, 0, 2, NULL, 0, &h->p3);
	}
	if (trace) { fprintf(stderr, "ctb_heff_prepare: plan1 %.2f ms, plan2 %.2f ms, transpose(l) %.2f ms, plan3 %.2f ms\n", tp1 - tp0, tp2 - tp1, tp3 - tp2, ctb_wall_ms() - tp3); }
	/* tracing out the two dummy bonds leaves the packed layout unchanged */
	struct ctb_tensor* bs = ctb_drop_dummy_axes(s, 1);
	ctb_tensor_free(s);
	if (h->world > 1) {
		CTB_REQUIRE(ctb_tensor_same_structure(bs, h->piece[h->rank]));
		ctb_tensor_free(bs);
		h->b = ctb_tensor_like(a, 0);
	}
	else {
		h->b = bs;
		CTB_REQUIRE(ctb_tensor_same_structure(h->b, a));
	}
	h->flops = h->p1.flops + h->p2.flops + h->p3.flops + (pair ? h->p2a.flops : 0.0);
	h->flops_total = h->flops * h->world;     /* the shards are balanced by construction; exact totals come from flops of world == 1 */
	h->n = a->nelem;
	h->nstore = a->nstore;
	return 0;
}

int ctb_heff_apply(struct ctb_heff* h, const void* a_data, void* b_data)
{
	CTB_CHECK(ctb_dot_exec(&h->p1, a_data, h->r->d, h->t1->d));
	if (h->w_second != NULL) {
		CTB_CHECK(ctb_dot_exec(&h->p2a, h->w_second->d, h->t1->d, h->tm->d));
		CTB_CHECK(ctb_dot_exec(&h->p2, h->w->d, h->tm->d, h->t2->d));
	}
	else { CTB_CHECK(ctb_dot_exec(&h->p2, h->w->d, h->t1->d, h->t2->d)); }
	CTB_CHECK(ctb_heff_step3(h, b_data));
	CTB_CHECK(ctb_heff_exchange(h, b_data));
	ctb_global_stats.heff_flops += h->flops;
	ctb_global_stats.heff_calls++;
	return 0;
}

/* third contraction; sharded: into the own all-gather slot, or (fused) into the landing buffers of all ranks */
int ctb_heff_step3(struct ctb_heff* h, void* b_data)
{
	if (h->world == 1) { return ctb_dot_exec(&h->p3, h->k->d, h->t2->d, b_data); }
	if (h->pull) { return ctb_dot_exec(&h->p3, h->k->d, h->t2->d, g_send_ptrs[g_send_flip][h->rank]); }
	if (h->push) {
		CTB_CHECK(ctb_dot_exec(&h->p3, h->k->d, h->t2->d, h->send));
		return ctbd_copy_plan_run_push(h->push_plan, h->send, h->world, (void* const*)g_land_ptrs[g_land_flip]);
	}
	if (!h->fused) { return ctb_dot_exec(&h->p3, h->k->d, h->t2->d, h->send); }
	if (g_land_mc[g_land_flip] != NULL) { return ctb_dot_exec_mc(&h->p3, h->k->d, h->t2->d, g_land_mc[g_land_flip]); }
	void* dst[8];
	void** ptrs = g_land_ptrs[g_land_flip];
	dst[0] = ptrs[h->rank];
	int n = 1;
	for (int p = 0; p < h->world; p++) { if (p != h->rank) { dst[n++] = ptrs[p]; } }
	return ctb_dot_exec_multi(&h->p3, h->k->d, h->t2->d, n, dst);
}

/* Where the NEXT application can deliver its result without a copy: the local landing buffer of the fused exchange (valid, and
 * free to be modified in place, until the application after next starts), or NULL when the caller must bring its own buffer. */
void* ctb_heff_result_buffer(const struct ctb_heff* h)
{
	return (h->world > 1 && (h->fused || h->push)) ? g_land_ptrs[g_land_flip][h->rank] : NULL;
}

int ctb_heff_exchange(struct ctb_heff* h, void* b_data)
{
	if (h->world > 1 && (h->fused || h->push))
	{
		/* every rank's slice has been stored into every landing buffer by the step-3 epilogues: wait for all ranks, then hand the
		 * full vector to the caller.  The two landing buffers alternate, so the next application may start writing at once. */
		const size_t esize = ctb_sizeof_dtype(h->b->dtype);
		CTB_CHECK(ctbd_barrier());
		/* a caller that asked for the landing buffer itself (ctb_heff_result_buffer) consumes the result in place */
		if (b_data != g_land_ptrs[g_land_flip][h->rank]) { CTB_CHECK(ctbd_d2d(b_data, g_land_ptrs[g_land_flip][h->rank], (size_t)h->nstore * esize)); }
		g_land_flip ^= 1;
		if (h->fused) { g_fused_count++; if (g_land_mc[g_land_flip ^ 1] != NULL) { g_mc_count++; } } else { g_push_count++; }
		return 0;
	}
	if (h->world > 1 && h->pull)
	{
		/* all slices are complete after the barrier; one kernel reads them from the peer-mapped send buffers of all ranks (NVLink loads,
		 * contiguous row runs) and writes the packed result.  The two send buffers alternate: a rank may run at most one application
		 * ahead of the slowest reader (the next barrier), so the buffer being read is never the one being written. */
		CTB_CHECK(ctbd_barrier());
		CTB_CHECK(ctbd_copy_plan_run_multi(h->scatter, h->world, (const void* const*)g_send_ptrs[g_send_flip], (int64_t)h->piece_cap, b_data));
		g_send_flip ^= 1;
		g_pull_count++;
		return 0;
	}
	if (h->world > 1)
	{
		/* exchange step: all-gather of the result slices over NVLink, then scatter into the packed layout of b */
		const size_t esize = ctb_sizeof_dtype(h->b->dtype);
		CTB_CHECK(ctbd_allgather(h->send, h->recv, (size_t)h->piece_cap * esize));
		CTB_CHECK(ctbd_copy_plan_run(h->scatter, h->recv, b_data));
		g_allgather_count++;
	}
	return 0;
}

void ctb_heff_free(struct ctb_heff* h)
{
	ctb_dot_plan_free(&h->p1); ctb_dot_plan_free(&h->p2); ctb_dot_plan_free(&h->p3); ctb_dot_plan_free(&h->p2a);
	ctb_tensor_free(h->t1); ctb_tensor_free(h->t2); ctb_tensor_free(h->tm); ctb_tensor_free(h->k); ctb_tensor_free(h->b);
	if (h->piece != NULL) { for (int p = 0; p < h->world; p++) { ctb_tensor_free(h->piece[p]); } ctb_free(h->piece); }
	if (h->ind != NULL) { for (int p = 0; p < h->world; p++) { ctb_free(h->ind[p]); } ctb_free(h->ind); ctb_free(h->nind); }
	ctb_tensor_free(h->r_own);
	if (h->scatter != NULL) { ctbd_copy_plan_destroy(h->scatter); }
	if (h->push_plan != NULL) { ctbd_copy_plan_destroy(h->push_plan); }
	ctb_tensor_free(h->bfull5);
	if (h->send != NULL) { ctbd_free(h->send); }
	if (h->recv != NULL) { ctbd_free(h->recv); }
	memset(h, 0, sizeof(*h));
}
