/*
 * su2_chain.c -- SU(2)-symmetric chain operations and DMRG on device-resident tensors (BASELINE configs[4]; SURVEY 8(a) last row).
 *
 * Reference: src/algorithm/su2_chain_ops.c (environment steps :160, :259, local Hamiltonian :496), src/algorithm/su2_dmrg.c
 * (single-site :155, two-site :262), src/state/su2_mps.c (local orthonormalisation :262, :307, split :582),
 * src/algorithm/su2_bond_ops.c:15 and src/algorithm/truncation.c:235 (selection with multiplicities).
 *
 * What is different from the reference, with identical results:
 *  - tensors stay on the device for the whole sweep; the launches of one effective-Hamiltonian application are recorded once per
 *    local solve and replayed for every Lanczos iteration (su2_internal.h);
 *  - the two-site problem is solved in PAIR form on the 4-leg tensor [Dl, d1, d2, Dr]: the two site MPO tensors are applied one
 *    after the other, the merged pair tensor of su2_mpo_merge_tensor_pair (src/operator/su2_mpo.c:341) and the fused physical
 *    axis of su2_mps_merge_tensor_pair (src/state/su2_mps.c:696) are never formed (fusing an axis only concatenates degeneracy
 *    tensors: same vector entries, same Krylov space, same energies);
 *  - QR / RQ / SVD of a site tensor work on the sector blocks directly: the matrix the reference builds with
 *    su2_tensor_fuse_axes_add_auxiliary (+ su2_tensor_reverse_axis_simple for the right-hand forms) is, per bond quantum number,
 *    the stack of the degeneracy tensors that share it, scaled by |reversal coefficient| = sqrt((j_right + 1) / (j_left + 1));
 *    the stacks of all quantum numbers go through ONE batched factorisation launch (ctbd_qr_batched / ctbd_svd_batched).
 */
#include <time.h>
#include <float.h>
#include "su2_internal.h"

static double g_stats[16];     /* [0..4] see ctb_su2_get_stats, [5] orthonormalisation + initial environments, [6] sweeps done, [7..15] seconds per sweep */
static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; }

void ctb_su2_get_stats(double* out16)
{
	g_stats[0] = (double)g_su2_launches;
	memcpy(out16, g_stats, sizeof g_stats);
}

/* ---------------------------------------------------------------------------------------------------------------------------
 * effective Hamiltonian and environment steps as compositions of the structural operations
 * ------------------------------------------------------------------------------------------------------------------------- */

/* su2_apply_local_hamiltonian (src/algorithm/su2_chain_ops.c:496-588) */
static struct su2t* su2_heff_single(struct su2t* a, struct su2t* w, struct su2t* l, struct su2t* r)
{
	CTB_REQUIRE(a->nl == 3 && w->nl == 4 && l->nl == 4 && r->nl == 4);
	struct su2t *s, *t;
	{ const int xa[1] = { 2 }, xr[1] = { 0 }; s = su2t_contract_simple(a, xa, r, xr, 1); }
	t = su2t_fmove(s, su2t_child_axis(s, 1, 0)); su2t_free(s);
	{ const int xw[2] = { 2, 3 }, xt[2] = { 1, 2 }; s = su2t_contract_simple(w, xw, t, xt, 2); } su2t_free(t);
	t = su2t_fmove(s, su2t_child_axis(s, 1, 1)); su2t_free(s);
	su2t_swap_tree_axes(t, 3, 4);
	su2t_reverse_axis_simple(t, 4);
	s = su2t_fmove(t, su2t_child_axis(t, 1, 1)); su2t_free(t);
	struct su2t* k = su2t_clone_meta(l);
	k->varies = 0;
	su2t_reverse_axis_simple(k, 0);
	struct su2t* b;
	{ const int xk[3] = { 0, 1, 2 }, xs[3] = { 4, 2, 0 }; b = su2t_contract_simple(k, xk, s, xs, 3); }
	su2t_free(s); su2t_free(k);
	return b;
}

/* the same for the two-site tensor a2 = [Dl, d1, d2, Dr] with tree (0, (1, 2)) -> 3 (the result of su2_mps_contract_tensor_pair,
 * src/state/su2_mps.c:673), the site operators w0, w1 applied one after the other; the result has the tree of a2 */
static struct su2t* su2_heff_pair(struct su2t* a2, struct su2t* w0, struct su2t* w1, struct su2t* l, struct su2t* r)
{
	CTB_REQUIRE(a2->nl == 4 && a2->na == 0 && w0->nl == 4 && w1->nl == 4 && l->nl == 4 && r->nl == 4);
	struct su2t *s, *t;
	/* ((0, 1), 2) -> 3 */
	t = su2t_fmove(a2, su2t_child_axis(a2, 1, 1));
	{ const int xa[1] = { 3 }, xr[1] = { 0 }; s = su2t_contract_simple(t, xa, r, xr, 1); } su2t_free(t);
	/* axes: 0 l, 1 p1, 2 p2, 3 w', 4 r', 5 outer; split (((0, 1), 2), 3) */
	t = su2t_fmove(s, su2t_child_axis(s, 1, 0)); su2t_free(s);
	{ const int xw[2] = { 2, 3 }, xt[2] = { 2, 3 }; s = su2t_contract_simple(w1, xw, t, xt, 2); } su2t_free(t);
	/* axes: 0 w-mid, 1 p2', 2 l, 3 p1, 4 r', 5 outer; split ((2, 3), (0, 1)) */
	t = su2t_fmove(s, su2t_child_axis(s, 1, 0)); su2t_free(s);          /* (2, (3, (0, 1))) */
	{
		const int q = su2t_child_axis(t, 1, 1);                     /* node (3, (0, 1)) */
		const int g = su2t_parent_axis(t, 1, 0);                    /* node (0, 1) */
		CTB_REQUIRE(g != q);
		s = su2t_fmove(t, g); su2t_free(t);                         /* (2, ((3, 0), 1)) */
	}
	{ const int xw[2] = { 2, 3 }, xs[2] = { 3, 0 }; t = su2t_contract_simple(w0, xw, s, xs, 2); } su2t_free(s);
	/* axes: 0 w-left, 1 p1', 2 p2', 3 l, 4 r', 5 outer; split (3, ((0, 1), 2)) */
	s = su2t_fmove(t, su2t_parent_axis(t, 1, 0)); su2t_free(t);         /* (3, (0, (1, 2))) */
	t = su2t_fmove(s, su2t_child_axis(s, 1, 1)); su2t_free(s);          /* ((3, 0), (1, 2)) */
	su2t_swap_tree_axes(t, 4, 5);
	su2t_reverse_axis_simple(t, 5);                                     /* split (5, ((3, 0), (1, 2))) root 4 */
	s = su2t_fmove(t, su2t_child_axis(t, 1, 1)); su2t_free(t);          /* ((5, (3, 0)), (1, 2)) */
	struct su2t* k = su2t_clone_meta(l);
	k->varies = 0;
	su2t_reverse_axis_simple(k, 0);
	struct su2t* b;
	{ const int xk[3] = { 0, 1, 2 }, xs[3] = { 5, 3, 0 }; b = su2t_contract_simple(k, xk, s, xs, 3); }
	su2t_free(s); su2t_free(k);
	return b;
}

/* su2_contraction_operator_step_right (src/algorithm/su2_chain_ops.c:160-232) */
static struct su2t* su2_step_right(struct su2t* a, struct su2t* b, struct su2t* w, struct su2t* r)
{
	struct su2t *s, *t;
	{ const int xa[1] = { 2 }, xr[1] = { 0 }; s = su2t_contract_simple(a, xa, r, xr, 1); }
	t = su2t_fmove(s, su2t_child_axis(s, 1, 0)); su2t_free(s);
	{ const int xw[2] = { 2, 3 }, xt[2] = { 1, 2 }; s = su2t_contract_simple(w, xw, t, xt, 2); } su2t_free(t);
	t = su2t_fmove(s, su2t_child_axis(s, 1, 1)); su2t_free(s);
	su2t_reverse_axis_simple(t, 1);
	su2t_swap_tree_axes(t, 3, 4);
	s = su2t_fmove(t, su2t_child_axis(t, 0, 0)); su2t_free(t);
	struct su2t* bd = su2t_clone_meta(b);
	su2t_reverse_axis_simple(bd, 1);
	su2t_flip_trees(bd);
	su2t_conjugate(bd);
	{ const int xs[2] = { 1, 3 }, xb[2] = { 1, 2 }; t = su2t_contract_simple(s, xs, bd, xb, 2); }
	su2t_free(s); su2t_free(bd);
	su2t_swap_tree_axes(t, 2, 3);
	const int perm[4] = { 1, 0, 3, 2 };
	struct su2t* rn = su2t_transpose_logical(t, perm);
	su2t_free(t);
	return rn;
}

/* su2_contraction_operator_step_left (src/algorithm/su2_chain_ops.c:259-362) */
static struct su2t* su2_step_left(struct su2t* a, struct su2t* b, struct su2t* w, struct su2t* l)
{
	struct su2t *s, *t;
	struct su2t* bd = su2t_clone_meta(b);
	su2t_flip_trees(bd);
	su2t_conjugate(bd);
	su2t_reverse_axis_simple(bd, 1);
	{ const int xl[1] = { 3 }, xb[1] = { 0 }; s = su2t_contract_simple(l, xl, bd, xb, 1); }
	su2t_free(bd);
	t = su2t_fmove(s, su2t_child_axis(s, 1, 1)); su2t_free(s);
	su2t_reverse_axis_simple(t, 3);
	s = su2t_fmove(t, su2t_child_axis(t, 0, 0)); su2t_free(t);
	{ const int xs[2] = { 2, 3 }, xw[2] = { 0, 1 }; t = su2t_contract_simple(s, xs, w, xw, 2); } su2t_free(s);
	s = su2t_fmove(t, su2t_child_axis(t, 0, 1)); su2t_free(t);
	{ const int xa[2] = { 0, 1 }, xs[2] = { 1, 3 }; t = su2t_contract_simple(a, xa, s, xs, 2); } su2t_free(s);
	const int perm[4] = { 1, 0, 3, 2 };
	struct su2t* ln = su2t_transpose_logical(t, perm);
	su2t_free(t);
	return ln;
}

/* ---- handles built from scratch ---- */
static struct su2_tree_node* mk_leaf(int ax) { struct su2_tree_node* n = ctb_malloc(sizeof *n); n->i_ax = ax; n->c[0] = n->c[1] = NULL; return n; }
static struct su2_tree_node* mk_node(int ax, struct su2_tree_node* a, struct su2_tree_node* b) { struct su2_tree_node* n = mk_leaf(ax); n->c[0] = a; n->c[1] = b; return n; }

/* host struct of a tensor with 4 logical axes, a single sector with a 1 x 1 x 1 x 1 degeneracy tensor holding 1 */
static void make_unit_block(enum numeric_type dtype, const qnumber j4[4], struct su2_tree_node* fuse, struct su2_tree_node* split, qnumber jroot, struct su2_tensor* t)
{
	memset(t, 0, sizeof *t);
	t->dtype = dtype; t->ndim_logical = 4; t->ndim_auxiliary = 0;
	t->tree.tree_fuse = fuse; t->tree.tree_split = split; t->tree.ndim = 5;
	t->outer_irreps = ctb_malloc(4 * sizeof *t->outer_irreps);
	t->dim_degen = ctb_malloc(4 * sizeof(ct_long*));
	for (int i = 0; i < 4; i++) {
		t->outer_irreps[i].num = 1;
		t->outer_irreps[i].jlist = ctb_malloc(sizeof(qnumber));
		t->outer_irreps[i].jlist[0] = j4[i];
		t->dim_degen[i] = ctb_calloc((size_t)j4[i] + 1, sizeof(ct_long));
		t->dim_degen[i][j4[i]] = 1;
	}
	t->charge_sectors.nsec = 1; t->charge_sectors.ndim = 5;
	t->charge_sectors.jlists = ctb_malloc(5 * sizeof(qnumber));
	for (int i = 0; i < 4; i++) { t->charge_sectors.jlists[i] = j4[i]; }
	t->charge_sectors.jlists[4] = jroot;
	t->degensors = ctb_malloc(sizeof(struct dense_tensor*));
	struct dense_tensor* d = ctb_calloc(1, sizeof *d);
	d->dtype = dtype; d->ndim = 4;
	d->dim = ctb_malloc(4 * sizeof(ct_long));
	for (int i = 0; i < 4; i++) { d->dim[i] = 1; }
	d->data = ctb_calloc(1, ctb_sizeof_dtype(dtype));
	if (dtype == CT_DOUBLE_REAL || dtype == CT_DOUBLE_COMPLEX) { ((double*)d->data)[0] = 1.0; } else { ((float*)d->data)[0] = 1.0f; }
	t->degensors[0] = d;
}

/* su2_create_dummy_operator_block_right (src/algorithm/su2_chain_ops.c:16-74) */
void su2_create_dummy_operator_block_right(const enum numeric_type dtype, const qnumber irrep_sector_state, struct su2_tensor* r)
{
	const qnumber j4[4] = { irrep_sector_state, 0, irrep_sector_state, 0 };
	make_unit_block(dtype, j4, mk_node(4, mk_leaf(2), mk_leaf(3)), mk_node(4, mk_leaf(0), mk_leaf(1)), irrep_sector_state, r);
}

/* su2_create_dummy_operator_block_left (src/algorithm/su2_chain_ops.c:82-131) */
void su2_create_dummy_operator_block_left(const enum numeric_type dtype, struct su2_tensor* l)
{
	const qnumber j4[4] = { 0, 0, 0, 0 };
	make_unit_block(dtype, j4, mk_node(4, mk_leaf(1), mk_leaf(2)), mk_node(4, mk_leaf(0), mk_leaf(3)), 0, l);
}

static void free_host_su2(struct su2_tensor* t)
{
	for (ct_long c = 0; c < t->charge_sectors.nsec; c++) { ctb_free(t->degensors[c]->data); ctb_free(t->degensors[c]->dim); ctb_free(t->degensors[c]); }
	ctb_free(t->degensors);
	for (int i = 0; i < t->ndim_logical; i++) { ctb_free(t->dim_degen[i]); }
	ctb_free(t->dim_degen);
	ctb_free(t->charge_sectors.jlists);
	for (int i = 0; i < t->ndim_logical + t->ndim_auxiliary; i++) { ctb_free(t->outer_irreps[i].jlist); }
	ctb_free(t->outer_irreps);
	struct su2_tree_node* stack[2 * SU2_MAXNODE]; int n = 0;
	stack[n++] = t->tree.tree_fuse; stack[n++] = t->tree.tree_split;
	while (n > 0) {
		struct su2_tree_node* nd = stack[--n];
		if (nd->c[0] != NULL) { stack[n++] = nd->c[0]; stack[n++] = nd->c[1]; }
		ctb_free(nd);
	}
}

static struct su2t* dummy_left_dev(int dtype) { struct su2_tensor t; su2_create_dummy_operator_block_left(dtype, &t); struct su2t* h = su2t_upload(&t); free_host_su2(&t); return h; }
static struct su2t* dummy_right_dev(int dtype, qnumber j) { struct su2_tensor t; su2_create_dummy_operator_block_right(dtype, j, &t); struct su2t* h = su2t_upload(&t); free_host_su2(&t); return h; }

/* ---------------------------------------------------------------------------------------------------------------------------
 * per-bond matrices and the factorisations of site tensors
 * ------------------------------------------------------------------------------------------------------------------------- */
struct su2_bondmat
{
	int dtype;
	int nj;
	qnumber* j;          /* bond quantum numbers, ascending */
	ct_long* rows;
	ct_long* cols;
	ct_long* off;        /* element offsets of the row-major matrices */
	void* dev;
};

static void bondmat_free(struct su2_bondmat* m)
{
	if (m == NULL) { return; }
	ctb_free(m->j); ctb_free(m->rows); ctb_free(m->cols); ctb_free(m->off);
	ctbd_free(m->dev);
	ctb_free(m);
}

static int bondmat_find(const struct su2_bondmat* m, qnumber j)
{
	for (int k = 0; k < m->nj; k++) { if (m->j[k] == j) { return k; } }
	return -1;
}

static int qn_cmp(const void* a, const void* b) { const qnumber x = *(const qnumber*)a, y = *(const qnumber*)b; return (x < y) ? -1 : (x > y); }

/* distinct values of axis 'ax' over the sectors of h, ascending */
static int sector_values(const struct su2t* h, int ax, qnumber** out)
{
	qnumber* v = ctb_malloc((size_t)(h->nsec + 1) * sizeof(qnumber));
	for (ct_long c = 0; c < h->nsec; c++) { v[c] = h->jl[c * h->ndim + ax]; }
	qsort(v, (size_t)h->nsec, sizeof(qnumber), qn_cmp);
	int n = 0;
	for (ct_long c = 0; c < h->nsec; c++) { if (c == 0 || v[c] != v[c - 1]) { v[n++] = v[c]; } }
	*out = v;
	return n;
}

static void set_axis_irreps(struct su2t* h, int ax, int nj, const qnumber* j, const ct_long* deg)
{
	ctb_free(h->irr[ax].jlist); ctb_free(h->dd[ax]);
	h->irr[ax].num = nj;
	h->irr[ax].jlist = ctb_malloc((size_t)(nj > 0 ? nj : 1) * sizeof(qnumber));
	qnumber jmax = 0;
	for (int k = 0; k < nj; k++) { h->irr[ax].jlist[k] = j[k]; if (j[k] > jmax) { jmax = j[k]; } }
	h->dd[ax] = ctb_calloc((size_t)jmax + 1, sizeof(ct_long));
	for (int k = 0; k < nj; k++) { h->dd[ax][j[k]] = deg[k]; }
}

static ct_long dd_at(const struct su2t* h, int ax, qnumber j) { return h->dd[ax][j]; }

/* |coefficient| of reversing the physical axis of a site tensor (j0, j1) -> j2 (su2_tensor_reverse_axis_simple, su2_tensor.c:975) */
static double reverse_weight(qnumber j0, qnumber j1, qnumber j2)
{
	return fabs(sqrt((double)j1 + 1.0) * ctb_su2_recoupling(j1, j0, j0, j1, j2, 0));
}

static void lc_block_2d(struct ctbd_lc_block* b, ct_long dst_off, ct_long rows, ct_long cols, ct_long dst_ld, ct_long src_ld, int tb, int te)
{
	memset(b, 0, sizeof *b);
	b->dst_off = dst_off; b->term_begin = tb; b->term_end = te; b->ndim = 2;
	b->dim[0] = (int32_t)rows; b->dim[1] = (int32_t)cols;
	b->dstride[0] = dst_ld; b->dstride[1] = 1;
	b->sstride[0] = src_ld; b->sstride[1] = 1;
}

/*
 * Left-orthonormalisation of a site tensor a = [Dl, d, Dr] with tree (0, 1) -> 2: a = q . R (su2_mps_local_orthonormalize_qr,
 * src/state/su2_mps.c:262-300, without the update of the neighbour).  side == 1: right-orthonormalisation a = R . q (:307-346).
 * 'a' is replaced by q; returns the matrices R per bond quantum number.
 */
static struct su2_bondmat* su2_site_qr(struct su2t** pa, int side)
{
	struct su2t* a = *pa;
	CTB_REQUIRE(a->nl == 3 && a->na == 0 && a->ndim == 3);
	const int bax = side ? 0 : 2;           /* bond that receives R */
	const size_t es = ctb_sizeof_dtype(a->dtype);
	qnumber* jv; const int nj = sector_values(a, bax, &jv);
	struct su2_bondmat* R = ctb_calloc(1, sizeof *R);
	R->dtype = a->dtype; R->nj = nj; R->j = jv;
	R->rows = ctb_calloc((size_t)nj + 1, sizeof(ct_long)); R->cols = ctb_calloc((size_t)nj + 1, sizeof(ct_long)); R->off = ctb_calloc((size_t)nj + 1, sizeof(ct_long));
	/* stacked matrices: side 0: (sum d0 d1) x D2 per j2;  side 1: D0 x (sum d1 d2) per j0 */
	ct_long* big = ctb_calloc((size_t)nj + 1, sizeof(ct_long));      /* stacked extent */
	ct_long* small = ctb_calloc((size_t)nj + 1, sizeof(ct_long));    /* bond degeneracy */
	ct_long* pos = ctb_calloc((size_t)a->nsec + 1, sizeof(ct_long)); /* start of the sector inside the stacked extent */
	int* jidx = ctb_calloc((size_t)a->nsec + 1, sizeof(int));
	for (ct_long c = 0; c < a->nsec; c++)
	{
		const qnumber* jl = &a->jl[c * 3];
		int k = -1;
		for (int q = 0; q < nj; q++) { if (jv[q] == jl[bax]) { k = q; break; } }
		jidx[c] = k;
		small[k] = dd_at(a, bax, jl[bax]);
		pos[c] = big[k];
		big[k] += side ? dd_at(a, 1, jl[1]) * dd_at(a, 2, jl[2]) : dd_at(a, 0, jl[0]) * dd_at(a, 1, jl[1]);
	}
	struct ctbd_mat_desc* desc = ctb_calloc((size_t)nj + 1, sizeof *desc);
	ct_long* moff = ctb_calloc((size_t)nj + 1, sizeof(ct_long));
	ct_long* qoff = ctb_calloc((size_t)nj + 1, sizeof(ct_long));
	ct_long* kk = ctb_calloc((size_t)nj + 1, sizeof(ct_long));
	ct_long mtot = 0, qtot = 0, rtot = 0;
	for (int k = 0; k < nj; k++)
	{
		const ct_long m = side ? small[k] : big[k], n = side ? big[k] : small[k];
		kk[k] = (m < n) ? m : n;
		moff[k] = mtot; mtot += m * n;
		desc[k].a_off = moff[k]; desc[k].m = (int32_t)m; desc[k].n = (int32_t)n;
		if (!side) { qoff[k] = qtot; qtot += m * kk[k]; R->off[k] = rtot; rtot += kk[k] * n; R->rows[k] = kk[k]; R->cols[k] = n;
			desc[k].o0_off = qoff[k]; desc[k].o1_off = R->off[k]; }
		else { R->off[k] = rtot; rtot += m * kk[k]; qoff[k] = qtot; qtot += kk[k] * n; R->rows[k] = m; R->cols[k] = kk[k];
			desc[k].o0_off = R->off[k]; desc[k].o1_off = qoff[k]; }
	}
	void *M = NULL, *Q = NULL;
	CTB_CHECK_ABORT(ctbd_malloc(&M, (size_t)(mtot > 0 ? mtot : 1) * es));
	CTB_CHECK_ABORT(ctbd_malloc(&Q, (size_t)(qtot > 0 ? qtot : 1) * es));
	CTB_CHECK_ABORT(ctbd_malloc(&R->dev, (size_t)(rtot > 0 ? rtot : 1) * es));
	/* stack */
	struct ctbd_lc_block* blocks = ctb_malloc((size_t)(a->nsec + 1) * sizeof *blocks);
	struct ctbd_lc_term* terms = ctb_malloc((size_t)(a->nsec + 1) * sizeof *terms);
	double* wgt = ctb_malloc((size_t)(a->nsec + 1) * sizeof(double));
	for (ct_long c = 0; c < a->nsec; c++)
	{
		const qnumber* jl = &a->jl[c * 3];
		const int k = jidx[c];
		const ct_long d0 = dd_at(a, 0, jl[0]), d1 = dd_at(a, 1, jl[1]), d2 = dd_at(a, 2, jl[2]);
		wgt[c] = side ? reverse_weight(jl[0], jl[1], jl[2]) : 1.0;
		if (!side) { lc_block_2d(&blocks[c], moff[k] + pos[c] * d2, d0 * d1, d2, d2, d2, (int)c, (int)c + 1); }
		else       { lc_block_2d(&blocks[c], moff[k] + pos[c], d0, d1 * d2, big[k], d1 * d2, (int)c, (int)c + 1); }
		terms[c].src_off = a->off[c];
		terms[c].coef = wgt[c] * (a->scale != NULL ? a->scale[c] : 1.0);
	}
	CTB_CHECK_ABORT(su2_dev_lc(a->dtype, a->conj, (int)a->nsec, blocks, (int)a->nsec, terms, a->dev, M, 0));
	const int dt = (a->dtype == CT_DOUBLE_COMPLEX) ? CTBD_C128 : CTBD_F64;
	if (!side) { CTB_CHECK_ABORT(ctbd_qr_batched(dt, 0, nj, desc, M, Q, R->dev)); }
	else       { CTB_CHECK_ABORT(ctbd_qr_batched(dt, 1, nj, desc, M, R->dev, Q)); }
	g_su2_launches++;
	/* the new site tensor */
	struct su2t* q = su2t_clone_meta(a);
	ctb_free(q->scale); q->scale = NULL; q->conj = 0; q->dev = NULL; q->own = 0;
	set_axis_irreps(q, bax, nj, jv, kk);
	ct_long o = 0;
	for (ct_long c = 0; c < q->nsec; c++)
	{
		const qnumber* jl = &q->jl[c * 3];
		q->nel[c] = dd_at(q, 0, jl[0]) * dd_at(q, 1, jl[1]) * dd_at(q, 2, jl[2]);
		q->off[c] = o; o += q->nel[c];
	}
	q->nstore = o;
	CTB_CHECK_ABORT(ctbd_malloc_noinit(&q->dev, (size_t)(o > 0 ? o : 1) * es)); q->own = 1;
	for (ct_long c = 0; c < q->nsec; c++)
	{
		const qnumber* jl = &q->jl[c * 3];
		const int k = jidx[c];
		const ct_long d0 = dd_at(a, 0, jl[0]), d1 = dd_at(a, 1, jl[1]), d2 = dd_at(a, 2, jl[2]);
		if (!side) { lc_block_2d(&blocks[c], q->off[c], d0 * d1, kk[k], kk[k], kk[k], (int)c, (int)c + 1); terms[c].src_off = qoff[k] + pos[c] * kk[k]; }
		else       { lc_block_2d(&blocks[c], q->off[c], kk[k], d1 * d2, d1 * d2, big[k], (int)c, (int)c + 1); terms[c].src_off = qoff[k] + pos[c]; }
		terms[c].coef = 1.0 / wgt[c];
	}
	CTB_CHECK_ABORT(su2_dev_lc(a->dtype, 0, (int)q->nsec, blocks, (int)q->nsec, terms, Q, q->dev, 0));
	ctbd_free(M); ctbd_free(Q);
	ctb_free(blocks); ctb_free(terms); ctb_free(wgt); ctb_free(desc); ctb_free(moff); ctb_free(qoff); ctb_free(kk);
	ctb_free(big); ctb_free(small); ctb_free(pos); ctb_free(jidx);
	su2t_free(a);
	*pa = q;
	return R;
}

/* b <- R . b on the left bond (side 0: rows of R become the new bond) or b <- b . R on the right bond (side 1); the update of the
 * neighbouring site tensor in su2_mps_local_orthonormalize_qr / _rq (src/state/su2_mps.c:285-299, :331-345): one grouped GEMM */
static void su2_site_absorb(struct su2t** pb, const struct su2_bondmat* R, int side)
{
	struct su2t* b = *pb;
	CTB_REQUIRE(b->nl == 3 && b->ndim == 3);
	CTB_CHECK_ABORT(su2t_materialize(b));
	const int bax = side ? 2 : 0;
	struct su2t* r = su2t_clone_meta(b);
	r->dev = NULL; r->own = 0;
	/* new bond: quantum numbers of R that occur in b */
	ct_long* deg = ctb_calloc((size_t)R->nj + 1, sizeof(ct_long));
	for (int k = 0; k < R->nj; k++) { deg[k] = side ? R->cols[k] : R->rows[k]; }
	set_axis_irreps(r, bax, R->nj, R->j, deg);
	ctb_free(deg);
	ct_long ns = 0, o = 0;
	ct_long* srcsec = ctb_malloc((size_t)(b->nsec + 1) * sizeof(ct_long));
	for (ct_long c = 0; c < b->nsec; c++)
	{
		const qnumber* jl = &b->jl[c * 3];
		const int k = bondmat_find(R, jl[bax]);
		if (k < 0) { continue; }
		CTB_REQUIRE((side ? R->rows[k] : R->cols[k]) == dd_at(b, bax, jl[bax]));
		memcpy(&r->jl[ns * 3], jl, 3 * sizeof(qnumber));
		r->nel[ns] = dd_at(r, 0, jl[0]) * dd_at(r, 1, jl[1]) * dd_at(r, 2, jl[2]);
		r->off[ns] = o; o += r->nel[ns];
		srcsec[ns] = c;
		ns++;
	}
	r->nsec = ns; r->nstore = o;
	CTB_CHECK_ABORT(ctbd_malloc_noinit(&r->dev, (size_t)(o > 0 ? o : 1) * ctb_sizeof_dtype(b->dtype))); r->own = 1;
	struct ctbd_gemm_out* outs = ctb_calloc((size_t)ns + 1, sizeof *outs);
	struct ctbd_gemm_seg* segs = ctb_calloc((size_t)ns + 1, sizeof *segs);
	ct_long tabcap = 1024, ntab = 0;
	int32_t* tab = malloc((size_t)tabcap * sizeof(int32_t));
	double flops = 0;
	for (ct_long q = 0; q < ns; q++)
	{
		const ct_long c = srcsec[q];
		const qnumber* jl = &b->jl[c * 3];
		const int k = bondmat_find(R, jl[bax]);
		ct_long M, N, K;
		if (!side) { M = R->rows[k]; K = R->cols[k]; N = dd_at(b, 1, jl[1]) * dd_at(b, 2, jl[2]); }
		else       { M = dd_at(b, 0, jl[0]) * dd_at(b, 1, jl[1]); K = R->rows[k]; N = R->cols[k]; }
		while (ntab + M + N > tabcap) { tabcap *= 2; tab = realloc(tab, (size_t)tabcap * sizeof(int32_t)); }
		outs[q].c_off = r->off[q]; outs[q].m = (int32_t)M; outs[q].n = (int32_t)N;
		outs[q].seg_begin = (int32_t)q; outs[q].seg_end = (int32_t)q + 1;
		outs[q].row_tab = (int32_t)ntab; for (ct_long i = 0; i < M; i++) { tab[ntab++] = (int32_t)(i * N); }
		outs[q].col_tab = (int32_t)ntab; for (ct_long i = 0; i < N; i++) { tab[ntab++] = (int32_t)i; }
		segs[q].a_off = side ? b->off[c] : R->off[k];
		segs[q].b_off = side ? R->off[k] : b->off[c];
		segs[q].k = (int32_t)K; segs[q].lda = (int32_t)K; segs[q].ldb = (int32_t)N;
		flops += 2.0 * (double)M * (double)N * (double)K;
	}
	struct ctbd_gemm_plan_host ph; memset(&ph, 0, sizeof ph);
	ph.dtype = (b->dtype == CT_DOUBLE_COMPLEX) ? CTBD_C128 : CTBD_F64;
	ph.a_kcontig = 1; ph.b_ncontig = 1;
	ph.nouts = (int32_t)ns; ph.nsegs = (int32_t)ns; ph.ntab = (int32_t)ntab;
	ph.outs = outs; ph.segs = segs; ph.tab = tab; ph.flops = flops;
	CTB_CHECK_ABORT(su2_dev_gemm(&ph, side ? b->dev : R->dev, side ? R->dev : b->dev, r->dev, 0));
	ctb_free(outs); ctb_free(segs); free(tab); ctb_free(srcsec);
	su2t_free(b);
	*pb = r;
}

/* ---- selection of singular values with multiplicities (retained_bond_indices_multiplicities, src/algorithm/truncation.c:235-373) ---- */
struct sv_rec { double v; ct_long i; };
static int sv_cmp(const void* a, const void* b)
{
	const struct sv_rec* x = a; const struct sv_rec* y = b;
	if (x->v < y->v) { return -1; }
	if (y->v < x->v) { return 1; }
	return (x->i > y->i) ? -1 : (x->i < y->i);      /* equal values: the later index counts as smaller, so a block keeps a prefix */
}

static void su2_retained(const double* sigma, const int* mult, ct_long n, double tol, int relative, ct_long max_vdim, char* keep, struct trunc_info* info)
{
	info->tol_eff = tol; info->norm_sigma = 0; info->entropy = 0;
	memset(keep, 0, (size_t)n);
	struct sv_rec* s = ctb_malloc((size_t)(n + 1) * sizeof *s);
	for (ct_long i = 0; i < n; i++) { s[i].v = sigma[i]; s[i].i = i; }
	qsort(s, (size_t)n, sizeof *s, sv_cmp);
	double sqsum = 0;
	for (ct_long i = 0; i < n; i++) { s[i].v = mult[s[i].i] * (s[i].v * s[i].v); sqsum += s[i].v; }
	if (sqsum == 0) { ctb_free(s); return; }
	if (relative) { for (ct_long i = 0; i < n; i++) { s[i].v /= sqsum; } }
	for (ct_long i = 1; i < n; i++) { s[i].v += s[i - 1].v; }
	ct_long n_logical = 0;
	for (ct_long i = 0; i < n; i++) { n_logical += mult[i]; }
	if (max_vdim < n_logical)
	{
		ct_long bare = 0, m = 0;
		for (ct_long i = 0; i < n; i++) {
			m += mult[s[i].i];
			if (n_logical - m <= max_vdim) { bare = n - i - 1; break; }
		}
		info->tol_eff = fmax(tol, s[n - bare - 1].v);
		for (ct_long i = 0; i < n - bare; i++) { s[i].v = 0; }
	}
	ct_long nret = 0;
	for (ct_long i = 0; i < n; i++) { if (s[i].v > tol) { keep[s[i].i] = 1; nret++; } }
	ctb_free(s);
	if (nret == 0) { return; }
	double nrm = 0;
	for (ct_long i = 0; i < n; i++) { if (keep[i]) { nrm += mult[i] * (sigma[i] * sigma[i]); } }
	info->norm_sigma = sqrt(nrm);
	double ent = 0;
	for (ct_long i = 0; i < n; i++) {
		if (keep[i]) {
			const double x = sigma[i] / info->norm_sigma;
			if (x > 0) { const double sq = x * x; ent -= (double)mult[i] * (sq * log(sq)); }
		}
	}
	info->entropy = ent;
}

struct rc_pair { qnumber a, b; };
static int rc_pair_cmp(const void* x, const void* y)
{
	const struct rc_pair* p = x; const struct rc_pair* q = y;
	if (p->a != q->a) { return p->a < q->a ? -1 : 1; }
	return (p->b < q->b) ? -1 : (p->b > q->b);
}

/*
 * Split of the two-site tensor th = [Dl, d1, d2, Dr] (tree (0, (1, 2)) -> 3) into two site tensors by a truncated SVD
 * (su2_mps_split_tensor_svd, src/state/su2_mps.c:582-665 with split_su2_matrix_svd, src/algorithm/su2_bond_ops.c:15-98).
 */
static int su2_split_pair(struct su2t* th, double tol, ct_long max_vdim, int distr, struct su2t** pa0, struct su2t** pa1, struct trunc_info* info)
{
	CTB_REQUIRE(th->nl == 4 && th->na == 0);
	struct su2t* T = su2t_fmove(th, su2t_child_axis(th, 1, 1));       /* ((0, 1) e, 2) -> 3, e = axis 4 */
	const int nd = 5;
	const size_t es = ctb_sizeof_dtype(T->dtype);
	qnumber* ev; const int ne = sector_values(T, 4, &ev);
	/* row blocks (j0, j1) and column blocks (j2, j3) per e */
	struct rc_pair** rb = ctb_calloc((size_t)ne + 1, sizeof *rb); int* nrb = ctb_calloc((size_t)ne + 1, sizeof(int));
	struct rc_pair** cb = ctb_calloc((size_t)ne + 1, sizeof *cb); int* ncb = ctb_calloc((size_t)ne + 1, sizeof(int));
	ct_long** rstart = ctb_calloc((size_t)ne + 1, sizeof *rstart); ct_long** cstart = ctb_calloc((size_t)ne + 1, sizeof *cstart);
	ct_long* rows = ctb_calloc((size_t)ne + 1, sizeof(ct_long)); ct_long* cols = ctb_calloc((size_t)ne + 1, sizeof(ct_long));
	for (int k = 0; k < ne; k++)
	{
		rb[k] = ctb_malloc((size_t)(T->nsec + 1) * sizeof **rb); cb[k] = ctb_malloc((size_t)(T->nsec + 1) * sizeof **cb);
		for (ct_long c = 0; c < T->nsec; c++) {
			const qnumber* jl = &T->jl[c * nd];
			if (jl[4] != ev[k]) { continue; }
			rb[k][nrb[k]].a = jl[0]; rb[k][nrb[k]].b = jl[1]; nrb[k]++;
			cb[k][ncb[k]].a = jl[2]; cb[k][ncb[k]].b = jl[3]; ncb[k]++;
		}
		qsort(rb[k], (size_t)nrb[k], sizeof **rb, rc_pair_cmp);
		qsort(cb[k], (size_t)ncb[k], sizeof **cb, rc_pair_cmp);
		int n = 0;
		for (int i = 0; i < nrb[k]; i++) { if (i == 0 || rc_pair_cmp(&rb[k][i], &rb[k][i - 1]) != 0) { rb[k][n++] = rb[k][i]; } }
		nrb[k] = n; n = 0;
		for (int i = 0; i < ncb[k]; i++) { if (i == 0 || rc_pair_cmp(&cb[k][i], &cb[k][i - 1]) != 0) { cb[k][n++] = cb[k][i]; } }
		ncb[k] = n;
		rstart[k] = ctb_calloc((size_t)nrb[k] + 1, sizeof(ct_long)); cstart[k] = ctb_calloc((size_t)ncb[k] + 1, sizeof(ct_long));
		for (int i = 0; i < nrb[k]; i++) { rstart[k][i] = rows[k]; rows[k] += dd_at(T, 0, rb[k][i].a) * dd_at(T, 1, rb[k][i].b); }
		for (int i = 0; i < ncb[k]; i++) { cstart[k][i] = cols[k]; cols[k] += dd_at(T, 2, cb[k][i].a) * dd_at(T, 3, cb[k][i].b); }
	}
	struct ctbd_mat_desc* desc = ctb_calloc((size_t)ne + 1, sizeof *desc);
	ct_long* moff = ctb_calloc((size_t)ne + 1, sizeof(ct_long)); ct_long* uoff = ctb_calloc((size_t)ne + 1, sizeof(ct_long));
	ct_long* voff = ctb_calloc((size_t)ne + 1, sizeof(ct_long)); ct_long* soff = ctb_calloc((size_t)ne + 1, sizeof(ct_long));
	ct_long* kk = ctb_calloc((size_t)ne + 1, sizeof(ct_long));
	ct_long mtot = 0, utot = 0, vtot = 0, stot = 0;
	for (int k = 0; k < ne; k++)
	{
		kk[k] = rows[k] < cols[k] ? rows[k] : cols[k];
		moff[k] = mtot; mtot += rows[k] * cols[k];
		uoff[k] = utot; utot += rows[k] * kk[k];
		voff[k] = vtot; vtot += kk[k] * cols[k];
		soff[k] = stot; stot += kk[k];
		desc[k].a_off = moff[k]; desc[k].m = (int32_t)rows[k]; desc[k].n = (int32_t)cols[k];
		desc[k].o0_off = uoff[k]; desc[k].o1_off = voff[k]; desc[k].s_off = soff[k];
	}
	void *M = NULL, *U = NULL, *V = NULL; double* S = NULL;
	CTB_CHECK(ctbd_malloc(&M, (size_t)(mtot > 0 ? mtot : 1) * es));           /* zero: blocks without a sector */
	CTB_CHECK(ctbd_malloc(&U, (size_t)(utot > 0 ? utot : 1) * es));
	CTB_CHECK(ctbd_malloc(&V, (size_t)(vtot > 0 ? vtot : 1) * es));
	CTB_CHECK(ctbd_malloc((void**)&S, (size_t)(stot > 0 ? stot : 1) * sizeof(double)));
	struct ctbd_lc_block* blocks = ctb_malloc((size_t)(T->nsec + 1) * sizeof *blocks);
	struct ctbd_lc_term* terms = ctb_malloc((size_t)(T->nsec + 1) * sizeof *terms);
	for (ct_long c = 0; c < T->nsec; c++)
	{
		const qnumber* jl = &T->jl[c * nd];
		int k = 0; while (ev[k] != jl[4]) { k++; }
		struct rc_pair pr = { jl[0], jl[1] }, pc = { jl[2], jl[3] };
		int ir = 0; while (rc_pair_cmp(&rb[k][ir], &pr) != 0) { ir++; }
		int ic = 0; while (rc_pair_cmp(&cb[k][ic], &pc) != 0) { ic++; }
		const ct_long nr = dd_at(T, 0, jl[0]) * dd_at(T, 1, jl[1]), nc = dd_at(T, 2, jl[2]) * dd_at(T, 3, jl[3]);
		lc_block_2d(&blocks[c], moff[k] + rstart[k][ir] * cols[k] + cstart[k][ic], nr, nc, cols[k], nc, (int)c, (int)c + 1);
		terms[c].src_off = T->off[c];
		terms[c].coef = reverse_weight(jl[4], jl[2], jl[3]) * (T->scale != NULL ? T->scale[c] : 1.0);
	}
	CTB_CHECK(su2_dev_lc(T->dtype, T->conj, (int)T->nsec, blocks, (int)T->nsec, terms, T->dev, M, 0));
	ctb_free(blocks); ctb_free(terms);
	const int dt = (T->dtype == CT_DOUBLE_COMPLEX) ? CTBD_C128 : CTBD_F64;
	int rc = ctbd_svd_batched(dt, ne, desc, M, U, V, S);
	g_su2_launches++;
	if (rc < 0) {
		fprintf(stderr, "chemtensor_b200: SU(2) split: batched SVD of %d sector matrices failed: %s\n", ne, ctbd_last_error());
		for (int k = 0; k < ne; k++) { fprintf(stderr, "  2j = %d: %lld x %lld\n", (int)ev[k], (long long)rows[k], (long long)cols[k]); }
		return rc;
	}
	double* sig = ctb_malloc((size_t)(stot + 1) * sizeof(double));
	CTB_CHECK(ctbd_d2h(sig, S, (size_t)stot * sizeof(double)));
	int* mult = ctb_malloc((size_t)(stot + 1) * sizeof(int));
	for (int k = 0; k < ne; k++) { for (ct_long i = 0; i < kk[k]; i++) { mult[soff[k] + i] = ev[k] + 1; } }
	char* keep = ctb_malloc((size_t)stot + 1);
	su2_retained(sig, mult, stot, tol, 1, max_vdim, keep, info);
	ct_long* nk = ctb_calloc((size_t)ne + 1, sizeof(ct_long));
	ct_long nret = 0;
	for (int k = 0; k < ne; k++) { for (ct_long i = 0; i < kk[k]; i++) { if (keep[soff[k] + i]) { nk[k]++; nret++; } } }
	if (nret == 0) { nk[0] = 1; sig[soff[0]] = 0; }     /* all truncated: one zero singular value on the smallest quantum number (su2_bond_ops.c:62-72) */
	/* new bond */
	qnumber* bj = ctb_malloc((size_t)(ne + 1) * sizeof(qnumber)); ct_long* bd = ctb_malloc((size_t)(ne + 1) * sizeof(ct_long));
	int nb = 0;
	for (int k = 0; k < ne; k++) { if (nk[k] > 0) { bj[nb] = ev[k]; bd[nb] = nk[k]; nb++; } }

	/* site tensors in MPS form: tree (0, 1) -> 2 */
	struct su2_tensor shell; memset(&shell, 0, sizeof shell);
	struct su2t* a0; struct su2t* a1;
	{
		/* a throw-away host description to create the handles' trees */
		struct su2_tree_node* f = mk_leaf(2); struct su2_tree_node* sp = mk_node(2, mk_leaf(0), mk_leaf(1));
		shell.dtype = T->dtype; shell.ndim_logical = 3; shell.ndim_auxiliary = 0;
		shell.tree.tree_fuse = f; shell.tree.tree_split = sp; shell.tree.ndim = 3;
		shell.outer_irreps = ctb_malloc(3 * sizeof *shell.outer_irreps);
		shell.dim_degen = ctb_malloc(3 * sizeof(ct_long*));
		for (int i = 0; i < 3; i++) {
			shell.outer_irreps[i].num = 1; shell.outer_irreps[i].jlist = ctb_calloc(1, sizeof(qnumber));
			shell.dim_degen[i] = ctb_calloc(1, sizeof(ct_long)); shell.dim_degen[i][0] = 1;
		}
		shell.charge_sectors.nsec = 0; shell.charge_sectors.ndim = 3; shell.charge_sectors.jlists = ctb_malloc(sizeof(qnumber));
		shell.degensors = ctb_malloc(sizeof(struct dense_tensor*));
		a0 = su2t_upload(&shell); a1 = su2t_upload(&shell);
		free_host_su2(&shell);
		ctbd_free(a0->dev); ctbd_free(a1->dev); a0->dev = a1->dev = NULL; a0->own = a1->own = 0;
	}
	{
		ct_long* dl = ctb_malloc((size_t)(T->irr[0].num + 1) * sizeof(ct_long));
		for (int i = 0; i < T->irr[0].num; i++) { dl[i] = dd_at(T, 0, T->irr[0].jlist[i]); }
		set_axis_irreps(a0, 0, T->irr[0].num, T->irr[0].jlist, dl); ctb_free(dl);
		dl = ctb_malloc((size_t)(T->irr[1].num + 1) * sizeof(ct_long));
		for (int i = 0; i < T->irr[1].num; i++) { dl[i] = dd_at(T, 1, T->irr[1].jlist[i]); }
		set_axis_irreps(a0, 1, T->irr[1].num, T->irr[1].jlist, dl); ctb_free(dl);
		set_axis_irreps(a0, 2, nb, bj, bd);
		set_axis_irreps(a1, 0, nb, bj, bd);
		dl = ctb_malloc((size_t)(T->irr[2].num + 1) * sizeof(ct_long));
		for (int i = 0; i < T->irr[2].num; i++) { dl[i] = dd_at(T, 2, T->irr[2].jlist[i]); }
		set_axis_irreps(a1, 1, T->irr[2].num, T->irr[2].jlist, dl); ctb_free(dl);
		dl = ctb_malloc((size_t)(T->irr[3].num + 1) * sizeof(ct_long));
		for (int i = 0; i < T->irr[3].num; i++) { dl[i] = dd_at(T, 3, T->irr[3].jlist[i]); }
		set_axis_irreps(a1, 2, T->irr[3].num, T->irr[3].jlist, dl); ctb_free(dl);
	}
	/* sectors: a0 (j0, j1, e) for every row block, a1 (e, j2, j3) for every column block of the kept e; one lc block per retained value */
	ct_long ns0 = 0, ns1 = 0;
	for (int k = 0; k < ne; k++) { if (nk[k] > 0) { ns0 += nrb[k]; ns1 += ncb[k]; } }
	for (int side = 0; side < 2; side++)
	{
		struct su2t* h = side ? a1 : a0;
		const ct_long ns = side ? ns1 : ns0;
		ctb_free(h->jl); h->jl = ctb_malloc((size_t)(3 * ns + 1) * sizeof(qnumber));
		ct_long q = 0;
		for (int k = 0; k < ne; k++) {
			if (nk[k] == 0) { continue; }
			const int nblk = side ? ncb[k] : nrb[k];
			for (int i = 0; i < nblk; i++) {
				if (!side) { h->jl[3 * q] = rb[k][i].a; h->jl[3 * q + 1] = rb[k][i].b; h->jl[3 * q + 2] = ev[k]; }
				else       { h->jl[3 * q] = ev[k]; h->jl[3 * q + 1] = cb[k][i].a; h->jl[3 * q + 2] = cb[k][i].b; }
				q++;
			}
		}
		/* lexicographic order: a1 is already sorted (e ascending, then (j2, j3)); a0 must be re-sorted */
		h->nsec = ns;
		if (!side) {
			struct rc3 { qnumber j[3]; }* tmp = (void*)h->jl;
			/* simple insertion sort on (j0, j1, j2): ns is small */
			for (ct_long i = 1; i < ns; i++) {
				struct rc3 x = tmp[i]; ct_long p = i - 1;
				while (p >= 0 && (tmp[p].j[0] > x.j[0] || (tmp[p].j[0] == x.j[0] && (tmp[p].j[1] > x.j[1] || (tmp[p].j[1] == x.j[1] && tmp[p].j[2] > x.j[2]))))) { tmp[p + 1] = tmp[p]; p--; }
				tmp[p + 1] = x;
			}
		}
		ctb_free(h->off); ctb_free(h->nel);
		h->off = ctb_malloc((size_t)(ns + 1) * sizeof(ct_long)); h->nel = ctb_malloc((size_t)(ns + 1) * sizeof(ct_long));
		ct_long o = 0;
		for (ct_long c = 0; c < ns; c++) {
			const qnumber* jl = &h->jl[3 * c];
			h->nel[c] = dd_at(h, 0, jl[0]) * dd_at(h, 1, jl[1]) * dd_at(h, 2, jl[2]);
			h->off[c] = o; o += h->nel[c];
		}
		h->nstore = o;
		CTB_CHECK(ctbd_malloc_noinit(&h->dev, (size_t)(o > 0 ? o : 1) * es)); h->own = 1;
		/* copy plan: one block per (sector, retained value) */
		ct_long nblocks = 0;
		for (ct_long c = 0; c < ns; c++) { nblocks += dd_at(h, side ? 0 : 2, h->jl[3 * c + (side ? 0 : 2)]); }
		struct ctbd_lc_block* bl = ctb_malloc((size_t)(nblocks + 1) * sizeof *bl);
		struct ctbd_lc_term* tm = ctb_malloc((size_t)(nblocks + 1) * sizeof *tm);
		ct_long nbq = 0;
		const int scaled = (side == 1) ? (distr == SU2_SVD_DISTR_RIGHT) : (distr == SU2_SVD_DISTR_LEFT);
		for (ct_long c = 0; c < ns; c++)
		{
			const qnumber* jl = &h->jl[3 * c];
			const qnumber e = side ? jl[0] : jl[2];
			int k = 0; while (ev[k] != e) { k++; }
			if (!side)
			{
				struct rc_pair pr = { jl[0], jl[1] }; int ir = 0; while (rc_pair_cmp(&rb[k][ir], &pr) != 0) { ir++; }
				const ct_long nr = dd_at(h, 0, jl[0]) * dd_at(h, 1, jl[1]);
				for (ct_long i = 0; i < nk[k]; i++) {
					/* column i of U_e restricted to the row block */
					memset(&bl[nbq], 0, sizeof *bl);
					bl[nbq].dst_off = h->off[c] + i; bl[nbq].term_begin = (int)nbq; bl[nbq].term_end = (int)nbq + 1; bl[nbq].ndim = 1;
					bl[nbq].dim[0] = (int32_t)nr; bl[nbq].dstride[0] = nk[k]; bl[nbq].sstride[0] = kk[k];
					tm[nbq].src_off = uoff[k] + rstart[k][ir] * kk[k] + i;
					tm[nbq].coef = scaled ? sig[soff[k] + i] : 1.0;
					nbq++;
				}
			}
			else
			{
				struct rc_pair pc = { jl[1], jl[2] }; int ic = 0; while (rc_pair_cmp(&cb[k][ic], &pc) != 0) { ic++; }
				const ct_long nc = dd_at(h, 1, jl[1]) * dd_at(h, 2, jl[2]);
				const double w = 1.0 / reverse_weight(e, jl[1], jl[2]);
				for (ct_long i = 0; i < nk[k]; i++) {
					/* row i of Vh_e restricted to the column block */
					memset(&bl[nbq], 0, sizeof *bl);
					bl[nbq].dst_off = h->off[c] + i * nc; bl[nbq].term_begin = (int)nbq; bl[nbq].term_end = (int)nbq + 1; bl[nbq].ndim = 1;
					bl[nbq].dim[0] = (int32_t)nc; bl[nbq].dstride[0] = 1; bl[nbq].sstride[0] = 1;
					tm[nbq].src_off = voff[k] + i * cols[k] + cstart[k][ic];
					tm[nbq].coef = w * (scaled ? sig[soff[k] + i] : 1.0);
					nbq++;
				}
			}
		}
		CTB_CHECK(su2_dev_lc(T->dtype, 0, (int)nbq, bl, (int)nbq, tm, side ? V : U, h->dev, 0));
		ctb_free(bl); ctb_free(tm);
	}
	ctbd_free(M); ctbd_free(U); ctbd_free(V); ctbd_free(S);
	for (int k = 0; k < ne; k++) { ctb_free(rb[k]); ctb_free(cb[k]); ctb_free(rstart[k]); ctb_free(cstart[k]); }
	ctb_free(rb); ctb_free(cb); ctb_free(nrb); ctb_free(ncb); ctb_free(rstart); ctb_free(cstart); ctb_free(rows); ctb_free(cols);
	ctb_free(desc); ctb_free(moff); ctb_free(uoff); ctb_free(voff); ctb_free(soff); ctb_free(kk); ctb_free(nk);
	ctb_free(sig); ctb_free(mult); ctb_free(keep); ctb_free(bj); ctb_free(bd); ctb_free(ev);
	su2t_free(T);
	*pa0 = a0; *pa1 = a1;
	return 0;
}

/* ---------------------------------------------------------------------------------------------------------------------------
 * local eigen-solver: Lanczos on the renormalised packed degeneracy tensors (su2_minimize_local_energy, src/algorithm/su2_dmrg.c:86-146,
 * with lanczos_iteration_d/z and eigensystem_krylov_*, src/util/krylov.c:24-345); the matvec is the replay of the recorded program
 * ------------------------------------------------------------------------------------------------------------------------- */
typedef struct su2t* (*su2_heff_fn)(struct su2t* a, void* ctx);

struct heff_single_ctx { struct su2t *w, *l, *r; };
struct heff_pair_ctx { struct su2t *w0, *w1, *l, *r; };
static struct su2t* heff_single_cb(struct su2t* a, void* ctx) { struct heff_single_ctx* c = ctx; return su2_heff_single(a, c->w, c->l, c->r); }
static struct su2t* heff_pair_cb(struct su2t* a, void* ctx) { struct heff_pair_ctx* c = ctx; return su2_heff_pair(a, c->w0, c->w1, c->l, c->r); }

static int su2_minimize(su2_heff_fn fn, void* ctx, struct su2t* a_start, int maxiter, double* en_min, struct su2t** a_opt)
{
	const double t0 = now_s();
	const int dtype = a_start->dtype;
	const int dt = (dtype == CT_DOUBLE_COMPLEX) ? CTBD_C128 : CTBD_F64;
	const size_t es = ctb_sizeof_dtype(dtype);
	/* canonical layout: every valid sector of the tree, packed in lexicographic order */
	struct su2t* can = su2t_clone_meta(a_start);
	can->dev = NULL; can->own = 0; can->varies = 0;
	su2t_set_sectors_all_valid(can);
	const ct_long n = can->nstore;
	CTB_REQUIRE(n > 0 && maxiter >= 1);
	void* V = NULL; void* w = NULL; double* scal = NULL;
	CTB_CHECK(ctbd_malloc_noinit(&V, (size_t)maxiter * (size_t)n * es));
	CTB_CHECK(ctbd_malloc_noinit(&w, (size_t)n * es));
	CTB_CHECK(ctbd_malloc((void**)&scal, (size_t)(3 * maxiter + 4) * sizeof(double)));
	double* d_alpha = scal; double* d_beta = scal + 2 * maxiter; double* d_tmp = scal + 3 * maxiter;
#define VJ(j) ((void*)((char*)V + (size_t)(j) * (size_t)n * es))
	/* v_0 = renormalised a_start / norm */
	can->dev = w;
	CTB_CHECK(su2t_embed(a_start, can, +1));
	CTB_CHECK(ctbd_nrm2(dt, n, w, d_tmp));
	CTB_CHECK(ctbd_rscale(dt, n, w, d_tmp, 1, VJ(0)));

	/* record the matvec: (renormalised vector in 'vin') -> tensor -> Heff -> renormalised vector in 'w' */
	void* vin = NULL;
	CTB_CHECK(ctbd_malloc_noinit(&vin, (size_t)n * es));
	CTB_CHECK(ctbd_d2d(vin, VJ(0), (size_t)n * es));
	struct su2_prog prog;
	su2_prog_begin(&prog);
	struct su2t* ain = su2t_clone_meta(can);
	ain->dev = vin; ain->own = 0; ain->varies = 1;
	struct su2t* aten = su2t_alloc_like(can, 0);
	aten->varies = 1;
	int rc = su2t_embed(ain, aten, -1);
	struct su2t* ha = NULL;
	if (rc == 0)
	{
		ha = fn(aten, ctx);
		struct su2t* wout = su2t_clone_meta(can);
		wout->dev = w; wout->own = 0; wout->varies = 1;
		rc = su2t_embed(ha, wout, +1);
		su2t_free(wout);
	}
	if (ha != NULL) { su2t_free(ha); }
	su2t_free(aten); su2t_free(ain);
	su2_prog_end();
	g_stats[1] += 1;
	if (rc == 0 && maxiter > 2) { su2_prog_capture(&prog); }

	int numiter = maxiter;
	double* host_scal = ctb_calloc((size_t)(3 * maxiter + 4), sizeof(double));
	double* alpha = ctb_calloc((size_t)maxiter + 1, sizeof(double)); double* beta = ctb_calloc((size_t)maxiter + 1, sizeof(double));
	for (int j = 0; j < maxiter && rc == 0; j++)
	{
		if (j > 0) {
			rc = ctbd_d2d(vin, VJ(j), (size_t)n * es);
			if (rc == 0) { rc = su2_prog_run(&prog); }
			g_stats[1] += 1;
			if (rc < 0) { break; }
		}
		rc = ctbd_dotc(dt, n, w, VJ(j), d_alpha + 2 * j);
		if (rc < 0 || j == maxiter - 1) { break; }
		rc = ctbd_lanczos_update(dt, n, w, VJ(j), j > 0 ? VJ(j - 1) : NULL, d_alpha + 2 * j, j > 0 ? d_beta + (j - 1) : NULL, d_beta + j);
		if (rc == 0) { rc = ctbd_rscale(dt, n, w, d_beta + j, 1, VJ(j + 1)); }
	}
	if (rc == 0) { rc = ctbd_d2h(host_scal, scal, (size_t)(3 * maxiter) * sizeof(double)); }
	if (rc == 0)
	{
		for (int j = 0; j < maxiter; j++) { alpha[j] = host_scal[2 * j]; }
		for (int j = 0; j < maxiter - 1; j++) {
			beta[j] = host_scal[2 * maxiter + j];
			if (!(beta[j] >= 100 * (double)n * DBL_EPSILON)) { numiter = j + 1; break; }     /* krylov.c:58 */
		}
		for (int j = 0; j < numiter; j++) {
			if (!isfinite(alpha[j])) {
				fprintf(stderr, "chemtensor_b200: SU(2) Lanczos produced a non-finite coefficient at iteration %d (alpha = %g, n = %lld)\n", j, alpha[j], (long long)n);
				rc = -1; break;
			}
		}
	}
	if (rc == 0)
	{
		double* z = ctb_malloc((size_t)numiter * numiter * sizeof(double));
		rc = ctb_tridiag_eig(numiter, alpha, beta, z);
		if (rc == 0)
		{
			*en_min = alpha[0];
			double* coef = ctb_malloc((size_t)numiter * sizeof(double));
			for (int j = 0; j < numiter; j++) { coef[j] = z[j * numiter]; }
			rc = ctbd_lincomb(dt, n, V, n, numiter, coef, w);
			ctb_free(coef);
			if (rc == 0)
			{
				struct su2t* wv = su2t_clone_meta(can);
				wv->dev = w; wv->own = 0;
				struct su2t* opt = su2t_alloc_like(can, 0);
				opt->varies = 0;
				rc = su2t_embed(wv, opt, -1);
				su2t_free(wv);
				*a_opt = opt;
			}
		}
		ctb_free(z);
	}
#undef VJ
	su2_prog_free(&prog);
	ctb_free(host_scal); ctb_free(alpha); ctb_free(beta);
	can->dev = NULL; su2t_free(can);
	ctbd_free(vin); ctbd_free(scal); ctbd_free(w); ctbd_free(V);
	g_stats[2] += now_s() - t0;
	return rc;
}

/* ---------------------------------------------------------------------------------------------------------------------------
 * MPS-level operations on device handles
 * ------------------------------------------------------------------------------------------------------------------------- */
static void su2_local_qr(struct su2t** a, struct su2t** a_next)
{
	struct su2_bondmat* R = su2_site_qr(a, 0);
	su2_site_absorb(a_next, R, 0);
	bondmat_free(R);
}

static void su2_local_rq(struct su2t** a, struct su2t** a_prev)
{
	struct su2_bondmat* R = su2_site_qr(a, 1);
	su2_site_absorb(a_prev, R, 1);
	bondmat_free(R);
}

/* right-orthonormalise site 0 against the fictitious head site; returns the (signed) scalar absorbed there (su2_mps.c:354, :527-575) */
static double su2_head_rq(struct su2t** a)
{
	struct su2_bondmat* R = su2_site_qr(a, 1);
	double nrm = 0;
	if (R->nj == 1 && R->rows[0] == 1 && R->cols[0] == 1) {
		double v[2] = { 0, 0 };
		CTB_CHECK_ABORT(ctbd_d2h(v, R->dev, ctb_sizeof_dtype(R->dtype)));
		nrm = v[0];
	}
	bondmat_free(R);
	return nrm;
}

static double su2_tail_qr(struct su2t** a)
{
	struct su2_bondmat* R = su2_site_qr(a, 0);
	double nrm = 0;
	if (R->nj == 1 && R->rows[0] == 1 && R->cols[0] == 1) {
		double v[2] = { 0, 0 };
		CTB_CHECK_ABORT(ctbd_d2h(v, R->dev, ctb_sizeof_dtype(R->dtype)));
		nrm = v[0];
	}
	bondmat_free(R);
	return nrm;
}

static void su2_negate(struct su2t* h)
{
	CTB_CHECK_ABORT(su2t_materialize(h));
	CTB_CHECK_ABORT(ctbd_scale_host(h->dtype == CT_DOUBLE_COMPLEX ? CTBD_C128 : CTBD_F64, h->nstore, h->dev, -1.0));
}

static double su2_orthonormalize_dev(struct su2t** a, int nsites, int mode)
{
	double nrm;
	if (mode == SU2_MPS_ORTHONORMAL_LEFT)
	{
		for (int l = 0; l < nsites - 1; l++) { su2_local_qr(&a[l], &a[l + 1]); }
		nrm = su2_tail_qr(&a[nsites - 1]);
		if (nrm < 0) { su2_negate(a[nsites - 1]); nrm = -nrm; }
	}
	else
	{
		for (int l = nsites - 1; l > 0; l--) { su2_local_rq(&a[l], &a[l - 1]); }
		nrm = su2_head_rq(&a[0]);
		if (nrm < 0) { su2_negate(a[0]); nrm = -nrm; }
	}
	return nrm;
}

static struct su2t** upload_sites(const struct su2_tensor* a, int n)
{
	struct su2t** h = ctb_calloc((size_t)n + 1, sizeof *h);
	for (int i = 0; i < n; i++) { h[i] = su2t_upload(&a[i]); }
	return h;
}

/* replace the host tensors of psi by the device ones (the host structs are released the way delete_su2_tensor does) */
static int download_sites(struct su2t** h, struct su2_mps* psi)
{
	for (int i = 0; i < psi->nsites; i++)
	{
		struct su2_tensor tmp;
		CTB_CHECK(su2t_download(h[i], &tmp));
		free_host_su2(&psi->a[i]);
		psi->a[i] = tmp;
	}
	return 0;
}

/* ---------------------------------------------------------------------------------------------------------------------------
 * exported entry points with the reference's signatures (host structs in and out)
 * ------------------------------------------------------------------------------------------------------------------------- */
void su2_tensor_contract_simple(const struct su2_tensor* s, const int* i_ax_s, const struct su2_tensor* t, const int* i_ax_t, const int ndim_mult, struct su2_tensor* r)
{
	struct su2t* hs = su2t_upload(s); struct su2t* ht = su2t_upload(t);
	struct su2t* hr = su2t_contract_simple(hs, i_ax_s, ht, i_ax_t, ndim_mult);
	CTB_CHECK_ABORT(su2t_download(hr, r));
	su2t_free(hs); su2t_free(ht); su2t_free(hr);
}

void su2_tensor_fmove(const struct su2_tensor* t, const int i_ax, struct su2_tensor* r)
{
	struct su2t* ht = su2t_upload(t);
	struct su2t* hr = su2t_fmove(ht, i_ax);
	CTB_CHECK_ABORT(su2t_download(hr, r));
	su2t_free(ht); su2t_free(hr);
}

void su2_contraction_operator_step_right(const struct su2_tensor* a, const struct su2_tensor* b, const struct su2_tensor* w, const struct su2_tensor* r, struct su2_tensor* r_next)
{
	struct su2t *ha = su2t_upload(a), *hb = su2t_upload(b), *hw = su2t_upload(w), *hr = su2t_upload(r);
	struct su2t* hn = su2_step_right(ha, hb, hw, hr);
	CTB_CHECK_ABORT(su2t_download(hn, r_next));
	su2t_free(ha); su2t_free(hb); su2t_free(hw); su2t_free(hr); su2t_free(hn);
}

void su2_contraction_operator_step_left(const struct su2_tensor* a, const struct su2_tensor* b, const struct su2_tensor* w, const struct su2_tensor* l, struct su2_tensor* l_next)
{
	struct su2t *ha = su2t_upload(a), *hb = su2t_upload(b), *hw = su2t_upload(w), *hl = su2t_upload(l);
	struct su2t* hn = su2_step_left(ha, hb, hw, hl);
	CTB_CHECK_ABORT(su2t_download(hn, l_next));
	su2t_free(ha); su2t_free(hb); su2t_free(hw); su2t_free(hl); su2t_free(hn);
}

void su2_apply_local_hamiltonian(const struct su2_tensor* a, const struct su2_tensor* w, const struct su2_tensor* l, const struct su2_tensor* r, struct su2_tensor* b)
{
	struct su2t *ha = su2t_upload(a), *hw = su2t_upload(w), *hl = su2t_upload(l), *hr = su2t_upload(r);
	struct su2t* hb = su2_heff_single(ha, hw, hl, hr);
	CTB_CHECK_ABORT(su2t_download(hb, b));
	su2t_free(ha); su2t_free(hw); su2t_free(hl); su2t_free(hr); su2t_free(hb);
}

void ctb_su2_apply_local_hamiltonian_pair(const struct su2_tensor* a2, const struct su2_tensor* w0, const struct su2_tensor* w1,
	const struct su2_tensor* l, const struct su2_tensor* r, struct su2_tensor* b2)
{
	struct su2t *ha = su2t_upload(a2), *h0 = su2t_upload(w0), *h1 = su2t_upload(w1), *hl = su2t_upload(l), *hr = su2t_upload(r);
	struct su2t* hb = su2_heff_pair(ha, h0, h1, hl, hr);
	CTB_CHECK_ABORT(su2t_download(hb, b2));
	su2t_free(ha); su2t_free(h0); su2t_free(h1); su2t_free(hl); su2t_free(hr); su2t_free(hb);
}

static struct su2t** right_blocks_dev(struct su2t** psi, struct su2t** chi, struct su2t** op, int nsites)
{
	struct su2t** rb = ctb_calloc((size_t)nsites + 1, sizeof *rb);
	const struct su2t* last = psi[nsites - 1];
	CTB_REQUIRE(last->irr[2].num == 1);
	rb[nsites - 1] = dummy_right_dev(last->dtype, last->irr[2].jlist[0]);
	for (int i = nsites - 1; i > 0; i--) { rb[i - 1] = su2_step_right(psi[i], chi[i], op[i], rb[i]); }
	return rb;
}

void su2_compute_right_operator_blocks(const struct su2_mps* psi, const struct su2_mps* chi, const struct su2_mpo* op, struct su2_tensor* r_list)
{
	const int n = op->nsites;
	struct su2t** hp = upload_sites(psi->a, n); struct su2t** hc = upload_sites(chi->a, n); struct su2t** ho = upload_sites(op->a, n);
	struct su2t** rb = right_blocks_dev(hp, hc, ho, n);
	for (int i = 0; i < n; i++) {
		CTB_CHECK_ABORT(su2t_download(rb[i], &r_list[i]));
		su2t_free(rb[i]); su2t_free(hp[i]); su2t_free(hc[i]); su2t_free(ho[i]);
	}
	ctb_free(rb); ctb_free(hp); ctb_free(hc); ctb_free(ho);
}

void su2_mpo_inner_product(const struct su2_mps* chi, const struct su2_mpo* op, const struct su2_mps* psi, void* ret)
{
	const int n = op->nsites;
	struct su2t** hp = upload_sites(psi->a, n); struct su2t** hc = upload_sites(chi->a, n); struct su2t** ho = upload_sites(op->a, n);
	struct su2t* r = dummy_right_dev(hp[n - 1]->dtype, hp[n - 1]->irr[2].jlist[0]);
	for (int i = n - 1; i >= 0; i--) {
		struct su2t* rn = su2_step_right(hp[i], hc[i], ho[i], r);
		su2t_free(r); r = rn;
	}
	CTB_CHECK_ABORT(su2t_materialize(r));
	CTB_REQUIRE(r->nsec == 1 && r->nel[0] == 1);
	CTB_CHECK_ABORT(ctbd_d2h(ret, (char*)r->dev + (size_t)r->off[0] * ctb_sizeof_dtype(r->dtype), ctb_sizeof_dtype(r->dtype)));
	su2t_free(r);
	for (int i = 0; i < n; i++) { su2t_free(hp[i]); su2t_free(hc[i]); su2t_free(ho[i]); }
	ctb_free(hp); ctb_free(hc); ctb_free(ho);
}

double su2_mps_orthonormalize_qr(struct su2_mps* mps, const enum su2_mps_orthonormalization_mode mode)
{
	struct su2t** h = upload_sites(mps->a, mps->nsites);
	const double nrm = su2_orthonormalize_dev(h, mps->nsites, (int)mode);
	CTB_CHECK_ABORT(download_sites(h, mps));
	for (int i = 0; i < mps->nsites; i++) { su2t_free(h[i]); }
	ctb_free(h);
	return nrm;
}

/* su2_dmrg_singlesite (src/algorithm/su2_dmrg.c:155-250) */
int su2_dmrg_singlesite(const struct su2_mpo* hamiltonian, const int num_sweeps, const int maxiter_lanczos, struct su2_mps* psi, double* en_sweeps)
{
	const int nsites = hamiltonian->nsites;
	CTB_REQUIRE(nsites == psi->nsites && nsites >= 1);
	memset(g_stats, 0, sizeof g_stats); g_su2_launches = 0;
	struct su2t** a = upload_sites(psi->a, nsites);
	struct su2t** w = upload_sites(hamiltonian->a, nsites);
	const int dtype = w[0]->dtype;
	double t0 = now_s();
	const double nrm = su2_orthonormalize_dev(a, nsites, SU2_MPS_ORTHONORMAL_RIGHT);
	if (nrm == 0) { printf("Warning: in 'su2_dmrg_singlesite': initial MPS has norm zero (possibly due to mismatching quantum numbers)\n"); }
	struct su2t** rb = right_blocks_dev(a, a, w, nsites);
	struct su2t** lb = ctb_calloc((size_t)nsites + 1, sizeof *lb);
	for (int i = 0; i < nsites; i++) { lb[i] = dummy_left_dev(dtype); }
	CTB_CHECK(ctbd_sync());
	g_stats[5] += now_s() - t0;
	int rc = 0;
	for (int n = 0; n < num_sweeps && rc == 0; n++)
	{
		double en = 0;
		const double t_sweep = now_s();
		for (int i = 0; i < nsites - 1 && rc == 0; i++)
		{
			struct heff_single_ctx ctx = { w[i], lb[i], rb[i] };
			struct su2t* opt = NULL;
			rc = su2_minimize(heff_single_cb, &ctx, a[i], maxiter_lanczos, &en, &opt);
			if (rc < 0) { break; }
			su2t_free(a[i]); a[i] = opt;
			t0 = now_s();
			su2_local_qr(&a[i], &a[i + 1]);
			g_stats[3] += now_s() - t0; t0 = now_s();
			su2t_free(lb[i + 1]);
			lb[i + 1] = su2_step_left(a[i], a[i], w[i], lb[i]);
			g_stats[4] += now_s() - t0;
		}
		for (int i = nsites - 1; i > 0 && rc == 0; i--)
		{
			struct heff_single_ctx ctx = { w[i], lb[i], rb[i] };
			struct su2t* opt = NULL;
			rc = su2_minimize(heff_single_cb, &ctx, a[i], maxiter_lanczos, &en, &opt);
			if (rc < 0) { break; }
			su2t_free(a[i]); a[i] = opt;
			t0 = now_s();
			su2_local_rq(&a[i], &a[i - 1]);
			g_stats[3] += now_s() - t0; t0 = now_s();
			su2t_free(rb[i - 1]);
			rb[i - 1] = su2_step_right(a[i], a[i], w[i], rb[i]);
			g_stats[4] += now_s() - t0;
		}
		if (rc == 0) {
			(void)su2_head_rq(&a[0]); en_sweeps[n] = en;
			g_stats[6] = n + 1;
			if (n < 9) { g_stats[7 + n] = now_s() - t_sweep; }
		}
	}
	if (rc == 0) { rc = download_sites(a, psi); }
	for (int i = 0; i < nsites; i++) { su2t_free(a[i]); su2t_free(w[i]); su2t_free(rb[i]); su2t_free(lb[i]); }
	ctb_free(a); ctb_free(w); ctb_free(rb); ctb_free(lb);
	return rc;
}

/* two-site tensor of neighbouring sites: su2_mps_contract_tensor_pair (src/state/su2_mps.c:673-688) */
static struct su2t* su2_contract_pair(struct su2t* a0, struct su2t* a1)
{
	const int x0[1] = { 2 }, x1[1] = { 0 };
	struct su2t* t = su2t_contract_simple(a0, x0, a1, x1, 1);
	struct su2t* th = su2t_fmove(t, su2t_child_axis(t, 1, 0));
	su2t_free(t);
	return th;
}

/* su2_dmrg_twosite (src/algorithm/su2_dmrg.c:262-429) */
int su2_dmrg_twosite(const struct su2_mpo* hamiltonian, const int num_sweeps, const int maxiter_lanczos, const double tol_split, const ct_long max_vdim,
	struct su2_mps* psi, double* en_sweeps, double* entropy)
{
	const int nsites = hamiltonian->nsites;
	CTB_REQUIRE(nsites == psi->nsites && nsites >= 2);
	memset(g_stats, 0, sizeof g_stats); g_su2_launches = 0;
	struct su2t** a = upload_sites(psi->a, nsites);
	struct su2t** w = upload_sites(hamiltonian->a, nsites);
	const int dtype = w[0]->dtype;
	double t0 = now_s();
	const double nrm = su2_orthonormalize_dev(a, nsites, SU2_MPS_ORTHONORMAL_RIGHT);
	if (nrm == 0) { printf("Warning: in 'su2_dmrg_twosite': initial MPS has norm zero (possibly due to mismatching quantum numbers)\n"); }
	struct su2t** rb = right_blocks_dev(a, a, w, nsites);
	struct su2t** lb = ctb_calloc((size_t)nsites + 1, sizeof *lb);
	for (int i = 0; i < nsites; i++) { lb[i] = dummy_left_dev(dtype); }
	CTB_CHECK(ctbd_sync());
	g_stats[5] += now_s() - t0;
	int rc = 0;
	for (int n = 0; n < num_sweeps && rc == 0; n++)
	{
		double en = 0;
		const double t_sweep = now_s();
		for (int dir = 0; dir < 2 && rc == 0; dir++)
		{
			const int i_begin = dir == 0 ? 0 : nsites - 2, i_end = dir == 0 ? nsites - 2 : -1, step = dir == 0 ? 1 : -1;
			for (int i = i_begin; i != i_end && rc == 0; i += step)
			{
				struct su2t* th = su2_contract_pair(a[i], a[i + 1]);
				struct heff_pair_ctx ctx = { w[i], w[i + 1], lb[i], rb[i + 1] };
				struct su2t* opt = NULL;
				rc = su2_minimize(heff_pair_cb, &ctx, th, maxiter_lanczos, &en, &opt);
				su2t_free(th);
				if (rc < 0) { break; }
				t0 = now_s();
				struct su2t *n0 = NULL, *n1 = NULL;
				struct trunc_info info;
				rc = su2_split_pair(opt, tol_split, max_vdim, dir == 0 ? SU2_SVD_DISTR_RIGHT : SU2_SVD_DISTR_LEFT, &n0, &n1, &info);
				su2t_free(opt);
				if (rc < 0) { break; }
				su2t_free(a[i]); su2t_free(a[i + 1]);
				a[i] = n0; a[i + 1] = n1;
				g_stats[3] += now_s() - t0; t0 = now_s();
				if (dir == 0) {
					su2t_free(lb[i + 1]);
					lb[i + 1] = su2_step_left(a[i], a[i], w[i], lb[i]);
				}
				else {
					entropy[i] = info.entropy;
					su2t_free(rb[i]);
					rb[i] = su2_step_right(a[i + 1], a[i + 1], w[i + 1], rb[i + 1]);
				}
				g_stats[4] += now_s() - t0;
			}
		}
		if (rc == 0) {
			(void)su2_head_rq(&a[0]); en_sweeps[n] = en;
			g_stats[6] = n + 1;
			if (n < 9) { g_stats[7 + n] = now_s() - t_sweep; }
		}
	}
	if (rc == 0) { rc = download_sites(a, psi); }
	for (int i = 0; i < nsites; i++) { su2t_free(a[i]); su2t_free(w[i]); su2t_free(rb[i]); su2t_free(lb[i]); }
	ctb_free(a); ctb_free(w); ctb_free(rb); ctb_free(lb);
	return rc;
}

/* ---- measurement aid: the block linear combination kernel alone against the HBM roofline ----
 * nblk destination blocks of nelem entries, each the sum of nterm source blocks (the shape of an F-move on a large tensor);
 * algorithmic bytes = (nterm + 1) x nblk x nelem x sizeof(T).  out = { best ms of 5 (L2 flushed before each), GB/s, bytes }. */
int ctb_su2_lc_benchmark(ct_long nelem, int nblk, int nterm, int cplx, double* out)
{
	CTB_CHECK(ctbd_init(-1));
	const size_t es = cplx ? 16 : 8;
	void *src = NULL, *dst = NULL, *flush = NULL;
	const size_t flush_bytes = (size_t)192 << 20;
	CTB_CHECK(ctbd_malloc(&src, (size_t)nelem * nblk * nterm * es));
	CTB_CHECK(ctbd_malloc(&dst, (size_t)nelem * nblk * es));
	CTB_CHECK(ctbd_malloc(&flush, flush_bytes));
	struct ctbd_lc_block* blocks = ctb_calloc((size_t)nblk, sizeof *blocks);
	struct ctbd_lc_term* terms = ctb_calloc((size_t)nblk * nterm, sizeof *terms);
	for (int b = 0; b < nblk; b++) {
		blocks[b].dst_off = (int64_t)b * nelem; blocks[b].term_begin = b * nterm; blocks[b].term_end = (b + 1) * nterm;
		blocks[b].ndim = 1; blocks[b].dim[0] = (int32_t)nelem; blocks[b].dstride[0] = 1; blocks[b].sstride[0] = 1;
		for (int t = 0; t < nterm; t++) { terms[b * nterm + t].src_off = ((int64_t)t * nblk + b) * nelem; terms[b * nterm + t].coef = 0.5 + t; }
	}
	void* plan = NULL;
	CTB_CHECK(ctbd_lc_plan_create(cplx ? CTBD_C128 : CTBD_F64, 0, nblk, blocks, nblk * nterm, terms, &plan));
	void *e0 = NULL, *e1 = NULL;
	CTB_CHECK(ctbd_event_create(&e0)); CTB_CHECK(ctbd_event_create(&e1));
	double best = 1e30;
	for (int rep = 0; rep < 6; rep++)
	{
		CTB_CHECK(ctbd_memset_zero(flush, flush_bytes));
		CTB_CHECK(ctbd_event_record(e0));
		CTB_CHECK(ctbd_lc_plan_run(plan, src, dst));
		CTB_CHECK(ctbd_event_record(e1));
		float ms = 0;
		CTB_CHECK(ctbd_event_elapsed_ms(e0, e1, &ms));
		if (rep > 0 && ms < best) { best = ms; }
	}
	const double bytes = (double)(nterm + 1) * (double)nblk * (double)nelem * (double)es;
	out[0] = best; out[1] = bytes / (best * 1e-3) / 1e9; out[2] = bytes;
	ctbd_event_destroy(e0); ctbd_event_destroy(e1);
	ctbd_lc_plan_destroy(plan);
	ctb_free(blocks); ctb_free(terms);
	ctbd_free(src); ctbd_free(dst); ctbd_free(flush);
	return 0;
}

/* ---------------------------------------------------------------------------------------------------------------------------
 * remaining entry points of SURVEY 8(a)'s SU(2) row with the reference's signatures
 * ------------------------------------------------------------------------------------------------------------------------- */

/* su2_tensor_num_elements_degensors (src/tensor/su2_tensor.c:4663) */
ct_long su2_tensor_num_elements_degensors(const struct su2_tensor* t)
{
	ct_long n = 0;
	for (ct_long c = 0; c < t->charge_sectors.nsec; c++) {
		ct_long m = 1;
		for (int i = 0; i < t->degensors[c]->ndim; i++) { m *= t->degensors[c]->dim[i]; }
		n += m;
	}
	return n;
}

/* su2_tensor_serialize_renormalized_entries / su2_tensor_deserialize_renormalized_entries (src/tensor/su2_tensor.c:4725-4942): the packed
 * vector the Lanczos iteration works on, every sector weighted by sqrt(2 j_root + 1) so that the Euclidean norm is the norm of the logical
 * tensor.  Host structs on both sides (inside the engine the same weights are folded into the first and last launch of the recorded matvec). */
static void su2_renormalized_copy(const struct su2_tensor* t, void* entries, int to_entries)
{
	const int root = t->tree.tree_fuse->i_ax;
	const int nd = t->charge_sectors.ndim;
	const int cplx = (t->dtype == CT_DOUBLE_COMPLEX || t->dtype == CT_SINGLE_COMPLEX);
	const int dbl = (t->dtype == CT_DOUBLE_REAL || t->dtype == CT_DOUBLE_COMPLEX);
	ct_long pos = 0;
	for (ct_long c = 0; c < t->charge_sectors.nsec; c++)
	{
		const qnumber j = t->charge_sectors.jlists[c * nd + root];
		ct_long n = 1;
		for (int i = 0; i < t->degensors[c]->ndim; i++) { n *= t->degensors[c]->dim[i]; }
		n *= cplx ? 2 : 1;
		if (dbl) {
			const double f = to_entries ? sqrt((double)j + 1.0) : 1.0 / sqrt((double)j + 1.0);
			double* d = t->degensors[c]->data; double* e = (double*)entries + pos;
			if (to_entries) { for (ct_long i = 0; i < n; i++) { e[i] = f * d[i]; } } else { for (ct_long i = 0; i < n; i++) { d[i] = f * e[i]; } }
		}
		else {
			const float f = to_entries ? sqrtf((float)j + 1.0f) : 1.0f / sqrtf((float)j + 1.0f);
			float* d = t->degensors[c]->data; float* e = (float*)entries + pos;
			if (to_entries) { for (ct_long i = 0; i < n; i++) { e[i] = f * d[i]; } } else { for (ct_long i = 0; i < n; i++) { d[i] = f * e[i]; } }
		}
		pos += n;
	}
}

void su2_tensor_serialize_renormalized_entries(const struct su2_tensor* t, void* entries) { su2_renormalized_copy(t, entries, 1); }
void su2_tensor_deserialize_renormalized_entries(struct su2_tensor* t, const void* entries) { su2_renormalized_copy(t, (void*)entries, 0); }

/* su2_tensor_svd (src/tensor/su2_tensor.c:4300-4440): economical SVD of an SU(2) symmetric matrix (two logical axes and one trivial auxiliary
 * axis): ONE batched launch over the degeneracy matrices of all charge sectors. */
int su2_tensor_svd(const struct su2_tensor* a, const bool copy_tree_left, struct su2_tensor* u, struct dense_tensor* s, int** multiplicities, struct su2_tensor* vh)
{
	CTB_REQUIRE(a->ndim_logical == 2 && a->ndim_auxiliary == 1 && a->charge_sectors.nsec > 0);
	struct su2t* ha = su2t_upload(a);
	const ct_long ns = ha->nsec;
	const size_t es = ctb_sizeof_dtype(ha->dtype);
	struct ctbd_mat_desc* desc = ctb_calloc((size_t)ns + 1, sizeof *desc);
	qnumber* js = ctb_malloc((size_t)(ns + 1) * sizeof(qnumber));
	ct_long* kk = ctb_malloc((size_t)(ns + 1) * sizeof(ct_long));
	ct_long utot = 0, vtot = 0, stot = 0;
	for (ct_long c = 0; c < ns; c++)
	{
		const qnumber j = ha->jl[c * 3];
		CTB_REQUIRE(ha->jl[c * 3 + 1] == j && ha->jl[c * 3 + 2] == 0);
		const ct_long m = dd_at(ha, 0, j), n = dd_at(ha, 1, j);
		js[c] = j; kk[c] = m < n ? m : n;
		desc[c].a_off = ha->off[c]; desc[c].m = (int32_t)m; desc[c].n = (int32_t)n;
		desc[c].o0_off = utot; utot += m * kk[c];
		desc[c].o1_off = vtot; vtot += kk[c] * n;
		desc[c].s_off = stot; stot += kk[c];
	}
	/* u and vh: structure of a with the shared bond in place of the second / first axis; the other factor carries the mirrored tree */
	struct su2t* hu = su2t_clone_meta(ha); struct su2t* hv = su2t_clone_meta(ha);
	hu->dev = hv->dev = NULL; hu->own = hv->own = 0;
	set_axis_irreps(hu, 1, (int)ns, js, kk);
	set_axis_irreps(hv, 0, (int)ns, js, kk);
	{
		struct su2t* flipped = copy_tree_left ? hv : hu;
		su2t_flip_trees(flipped);
		for (int i = 0; i < flipped->nn; i++) { if (flipped->node[i].ax == 0) { flipped->node[i].ax = 1; } else if (flipped->node[i].ax == 1) { flipped->node[i].ax = 0; } }
	}
	for (ct_long c = 0; c < ns; c++) {
		hu->off[c] = desc[c].o0_off; hu->nel[c] = (ct_long)desc[c].m * kk[c];
		hv->off[c] = desc[c].o1_off; hv->nel[c] = kk[c] * (ct_long)desc[c].n;
	}
	hu->nstore = utot; hv->nstore = vtot;
	double* S = NULL;
	CTB_CHECK(ctbd_malloc_noinit(&hu->dev, (size_t)(utot > 0 ? utot : 1) * es)); hu->own = 1;
	CTB_CHECK(ctbd_malloc_noinit(&hv->dev, (size_t)(vtot > 0 ? vtot : 1) * es)); hv->own = 1;
	CTB_CHECK(ctbd_malloc((void**)&S, (size_t)(stot > 0 ? stot : 1) * sizeof(double)));
	int rc = ctbd_svd_batched(ha->dtype == CT_DOUBLE_COMPLEX ? CTBD_C128 : CTBD_F64, (int)ns, desc, ha->dev, hu->dev, hv->dev, S);
	if (rc == 0)
	{
		s->dtype = CT_DOUBLE_REAL; s->ndim = 1;
		s->dim = ctb_malloc(sizeof(ct_long)); s->dim[0] = stot;
		s->data = ctb_malloc((size_t)(stot > 0 ? stot : 1) * sizeof(double));
		rc = ctbd_d2h(s->data, S, (size_t)stot * sizeof(double));
		*multiplicities = ctb_malloc((size_t)(stot > 0 ? stot : 1) * sizeof(int));
		for (ct_long c = 0; c < ns; c++) { for (ct_long i = 0; i < kk[c]; i++) { (*multiplicities)[desc[c].s_off + i] = js[c] + 1; } }
	}
	if (rc == 0) { rc = su2t_download(hu, u); }
	if (rc == 0) { rc = su2t_download(hv, vh); }
	ctbd_free(S);
	su2t_free(ha); su2t_free(hu); su2t_free(hv);
	ctb_free(desc); ctb_free(js); ctb_free(kk);
	return rc < 0 ? -1 : 0;
}

/* su2_mps_local_orthonormalize_qr / _rq (src/state/su2_mps.c:262-346): both tensors are replaced (the old host tensors are released the
 * way delete_su2_tensor does).  A bond quantum number without any charge sector in 'a' is dropped from the new bond (the reference keeps
 * it with an identity block in q and no block in r: the same state). */
static void su2_local_orthonormalize_host(struct su2_tensor* a, struct su2_tensor* a_nb, int rq)
{
	struct su2t* ha = su2t_upload(a); struct su2t* hn = su2t_upload(a_nb);
	if (rq) { su2_local_rq(&ha, &hn); } else { su2_local_qr(&ha, &hn); }
	struct su2_tensor ta, tn;
	CTB_CHECK_ABORT(su2t_download(ha, &ta)); CTB_CHECK_ABORT(su2t_download(hn, &tn));
	free_host_su2(a); free_host_su2(a_nb);
	*a = ta; *a_nb = tn;
	su2t_free(ha); su2t_free(hn);
}
void su2_mps_local_orthonormalize_qr(struct su2_tensor* a, struct su2_tensor* a_next) { su2_local_orthonormalize_host(a, a_next, 0); }
void su2_mps_local_orthonormalize_rq(struct su2_tensor* a, struct su2_tensor* a_prev) { su2_local_orthonormalize_host(a, a_prev, 1); }
