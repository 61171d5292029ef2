/*
 * assembly.c -- MPO tensors straight from the operator graph, and the coefficient gradient built on the device primitives.
 *
 *   mpo_from_assembly                        <- src/operator/mpo.c:59-172
 *       The reference accumulates every site tensor in a DENSE Dw x d x d x Dw' array first (mpo.c:77-79) and converts it to block
 *       form afterwards: for the molecular Hamiltonian of BASELINE.json configs[3] (Dw ~ 7 000) that is ~ 7 GB per site.  Here the
 *       edges of the graph are scattered directly into the packed block-sparse storage (memory proportional to the stored
 *       entries); entries that violate the quantum-number pattern are ignored, as in the reference (mpo.c:147).
 *   operator_average_coefficient_gradient    <- src/algorithm/gradient.c:15-236
 *       value and gradient of <chi| op |psi> with respect to the MPO coefficients: environments and the MPO-tensor environment
 *       of every site (compute_local_hamiltonian_environment, chain_ops.c:424) are device contractions; only each dw tensor
 *       (Dw x d x d x Dw', block-sparse) crosses back for the edge loop of gradient.c:101-216.
 */
#include <complex.h>
#include "ctb_internal.h"
#include "chemtensor_b200.h"

/* packed offset of a logical entry, or -1 if its block does not conserve the quantum numbers */
static ct_long packed_offset(const struct ctb_tensor* t, const ct_long* index)
{
	int sec[CTB_MAXDIM];
	for (int i = 0; i < t->ndim; i++) { sec[i] = t->ax[i].sec_of[index[i]]; }
	const ct_long base = ctb_grid_offset(t, ctb_grid_ravel(t, sec));
	if (base < 0) { return -1; }
	ct_long off = 0;
	for (int i = 0; i < t->ndim; i++) { off = off * t->ax[i].secdim[sec[i]] + t->ax[i].pos_of[index[i]]; }
	return base + off;
}

/* weighted sum of local operators of one edge (construct_local_operator, src/operator/local_op.c:49-178), d x d, as complex numbers */
static void edge_operator(const struct mpo_assembly* as, const struct mpo_graph_edge* edge, double complex* op)
{
	const ct_long d = as->d;
	const bool cplx = ctb_is_complex(as->dtype);
	for (ct_long j = 0; j < d * d; j++) { op[j] = 0; }
	for (int k = 0; k < edge->nopics; k++)
	{
		const int cid = edge->opics[k].cid, oid = edge->opics[k].oid;
		CTB_REQUIRE(0 <= cid && cid < as->num_coeffs);
		const double complex c = cplx ? ((const double complex*)as->coeffmap)[cid] : ((const double*)as->coeffmap)[cid];
		if (oid == OID_NOP) {
			CTB_REQUIRE(d == 1);      /* dummy identity on a one-dimensional site */
			op[0] += c;
			continue;
		}
		CTB_REQUIRE(0 <= oid && oid < as->num_local_ops);
		const struct dense_tensor* m = &as->opmap[oid];
		CTB_REQUIRE(m->ndim == 2 && m->dim[0] == d && m->dim[1] == d && m->dtype == as->dtype);
		for (ct_long j = 0; j < d * d; j++) {
			op[j] += c * (cplx ? ((const double complex*)m->data)[j] : ((const double*)m->data)[j]);
		}
	}
}

static struct ctb_tensor* site_structure(const struct mpo_assembly* as, int l)
{
	const struct mpo_graph* g = &as->graph;
	qnumber* qb[2];
	for (int i = 0; i < 2; i++) {
		qb[i] = ctb_malloc((size_t)g->num_verts[l + i] * sizeof(qnumber));
		for (int j = 0; j < g->num_verts[l + i]; j++) { qb[i][j] = g->verts[l + i][j].qnum; }
	}
	const ct_long dim[4] = { g->num_verts[l], as->d, as->d, g->num_verts[l + 1] };
	const int dirs[4] = { TENSOR_AXIS_OUT, TENSOR_AXIS_OUT, TENSOR_AXIS_IN, TENSOR_AXIS_IN };
	const qnumber* qn[4] = { qb[0], as->qsite, as->qsite, qb[1] };
	struct ctb_tensor* t = ctb_tensor_create(as->dtype, 4, dim, dirs, qn, 0);      /* structure only */
	ctb_free(qb[0]); ctb_free(qb[1]);
	return t;
}

void mpo_from_assembly(const struct mpo_assembly* assembly, struct mpo* mpo)
{
	const struct mpo_graph* g = &assembly->graph;
	CTB_REQUIRE(g->nsites >= 1 && assembly->d >= 1);
	CTB_REQUIRE(assembly->dtype == CT_DOUBLE_REAL || assembly->dtype == CT_DOUBLE_COMPLEX);
	const ct_long d = assembly->d;
	const bool cplx = ctb_is_complex(assembly->dtype);
	mpo->nsites = g->nsites;
	mpo->d = d;
	mpo->qsite = ctb_malloc((size_t)d * sizeof(qnumber));
	memcpy(mpo->qsite, assembly->qsite, (size_t)d * sizeof(qnumber));
	mpo->a = ctb_calloc((size_t)g->nsites, sizeof(struct block_sparse_tensor));
	double complex* op = ctb_malloc((size_t)(d * d) * sizeof(double complex));
	for (int l = 0; l < g->nsites; l++)
	{
		struct ctb_tensor* t = site_structure(assembly, l);
		const size_t esize = ctb_sizeof_dtype(assembly->dtype);
		void* packed = ctb_calloc((size_t)(t->nstore > 0 ? t->nstore : 1), esize);
		for (int e = 0; e < g->num_edges[l]; e++)
		{
			const struct mpo_graph_edge* edge = &g->edges[l][e];
			CTB_REQUIRE(0 <= edge->vids[0] && edge->vids[0] < g->num_verts[l] && 0 <= edge->vids[1] && edge->vids[1] < g->num_verts[l + 1]);
			edge_operator(assembly, edge, op);
			for (ct_long x = 0; x < d; x++) {
				for (ct_long y = 0; y < d; y++) {
					const double complex v = op[x * d + y];
					if (v == 0) { continue; }
					const ct_long index[4] = { edge->vids[0], x, y, edge->vids[1] };
					const ct_long off = packed_offset(t, index);
					if (off < 0) { continue; }      /* outside the quantum-number pattern: ignored (mpo.c:147) */
					if (cplx) { ((double complex*)packed)[off] += v; } else { ((double*)packed)[off] += creal(v); }
				}
			}
		}
		/* host struct with the reference's layout, filled from the packed entries (serialisation order) */
		ct_long dim[4]; enum tensor_axis_direction dirs[4]; const qnumber* qn[4];
		for (int i = 0; i < 4; i++) { dim[i] = t->ax[i].dim; dirs[i] = (enum tensor_axis_direction)t->ax[i].dir; qn[i] = t->ax[i].qlog; }
		allocate_block_sparse_tensor(assembly->dtype, 4, dim, dirs, qn, &mpo->a[l]);
		if (t->nstore > 0) { block_sparse_tensor_deserialize_entries(&mpo->a[l], packed); }
		ctb_free(packed);
		ctb_tensor_free(t);
	}
	ctb_free(op);
}

/* ---- gradient ---- */

/* the MPO-tensor environment of one site on device-resident operands (the body of compute_local_hamiltonian_environment) */
static struct ctb_tensor* local_environment(const struct ctb_tensor* ad, const struct ctb_tensor* bd, struct ctb_tensor* ld, const struct ctb_tensor* rd)
{
	const int perm0[5] = { 0, 1, 2, 4, 3 };
	struct ctb_tensor* s = ctb_dot(ad, TENSOR_AXIS_RANGE_TRAILING, 0, rd, TENSOR_AXIS_RANGE_LEADING, 0, 1, perm0);
	struct ctb_tensor* br = ctb_view_reversed_dirs(bd);
	const int perm2[6] = { 2, 0, 1, 3, 4, 5 };
	struct ctb_tensor* t = ctb_dot(br, TENSOR_AXIS_RANGE_TRAILING, ctb_is_complex(bd->dtype), s, TENSOR_AXIS_RANGE_TRAILING, 0, 1, perm2);
	ctb_tensor_free(br); ctb_tensor_free(s);
	const int perm1[4] = { 0, 2, 1, 3 };
	struct ctb_tensor* k = ctb_transpose(ld, perm1, 0);
	struct ctb_tensor* u = ctb_dot(k, TENSOR_AXIS_RANGE_TRAILING, 0, t, TENSOR_AXIS_RANGE_LEADING, 0, 2, NULL);
	ctb_tensor_free(k); ctb_tensor_free(t);
	struct ctb_tensor* dw = ctb_drop_dummy_axes(u, 1);
	ctb_tensor_free(u);
	return dw;
}

void operator_average_coefficient_gradient(const struct mpo_assembly* assembly, const struct mps* psi, const struct mps* chi, void* avr, void* dcoeff)
{
	CTB_CHECK_ABORT(ctbd_init(-1));
	const int L = assembly->graph.nsites;
	CTB_REQUIRE(assembly->d == psi->d && assembly->d == chi->d && psi->nsites == L && chi->nsites == L && L >= 1);
	CTB_REQUIRE(assembly->dtype == psi->a[0].dtype && assembly->dtype == chi->a[0].dtype);
	const bool cplx = ctb_is_complex(assembly->dtype);
	const size_t esize = ctb_sizeof_dtype(assembly->dtype);
	const ct_long d = assembly->d;
	memset(dcoeff, 0, (size_t)assembly->num_coeffs * esize);
	memset(avr, 0, esize);

	struct mpo mpo;
	mpo_from_assembly(assembly, &mpo);
	struct ctb_tensor** W = calloc((size_t)L, sizeof(*W));
	struct ctb_tensor** A = calloc((size_t)L, sizeof(*A));
	struct ctb_tensor** B = calloc((size_t)L, sizeof(*B));
	for (int l = 0; l < L; l++) {
		W[l] = ctb_upload(&mpo.a[l]);
		A[l] = ctb_upload(&psi->a[l]);
		B[l] = (chi == psi) ? A[l] : ctb_upload(&chi->a[l]);
	}
	delete_mpo(&mpo);
	/* right operator blocks (compute_right_operator_blocks, chain_ops.c:253), all resident */
	struct ctb_tensor** Rb = calloc((size_t)L, sizeof(*Rb));
	Rb[L - 1] = ctb_dummy_block_right(A[L - 1], B[L - 1], W[L - 1]);
	for (int l = L - 1; l > 0; l--) { Rb[l - 1] = ctb_env_step_right(A[l], B[l], W[l], Rb[l]); }
	struct ctb_tensor* lblock = ctb_dummy_block_left(A[0], B[0], W[0]);

	/* expectation value: one more step to the left of site 0, a 1 x 1 x 1 x 1 tensor remains (gradient.c:43-78) */
	bool nonzero = false;
	{
		struct ctb_tensor* r = ctb_env_step_right(A[0], B[0], W[0], Rb[0]);
		CTB_REQUIRE(r->nelem <= 1);
		if (r->nstore >= 1) { CTB_CHECK_ABORT(ctbd_d2h(avr, r->d, esize)); nonzero = true; }
		ctb_tensor_free(r);
	}
	for (int l = 0; l < L && nonzero; l++)
	{
		struct ctb_tensor* dw = local_environment(A[l], B[l], lblock, Rb[l]);
		CTB_REQUIRE(dw->ndim == 4 && dw->ax[1].dim == d && dw->ax[2].dim == d);
		CTB_REQUIRE(dw->ax[0].dim == W[l]->ax[0].dim && dw->ax[3].dim == W[l]->ax[3].dim);
		void* entries = ctb_malloc((size_t)(dw->nstore > 0 ? dw->nstore : 1) * esize);
		CTB_CHECK_ABORT(ctb_download_entries(dw, entries));
		for (int e = 0; e < assembly->graph.num_edges[l]; e++)
		{
			const struct mpo_graph_edge* edge = &assembly->graph.edges[l][e];
			for (int j = 0; j < edge->nopics; j++)
			{
				const int cid = edge->opics[j].cid, oid = edge->opics[j].oid;
				CTB_REQUIRE(0 <= cid && cid < assembly->num_coeffs && 0 <= oid && oid < assembly->num_local_ops);
				const struct dense_tensor* m = &assembly->opmap[oid];
				for (ct_long x = 0; x < d; x++) {
					for (ct_long y = 0; y < d; y++) {
						const double complex o = cplx ? ((const double complex*)m->data)[x * d + y] : ((const double*)m->data)[x * d + y];
						if (o == 0) { continue; }
						const ct_long index[4] = { edge->vids[0], x, y, edge->vids[1] };
						const ct_long off = packed_offset(dw, index);
						if (off < 0) { continue; }
						if (cplx) { ((double complex*)dcoeff)[cid] += o * ((const double complex*)entries)[off]; }
						else      { ((double*)dcoeff)[cid] += creal(o) * ((const double*)entries)[off]; }
					}
				}
			}
		}
		ctb_free(entries);
		ctb_tensor_free(dw);
		struct ctb_tensor* lnext = ctb_env_step_left(A[l], B[l], W[l], lblock);
		ctb_tensor_free(lblock);
		lblock = lnext;
	}
	ctb_tensor_free(lblock);
	for (int l = 0; l < L; l++) {
		ctb_tensor_free(Rb[l]); ctb_tensor_free(W[l]);
		if (B[l] != A[l]) { ctb_tensor_free(B[l]); }
		ctb_tensor_free(A[l]);
	}
	free(Rb); free(W); free(A); free(B);
}
