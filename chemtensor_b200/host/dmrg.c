/*
 * dmrg.c -- single- and two-site DMRG sweeps with all tensors resident on the device.
 *
 * Same sweep logic, ordering and recorded quantities as the reference src/algorithm/dmrg.c
 * (dmrg_singlesite :155-258, dmrg_twosite :262-399): the MPO and MPS are uploaded
 * once, environments / merged two-site MPO tensors / Krylov vectors never leave the device, and the
 * optimised MPS is downloaded at the end into genuine host structs (the caller's 'psi' is updated
 * in place: old site tensors are released and replaced).
 */
#include "ctb_internal.h"
#include "chemtensor_b200.h"

static void free_tensor_array(struct ctb_tensor** arr, int n)
{
	if (arr == NULL) { return; }
	for (int i = 0; i < n; i++) { ctb_tensor_free(arr[i]); }
	free(arr);
}

/* right-orthonormalise a device MPS (reference mps_orthonormalize_qr, src/state/mps.c:609-757, RIGHT branch) */
static int orthonormalize_right(struct ctb_tensor** A, int nsites, double* norm_out)
{
	for (int i = nsites - 1; i > 0; i--) {
		int rc = ctb_mps_local_rq(&A[i], &A[i - 1]);
		if (rc < 0) { return rc; }
	}
	CTB_REQUIRE(A[0]->ax[0].dim == 1);
	const ct_long dim_head[3] = { 1, 1, 1 };
	const int dir_head[3] = { TENSOR_AXIS_OUT, TENSOR_AXIS_OUT, TENSOR_AXIS_IN };
	const qnumber qzero[1] = { 0 };
	const qnumber* qn_head[3] = { A[0]->ax[0].qlog, qzero, A[0]->ax[0].qlog };
	struct ctb_tensor* head = ctb_tensor_create(A[0]->dtype, 3, dim_head, dir_head, qn_head, 1);
	CTB_REQUIRE(head->nblk == 1);
	CTB_CHECK(ctb_set_entry(head, 0, 1.0, 0.0));
	int rc = ctb_mps_local_rq(&A[0], &head);
	if (rc < 0) { ctb_tensor_free(head); return rc; }
	double norm = 0;
	if (head->ngrid > 0 && ctb_grid_offset(head, 0) >= 0)
	{
		double v[2] = { 0, 0 };
		CTB_CHECK(ctbd_d2h(v, (char*)head->d + (size_t)ctb_grid_offset(head, 0) * ctb_sizeof_dtype(head->dtype), ctb_sizeof_dtype(head->dtype)));
		norm = v[0];
		if (norm < 0)
		{
			/* keep the normalisation factor non-negative (reference mps.c:745-750) */
			CTB_CHECK(ctbd_scale_host(A[0]->dtype, A[0]->nstore, A[0]->d, -1.0));
			norm = -norm;
		}
	}
	ctb_tensor_free(head);
	*norm_out = norm;
	return 0;
}

/* re-normalise the leftmost tensor at the end of a sweep (reference dmrg.c:366-378) */
static int normalize_first_site(struct ctb_tensor** A0)
{
	const ct_long d0 = (*A0)->ax[0].dim;
	const ct_long dim[3] = { d0, 1, d0 };
	const int dirs[3] = { TENSOR_AXIS_OUT, TENSOR_AXIS_OUT, TENSOR_AXIS_IN };
	const qnumber qzero[1] = { 0 };
	const qnumber* qn[3] = { (*A0)->ax[0].qlog, qzero, (*A0)->ax[0].qlog };
	struct ctb_tensor* t = ctb_tensor_create((*A0)->dtype, 3, dim, dirs, qn, 1);
	int rc = ctb_mps_local_rq(A0, &t);
	ctb_tensor_free(t);
	return rc;
}

/* The merged pair tensor of the reference (dmrg.c:289-292) has up to Dw d^4 Dw'' entries.  Beyond this bound (dense count), or
 * with CTB_HEFF_PAIR=1, the two site tensors are applied one after the other instead and no merged tensor is ever built. */
#define CTB_PAIR_MERGE_LIMIT ((double)((ct_long)1 << 28))
static bool use_pair_form(const struct ctb_tensor* w0, const struct ctb_tensor* w1)
{
	const char* env = getenv("CTB_HEFF_PAIR");
	if (env != NULL) { return atoi(env) != 0; }
	const double d0 = (double)w0->ax[1].dim, d1 = (double)w1->ax[1].dim;
	return (double)w0->ax[0].dim * d0 * d0 * d1 * d1 * (double)w1->ax[3].dim > CTB_PAIR_MERGE_LIMIT;
}

/* w_second == NULL: w is the merged pair tensor and a_start [Dl, dd, Dr]; else w, w_second are the site tensors and a_start [Dl, d1, d2, Dr] */
static int minimize_local_energy(const struct ctb_tensor* w, const struct ctb_tensor* w_second, struct ctb_tensor* l, const struct ctb_tensor* r,
	const struct ctb_tensor* a_start, int maxiter, double* en, struct ctb_tensor** a_opt)
{
	struct ctb_heff h;
	const double t0 = ctb_wall_ms();
	if (w_second != NULL) { CTB_CHECK(ctb_heff_prepare_pair(a_start, w, w_second, l, r, &h)); }
	else { CTB_CHECK(ctb_heff_prepare(a_start, w, l, r, &h)); }
	int rc = ctb_lanczos_min(&h, a_start, maxiter, en, a_opt, NULL);
	ctb_heff_free(&h);
	if (a_start->nelem > ctb_global_stats.max_vector_len) { ctb_global_stats.max_vector_len = a_start->nelem; }
	ctb_global_stats.lanczos_ms += ctb_wall_ms() - t0;
	return rc;
}

static int upload_chain(const struct block_sparse_tensor* host, int n, struct ctb_tensor*** dev)
{
	*dev = calloc((size_t)n, sizeof(struct ctb_tensor*));
	ctb_collective_upload++;      /* dmrg_* is called by all ranks together */
	for (int i = 0; i < n; i++) { (*dev)[i] = ctb_upload(&host[i]); }
	ctb_collective_upload--;
	return 0;
}

static int download_mps(struct ctb_tensor** A, struct mps* psi)
{
	for (int i = 0; i < psi->nsites; i++)
	{
		/* into a temporary first: a failed copy must not leave psi holding a released tensor */
		struct block_sparse_tensor tmp;
		CTB_CHECK(ctb_download(A[i], &tmp));
		delete_block_sparse_tensor(&psi->a[i]);
		psi->a[i] = tmp;
	}
	return 0;
}

int dmrg_twosite(const struct mpo* hamiltonian, const int num_sweeps, const int maxiter_lanczos, const double tol_split, const ct_long max_vdim,
	struct mps* psi, double* en_sweeps, double* entropy)
{
	const int nsites = hamiltonian->nsites;
	CTB_REQUIRE(nsites == psi->nsites && nsites >= 2);
	CTB_REQUIRE(hamiltonian->a[0].dtype == CT_DOUBLE_REAL || hamiltonian->a[0].dtype == CT_DOUBLE_COMPLEX);
	CTB_CHECK(ctbd_init(-1));
	memset(&ctb_global_stats, 0, sizeof(ctb_global_stats));
	const double t_begin = ctb_wall_ms();

	struct ctb_tensor **W = NULL, **A = NULL;
	upload_chain(hamiltonian->a, nsites, &W);
	upload_chain(psi->a, nsites, &A);

	int ret = 0;
	double nrm = 0;
	ret = orthonormalize_right(A, nsites, &nrm);
	if (ret < 0) { goto cleanup_early; }
	if (nrm == 0) {
		printf("Warning: in 'dmrg_twosite': initial MPS has norm zero (possibly due to mismatching quantum numbers)\n");
	}

	struct ctb_tensor** Lb = calloc((size_t)nsites, sizeof(struct ctb_tensor*));
	struct ctb_tensor** Rb = calloc((size_t)nsites, sizeof(struct ctb_tensor*));
	struct ctb_tensor** h2 = calloc((size_t)nsites, sizeof(struct ctb_tensor*));
	{
		const double t0 = ctb_wall_ms();
		Rb[nsites - 1] = ctb_dummy_block_right(A[nsites - 1], A[nsites - 1], W[nsites - 1]);
		for (int i = nsites - 1; i > 0; i--) {
			Rb[i - 1] = ctb_env_step_right(A[i], A[i], W[i], Rb[i]);
		}
		Lb[0] = ctb_dummy_block_left(A[0], A[0], W[0]);
		for (int i = 0; i < nsites - 1; i++) {
			if (!use_pair_form(W[i], W[i + 1])) { h2[i] = ctb_mpo_merge_pair(W[i], W[i + 1]); }
		}
		ctb_global_stats.env_ms += ctb_wall_ms() - t0;
	}

	const ct_long d_pair[2] = { psi->d, psi->d };
	const qnumber* qsite_pair[2] = { psi->qsite, psi->qsite };

	for (int n = 0; n < num_sweeps && ret == 0; n++)
	{
		double en = 0;
		const double t_sweep = ctb_wall_ms();
		for (int pass = 0; pass < 2 && ret == 0; pass++)
		{
			/* pass 0: left to right over pairs 0..L-3; pass 1: right to left over pairs L-2..0 */
			const int i_begin = (pass == 0 ? 0 : nsites - 2);
			const int i_end   = (pass == 0 ? nsites - 2 : -1);
			const int step    = (pass == 0 ? 1 : -1);
			for (int i = i_begin; i != i_end; i += step)
			{
				struct ctb_tensor* a_cur = NULL;
				struct ctb_tensor* a_opt = NULL;
				if (h2[i] != NULL)
				{
					a_cur = ctb_mps_merge_pair(A[i], A[i + 1]);
					ret = minimize_local_energy(h2[i], NULL, Lb[i], Rb[i + 1], a_cur, maxiter_lanczos, &en, &a_opt);
				}
				else
				{
					/* pair form: the two-site tensor keeps its physical legs apart during the local solve (the Lanczos vector holds the
					 * same entries in the packed order of the 4-leg tensor) and is fused afterwards for the split */
					a_cur = ctb_dot(A[i], TENSOR_AXIS_RANGE_TRAILING, 0, A[i + 1], TENSOR_AXIS_RANGE_LEADING, 0, 1, NULL);
					struct ctb_tensor* a_opt4 = NULL;
					ret = minimize_local_energy(W[i], W[i + 1], Lb[i], Rb[i + 1], a_cur, maxiter_lanczos, &en, &a_opt4);
					if (ret == 0) { a_opt = ctb_flatten_axes(a_opt4, 1, TENSOR_AXIS_OUT); }
					ctb_tensor_free(a_opt4);
				}
				ctb_tensor_free(A[i]);     A[i] = NULL;
				ctb_tensor_free(A[i + 1]); A[i + 1] = NULL;
				ctb_tensor_free(a_cur);
				if (ret < 0) { break; }

				struct trunc_info info;
				const double t0 = ctb_wall_ms();
				ret = ctb_mps_split_svd(a_opt, d_pair, qsite_pair, tol_split, max_vdim, false,
					pass == 0 ? SVD_DISTR_RIGHT : SVD_DISTR_LEFT, &A[i], &A[i + 1], &info);
				ctb_tensor_free(a_opt);
				ctb_global_stats.svd_ms += ctb_wall_ms() - t0;
				if (ret < 0) { break; }
				if (A[i]->ax[2].dim > ctb_global_stats.max_bond_dim) { ctb_global_stats.max_bond_dim = A[i]->ax[2].dim; }

				const double t1 = ctb_wall_ms();
				if (pass == 0)
				{
					ctb_tensor_free(Lb[i + 1]);
					Lb[i + 1] = ctb_env_step_left(A[i], A[i], W[i], Lb[i]);
				}
				else
				{
					entropy[i] = info.entropy;
					ctb_tensor_free(Rb[i]);
					Rb[i] = ctb_env_step_right(A[i + 1], A[i + 1], W[i + 1], Rb[i + 1]);
				}
				ctb_global_stats.env_ms += ctb_wall_ms() - t1;
			}
		}
		if (ret < 0) { break; }
		ret = normalize_first_site(&A[0]);
		en_sweeps[n] = en;
		if (n < 8) { CTB_CHECK(ctbd_sync()); ctb_global_stats.sweep_ms[n] = ctb_wall_ms() - t_sweep; }
	}

	if (ret == 0) {
		CTB_CHECK(ctbd_sync());
		ctb_global_stats.total_ms = ctb_wall_ms() - t_begin;
		ret = download_mps(A, psi);
	}
	if (getenv("CTB_TRACE_PLAN") != NULL) {
		fprintf(stderr, "contraction plans: %.0f plans, %.0f output blocks, %.0f table entries; result tensors %.1f ms, host lists (incl.) %.1f ms, device plans %.1f ms\n",
			ctb_plan_profile[3], ctb_plan_profile[4], ctb_plan_profile[5], ctb_plan_profile[0], ctb_plan_profile[1], ctb_plan_profile[2]);
		fprintf(stderr, "  of the host lists: enumeration of the contracted sector tuples %.1f ms (plain) + %.1f ms (merged rows)\n", ctb_plan_profile[6], ctb_plan_profile[7]);
		fprintf(stderr, "  merged rows: row descriptors / tables %.1f ms, gather lists %.1f ms, packed-matrix reuse search %.1f ms; plain: offset tables %.1f ms\n",
			ctb_plan_profile[8], ctb_plan_profile[9], ctb_plan_profile[10], ctb_plan_profile[11]);
	}

	free_tensor_array(h2, nsites);
	free_tensor_array(Lb, nsites);
	free_tensor_array(Rb, nsites);
cleanup_early:
	free_tensor_array(A, nsites);
	free_tensor_array(W, nsites);
	return ret;
}

int dmrg_singlesite(const struct mpo* hamiltonian, const int num_sweeps, const int maxiter_lanczos, struct mps* psi, double* en_sweeps)
{
	const int nsites = hamiltonian->nsites;
	CTB_REQUIRE(nsites == psi->nsites && nsites >= 1);
	CTB_REQUIRE(hamiltonian->a[0].dtype == CT_DOUBLE_REAL || hamiltonian->a[0].dtype == CT_DOUBLE_COMPLEX);
	CTB_CHECK(ctbd_init(-1));
	memset(&ctb_global_stats, 0, sizeof(ctb_global_stats));
	const double t_begin = ctb_wall_ms();

	struct ctb_tensor **W = NULL, **A = NULL;
	upload_chain(hamiltonian->a, nsites, &W);
	upload_chain(psi->a, nsites, &A);

	int ret = 0;
	double nrm = 0;
	ret = orthonormalize_right(A, nsites, &nrm);
	if (ret < 0) { goto cleanup_early; }
	if (nrm == 0) {
		printf("Warning: in 'dmrg_singlesite': initial MPS has norm zero (possibly due to mismatching quantum numbers)\n");
	}

	struct ctb_tensor** Lb = calloc((size_t)nsites, sizeof(struct ctb_tensor*));
	struct ctb_tensor** Rb = calloc((size_t)nsites, sizeof(struct ctb_tensor*));
	Rb[nsites - 1] = ctb_dummy_block_right(A[nsites - 1], A[nsites - 1], W[nsites - 1]);
	for (int i = nsites - 1; i > 0; i--) {
		Rb[i - 1] = ctb_env_step_right(A[i], A[i], W[i], Rb[i]);
	}
	Lb[0] = ctb_dummy_block_left(A[0], A[0], W[0]);

	for (int n = 0; n < num_sweeps && ret == 0; n++)
	{
		double en = 0;
		for (int i = 0; i < nsites - 1 && ret == 0; i++)
		{
			struct ctb_tensor* a_opt = NULL;
			ret = minimize_local_energy(W[i], NULL, Lb[i], Rb[i], A[i], maxiter_lanczos, &en, &a_opt);
			if (ret < 0) { break; }
			ctb_tensor_free(A[i]);
			A[i] = a_opt;
			ret = ctb_mps_local_qr(&A[i], &A[i + 1]);
			if (ret < 0) { break; }
			ctb_tensor_free(Lb[i + 1]);
			Lb[i + 1] = ctb_env_step_left(A[i], A[i], W[i], Lb[i]);
		}
		for (int i = nsites - 1; i > 0 && ret == 0; i--)
		{
			struct ctb_tensor* a_opt = NULL;
			ret = minimize_local_energy(W[i], NULL, Lb[i], Rb[i], A[i], maxiter_lanczos, &en, &a_opt);
			if (ret < 0) { break; }
			ctb_tensor_free(A[i]);
			A[i] = a_opt;
			ret = ctb_mps_local_rq(&A[i], &A[i - 1]);
			if (ret < 0) { break; }
			ctb_tensor_free(Rb[i - 1]);
			Rb[i - 1] = ctb_env_step_right(A[i], A[i], W[i], Rb[i]);
		}
		if (ret < 0) { break; }
		ret = normalize_first_site(&A[0]);
		en_sweeps[n] = en;
	}

	if (ret == 0) {
		CTB_CHECK(ctbd_sync());
		ctb_global_stats.total_ms = ctb_wall_ms() - t_begin;
		ret = download_mps(A, psi);
	}
	free_tensor_array(Lb, nsites);
	free_tensor_array(Rb, nsites);
cleanup_early:
	free_tensor_array(A, nsites);
	free_tensor_array(W, nsites);
	return ret;
}
