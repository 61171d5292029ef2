/*
 * krylov.c -- Lanczos iteration with device-resident Krylov vectors.
 *
 * Semantics of the reference src/util/krylov.c:
 *   plain three-term Lanczos without re-orthogonalisation (:24-89 real, :96-167 complex),
 *   alpha_j = Re <w, v_j>, breakdown test beta_j < 100 n eps (:58) ending with numiter = j + 1,
 *   the last iteration only computes alpha; tridiagonal eigenproblem on 'numiter' (dsteqr "I", :230),
 *   Ritz vector u = V^T U[:, :numeig] without re-normalisation (:242, :335).
 * The level-1 work (dot, fused update + norm, scaling, Ritz combination) runs in fused device kernels;
 * only alpha_j and beta_j (two doubles) return to the host per iteration for the breakdown test.
 * The tridiagonal problem (<= maxiter x maxiter) is solved on the host by implicit QL.
 */
#include <float.h>
#include "ctb_internal.h"

/* ---- symmetric tridiagonal eigenproblem, implicit QL with Wilkinson shifts ----
 * d[0..n) diagonal, e[0..n-1) off-diagonal (destroyed); z (n x n, row-major) receives the eigenvectors
 * as columns; eigenvalues returned ascending in d.  Returns 0, or -2 if an eigenvalue fails to converge. */
int ctb_tridiag_eig(int n, double* d, double* e_in, double* z)
{
	double* e = ctb_calloc((size_t)n + 1, sizeof(double));
	for (int i = 0; i + 1 < n; i++) { e[i] = e_in[i]; }
	for (int i = 0; i < n; i++) { for (int j = 0; j < n; j++) { z[i * n + j] = (i == j) ? 1.0 : 0.0; } }

	for (int l = 0; l < n; l++)
	{
		int iter = 0;
		int m;
		do
		{
			for (m = l; m < n - 1; m++)
			{
				const double dd = fabs(d[m]) + fabs(d[m + 1]);
				if (fabs(e[m]) <= DBL_EPSILON * dd) { break; }
			}
			if (m != l)
			{
				if (iter++ == 60) { ctb_free(e); return -2; }
				double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
				double r = hypot(g, 1.0);
				g = d[m] - d[l] + e[l] / (g + (g >= 0 ? fabs(r) : -fabs(r)));
				double s = 1.0, c = 1.0, p = 0.0;
				int i;
				for (i = m - 1; i >= l; i--)
				{
					double f = s * e[i];
					const double b = c * e[i];
					r = hypot(f, g);
					e[i + 1] = r;
					if (r == 0.0)
					{
						d[i + 1] -= p;
						e[m] = 0.0;
						break;
					}
					s = f / r;
					c = g / r;
					g = d[i + 1] - p;
					r = (d[i] - g) * s + 2.0 * c * b;
					p = s * r;
					d[i + 1] = g + p;
					g = c * r - b;
					for (int k = 0; k < n; k++)
					{
						f = z[k * n + i + 1];
						z[k * n + i + 1] = s * z[k * n + i] + c * f;
						z[k * n + i]     = c * z[k * n + i] - s * f;
					}
				}
				if (r == 0.0 && i >= l) { continue; }
				d[l] -= p;
				e[l] = g;
				e[m] = 0.0;
			}
		} while (m != l);
	}
	ctb_free(e);

	/* selection sort: ascending eigenvalues, columns follow */
	for (int i = 0; i < n - 1; i++)
	{
		int k = i;
		for (int j = i + 1; j < n; j++) { if (d[j] < d[k]) { k = j; } }
		if (k != i)
		{
			const double t = d[i]; d[i] = d[k]; d[k] = t;
			for (int r = 0; r < n; r++) {
				const double u = z[r * n + i]; z[r * n + i] = z[r * n + k]; z[r * n + k] = u;
			}
		}
	}
	return 0;
}

/* ---- device Lanczos driving the cached Heff plans ---- */
int ctb_lanczos_min(struct ctb_heff* h, const struct ctb_tensor* a_start, int maxiter, double* en_min, struct ctb_tensor** a_opt, int* numiter_out)
{
	CTB_REQUIRE(maxiter >= 1);
	const int dtype = a_start->dtype;
	const size_t esize = ctb_sizeof_dtype(dtype);
	const ct_long ns = a_start->nstore;     /* stored vector length on the device */
	const ct_long n  = a_start->nelem;      /* logical length (reference 'n') */
	CTB_REQUIRE(ns > 0);

	void* V = NULL; void* w_own = NULL; double* scal = NULL;
	double *alpha = NULL, *beta = NULL, *host_scal = NULL;
	int rc = 0, numiter = maxiter;
#define LZ(call) do { rc = (call); if (rc < 0) { goto done; } } while (0)
	/* not zero-filled: every Krylov vector and the matvec result are written completely before they are read (the packed layout has
	 * no padding between blocks, CTB_BLOCK_ALIGN == 1) -- at D = 4096 the fill was a 1.4 GB memset per bond */
	LZ(ctbd_malloc_noinit(&V, (size_t)maxiter * (size_t)ns * esize));
	LZ(ctbd_malloc_noinit(&w_own, (size_t)ns * esize));
	/* scal: [0..maxiter) alpha (2 doubles each), [..] beta, then scratch */
	LZ(ctbd_malloc((void**)&scal, (size_t)(3 * maxiter + 4) * sizeof(double)));
	double* d_alpha = scal;                 /* stride 2 (re, im) */
	double* d_beta  = scal + 2 * maxiter;
	double* d_tmp   = scal + 3 * maxiter;
	alpha = ctb_calloc((size_t)maxiter, sizeof(double));
	beta  = ctb_calloc((size_t)maxiter, sizeof(double));
	host_scal = ctb_calloc((size_t)(3 * maxiter + 4), sizeof(double));
#define VJ(j) ((void*)((char*)V + (size_t)(j) * (size_t)ns * esize))

	/* v_0 = vstart / ||vstart|| */
	LZ(ctbd_nrm2(dtype, ns, a_start->d, d_tmp));
	LZ(ctbd_rscale(dtype, ns, a_start->d, d_tmp, 1, VJ(0)));

	/* All iterations are enqueued without waiting for the device: the breakdown test of the reference (beta_j < 100 n eps ends the
	 * iteration with numiter = j + 1, krylov.c:58) is evaluated afterwards on the coefficients -- whatever was computed past a
	 * breakdown is never read.  One synchronising copy per local solve instead of one per iteration (the small-bond regime is
	 * launch-bound: a matvec there takes tens of microseconds). */
	for (int j = 0; j < maxiter; j++)
	{
		/* sharded with the fused exchange: work in the landing buffer the peers have stored the result into (no copy) */
		void* wl = ctb_heff_result_buffer(h);
		void* w = (wl != NULL) ? wl : w_own;
		LZ(ctb_heff_apply(h, VJ(j), w));
		LZ(ctbd_dotc(dtype, ns, w, VJ(j), d_alpha + 2 * j));
		if (j == maxiter - 1) { break; }      /* the last iteration only contributes alpha */
		LZ(ctbd_lanczos_update(dtype, ns, w, VJ(j), j > 0 ? VJ(j - 1) : NULL, d_alpha + 2 * j, j > 0 ? d_beta + (j - 1) : NULL, d_beta + j));
		LZ(ctbd_rscale(dtype, ns, w, d_beta + j, 1, VJ(j + 1)));
	}
	LZ(ctbd_d2h(host_scal, scal, (size_t)(3 * maxiter) * sizeof(double)));
	for (int j = 0; j < maxiter; j++) { alpha[j] = host_scal[2 * j]; }
	for (int j = 0; j < maxiter - 1; j++) {
		beta[j] = host_scal[2 * maxiter + j];
		if (!(beta[j] >= 100 * (double)n * DBL_EPSILON)) { numiter = j + 1; break; }      /* also catches a NaN */
	}
	if (numiter_out != NULL) { *numiter_out = numiter; }

	if (numiter < 1) { rc = -1; }
	for (int j = 0; j < numiter && rc == 0; j++) {
		if (!isfinite(alpha[j]) || (j + 1 < numiter && !isfinite(beta[j]))) {
			fprintf(stderr, "chemtensor_b200: Lanczos produced a non-finite coefficient at iteration %d (alpha = %g, beta = %g, n = %lld)\n",
				j, alpha[j], j + 1 < numiter ? beta[j] : 0.0, (long long)n);
			rc = -1;
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
			for (int j = 0; j < numiter; j++) { coef[j] = z[j * numiter + 0]; }
			*a_opt = ctb_tensor_like(a_start, 1);
			rc = ctbd_lincomb(dtype, ns, V, ns, numiter, coef, (*a_opt)->d);
			ctb_free(coef);
		}
		else {
			fprintf(stderr, "chemtensor_b200: tridiagonal eigensolver failed to converge\n");
		}
		ctb_free(z);
	}
done:
#undef VJ
#undef LZ
	ctb_free(alpha); ctb_free(beta); ctb_free(host_scal);
	ctbd_free(scal);
	ctbd_free(w_own);
	ctbd_free(V);
	return rc;
}

/* ---- reference-signature Lanczos with a host callback (include/util/krylov.h:10-30) ----
 * The Krylov vectors live on the device; v_j is handed to the caller's 'afunc' through the host array 'v'
 * that the interface exposes anyway, and A v_j is pushed back for the fused device update. */
static void lanczos_host_callback(int dtype, const ct_long n, void (*afunc)(const ct_long, const void*, const void*, void*), const void* adata,
	const void* vstart, const int maxiter, double* alpha, double* beta, void* v, int* numiter)
{
	const size_t esize = ctb_sizeof_dtype(dtype);
	CTB_CHECK_ABORT(ctbd_init(-1));
	void* V = NULL; void* w = NULL; double* scal = NULL;
	CTB_CHECK_ABORT(ctbd_malloc(&V, (size_t)maxiter * (size_t)n * esize));
	CTB_CHECK_ABORT(ctbd_malloc(&w, (size_t)n * esize));
	CTB_CHECK_ABORT(ctbd_malloc((void**)&scal, (size_t)(3 * maxiter + 4) * sizeof(double)));
	double* d_alpha = scal; double* d_beta = scal + 2 * maxiter; double* d_tmp = scal + 3 * maxiter;
	void* w_host = ctb_malloc((size_t)n * esize);
#define VJ(j) ((void*)((char*)V + (size_t)(j) * (size_t)n * esize))
#define HV(j) ((void*)((char*)v + (size_t)(j) * (size_t)n * esize))
	CTB_CHECK_ABORT(ctbd_h2d(w, vstart, (size_t)n * esize));
	CTB_CHECK_ABORT(ctbd_nrm2(dtype, n, w, d_tmp));
	CTB_CHECK_ABORT(ctbd_rscale(dtype, n, w, d_tmp, 1, VJ(0)));
	CTB_CHECK_ABORT(ctbd_d2h(HV(0), VJ(0), (size_t)n * esize));
	*numiter = maxiter;
	for (int j = 0; j < maxiter; j++)
	{
		afunc(n, adata, HV(j), w_host);
		CTB_CHECK_ABORT(ctbd_h2d(w, w_host, (size_t)n * esize));
		CTB_CHECK_ABORT(ctbd_dotc(dtype, n, w, VJ(j), d_alpha + 2 * j));
		double a2[2];
		CTB_CHECK_ABORT(ctbd_d2h(a2, d_alpha + 2 * j, 2 * sizeof(double)));
		alpha[j] = a2[0];
		if (j == maxiter - 1) { break; }
		CTB_CHECK_ABORT(ctbd_lanczos_update(dtype, n, w, VJ(j), j > 0 ? VJ(j - 1) : NULL, d_alpha + 2 * j, j > 0 ? d_beta + (j - 1) : NULL, d_beta + j));
		CTB_CHECK_ABORT(ctbd_d2h(&beta[j], d_beta + j, sizeof(double)));
		if (beta[j] < 100 * (double)n * DBL_EPSILON) { *numiter = j + 1; break; }
		CTB_CHECK_ABORT(ctbd_rscale(dtype, n, w, d_beta + j, 1, VJ(j + 1)));
		CTB_CHECK_ABORT(ctbd_d2h(HV(j + 1), VJ(j + 1), (size_t)n * esize));
	}
#undef VJ
#undef HV
	ctb_free(w_host);
	CTB_CHECK_ABORT(ctbd_free(scal));
	CTB_CHECK_ABORT(ctbd_free(w));
	CTB_CHECK_ABORT(ctbd_free(V));
}

void lanczos_iteration_d(const ct_long n, lanczos_linear_func_d afunc, const void* adata, const double* vstart, const int maxiter,
	double* alpha, double* beta, double* v, int* numiter)
{
	lanczos_host_callback(CT_DOUBLE_REAL, n, (void (*)(const ct_long, const void*, const void*, void*))afunc, adata, vstart, maxiter, alpha, beta, v, numiter);
}

void lanczos_iteration_z(const ct_long n, lanczos_linear_func_z afunc, const void* adata, const void* vstart, const int maxiter,
	double* alpha, double* beta, void* v, int* numiter)
{
	lanczos_host_callback(CT_DOUBLE_COMPLEX, n, (void (*)(const ct_long, const void*, const void*, void*))afunc, adata, vstart, maxiter, alpha, beta, v, numiter);
}

static int eigensystem_krylov(int dtype, const ct_long n, void (*afunc)(const ct_long, const void*, const void*, void*), const void* adata,
	const void* vstart, const int maxiter, const int numeig, double* lambda, void* u_ritz)
{
	CTB_REQUIRE(numeig <= maxiter);
	const size_t esize = ctb_sizeof_dtype(dtype);
	double* alpha = ctb_calloc((size_t)maxiter, sizeof(double));
	double* beta  = ctb_calloc((size_t)maxiter, sizeof(double));
	void* v = ctb_malloc((size_t)maxiter * (size_t)n * esize);
	int numiter = 0;
	lanczos_host_callback(dtype, n, afunc, adata, vstart, maxiter, alpha, beta, v, &numiter);
	if (numiter < numeig) {
		fprintf(stderr, "Lanczos iteration stopped after %i iterations, cannot compute %i eigenvalues\n", numiter, numeig);
		ctb_free(v); ctb_free(beta); ctb_free(alpha);
		return -1;
	}
	double* z = ctb_malloc((size_t)numiter * numiter * sizeof(double));
	int rc = ctb_tridiag_eig(numiter, alpha, beta, z);
	if (rc < 0) {
		fprintf(stderr, "tridiagonal eigensolver failed, return value: %i\n", rc);
		ctb_free(z); ctb_free(v); ctb_free(beta); ctb_free(alpha);
		return -2;
	}
	memcpy(lambda, alpha, (size_t)numeig * sizeof(double));
	/* Ritz vectors u[:, e] = sum_j z[j, e] v_j on the device; output row-major n x numeig */
	void* V = NULL; void* col = NULL;
	CTB_CHECK(ctbd_malloc(&V, (size_t)numiter * (size_t)n * esize));
	CTB_CHECK(ctbd_malloc(&col, (size_t)n * esize));
	CTB_CHECK(ctbd_h2d(V, v, (size_t)numiter * (size_t)n * esize));
	double* coef = ctb_malloc((size_t)numiter * sizeof(double));
	void* col_host = ctb_malloc((size_t)n * esize);
	for (int e = 0; e < numeig; e++)
	{
		for (int j = 0; j < numiter; j++) { coef[j] = z[j * numiter + e]; }
		CTB_CHECK(ctbd_lincomb(dtype, n, V, n, numiter, coef, col));
		CTB_CHECK(ctbd_d2h(col_host, col, (size_t)n * esize));
		for (ct_long i = 0; i < n; i++) {
			memcpy((char*)u_ritz + ((size_t)i * numeig + e) * esize, (char*)col_host + (size_t)i * esize, esize);
		}
	}
	ctb_free(col_host); ctb_free(coef);
	CTB_CHECK(ctbd_free(col));
	CTB_CHECK(ctbd_free(V));
	ctb_free(z); ctb_free(v); ctb_free(beta); ctb_free(alpha);
	return 0;
}

int eigensystem_krylov_symmetric(const ct_long n, lanczos_linear_func_d afunc, const void* adata,
	const double* vstart, const int maxiter, const int numeig, double* lambda, double* u_ritz)
{
	return eigensystem_krylov(CT_DOUBLE_REAL, n, (void (*)(const ct_long, const void*, const void*, void*))afunc, adata, vstart, maxiter, numeig, lambda, u_ritz);
}

int eigensystem_krylov_hermitian(const ct_long n, lanczos_linear_func_z afunc, const void* adata,
	const void* vstart, const int maxiter, const int numeig, double* lambda, void* u_ritz)
{
	return eigensystem_krylov(CT_DOUBLE_COMPLEX, n, (void (*)(const ct_long, const void*, const void*, void*))afunc, adata, vstart, maxiter, numeig, lambda, u_ritz);
}
