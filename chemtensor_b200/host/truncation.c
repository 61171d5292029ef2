/*
 * truncation.c -- singular-value selection rule (host, tiny data, semantics critical).
 *
 * Restates reference src/algorithm/truncation.c:110-223 (retained_bond_indices) and :13-27
 * (von_neumann_entropy), SURVEY.md §9.7:
 *   sort sigma ascending, square, optionally normalise by the total, running sum from the smallest;
 *   with max_vdim < n the n - max_vdim smallest running sums are zeroed and tol_eff is raised;
 *   keep index i (original order) iff its running sum > tol.
 * The only data that leaves the device for this step is the vector of singular values.
 * Ties: the reference uses libc qsort (unspecified order among equal values); here equal values
 * are ordered by index, which is one of the orders qsort may produce.
 */
#include "ctb_internal.h"

struct sv_item { double v; ct_long i; };

static int cmp_sv_item(const void* a, const void* b)
{
	const struct sv_item* x = a; const struct sv_item* y = b;
	if (x->v < y->v) { return -1; }
	if (x->v > y->v) { return  1; }
	return (x->i > y->i) - (x->i < y->i);
}

double ctb_von_neumann_entropy(const double* sigma, ct_long n)
{
	double s = 0;
	for (ct_long i = 0; i < n; i++) {
		if (sigma[i] > 0) {
			const double sq = sigma[i] * sigma[i];
			s -= sq * log(sq);
		}
	}
	return s;
}

void ctb_retained_bond_indices(const double* sigma, ct_long n, double tol, bool relative_thresh, ct_long max_vdim,
	struct index_list* list, struct trunc_info* info)
{
	info->tol_eff = tol;
	list->ind = NULL;
	list->num = 0;
	info->norm_sigma = 0;
	info->entropy = 0;

	struct sv_item* srt = ctb_malloc((size_t)n * sizeof(struct sv_item));
	for (ct_long i = 0; i < n; i++) { srt[i].v = sigma[i]; srt[i].i = i; }
	qsort(srt, (size_t)n, sizeof(struct sv_item), cmp_sv_item);

	double sqsum = 0;
	for (ct_long i = 0; i < n; i++) {
		srt[i].v = srt[i].v * srt[i].v;
		sqsum += srt[i].v;
	}
	if (sqsum == 0) { ctb_free(srt); return; }
	if (relative_thresh) {
		for (ct_long i = 0; i < n; i++) { srt[i].v /= sqsum; }
	}
	for (ct_long i = 1; i < n; i++) { srt[i].v += srt[i - 1].v; }
	if (max_vdim < n)
	{
		info->tol_eff = fmax(tol, srt[n - max_vdim - 1].v);
		for (ct_long i = 0; i < n - max_vdim; i++) { srt[i].v = 0; }
	}
	double* accum = ctb_malloc((size_t)n * sizeof(double));
	for (ct_long i = 0; i < n; i++) { accum[srt[i].i] = srt[i].v; }
	ctb_free(srt);

	list->ind = ctb_malloc((size_t)n * sizeof(ct_long));
	for (ct_long i = 0; i < n; i++) {
		if (accum[i] > tol) { list->ind[list->num++] = i; }
	}
	ctb_free(accum);
	if (list->num == 0) { ctb_free(list->ind); list->ind = NULL; return; }

	/* norm and entropy of the retained, normalised singular values */
	double* kept = ctb_malloc((size_t)list->num * sizeof(double));
	double scale = 0, ssq = 1;   /* scaled 2-norm as in BLAS nrm2 */
	for (ct_long i = 0; i < list->num; i++)
	{
		kept[i] = sigma[list->ind[i]];
		const double a = fabs(kept[i]);
		if (a > 0) {
			if (scale < a) { ssq = 1 + ssq * (scale / a) * (scale / a); scale = a; }
			else { ssq += (a / scale) * (a / scale); }
		}
	}
	info->norm_sigma = scale * sqrt(ssq);
	for (ct_long i = 0; i < list->num; i++) { kept[i] /= info->norm_sigma; }
	info->entropy = ctb_von_neumann_entropy(kept, list->num);
	ctb_free(kept);
}
