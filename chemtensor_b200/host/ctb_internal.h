/*
 * ctb_internal.h -- host-side (plain C) engine internals.
 *
 * The engine keeps every tensor of the DMRG hot path resident on the device in
 * a PACKED block layout: the stored (quantum-number conserving) blocks are
 * concatenated in row-major order of the sector grid, each block row-major.
 * With CTB_BLOCK_ALIGN == 1 this is exactly the order produced by the
 * reference's block_sparse_tensor_serialize_entries
 * (src/tensor/block_sparse_tensor.c:3131-3150), so a Lanczos vector IS the
 * storage of the two-site tensor and no (de)serialisation happens per matvec.
 *
 * All sector bookkeeping below is integer-only and reproduces the reference's
 * structural rules bit-exactly.
 */
#ifndef CTB_INTERNAL_H
#define CTB_INTERNAL_H

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "ctb_types.h"
#include "ctb_device.h"

#define CTB_MAXDIM CTBD_MAXDIM
/* alignment (in elements) of each stored block inside the packed device buffer */
#define CTB_BLOCK_ALIGN 1

/* largest total contracted extent for which a merged-row contraction uses the streaming "mixing" kernel instead of the tensor-pipe GEMM */
#define CTB_MIX_KMAX 128

#define CTB_CHECK(call) do { int ctb_rc_ = (call); if (ctb_rc_ < 0) { \
	fprintf(stderr, "chemtensor_b200: %s failed at %s:%d: %s\n", #call, __FILE__, __LINE__, ctbd_last_error()); \
	return ctb_rc_; } } while (0)

#define CTB_CHECK_ABORT(call) do { int ctb_rc_ = (call); if (ctb_rc_ < 0) { \
	fprintf(stderr, "chemtensor_b200: %s failed at %s:%d: %s\n", #call, __FILE__, __LINE__, ctbd_last_error()); \
	abort(); } } while (0)

#define CTB_REQUIRE(cond) do { if (!(cond)) { \
	fprintf(stderr, "chemtensor_b200: requirement '%s' violated at %s:%d\n", #cond, __FILE__, __LINE__); \
	abort(); } } while (0)

static inline size_t ctb_sizeof_dtype(int dtype)
{
	switch (dtype) {
		case CT_SINGLE_REAL:    return 4;
		case CT_DOUBLE_REAL:    return 8;
		case CT_SINGLE_COMPLEX: return 8;
		case CT_DOUBLE_COMPLEX: return 16;
		default: return 0;
	}
}

static inline int ctb_is_complex(int dtype) { return dtype == CT_SINGLE_COMPLEX || dtype == CT_DOUBLE_COMPLEX; }

/* 16-byte aligned host allocation, compatible with the reference's ct_malloc / free() pairing */
static inline void* ctb_malloc(size_t size)
{
	if (size == 0) { size = 1; }
	return aligned_alloc(16, (size + 15) & ~(size_t)15);
}
static inline void* ctb_calloc(size_t num, size_t size)
{
	void* p = ctb_malloc(num * size);
	if (p != NULL) { memset(p, 0, num * size); }
	return p;
}
static inline void ctb_free(void* p) { free(p); }

/* ---- one tensor leg: logical quantum numbers and the derived sector structure ---- */
struct ctb_axis
{
	ct_long dim;        /* logical dimension */
	int dir;            /* +1 out, -1 in */
	int nsec;           /* number of sectors = distinct quantum numbers */
	qnumber* qlog;      /* [dim]    logical quantum numbers */
	qnumber* qsec;      /* [nsec]   sorted distinct quantum numbers */
	int32_t* secdim;    /* [nsec]   multiplicities */
	int32_t* sec_of;    /* [dim]    sector of each logical index */
	int32_t* pos_of;    /* [dim]    position within the sector (order of appearance) */
	int32_t* secstart;  /* [nsec+1] prefix sums of secdim */
	int32_t* log_of;    /* [dim]    logical indices grouped by sector */
};

/* ---- device-resident block-sparse tensor ---- */
struct ctb_tensor
{
	int dtype;
	int ndim;
	struct ctb_axis ax[CTB_MAXDIM];
	ct_long ngrid;       /* cells of the sector grid */
	ct_long* grid_off;   /* [ngrid] element offset of the block or -1; NULL for a large grid, which keeps the hash index below instead */
	ct_long* gh_key;     /* large grids (ngrid > CTB_GRID_DENSE_MAX): open-addressing index cell -> element offset over the stored blocks */
	ct_long* gh_val;
	ct_long  gh_cap;     /* power of two, 0 when the dense table is used */
	int nblk;            /* stored blocks */
	ct_long* blk_grid;   /* [nblk] grid cell per stored block, ascending */
	ct_long* blk_off;    /* [nblk+1] */
	ct_long nelem;       /* stored entries without padding (reference: num_elements_blocks) */
	ct_long nstore;      /* elements of the device buffer */
	void* d;             /* device buffer (owned unless 'borrowed') */
	int borrowed;
	void* layout;        /* lazily created device-side layout tables */
};

/* tensor_meta.c */
void ctb_axis_init(struct ctb_axis* ax, ct_long dim, int dir, const qnumber* qlog);
void ctb_axis_copy(struct ctb_axis* dst, const struct ctb_axis* src);
void ctb_axis_free(struct ctb_axis* ax);
int  ctb_axis_find_sector(const struct ctb_axis* ax, qnumber q);
bool ctb_axis_same_qnums(const struct ctb_axis* a, const struct ctb_axis* b);

/* build a tensor from axes (axes are MOVED into the tensor); allocates a zeroed device buffer iff alloc != 0 */
struct ctb_tensor* ctb_tensor_from_axes(int dtype, int ndim, struct ctb_axis* axes, int alloc);
struct ctb_tensor* ctb_tensor_create(int dtype, int ndim, const ct_long* dim, const int* dirs, const qnumber* const* qnums, int alloc);
struct ctb_tensor* ctb_tensor_like(const struct ctb_tensor* t, int alloc);
struct ctb_tensor* ctb_tensor_clone(const struct ctb_tensor* t);
void ctb_tensor_free(struct ctb_tensor* t);
void* ctb_tensor_layout(struct ctb_tensor* t);
void ctb_grid_unravel(const struct ctb_tensor* t, ct_long cell, int* idx);
ct_long ctb_grid_ravel(const struct ctb_tensor* t, const int* idx);
/* sector grids beyond this many cells are not tabulated: the 6-leg intermediates of a molecular bond have 10^7 cells and 10^3 - 10^4
 * stored blocks (allocating and clearing a dense table per plan was a third of the plan-building time at 24 orbitals).  The 5-leg
 * intermediates of the D = 4096 Fermi-Hubbard sweep (1.2 M cells) stay below the bound and keep the table. */
#define CTB_GRID_DENSE_MAX ((ct_long)1 << 22)
/* element offset of the stored block of grid cell 'cell', or -1 */
static inline ct_long ctb_grid_offset(const struct ctb_tensor* t, ct_long cell)
{
	if (t->grid_off != NULL) { return t->grid_off[cell]; }
	if (t->gh_cap == 0) { return -1; }
	uint64_t h = (uint64_t)cell * 0x9E3779B97F4A7C15ull;
	ct_long q = (ct_long)(h >> 20) & (t->gh_cap - 1);
	while (t->gh_key[q] >= 0) {
		if (t->gh_key[q] == cell) { return t->gh_val[q]; }
		q = (q + 1) & (t->gh_cap - 1);
	}
	return -1;
}
bool ctb_tensor_same_structure(const struct ctb_tensor* a, const struct ctb_tensor* b);

/* host struct <-> device tensor */
struct ctb_tensor* ctb_upload(const struct block_sparse_tensor* h);
int  ctb_download(const struct ctb_tensor* t, struct block_sparse_tensor* h);   /* allocates the host payload */
/* the two halves of ctb_upload / ctb_download, for callers that overlap the payload copies with other work */
struct ctb_tensor* ctb_upload_begin(const struct block_sparse_tensor* h);
int  ctb_upload_data(struct ctb_tensor* t, const struct block_sparse_tensor* h);
void ctb_download_begin(const struct ctb_tensor* t, struct block_sparse_tensor* h, int prefault);
int  ctb_download_data(const struct ctb_tensor* t, struct block_sparse_tensor* h);
int  ctb_upload_entries(struct ctb_tensor* t, const void* entries);     /* packed entries, serialize order */
int  ctb_download_entries(const struct ctb_tensor* t, void* entries);
int  ctb_set_entry(struct ctb_tensor* t, ct_long offset, double re, double im);

/* host struct helpers (host_structs.c) */
void ctb_host_allocate_bst(int dtype, int ndim, const ct_long* dim, const enum tensor_axis_direction* axis_dir, const qnumber* const* qnums, struct block_sparse_tensor* t);

/* ---- contraction plans (tensor_ops.c) ---- */
struct ctb_dot_plan
{
	void* dev;          /* device-resident plan (ctbd_gemm_plan_create) */
	double flops;       /* algorithmic flops per execution */
	int ntiles, nouts, nsegs;
};

/* r = transpose(dot(s, t), perm); perm == NULL means identity.  conj_s/conj_t conjugate an operand on load. */
struct ctb_tensor* ctb_dot_prepare(const struct ctb_tensor* s, int axrange_s, int conj_s,
	const struct ctb_tensor* t, int axrange_t, int conj_t, int ndim_mult, const int* perm,
	int alloc_result, struct ctb_dot_plan* plan);
/* flags of ctb_dot_prepare_ex */
#define CTB_DOT_MERGE_ROWS 1   /* s is small and constant for the life of the plan: all result blocks that share the free sectors of t
                                  become ONE tall-skinny GEMM against a packed copy of s, so every block of t is read once */
struct ctb_tensor* ctb_dot_prepare_ex(const struct ctb_tensor* s, int axrange_s, int conj_s,
	const struct ctb_tensor* t, int axrange_t, int conj_t, int ndim_mult, const int* perm,
	int alloc_result, int flags, struct ctb_dot_plan* plan);
/* the result (natural axis order, no buffer) is written into the packed layout of the larger tensor 'full', which has the same axes
 * except 'axis', whose logical index j of the result corresponds to logical index ind[j] of 'full' */
struct ctb_embed { int axis; const struct ctb_tensor* full; const ct_long* ind; };
struct ctb_tensor* ctb_dot_prepare_embed(const struct ctb_tensor* s, int axrange_s, int conj_s,
	const struct ctb_tensor* t, int axrange_t, int conj_t, int ndim_mult, const struct ctb_embed* emb, struct ctb_dot_plan* plan);
int  ctb_dot_exec(const struct ctb_dot_plan* plan, const void* s_data, const void* t_data, void* r_data);
/* the same with the result stored to 'ndst' buffers (peer-mapped result buffers of all GPUs: GEMM fused with its all-gather) */
int  ctb_dot_exec_multi(const struct ctb_dot_plan* plan, const void* s_data, const void* t_data, int ndst, void* const* r_datas);
/* the same with one multimem store per element to the NVSwitch multicast address r_mc (ctbd_mc_buffer_create) */
int  ctb_dot_exec_mc(const struct ctb_dot_plan* plan, const void* s_data, const void* t_data, void* r_mc);
void ctb_dot_plan_free(struct ctb_dot_plan* plan);
struct ctb_tensor* ctb_dot(const struct ctb_tensor* s, int axrange_s, int conj_s,
	const struct ctb_tensor* t, int axrange_t, int conj_t, int ndim_mult, const int* perm);

/* logical-index remaps */
struct ctb_tensor* ctb_transpose(struct ctb_tensor* t, const int* perm, int conj);
struct ctb_tensor* ctb_flatten_axes(struct ctb_tensor* t, int i_ax, int new_dir);
struct ctb_tensor* ctb_split_axis(struct ctb_tensor* t, int i_ax, const ct_long new_dim[2], const int new_dir[2], const qnumber* const new_qnums[2]);
struct ctb_tensor* ctb_slice(struct ctb_tensor* t, int i_ax, const ct_long* ind, ct_long nind);
struct ctb_tensor* ctb_scale_axis(struct ctb_tensor* t, int i_ax, const double* scale_dev);
/* drop 'ntrace' leading and trailing axes, all of logical dimension 1 (cyclic partial trace of dummy bonds) */
struct ctb_tensor* ctb_drop_dummy_axes(const struct ctb_tensor* t, int ntrace);
struct ctb_tensor* ctb_cyclic_partial_trace(struct ctb_tensor* t, int ntrace);

/* block-wise factorizations of a block-sparse matrix */
int ctb_svd(struct ctb_tensor* a, struct ctb_tensor** u, double** s_dev, ct_long* ns, struct ctb_tensor** vh);
int ctb_qr(struct ctb_tensor* a, struct ctb_tensor** q, struct ctb_tensor** r);
int ctb_rq(struct ctb_tensor* a, struct ctb_tensor** r, struct ctb_tensor** q);

/* ---- algorithms on device tensors (chain.c, krylov.c, dmrg.c) ---- */
int ctb_split_matrix_svd(struct ctb_tensor* a, double tol, bool relative_thresh, ct_long max_vdim, bool renormalize,
	int svd_distr, struct ctb_tensor** a0, struct ctb_tensor** a1, struct trunc_info* info);
struct ctb_tensor* ctb_mps_merge_pair(const struct ctb_tensor* a0, const struct ctb_tensor* a1);
int ctb_mps_split_svd(struct ctb_tensor* a, const ct_long d[2], const qnumber* const new_qsite[2], double tol, ct_long max_vdim,
	bool renormalize, int svd_distr, struct ctb_tensor** a0, struct ctb_tensor** a1, struct trunc_info* info);
int ctb_mps_local_qr(struct ctb_tensor** a, struct ctb_tensor** a_next);
int ctb_mps_local_rq(struct ctb_tensor** a, struct ctb_tensor** a_prev);
/* metadata copy that shares the device buffer, with all axis directions reversed (the bra side of a contraction) */
struct ctb_tensor* ctb_view_reversed_dirs(const struct ctb_tensor* t);
struct ctb_tensor* ctb_mpo_merge_pair(const struct ctb_tensor* w0, const struct ctb_tensor* w1);
struct ctb_tensor* ctb_dummy_block_right(const struct ctb_tensor* a, const struct ctb_tensor* b, const struct ctb_tensor* w);
struct ctb_tensor* ctb_dummy_block_left(const struct ctb_tensor* a, const struct ctb_tensor* b, const struct ctb_tensor* w);
struct ctb_tensor* ctb_env_step_right(const struct ctb_tensor* a, const struct ctb_tensor* b, const struct ctb_tensor* w, const struct ctb_tensor* r);
struct ctb_tensor* ctb_env_step_left(const struct ctb_tensor* a, const struct ctb_tensor* b, const struct ctb_tensor* w, const struct ctb_tensor* l);

/* effective Hamiltonian of one bond: three cached grouped-GEMM plans + workspaces */
struct ctb_heff
{
	struct ctb_dot_plan p1, p2, p3;
	struct ctb_tensor* t1;   /* [dd, Dw', Dl, Dr', 1] */
	struct ctb_tensor* t2;   /* [Dl, Dw, dd, Dr', 1]  */
	/* pair form (ctb_heff_prepare_pair): the two site tensors w_second (site i+1) and w (site i) are applied one after the other,
	 * a and the result keep the physical legs apart: t1 [d2, Dw'', Dl, d1, Dr', 1] -> tm [d1, Dw', Dl, d2, Dr', 1] -> t2 [Dl, Dw, d1, d2, Dr', 1] */
	struct ctb_dot_plan p2a;
	struct ctb_tensor* tm;
	const struct ctb_tensor* w_second;
	struct ctb_tensor* k;    /* transpose(l, [0,3,1,2]) computed once per bond */
	struct ctb_tensor* b;    /* structure of the result (no buffer) */
	const struct ctb_tensor* w;
	const struct ctb_tensor* r;
	double flops;            /* algorithmic flops per matvec (of this rank's shard when sharded) */
	ct_long n;               /* logical number of entries of a (Lanczos vector length) */
	ct_long nstore;
	/* one process per GPU: the bra bond of r is cut into 'world' balanced index sets (every sector split evenly), rank p
	 * contracts with its column slice of r and owns the matching slice of the result; one all-gather + scatter rebuilds b */
	int world, rank;
	struct ctb_tensor* r_own;     /* slice(r, axis 2, ind[rank]) (owned) */
	struct ctb_tensor** piece;    /* [world] structure of every rank's result slice (no buffers) */
	ct_long** ind;                /* [world] logical indices of the bra bond owned by each rank, ascending */
	ct_long* nind;
	ct_long piece_cap;            /* elements of one all-gather slot (largest piece) */
	void* send; void* recv;       /* device: own piece; all pieces */
	void* scatter;                /* device copy plan: gathered pieces -> packed layout of b */
	int push;                     /* 1: local step 3, then one copy kernel stores the slice into the landing buffers of all ranks */
	void* push_plan;              /* device copy plan: own piece -> packed layout of b */
	int pull;                     /* 1: slices stay in peer-mapped send buffers, one kernel per rank pulls them over NVLink after the barrier */
	int fused;                    /* 1: step 3 stores straight into the peer-mapped result buffers of all ranks (no all-gather) */
	struct ctb_tensor* bfull5;    /* fused: 5-leg view of the full result the embedded step-3 plan writes into */
	double flops_total;           /* algorithmic flops of the whole matvec (all ranks) */
};
/* rank / world of this process (ctb_dist_init); world == 1 means no sharding */
extern int ctb_dist_rank, ctb_dist_world;
extern int ctb_collective_upload;      /* see ctb_upload_data */
int  ctb_heff_prepare(const struct ctb_tensor* a, const struct ctb_tensor* w, struct ctb_tensor* l, const struct ctb_tensor* r, struct ctb_heff* h);
/* the same; 'l_ready(ctx)' (may be NULL) is called right before the first device use of l's payload, so that a caller can build the
 * plans of steps 1 and 2 (which need structure only, and the payload of w) while the payloads of a, l and r are still in flight */
int  ctb_heff_prepare_ex(const struct ctb_tensor* a, const struct ctb_tensor* w, struct ctb_tensor* l, const struct ctb_tensor* r, struct ctb_heff* h,
	void (*l_ready)(void*), void* ctx);
int  ctb_heff_prepare_pair(const struct ctb_tensor* a4, const struct ctb_tensor* w0, const struct ctb_tensor* w1, struct ctb_tensor* l, const struct ctb_tensor* r, struct ctb_heff* h);
int  ctb_heff_apply(struct ctb_heff* h, const void* a_data, void* b_data);
int  ctb_heff_step3(struct ctb_heff* h, void* b_data);
void* ctb_heff_result_buffer(const struct ctb_heff* h);
void ctb_dist_release_buffers(void);
void ctb_dist_counters(long long* fused, long long* allgather);
long long ctb_dist_multicast_count(void);
long long ctb_dist_pull_count(void);
long long ctb_dist_push_count(void);
/* sharded case: all-gather of the result slices (h->send) and scatter into b_data, or (fused) barrier + local copy; no-op on one rank */
int  ctb_heff_exchange(struct ctb_heff* h, void* b_data);
void ctb_heff_free(struct ctb_heff* h);

/* symmetric tridiagonal eigen-decomposition (implicit QL); eigenvalues ascending in d, vectors in columns of z (row-major n x n) */
int ctb_tridiag_eig(int n, double* d, double* e, double* z);
/* Lanczos ground state of Heff on the device: a_opt has the structure of a_start */
int ctb_lanczos_min(struct ctb_heff* h, const struct ctb_tensor* a_start, int maxiter, double* en_min, struct ctb_tensor** a_opt, int* numiter_out);

/* truncation.c */
double ctb_von_neumann_entropy(const double* sigma, ct_long n);
void ctb_retained_bond_indices(const double* sigma, ct_long n, double tol, bool relative_thresh, ct_long max_vdim,
	struct index_list* list, struct trunc_info* info);

/* statistics of the last dmrg call (read by bench through ctb_get_stats) */
struct ctb_stats
{
	double heff_flops;       /* algorithmic flops spent in Heff matvecs */
	long long heff_calls;
	double heff_ms;          /* device time in Heff matvecs (only when CTB_PROFILE=1: adds event syncs) */
	double env_flops;
	double svd_ms, lanczos_ms, env_ms, total_ms;   /* host wall-clock per phase (includes device sync at phase end) */
	long long max_vector_len;
	long long max_bond_dim;
	double plan_ms;          /* host time spent building contraction plans (work lists + their upload), part of the phases above */
	double remap_ms;         /* host time spent in re-blocking calls (transpose / flatten / split / slice), part of the phases above */
	double sweep_ms[8];      /* wall-clock of the first eight sweeps of the last dmrg_twosite / dmrg_singlesite call (device synchronised at the end of each) */
};
extern struct ctb_stats ctb_global_stats;
extern double ctb_plan_profile[16];

double ctb_wall_ms(void);

#endif
