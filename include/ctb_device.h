/*
 * ctb_device.h -- the thin C-ABI CUDA layer ("ctbd_*").
 *
 * Everything below is extern "C", plain pointers and sizes: no C++ or torch
 * types cross this boundary.  The host side of the engine (plain C, see
 * chemtensor_b200/host/) builds integer-only work lists ("plans") from
 * quantum-number sector metadata and hands them to this layer, which owns
 * device memory, the stream and the hand-written sm_100a kernels.
 *
 * Implemented by the .cu files under chemtensor_b200/csrc (the product).  A CPU test double
 * with the same symbols lives in tests/emu/ and is linked ONLY into the
 * host-logic test library (never into libchemtensor_b200.so).
 *
 * What each group replaces in the reference (file:line under the reference tree):
 *   grouped GEMM   : the per-block cblas_?gemm swarm issued by
 *                    block_sparse_tensor_dot, src/tensor/block_sparse_tensor.c:1935-1994
 *                    -> dense_tensor_dot_update, src/tensor/dense_tensor.c:1761-1828;
 *                    with the output permutation of the following
 *                    block_sparse_tensor_transpose (:785) fused into the epilogue
 *   remap          : block_sparse_tensor_transpose :785, _flatten_axes :950,
 *                    _split_axis :1123, _slice :1446, _multiply_pointwise_vector :1654
 *   level-1        : cblas_dnrm2/dscal/ddot/zdotc + the axpy loops of
 *                    lanczos_iteration_d/z, src/util/krylov.c:24-167, and the Ritz
 *                    vector GEMM at krylov.c:242/:335
 *   batched SVD/QR : LAPACK ?gesvd / ?geqrf+?orgqr / ?gerqf+?orgrq per block,
 *                    src/tensor/dense_tensor.c:2253, :2680, :3538
 *
 * All functions return 0 on success and a negative value on failure (the
 * reference's "<0" convention); ctbd_last_error() gives the message.
 */
#ifndef CTB_DEVICE_H
#define CTB_DEVICE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTBD_MAXDIM 8

/* dtype codes follow enum numeric_type; only the double types are computed on */
#define CTBD_F64 1
#define CTBD_C128 3

/* ---- runtime ------------------------------------------------------------------------------- */

/* select device (negative: keep current / use CTB_DEVICE env or 0), create the stream; idempotent */
int ctbd_init(int device);
int ctbd_shutdown(void);
/* 1 = CUDA sm_100a kernels, 2 = host test double */
int ctbd_backend(void);
const char* ctbd_last_error(void);
/* number of kernels this layer has launched since start (bench.py's gpu_launches) */
long long ctbd_launch_count(void);
/* number of SMs of the active device */
int ctbd_sm_count(void);
/* the cudaStream_t all work is enqueued on (opaque), for CUDA-event timing by the caller */
void* ctbd_stream(void);

/* device timing helpers (CUDA events on the layer's stream) */
int ctbd_event_create(void** ev);
int ctbd_event_record(void* ev);
int ctbd_event_elapsed_ms(void* ev_start, void* ev_stop, float* ms);   /* synchronises on ev_stop */
int ctbd_event_destroy(void* ev);

/* ---- memory -------------------------------------------------------------------------------- */

int ctbd_malloc(void** dptr, size_t bytes);          /* zero-initialised device memory */
int ctbd_malloc_noinit(void** dptr, size_t bytes);   /* the same without the zero fill: for buffers the caller overwrites completely */
int ctbd_free(void* dptr);
int ctbd_memset_zero(void* dptr, size_t bytes);
int ctbd_h2d(void* dptr, const void* hptr, size_t bytes);
int ctbd_d2h(void* hptr, const void* dptr, size_t bytes);
int ctbd_d2d(void* dst, const void* src, size_t bytes);
/* pack 'nblk' separately allocated host blocks into one device buffer (block b -> dptr + dst_off[b], byte units) and back,
 * streamed through the layer's persistent pinned staging ring; h2d returns once the host blocks have been consumed */
int ctbd_h2d_blocks(void* dptr, int nblk, const void* const* hptrs, const int64_t* dst_off, const int64_t* nbytes);
int ctbd_d2h_blocks(const void* dptr, int nblk, void* const* hptrs, const int64_t* src_off, const int64_t* nbytes);
/* fault in the pages of freshly allocated (not yet written) host blocks in parallel; their contents are undefined afterwards.  Lets
 * the caller take the page faults of a result download while the device is still computing that result */
int ctbd_host_prefault(int nblk, void* const* hptrs, const int64_t* nbytes);
int ctbd_sync(void);
int ctbd_host_alloc(void** hptr, size_t bytes);      /* pinned host staging memory */
int ctbd_host_free(void* hptr);
/* bytes currently allocated through ctbd_malloc */
long long ctbd_bytes_in_use(void);

/* ---- one process per GPU: the exchange step of the sharded effective Hamiltonian -------------------------------------
 * The collective is NCCL (all-gather over NVLink / NVSwitch) enqueued on the layer's stream.  libnccl is bound at run time
 * (dlopen "libnccl.so.2"), so a single-GPU host never needs it.  A host that owns its own communicator may instead register
 * a callback; the CPU test double only knows the callback form (gloo in the tests). */
#define CTBD_UNIQUE_ID_BYTES 128
typedef int (*ctbd_allgather_fn)(void* ctx, const void* sendbuf, void* recvbuf, size_t bytes_per_rank, void* stream);
int ctbd_dist_unique_id(void* id_out);                              /* rank 0: ncclGetUniqueId */
int ctbd_dist_init(int rank, int world, const void* unique_id);     /* ncclCommInitRank on the active device; unique_id may be NULL with a callback */
int ctbd_dist_set_allgather(ctbd_allgather_fn fn, void* ctx);       /* optional host-provided collective */
int ctbd_dist_finalize(void);
/* recv[p * bytes_per_rank ...] = send of rank p, for all ranks p; device buffers; ordered on the layer's stream */
int ctbd_allgather(const void* sendbuf, void* recvbuf, size_t bytes_per_rank);
/* all ranks have completed (and made visible) everything enqueued before the barrier; ordered on the layer's stream */
int ctbd_barrier(void);
/* peer-mapped buffer: every rank allocates 'bytes' and maps the buffers of all other ranks of the box into its address space
 * (CUDA IPC over NVLink peer access).  ptrs[p] is rank p's buffer as seen from this process (ptrs[rank] = the local one).  Collective
 * call.  Returns < 0 when peer mapping is not available (no NCCL communicator, no P2P, test double): callers then fall back to the
 * all-gather form of the exchange. */
int ctbd_peer_buffer_create(size_t bytes, void** handle);
int ctbd_peer_buffer_ptrs(void* handle, void** ptrs /* [world] */);
int ctbd_peer_buffer_destroy(void* handle);

/* NVSwitch multicast buffer: every rank allocates 'bytes'; the buffers of all ranks are bound to one multicast object.  *local_ptr is
 * this rank's buffer, *mc_ptr the multicast address: a (multimem) store to it lands in the buffers of ALL ranks, replicated by the
 * switch.  Collective call; < 0 on every rank when multicast is not available (callers fall back to the peer-mapped form). */
int ctbd_mc_buffer_create(size_t bytes, void** handle, void** local_ptr, void** mc_ptr);
int ctbd_mc_buffer_destroy(void* handle);

/* ---- grouped block GEMM -------------------------------------------------------------------- */

/* one contracted sector tuple of one output block: C += op(A_seg) * op(B_seg), inner extent k */
struct ctbd_gemm_seg
{
	int64_t a_off;    /* element offset of the A block in the A buffer */
	int64_t b_off;    /* element offset of the B block in the B buffer */
	int32_t k;        /* contracted extent (product of contracted sector multiplicities) */
	int32_t lda;      /* a_kcontig: A(i,kk) = a[a_off + i*lda + kk];  else A(i,kk) = a[a_off + kk*lda + i] */
	int32_t ldb;      /* b_ncontig: B(kk,j) = b[b_off + kk*ldb + j];  else B(kk,j) = b[b_off + j*ldb + kk] */
	int32_t pad_;
};

/* one output block: m x n, C(i,j) stored at c[c_off + tab[row_tab + i] + tab[col_tab + j]];
 * merged-row form (col_tab < 0): rows of several result blocks stacked, C(i,j) at c[c_off + tab[row_tab + i] + tab[tab[row_tab + m + i] + j]] */
struct ctbd_gemm_out
{
	int64_t c_off;
	int32_t m, n;
	int32_t seg_begin, seg_end;   /* [begin,end) into the segment array, accumulated in this order */
	int32_t row_tab, col_tab;     /* start indices into the int32 offset table */
};

/* ---- "mixing" form of a merged-row contraction with a small constant left operand (the MPO tensor) ----
 * One group = all result rows that share the free sectors of the right operand t.  For column j of the group
 * (a multi-index over the free axes of t, row-major, extents dig_dim[]) and row i:
 *     C[row.c_off + sum_a digit_a(j) * row.cs[a]]  =  sum_k  Apacked[row.a_off + k] * B[b_rowtab[group.brow_begin + k] + j],   k < kp
 * Exact zeros of the packed operand are skipped (x + 0 * y == x), so the work is proportional to the non-zero operator
 * entries; the traffic is one read of every B row block and one write of every C row block (HBM-bound by construction). */
struct ctbd_mix_group
{
	int64_t brow_begin;          /* first entry of this group in b_rowtab */
	int32_t n;                   /* number of columns = product of dig_dim */
	int32_t kp;                  /* contracted extent K' */
	int32_t row_begin, row_end;  /* rows of this group in the row array */
	int32_t ndig;                /* number of column digits (free axes of t), <= 4 */
	int32_t dig_dim[4];
	int32_t pad_;
};
struct ctbd_mix_row
{
	int64_t c_off;               /* element offset of C(i, column 0) */
	int64_t a_off;               /* start of the kp packed operand entries of this row */
	int32_t cs[4];               /* stride of each column digit in the result block of this row */
};

struct ctbd_gemm_plan_host
{
	int32_t dtype;                /* CTBD_F64 or CTBD_C128 */
	int32_t a_kcontig, b_ncontig; /* operand layouts, see ctbd_gemm_seg */
	int32_t conj_a, conj_b;       /* complex only: conjugate operand on load */
	int32_t nouts, nsegs, ntab;
	const struct ctbd_gemm_out*  outs;    /* host arrays; copied to the device by plan_create */
	const struct ctbd_gemm_seg*  segs;
	const int32_t* tab;
	double flops;                 /* algorithmic flops of one run: sum 2*m*n*k (x4 complex) */
	/* optional: the plan owns a packed copy of the (constant) A operand, a_packed[i] = a_gather[i] >= 0 ? a_src[a_gather[i]] : 0,
	 * built once at plan creation; segment a_off then index a_packed and the A argument of ctbd_gemm_run is ignored */
	const int64_t* a_gather;      /* host array or NULL */
	int64_t n_a_gather;
	const void* a_src;            /* device buffer the gather reads from */
	/* optional (needs b_ncontig): the rows of B are gathered through a row table, B(kk, j) = b[b_rowtab[seg.b_off + kk] + j];
	 * lets ONE pipeline step span the rows of many small blocks (the MPO-mixing contraction has 1-4 rows per block) */
	const int64_t* b_rowtab;      /* host array or NULL */
	int64_t n_b_rowtab;
	/* optional: mixing form (needs a_gather and b_rowtab); when present outs/segs/tab are empty and the run is one mix launch */
	const struct ctbd_mix_group* mix_groups;
	int32_t n_mix_groups;
	const struct ctbd_mix_row* mix_rows;
	int32_t n_mix_rows;
};

/* builds the device-resident work list: every output block is cut into tiles of the kernel variant that
 * fits it best (tile classes are an internal matter of the CUDA layer), heaviest tiles first */
int ctbd_gemm_plan_create(const struct ctbd_gemm_plan_host* h, void** plan);
int ctbd_gemm_plan_destroy(void* plan);
/* C (every output block of the plan is overwritten) = sum over segments op(A) op(B); A, B, C device buffers */
int ctbd_gemm_run(void* plan, const void* A, const void* B, void* C);
/* the same with the epilogue storing every output element to 'ndst' (<= 8) destination buffers: the GEMM fused with the all-gather
 * of its result over NVLink peer memory (Cs[p] = peer-mapped buffer of rank p) */
int ctbd_gemm_run_multi(void* plan, const void* A, const void* B, int ndst, void* const* Cs);
/* the same with every output element stored ONCE to a multicast address (ctbd_mc_buffer_create): the switch delivers it to all ranks */
int ctbd_gemm_run_mc(void* plan, const void* A, const void* B, void* C_mc);
/* number of tiles / kernel launches one run of the plan issues */
int ctbd_gemm_plan_info(void* plan, int* ntiles, int* nlaunches);

/* ---- packed block-sparse layout resident on the device, and logical-index remaps ------------ */

struct ctbd_layout_host
{
	int32_t ndim;
	int32_t dtype;
	int64_t dim[CTBD_MAXDIM];          /* logical dimension per axis */
	int32_t nsec[CTBD_MAXDIM];         /* number of sectors per axis */
	const int32_t* sec_of[CTBD_MAXDIM];   /* [dim]   sector index of a logical index */
	const int32_t* pos_of[CTBD_MAXDIM];   /* [dim]   position inside its sector */
	const int32_t* secstart[CTBD_MAXDIM]; /* [nsec+1] prefix sums of sector multiplicities */
	const int32_t* log_of[CTBD_MAXDIM];   /* [dim]   logical indices grouped by sector */
	int64_t ngrid;                     /* number of cells of the sector grid */
	const int64_t* grid_off;           /* [ngrid] element offset of the block, -1 if not conserved */
	int32_t nblk;                      /* number of stored blocks */
	const int64_t* blk_grid;           /* [nblk] grid cell of each stored block (ascending) */
	const int64_t* blk_off;            /* [nblk+1] element offsets (blk_off[nblk] = nstore) */
	int64_t nstore;                    /* stored elements incl. alignment padding */
};

int ctbd_layout_create(const struct ctbd_layout_host* h, void** layout);
int ctbd_layout_destroy(void* layout);

#define CTBD_REMAP_TRANSPOSE 0   /* dst axis i = src axis perm[i] */
#define CTBD_REMAP_FLATTEN   1   /* dst axis i_ax = src axes (i_ax, i_ax+1) fused row-major */
#define CTBD_REMAP_SPLIT     2   /* dst axes (i_ax, i_ax+1) = src axis i_ax split row-major */
#define CTBD_REMAP_SLICE     3   /* dst index j on axis i_ax = src index ind[j] */
#define CTBD_REMAP_IDENTITY  4   /* same logical index (used with 'scale') */
#define CTBD_REMAP_UNSLICE   5   /* scatter: dst index ind[j] on axis i_ax = src index j (other dst entries untouched) */

struct ctbd_remap_args
{
	int32_t op;
	int32_t i_ax;
	int32_t perm[CTBD_MAXDIM];
	const int64_t* ind;      /* host array of length dst dim[i_ax] (SLICE) */
	int32_t conj;            /* complex: conjugate while copying */
	int32_t scale_ax;        /* >= 0: multiply by scale[logical index on this dst axis] */
	const double* scale;     /* device vector (real), length dst dim[scale_ax] */
	void* dst_layout; void* dst;
	void* src_layout; const void* src;
};

int ctbd_remap(const struct ctbd_remap_args* args);

/* batched strided 2-d copies with a device-resident descriptor list (built once, run many times):
 * dst[d.dst_off + i * d.dst_ld + j] = src[d.src_off + i * d.src_ld + j], i < rows, j < cols.
 * Used to scatter the all-gathered column slices of the sharded effective Hamiltonian into the packed result. */
struct ctbd_copy2d
{
	int64_t src_off, dst_off;
	int32_t rows, cols, src_ld, dst_ld;
};
int ctbd_copy_plan_create(int dtype, int n, const struct ctbd_copy2d* descs_host, void** plan);
int ctbd_copy_plan_run(void* plan, const void* src, void* dst);
/* the same with the source split over 'nsrc' (<= 8) buffers: element offset o of the virtual source lives in srcs[o / src_stride] at
 * o % src_stride.  With peer-mapped buffers (ctbd_peer_buffer_*) this is the "pull" form of the exchange: one kernel reads the result
 * slices of all GPUs over NVLink (coalesced row runs) and writes the packed local result */
int ctbd_copy_plan_run_multi(void* plan, int nsrc, const void* const* srcs, int64_t src_stride, void* dst);
/* the "push" form: every copy is read once from the local 'src' and written to all 'ndst' (<= 8) destination buffers (peer-mapped
 * result buffers of the GPUs of the box): wide posted NVLink stores, whole row runs per warp */
int ctbd_copy_plan_run_push(void* plan, const void* src, int ndst, void* const* dsts);
int ctbd_copy_plan_destroy(void* plan);

/* ---- batched strided block linear combinations ------------------------------------------------
 * The data movement of the SU(2) layer (host/su2.c): F-moves (a result degeneracy tensor is a recoupling-coefficient weighted sum
 * of source degeneracy tensors of the same shape, reference src/tensor/su2_tensor.c:770-886), axis permutations of degeneracy tensors
 * (:647), the per-sector scalings of axis reversals and swaps (:916, :467), stacking of sector blocks into the matrices the batched
 * QR / SVD kernels factorise, and the (de)normalisation of Lanczos vectors (:4725).  For every block and every multi-index i over
 * dim[] (row-major enumeration):
 *     dst[dst_off + sum_a i_a * dstride[a]]  =  sum_{t in [term_begin, term_end)}  coef_t * op(src[src_off_t + sum_a i_a * sstride[a]])
 * op = complex conjugation when the plan was created with conj != 0.  A block without terms is zero-filled.  src may equal dst
 * only when both index maps are the same (in-place scaling).  HBM-bound: (terms + 1) x block bytes. */
#define CTBD_LC_MAXDIM 8
struct ctbd_lc_term
{
	int64_t src_off;
	double coef;
};
struct ctbd_lc_block
{
	int64_t dst_off;
	int32_t term_begin, term_end;
	int32_t ndim;                        /* 1 .. CTBD_LC_MAXDIM */
	int32_t pad_;
	int32_t dim[CTBD_LC_MAXDIM];
	int64_t dstride[CTBD_LC_MAXDIM];
	int64_t sstride[CTBD_LC_MAXDIM];
};
int ctbd_lc_plan_create(int dtype, int conj, int nblk, const struct ctbd_lc_block* blocks_host, int nterm, const struct ctbd_lc_term* terms_host, void** plan);
int ctbd_lc_plan_run(void* plan, const void* src, void* dst);
int ctbd_lc_plan_destroy(void* plan);

/* ---- CUDA graph of a launch sequence ---------------------------------------------------------
 * The recorded effective-Hamiltonian program of the SU(2) layer (host/su2_core.c) is ~15 launches of a few microseconds each in the
 * small-bond regime; captured once per local solve it is replayed with ONE graph launch per Lanczos iteration.  Only kernel launches
 * of this layer may be issued between begin and end (ctbd_gemm_run, ctbd_lc_plan_run): no allocation, copy to the host or synchronisation.
 * Return value > 0 means "not available / capture discarded": the caller then replays launch by launch (the same kernels). */
int ctbd_graph_capture_begin(void);
int ctbd_graph_capture_end(void** graph);
int ctbd_graph_launch(void* graph);
int ctbd_graph_destroy(void* graph);

/* ---- level-1 kernels for the Lanczos iteration (scalars stay on the device) ----------------- */

/* out[0] = Re sum conj(x_i) y_i, out[1] = Im (0 for real) */
int ctbd_dotc(int dtype, int64_t n, const void* x, const void* y, double* out_dev);
/* out[0] = sqrt(sum |x_i|^2) */
int ctbd_nrm2(int dtype, int64_t n, const void* x, double* out_dev);
/* y = x / s[0]   (divide != 0) or y = x * s[0]; x may equal y */
int ctbd_rscale(int dtype, int64_t n, const void* x, const double* s_dev, int divide, void* y);
/* w -= alpha[0]*vj + beta_prev[0]*vjm1 (vjm1 may be NULL); out[0] = ||w|| afterwards */
int ctbd_lanczos_update(int dtype, int64_t n, void* w, const void* vj, const void* vjm1,
	const double* alpha_dev, const double* beta_prev_dev, double* out_dev);
/* out = sum_{j<m} coef[j] * V[j*ldv ...]; coef is a HOST array of m reals */
int ctbd_lincomb(int dtype, int64_t n, const void* V, int64_t ldv, int m, const double* coef_host, void* out);
/* x *= alpha (host scalar, real) */
int ctbd_scale_host(int dtype, int64_t n, void* x, double alpha);
/* x *= (re + i im) for complex128 entries, x *= re for real ones (phase absorption of mps_compress, reference src/state/mps.c:938-977) */
int ctbd_zscale_host(int dtype, int64_t n, void* x, double re, double im);

/* ---- batched dense factorizations of the sector blocks -------------------------------------- */

struct ctbd_mat_desc
{
	int64_t a_off;      /* m x n row-major input block */
	int32_t m, n;
	int64_t o0_off;     /* SVD: U (m x k);  QR: Q (m x k);  RQ: R (m x k)   [k = min(m,n)] */
	int64_t o1_off;     /* SVD: Vh (k x n); QR: R (k x n);  RQ: Q (k x n) */
	int64_t s_off;      /* SVD: index of the first singular value of this block in S */
};

/* per block: A = U diag(S) Vh, singular values descending */
int ctbd_svd_batched(int dtype, int nmat, const struct ctbd_mat_desc* descs_host,
	const void* A, void* U, void* Vh, double* S_dev);
/* The big-block SVD in pieces, for the GEMM-driven block-Jacobi stage of the host (host/svd_block.c):
 *   create : every block gets a work matrix [G | W] (R x (C + R) row-major, R = min(m, n), C = max(m, n)): G = the block (its
 *            conjugate transpose when m > n) normalised by a power of two, W = identity.  The work matrices are packed in
 *            descriptor order into ONE device buffer (*G, *g_total elements); the host applies unitary row operations to them
 *            (the same on G and W), possibly into a second buffer of the same size;
 *   finish : takes the buffer that holds the current work matrices, optionally polishes them with the in-kernel block-Jacobi
 *            tournament until no rotation is left, then sorts / normalises and writes U, Vh, S as ctbd_svd_batched does; releases ws.
 *   gram_offdiag : convergence measure over a list of small square matrices G_k (element offsets / dimensions on the device):
 *            out[1] = max(out[1], max_i |G_ii|), out[0] = max(out[0], max_{i != j} |G_ij|^2 / max(|G_ii| |G_jj|, (floor_rel out[1])^2)) */
int ctbd_svdws_create(int dtype, int nmat, const struct ctbd_mat_desc* descs_host, const void* A, void** ws, void** G, int64_t* g_total);
int ctbd_svdws_finish(void* ws, const void* G_cur, int polish, void* U, void* Vh, double* S_dev);
int ctbd_gram_offdiag(int dtype, int ngram, const int64_t* off_dev, const int32_t* dim_dev, const void* G, double floor_rel, double* out_dev);
/* Singular-value selection of a split on the device (reference src/algorithm/truncation.c:110-223, same integer / ordering and
 * floating-point semantics, equal values ordered by index): S_dev holds n singular values.  Returns on the HOST what the tensor
 * metadata needs -- the ascending list of retained indices (ind_host, capacity n), their number, and info3 = { norm_sigma, entropy,
 * tol_eff } -- and on the DEVICE the retained values, rescaled to the norm of all values when renormalize != 0 (s_ret_dev,
 * capacity n).  The singular values themselves do not cross to the host. */
int ctbd_truncate_select(int64_t n, const double* S_dev, double tol, int relative, int64_t max_vdim, int renormalize,
	int64_t* nret, int64_t* ind_host, double* info3_host, double* s_ret_dev);
/* rq == 0: A = Q R (Q: m x k isometry, R upper triangular);  rq != 0: A = R Q (Q: k x n) */
int ctbd_qr_batched(int dtype, int rq, int nmat, const struct ctbd_mat_desc* descs_host,
	const void* A, void* O0, void* O1);

#ifdef __cplusplus
}
#endif

#endif
