/*
 * chemtensor_b200.h -- the drop-in C boundary of the B200 engine.
 *
 * Every function below has the SAME name, signature, argument meaning, ownership
 * and error convention (0 ok, <0 failure, message on stderr) as the function of
 * qc-tum/chemtensor it replaces, so that the reference's own callers
 * (perf/perf_dmrg.c:96, python/pymodule.c:3317, test/algorithm/test_dmrg.c:369, ...)
 * link against libchemtensor_b200.so unchanged.  Inputs and outputs are genuine
 * host-memory structs (ctb_types.h); the device is an implementation detail.
 * All arithmetic runs in hand-written sm_100a kernels behind ctb_device.h; there
 * is no CPU fallback: without a CUDA device every compute entry point fails loudly.
 *
 * Citations are file:line in the reference tree (declaration; definition).
 */
#ifndef CHEMTENSOR_B200_H
#define CHEMTENSOR_B200_H

#include "ctb_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- host containers (allocation helpers; no arithmetic) ---------------------------------------------------- */
/* include/tensor/dense_tensor.h:29-35; src/tensor/dense_tensor.c */
void allocate_dense_tensor(const enum numeric_type dtype, const int ndim, const ct_long* dim, struct dense_tensor* t);
void allocate_zero_dense_tensor(const enum numeric_type dtype, const int ndim, const ct_long* dim, struct dense_tensor* t);
void delete_dense_tensor(struct dense_tensor* t);
/* include/tensor/block_sparse_tensor.h:35-41; src/tensor/block_sparse_tensor.c:37, :146, :156, :205 */
void allocate_block_sparse_tensor(const enum numeric_type dtype, const int ndim, const ct_long* dim, const enum tensor_axis_direction* axis_dir, const qnumber** qnums, struct block_sparse_tensor* t);
void allocate_block_sparse_tensor_like(const struct block_sparse_tensor* s, struct block_sparse_tensor* t);
void delete_block_sparse_tensor(struct block_sparse_tensor* t);
void copy_block_sparse_tensor(const struct block_sparse_tensor* src, struct block_sparse_tensor* dst);
/* include/state/mps.h:28-31, include/operator/mpo.h:46-49 */
void allocate_mps(const enum numeric_type dtype, const int nsites, const ct_long d, const qnumber* qsite, const ct_long* dim_bonds, const qnumber** qbonds, struct mps* mps);
void delete_mps(struct mps* mps);
/* reference include/state/mps.h:129-131, src/state/mps.c:1219 / :1309: the MPS in the reference's HDF5 layout (attributes nsites, qsite,
 * qbond_<i>; dense datasets tensor_<i>); written and parsed without libhdf5 (chemtensor_b200/host/mps_io.c); 0 on success, < 0 otherwise */
int save_mps(const char* filename, const struct mps* mps);
int load_mps(const char* filename, struct mps* mps);
void allocate_mpo(const enum numeric_type dtype, const int nsites, const ct_long d, const qnumber* qsite, const ct_long* dim_bonds, const qnumber** qbonds, struct mpo* mpo);
void delete_mpo(struct mpo* mpo);
/* include/algorithm/truncation.h:38 */
void delete_index_list(struct index_list* list);

/* ---- block-sparse tensor primitives on the hot path --------------------------------------------------------- */
/* include/tensor/block_sparse_tensor.h:218-222; src/tensor/block_sparse_tensor.c:3111, :3131, :3157 */
ct_long block_sparse_tensor_num_elements_blocks(const struct block_sparse_tensor* t);
void block_sparse_tensor_serialize_entries(const struct block_sparse_tensor* t, void* entries);
void block_sparse_tensor_deserialize_entries(struct block_sparse_tensor* t, const void* entries);
/* include/tensor/block_sparse_tensor.h:99-101; src/tensor/block_sparse_tensor.c:785 */
void block_sparse_tensor_transpose(const int* perm, const struct block_sparse_tensor* t, struct block_sparse_tensor* r);
void block_sparse_tensor_conjugate_transpose(const int* perm, const struct block_sparse_tensor* t, struct block_sparse_tensor* r);
/* include/tensor/block_sparse_tensor.h:109-111; src/tensor/block_sparse_tensor.c:950, :1123 */
void block_sparse_tensor_flatten_axes(const struct block_sparse_tensor* t, const int i_ax, const enum tensor_axis_direction new_axis_dir, struct block_sparse_tensor* r);
void block_sparse_tensor_split_axis(const struct block_sparse_tensor* t, const int i_ax, const ct_long new_dim_logical[2], const enum tensor_axis_direction new_axis_dir[2], const qnumber* new_qnums_logical[2], struct block_sparse_tensor* r);
/* include/tensor/block_sparse_tensor.h:153; src/tensor/block_sparse_tensor.c:1446 */
void block_sparse_tensor_slice(const struct block_sparse_tensor* t, const int i_ax, const ct_long* ind, const ct_long nind, struct block_sparse_tensor* r);
/* include/tensor/block_sparse_tensor.h:57; src/tensor/block_sparse_tensor.c:1560.
 * Restriction: the traced legs must have logical dimension 1 (the dummy outer bonds of the DMRG path). */
void block_sparse_tensor_cyclic_partial_trace(const struct block_sparse_tensor* t, const int ndim_trace, struct block_sparse_tensor* r);
/* include/tensor/block_sparse_tensor.h:163; src/tensor/block_sparse_tensor.c:1654 */
void block_sparse_tensor_multiply_pointwise_vector(const struct block_sparse_tensor* s, const struct dense_tensor* t, const enum tensor_axis_range axrange, struct block_sparse_tensor* r);
/* include/tensor/block_sparse_tensor.h:167; src/tensor/block_sparse_tensor.c:1826 */
void block_sparse_tensor_dot(const struct block_sparse_tensor* s, const enum tensor_axis_range axrange_s, const struct block_sparse_tensor* t, const enum tensor_axis_range axrange_t, const int ndim_mult, struct block_sparse_tensor* r);
/* include/tensor/block_sparse_tensor.h:179-181; src/tensor/block_sparse_tensor.c:2402, :2544.  mode must be QR_REDUCED. */
int block_sparse_tensor_qr(const struct block_sparse_tensor* a, const enum qr_mode mode, struct block_sparse_tensor* q, struct block_sparse_tensor* r);
int block_sparse_tensor_rq(const struct block_sparse_tensor* a, const enum qr_mode mode, struct block_sparse_tensor* r, struct block_sparse_tensor* q);
/* include/tensor/block_sparse_tensor.h:189; src/tensor/block_sparse_tensor.c:2686 */
int block_sparse_tensor_svd(const struct block_sparse_tensor* a, struct block_sparse_tensor* u, struct dense_tensor* s, struct block_sparse_tensor* vh);

/* ---- truncation and bond operations ------------------------------------------------------------------------ */
/* include/algorithm/truncation.h:10, :40; src/algorithm/truncation.c:13, :110 */
double von_neumann_entropy(const double* sigma, const ct_long n);
void retained_bond_indices(const double* sigma, const ct_long n, const double tol, const bool relative_thresh, const ct_long max_vdim, struct index_list* list, struct trunc_info* info);
/* extension: the same rule evaluated by the device kernels the SVD split uses (rank sort + sequential sums on the GPU) */
int ctb_retained_bond_indices_device(const double* sigma, const ct_long n, const double tol, const bool relative_thresh, const ct_long max_vdim, struct index_list* list, struct trunc_info* info);
/* include/algorithm/bond_ops.h:22; src/algorithm/bond_ops.c:15 */
int split_block_sparse_matrix_svd(const struct block_sparse_tensor* a, const double tol, const bool relative_thresh, const ct_long max_vdim, const bool renormalize, const enum singular_value_distr svd_distr, struct block_sparse_tensor* a0, struct block_sparse_tensor* a1, struct trunc_info* info);

/* ---- MPS / MPO pieces used by the sweep ------------------------------------------------------------------- */
/* include/state/mps.h:68-70; src/state/mps.c:513, :560, :609 */
void mps_local_orthonormalize_qr(struct block_sparse_tensor* a, struct block_sparse_tensor* a_next);
void mps_local_orthonormalize_rq(struct block_sparse_tensor* a, struct block_sparse_tensor* a_prev);
double mps_orthonormalize_qr(struct mps* mps, const enum mps_orthonormalization_mode mode);
/* include/state/mps.h:92-97; src/state/mps.c:1119, :1166 */
int mps_split_tensor_svd(const struct block_sparse_tensor* a, const ct_long d[2], const qnumber* new_qsite[2], const double tol, const ct_long max_vdim, const bool renormalize, const enum singular_value_distr svd_distr, struct block_sparse_tensor* a0, struct block_sparse_tensor* a1, struct trunc_info* info);
void mps_merge_tensor_pair(const struct block_sparse_tensor* a0, const struct block_sparse_tensor* a1, struct block_sparse_tensor* a);
/* include/operator/mpo.h:72; src/operator/mpo.c:255 */
void mpo_merge_tensor_pair(const struct block_sparse_tensor* a0, const struct block_sparse_tensor* a1, struct block_sparse_tensor* a);

/* ---- chain operations -------------------------------------------------------------------------------------- */
/* include/algorithm/chain_ops.h:10-24; src/algorithm/chain_ops.c:14, :53, :116, :196, :253 */
void create_dummy_operator_block_right(const struct block_sparse_tensor* a, const struct block_sparse_tensor* b, const struct block_sparse_tensor* w, struct block_sparse_tensor* r);
void create_dummy_operator_block_left(const struct block_sparse_tensor* a, const struct block_sparse_tensor* b, const struct block_sparse_tensor* w, struct block_sparse_tensor* l);
void contraction_operator_step_right(const struct block_sparse_tensor* a, const struct block_sparse_tensor* b, const struct block_sparse_tensor* w, const struct block_sparse_tensor* r, struct block_sparse_tensor* r_next);
void contraction_operator_step_left(const struct block_sparse_tensor* a, const struct block_sparse_tensor* b, const struct block_sparse_tensor* w, const struct block_sparse_tensor* l, struct block_sparse_tensor* l_next);
void compute_right_operator_blocks(const struct mps* psi, const struct mps* chi, const struct mpo* op, struct block_sparse_tensor* r_list);
/* include/algorithm/chain_ops.h:29-30; src/algorithm/chain_ops.c:353 */
void apply_local_hamiltonian(const struct block_sparse_tensor* a, const struct block_sparse_tensor* w, const struct block_sparse_tensor* l, const struct block_sparse_tensor* r, struct block_sparse_tensor* b);

/* ---- Krylov interface --------------------------------------------------------------------------------------- */
/* include/util/krylov.h:15-30; src/util/krylov.c:24, :96, :172, :258 */
void lanczos_iteration_d(const ct_long n, lanczos_linear_func_d afunc, const void* adata, const double* vstart, const int maxiter, double* alpha, double* beta, double* v, int* numiter);
void lanczos_iteration_z(const ct_long n, lanczos_linear_func_z afunc, const void* adata, const void* vstart, const int maxiter, double* alpha, double* beta, void* v, int* numiter);
int eigensystem_krylov_symmetric(const ct_long n, lanczos_linear_func_d afunc, const void* adata, const double* vstart, const int maxiter, const int numeig, double* lambda, double* u_ritz);
int eigensystem_krylov_hermitian(const ct_long n, lanczos_linear_func_z afunc, const void* adata, const void* vstart, const int maxiter, const int numeig, double* lambda, void* u_ritz);

/* ---- DMRG drivers ------------------------------------------------------------------------------------------- */
/* include/algorithm/dmrg.h:10-13; src/algorithm/dmrg.c:155, :262 */
int dmrg_singlesite(const struct mpo* hamiltonian, const int num_sweeps, const int maxiter_lanczos, struct mps* psi, double* en_sweeps);
int dmrg_twosite(const struct mpo* hamiltonian, const int num_sweeps, const int maxiter_lanczos, const double tol_split, const ct_long max_vdim, struct mps* psi, double* en_sweeps, double* entropy);

/* ---- other callers of the same primitives ------------------------------------ */
/* include/state/mps.h:73-75 (src/state/mps.c:314, :375) */
void mps_vdot(const struct mps* chi, const struct mps* psi, void* ret);
double mps_norm(const struct mps* psi);
/* include/algorithm/chain_ops.h (src/algorithm/chain_ops.c:274, :485, :424) */
void mpo_inner_product(const struct mps* chi, const struct mpo* op, const struct mps* psi, void* ret);
void apply_mpo(const struct mpo* op, const struct mps* psi, struct mps* op_psi);
void compute_local_hamiltonian_environment(const struct block_sparse_tensor* a, const struct block_sparse_tensor* b, const struct block_sparse_tensor* l, const struct block_sparse_tensor* r, struct block_sparse_tensor* dw);
/* include/algorithm/bond_ops.h (src/algorithm/bond_ops.c:146) */
int split_block_sparse_matrix_svd_isometry(const struct block_sparse_tensor* a, const double tol, const bool relative_thresh, const ct_long max_vdim, struct block_sparse_tensor* u, struct trunc_info* info);
/* include/state/mps.h:93-103 (src/state/mps.c:764, :815, :868, :1071) */
int mps_local_orthonormalize_left_svd(const double tol, const ct_long max_vdim, const bool renormalize, struct block_sparse_tensor* a, struct block_sparse_tensor* a_next, struct trunc_info* info);
int mps_local_orthonormalize_right_svd(const double tol, const ct_long max_vdim, const bool renormalize, struct block_sparse_tensor* a, struct block_sparse_tensor* a_prev, struct trunc_info* info);
int mps_compress(const double tol, const ct_long max_vdim, const enum mps_orthonormalization_mode mode, struct mps* mps, double* norm, double* trunc_scale, struct trunc_info* info);
int mps_compress_rescale(const double tol, const ct_long max_vdim, const enum mps_orthonormalization_mode mode, struct mps* mps, double* trunc_scale, struct trunc_info* info);

/* include/operator/mpo.h:50 (src/operator/mpo.c:59) -- sparse-direct: no dense Dw x d x d x Dw' intermediate */
void mpo_from_assembly(const struct mpo_assembly* assembly, struct mpo* mpo);
/* include/algorithm/gradient.h (src/algorithm/gradient.c:15) */
void operator_average_coefficient_gradient(const struct mpo_assembly* assembly, const struct mps* psi, const struct mps* chi, void* avr, void* dcoeff);

/* ---- engine extensions (no reference counterpart; measurement and lifecycle) -------------------------------- */
/* explicit device selection / start-up; returns <0 when no CUDA device is usable */
int ctb_init(int device);
/* one process per GPU: after ctb_dist_init every effective-Hamiltonian application inside dmrg_* / the Lanczos
 * driver is sharded over the ranks -- the bra bond of the right environment is cut into balanced index sets, each rank
 * contracts its slice and one NCCL all-gather over NVLink rebuilds the result on every rank; everything else (level-1
 * Lanczos work, SVD, environment updates) runs replicated and deterministic, so all ranks hold bit-identical states.
 * unique_id: the 128-byte NCCL id obtained on rank 0 with ctb_dist_unique_id and distributed by the host (MPI, torchrun, ...);
 * alternatively pass NULL and register the host's own all-gather with ctb_dist_set_allgather. */
int ctb_dist_unique_id(void* id_out_128_bytes);
int ctb_dist_init(int rank, int world, const void* unique_id);
int ctb_dist_set_allgather(int (*fn)(void* ctx, const void* sendbuf, void* recvbuf, size_t bytes_per_rank, void* stream), void* ctx);
int ctb_dist_finalize(void);
/* out[0] = rank, out[1] = world, out[2] = exchanges done by the fused peer-store path, out[3] = exchanges done by all-gather + scatter */
int ctb_dist_info(long long* out);
long long ctb_dist_pull_exchanges(void);
/* two-site effective Hamiltonian from the two single-site MPO tensors, no merged pair tensor */
int ctb_apply_local_hamiltonian_pair(const struct block_sparse_tensor* a, const struct block_sparse_tensor* w0, const struct block_sparse_tensor* w1, const struct block_sparse_tensor* l, const struct block_sparse_tensor* r, struct block_sparse_tensor* b);
long long ctb_dist_push_exchanges(void);
/* exchanges of the fused form that went through NVSwitch multicast stores (one store per element instead of one per peer) */
long long ctb_dist_multicast_exchanges(void);
/* 1 = CUDA kernels, 2 = host test double (tests/emu only) */
int ctb_backend(void);
/* kernels launched by the engine so far */
long long ctb_launch_count(void);
/* Heff micro-benchmark on one bond: uploads (a, w, l, r), builds the plans once, runs 'warmup' + 'reps' matvecs
 * device-resident and reports the mean device time per matvec (CUDA events) and the algorithmic flops per matvec
 * (sum of 2 m n k over the block GEMMs of the three contractions, x4 for complex).  per_step_ms (optional, 3 doubles)
 * receives the mean time of each of the three grouped-GEMM launches. */
int ctb_heff_benchmark(const struct block_sparse_tensor* a, const struct block_sparse_tensor* w, const struct block_sparse_tensor* l, const struct block_sparse_tensor* r,
	int warmup, int reps, int flush_l2, double* ms_per_matvec, double* flops_per_matvec, double* per_step_ms, double* per_step_flops);
/* micro-benchmark of one contraction r = dot(s, t) (one grouped-GEMM launch), operands device-resident, plan built once */
int ctb_dot_benchmark(const struct block_sparse_tensor* s, const int axrange_s, const struct block_sparse_tensor* t, const int axrange_t, const int ndim_mult,
	int warmup, int reps, int flush_l2, double* ms_per_run, double* flops_per_run);
/* plan-only query of the effective Hamiltonian as rank 'rank' of 'world' would build it (no communicator needed): out[0] flops of the
 * shard, out[1..3] tiles of the three launches, out[4] entries of the result slice, out[5] entries of an all-gather slot, out[6..7] entries of t1, t2 */
int ctb_heff_plan_info(const struct block_sparse_tensor* a, const struct block_sparse_tensor* w, const struct block_sparse_tensor* l, const struct block_sparse_tensor* r,
	int rank, int world, double* out);
/* statistics of the last dmrg_* call: fills up to 'n' doubles:
 * [0] heff flops, [1] heff calls, [2] env flops, [3] lanczos ms, [4] svd ms, [5] env ms, [6] total ms,
 * [7] longest Lanczos vector, [8] largest bond dimension */
int ctb_get_stats(double* out, int n);
/* device time and algorithmic bytes of the re-blocking kernel on t: out = { ms transpose, bytes, ms flatten(0,1), bytes } */
int ctb_remap_benchmark(const struct block_sparse_tensor* t, double* out);

#ifdef __cplusplus
}
#endif

#endif
