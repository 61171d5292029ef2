/*
 * ctb_su2.h -- SU(2)-symmetric tensors across the drop-in boundary (SURVEY.md section 8(a) last row, 8(f) rank 2; BASELINE configs[4]).
 *
 * Public structs of the reference restated with the SAME tags, field order and integer widths, so that host objects built by
 * the reference's generators (construct_heisenberg_1d_su2_mpo, construct_random_su2_mps) cross the boundary unchanged and
 * objects returned by this library are released by the reference's delete_su2_tensor / delete_su2_mps (every node, list and
 * degeneracy tensor is a separate 16-byte aligned malloc, as in the reference):
 *   struct su2_tree_node, su2_fuse_split_tree        include/tensor/su2_tree.h:16-20, 77-82
 *   struct su2_irreducible_list, charge_sectors      include/tensor/su2_irreps.h:14-18, 60-65
 *   struct su2_tensor                                include/tensor/su2_tensor.h:16-26
 *   struct su2_mps                                   include/state/su2_mps.h:15-19
 *   struct su2_mpo                                   include/operator/su2_mpo.h:14-18
 *   enum su2_singular_value_distr                    include/algorithm/su2_bond_ops.h:15-19
 *
 * Entry points (same names, argument meaning and error convention as the reference):
 *   su2_tensor_contract_simple                       include/tensor/su2_tensor.h:168   (src/tensor/su2_tensor.c:2387)
 *   su2_tensor_fmove                                 include/tensor/su2_tensor.h:134   (src/tensor/su2_tensor.c:770)
 *   su2_apply_local_hamiltonian                      include/algorithm/su2_chain_ops.h (src/algorithm/su2_chain_ops.c:496)
 *   su2_contraction_operator_step_left / _right      (src/algorithm/su2_chain_ops.c:259, :160)
 *   su2_compute_right_operator_blocks                (src/algorithm/su2_chain_ops.c:370)
 *   su2_create_dummy_operator_block_left / _right    (src/algorithm/su2_chain_ops.c:82, :16)
 *   su2_mpo_inner_product                            (src/algorithm/su2_chain_ops.c:399)
 *   su2_mps_orthonormalize_qr                        include/state/su2_mps.h:79        (src/state/su2_mps.c:450)
 *   su2_tensor_svd                                   include/tensor/su2_tensor.h:197   (src/tensor/su2_tensor.c:4300)
 *   su2_tensor_(de)serialize_renormalized_entries    include/tensor/su2_tensor.h:221   (src/tensor/su2_tensor.c:4725, :4837)
 *   su2_dmrg_singlesite / su2_dmrg_twosite           include/algorithm/su2_dmrg.h:10-13 (src/algorithm/su2_dmrg.c:155, :262)
 */
#ifndef CTB_SU2_H
#define CTB_SU2_H

#include "ctb_types.h"

#ifdef __cplusplus
extern "C" {
#endif

struct su2_tree_node
{
	int i_ax;                    /* tensor axis index (outer for leaves, internal otherwise) */
	struct su2_tree_node* c[2];  /* children; NULL for a leaf */
};

struct su2_fuse_split_tree
{
	struct su2_tree_node* tree_fuse;
	struct su2_tree_node* tree_split;
	int ndim;
};

struct su2_irreducible_list
{
	qnumber* jlist;   /* 'j' quantum numbers times 2 */
	int num;
};

struct charge_sectors
{
	qnumber* jlists;  /* nsec x ndim, sorted lexicographically */
	ct_long nsec;
	int ndim;
};

struct su2_tensor
{
	struct su2_fuse_split_tree tree;
	struct su2_irreducible_list* outer_irreps;
	struct charge_sectors charge_sectors;
	struct dense_tensor** degensors;
	ct_long** dim_degen;
	enum numeric_type dtype;
	int ndim_logical;
	int ndim_auxiliary;
};

struct su2_mps
{
	struct su2_tensor* a;
	int nsites;
};

struct su2_mpo
{
	struct su2_tensor* a;
	int nsites;
};

enum su2_singular_value_distr
{
	SU2_SVD_DISTR_LEFT  = 0,
	SU2_SVD_DISTR_RIGHT = 1,
};

enum su2_mps_orthonormalization_mode
{
	SU2_MPS_ORTHONORMAL_LEFT  = 0,
	SU2_MPS_ORTHONORMAL_RIGHT = 1,
};

double su2_recoupling_coefficient(const qnumber ja, const qnumber jb, const qnumber jc, const qnumber js, const qnumber je, const qnumber jf);

void su2_tensor_contract_simple(const struct su2_tensor* s, const int* i_ax_s, const struct su2_tensor* t, const int* i_ax_t, const int ndim_mult, struct su2_tensor* r);
void su2_tensor_fmove(const struct su2_tensor* t, const int i_ax, struct su2_tensor* r);

void su2_create_dummy_operator_block_right(const enum numeric_type dtype, const qnumber irrep_sector_state, struct su2_tensor* r);
void su2_create_dummy_operator_block_left(const enum numeric_type dtype, struct su2_tensor* l);
void su2_contraction_operator_step_right(const struct su2_tensor* a, const struct su2_tensor* b, const struct su2_tensor* w, const struct su2_tensor* r, struct su2_tensor* r_next);
void su2_contraction_operator_step_left(const struct su2_tensor* a, const struct su2_tensor* b, const struct su2_tensor* w, const struct su2_tensor* l, struct su2_tensor* l_next);
void su2_compute_right_operator_blocks(const struct su2_mps* psi, const struct su2_mps* chi, const struct su2_mpo* op, struct su2_tensor* r_list);
void su2_mpo_inner_product(const struct su2_mps* chi, const struct su2_mpo* op, const struct su2_mps* psi, void* ret);
void su2_apply_local_hamiltonian(const struct su2_tensor* a, const struct su2_tensor* w, const struct su2_tensor* l, const struct su2_tensor* r, struct su2_tensor* b);

double su2_mps_orthonormalize_qr(struct su2_mps* mps, const enum su2_mps_orthonormalization_mode mode);
/* include/state/su2_mps.h:75-76 (src/state/su2_mps.c:262, :307) */
void su2_mps_local_orthonormalize_qr(struct su2_tensor* a, struct su2_tensor* a_next);
void su2_mps_local_orthonormalize_rq(struct su2_tensor* a, struct su2_tensor* a_prev);

/* include/tensor/su2_tensor.h:197, :217-222 */
int su2_tensor_svd(const struct su2_tensor* a, const bool copy_tree_left, struct su2_tensor* u, struct dense_tensor* s, int** multiplicities, struct su2_tensor* vh);
ct_long su2_tensor_num_elements_degensors(const struct su2_tensor* t);
void su2_tensor_serialize_renormalized_entries(const struct su2_tensor* t, void* entries);
void su2_tensor_deserialize_renormalized_entries(struct su2_tensor* t, const void* entries);

int su2_dmrg_singlesite(const struct su2_mpo* hamiltonian, const int num_sweeps, const int maxiter_lanczos, struct su2_mps* psi, double* en_sweeps);
int su2_dmrg_twosite(const struct su2_mpo* hamiltonian, const int num_sweeps, const int maxiter_lanczos, const double tol_split, const ct_long max_vdim,
	struct su2_mps* psi, double* en_sweeps, double* entropy);

/* extension: application of the two-site effective Hamiltonian in PAIR form (the two site MPO tensors one after the other on the
 * 4-leg two-site tensor [Dl, d1, d2, Dr]; no merged pair tensor, cf. su2_mpo_merge_tensor_pair src/operator/su2_mpo.c:341).
 * a2 = su2_mps_contract_tensor_pair(a_i, a_{i+1}) (src/state/su2_mps.c:673); b2 has its structure. */
void ctb_su2_apply_local_hamiltonian_pair(const struct su2_tensor* a2, const struct su2_tensor* w0, const struct su2_tensor* w1,
	const struct su2_tensor* l, const struct su2_tensor* r, struct su2_tensor* b2);

/* statistics of the last su2_dmrg_* call, 16 doubles: [0] device launches of the SU(2) layer, [1] Heff applications, [2] seconds in local solves,
 * [3] seconds in splits / QR, [4] seconds in environment steps, [5] seconds of the initial orthonormalisation + right environments,
 * [6] sweeps completed, [7..15] seconds of each sweep */
void ctb_su2_get_stats(double* out16);
/* measurement aid: the block linear combination kernel (csrc/ctbd_blocklc.cu) alone; out = { ms, GB/s, algorithmic bytes } */
int ctb_su2_lc_benchmark(ct_long nelem, int nblk, int nterm, int cplx, double* out);

#ifdef __cplusplus
}
#endif

#endif
