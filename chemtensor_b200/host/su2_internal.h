/*
 * su2_internal.h -- device-resident SU(2)-symmetric tensors (host side of the SU(2) layer).
 *
 * The reference keeps an SU(2) tensor as a fusion-splitting tree, the list of charge sectors ('j' configurations of all axes,
 * sorted lexicographically) and one separately allocated dense "degeneracy" tensor per sector (include/tensor/su2_tensor.h:16-26);
 * every operation walks the sectors on the host (src/tensor/su2_tensor.c).  Here the structural part (tree, irreducible lists,
 * sector table, offsets) is integer metadata on the host, built once per operation of a bond, and all degeneracy tensors of one
 * SU(2) tensor are ONE packed device buffer.  An operation is a structural step on the host plus at most one launch:
 *   contraction  -> one grouped GEMM over all matching sector pairs (ctbd_gemm_*, csrc/ctbd_gemm.cu)
 *   F-move, transposition, pending scalings of axis reversals / swaps -> one block linear combination (ctbd_lc_*, csrc/ctbd_blocklc.cu)
 * The launches of an effective-Hamiltonian application are RECORDED the first time (struct su2_prog) and replayed for the
 * remaining Lanczos iterations of the local solve: no structural work, no allocation and no host synchronisation per matvec.
 */
#ifndef CTB_SU2_INTERNAL_H
#define CTB_SU2_INTERNAL_H

#include "ctb_internal.h"
#include "ctb_su2.h"

#define SU2_MAXAX 24          /* axes of one tensor: outer + internal */
#define SU2_MAXNODE (2 * SU2_MAXAX)

struct su2_node { int ax; int c[2]; };    /* c[k] = node index or -1 (leaf) */

struct su2t
{
	int dtype;
	int nl, na;                 /* logical and auxiliary outer axes */
	int ndim;                   /* all axes: 2 (nl + na) - 3 */
	struct su2_node node[SU2_MAXNODE];
	int nn;
	int rf, rs;                 /* root nodes of the fusion and the splitting tree (same axis) */
	struct su2_irreducible_list irr[SU2_MAXAX];   /* outer irreducible lists (owned) */
	ct_long* dd[SU2_MAXAX];     /* degeneracy dimension per logical axis, indexed by j (owned, length j_max + 1) */
	ct_long nsec;
	qnumber* jl;                /* [nsec x ndim] sorted lexicographically */
	ct_long* off;               /* [nsec] element offset of the degeneracy tensor in the device buffer */
	ct_long* nel;               /* [nsec] number of elements */
	double* scale;              /* [nsec] pending real factor per sector, NULL = 1 */
	int conj;                   /* pending complex conjugation */
	ct_long nstore;             /* elements of the device buffer */
	void* dev;
	int own;                    /* buffer is released with the handle */
	int varies;                 /* depends on the input of the program being recorded */
};

/* one recorded launch */
struct su2_op { int kind; void* plan; void* a; void* b; void* c; };   /* kind 0: gemm C = A B, 1: lc (a -> c) */

struct su2_prog
{
	struct su2_op* ops; int nops, cap;
	void** keep; int nkeep, keepcap;     /* device buffers of the intermediates, released with the program */
	void* in;                            /* device buffer the recorded launches read the input from */
	void* out;                           /* device buffer of the result */
	void* graph;                         /* the launches captured as ONE CUDA graph (NULL: replay launch by launch) */
	ct_long n;                           /* vector length */
};

/* su2_core.c */
double ctb_su2_recoupling(qnumber ja, qnumber jb, qnumber jc, qnumber js, qnumber je, qnumber jf);
struct su2t* su2t_upload(const struct su2_tensor* t);
int su2t_download(struct su2t* h, struct su2_tensor* t);
struct su2t* su2t_clone_meta(const struct su2t* h);              /* metadata copy sharing the device buffer (not owned) */
struct su2t* su2t_alloc_like(const struct su2t* h, int zero);
void su2t_free(struct su2t* h);
int su2t_materialize(struct su2t* h);
void su2t_set_sectors_all_valid(struct su2t* h);                 /* sector table = every valid configuration; offsets sequential; no buffer */
ct_long su2t_find_sector(const struct su2t* h, const qnumber* jl);
int su2t_embed(struct su2t* src, struct su2t* dst, int weight_mode);   /* dst sectors filled from src (missing -> 0); weight_mode +1: * sqrt(j_root + 1), -1: / sqrt(j_root + 1) */

void su2t_flip_trees(struct su2t* h);
void su2t_conjugate(struct su2t* h);
void su2t_swap_tree_axes(struct su2t* h, int a0, int a1);
void su2t_reverse_axis_simple(struct su2t* h, int ax);
struct su2t* su2t_fmove(struct su2t* t, int ax);
struct su2t* su2t_transpose_logical(struct su2t* t, const int* perm);
struct su2t* su2t_contract_simple(struct su2t* s, const int* axs, struct su2t* t, const int* axt, int nmult);
int su2t_parent_axis(const struct su2t* h, int split_tree, int ax);    /* axis of the parent node of 'ax' in the given tree, -1 if none */
int su2t_child_axis(const struct su2t* h, int split_tree, int k);      /* axis of child k of the root of the given tree (-1: root is a leaf) */

void su2_prog_begin(struct su2_prog* p);
void su2_prog_end(void);
int su2_prog_run(struct su2_prog* p);
void su2_prog_capture(struct su2_prog* p);      /* after recording: capture the launches into a CUDA graph where the device layer offers it */
void su2_prog_free(struct su2_prog* p);
int su2_dev_lc(int dtype, int conj, int nblk, struct ctbd_lc_block* blocks, int nterm, struct ctbd_lc_term* terms, const void* src, void* dst, int varies);
int su2_dev_gemm(const struct ctbd_gemm_plan_host* ph, const void* A, const void* B, void* C, int varies);
void su2_keep_or_free(void* dev, int varies);
extern long long g_su2_launches;

#endif
