/*
 * ctb_types.h -- plain-C data types shared across the drop-in boundary.
 *
 * These are the public structs of the reference C API, restated with the SAME
 * tags, field order and integer widths so that host objects can cross the
 * boundary unchanged (the reference's callers -- its test-suite, perf/perf_dmrg.c
 * and python/pymodule.c -- read the fields directly, SURVEY.md §8(b) "struct ABI").
 *
 * Reference interface restated here (file:line under the reference tree):
 *   enum numeric_type                 include/numeric.h:15-22
 *   ct_long (= int64_t)               include/util/util.h:12
 *   qnumber (= int32_t)               include/tensor/qnumber.h:17
 *   enum tensor_axis_direction        include/tensor/qnumber.h:76-80
 *   struct dense_tensor               include/tensor/dense_tensor.h:17-23
 *   enum tensor_axis_range            include/tensor/dense_tensor.h:152-157
 *   enum qr_mode                      include/tensor/dense_tensor.h:183-188
 *   struct block_sparse_tensor        include/tensor/block_sparse_tensor.h:18-28
 *   struct trunc_info / index_list    include/algorithm/truncation.h:21-36
 *   enum singular_value_distr         include/algorithm/bond_ops.h:15-19
 *   struct mps                        include/state/mps.h:14-20
 *   enum mps_orthonormalization_mode  include/state/mps.h:61-65
 *   struct mpo                        include/operator/mpo.h:34-40
 *   lanczos_linear_func_d/_z          include/util/krylov.h:10-12
 *   struct local_op_ref               include/operator/local_op.h:38-42
 *   struct mpo_graph_vertex/edge, mpo_graph   include/operator/mpo_graph.h:16-50
 *   struct mpo_assembly               include/operator/mpo.h:14-24
 *
 * Ownership convention (reference include/aligned_memory.h): output structs
 * are caller-provided shells; every payload is allocated by the callee with a
 * 16-byte aligned malloc and released by the matching delete_* function, which
 * uses plain free().  The replacement keeps exactly this, so objects created by
 * either library can be released by the other.
 */
#ifndef CTB_TYPES_H
#define CTB_TYPES_H

#include <stdint.h>
#include <stdbool.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t ct_long;
typedef int32_t qnumber;

enum numeric_type
{
	CT_SINGLE_REAL       = 0,
	CT_DOUBLE_REAL       = 1,
	CT_SINGLE_COMPLEX    = 2,
	CT_DOUBLE_COMPLEX    = 3,
	CT_NUM_NUMERIC_TYPES = 4,
};

enum tensor_axis_direction
{
	TENSOR_AXIS_IN  = -1,
	TENSOR_AXIS_OUT =  1,
};

enum tensor_axis_range
{
	TENSOR_AXIS_RANGE_LEADING  = 0,
	TENSOR_AXIS_RANGE_TRAILING = 1,
	TENSOR_AXIS_RANGE_NUM      = 2,
};

enum qr_mode
{
	QR_REDUCED   = 0,
	QR_COMPLETE  = 1,
	QR_NUM_MODES = 2,
};

enum singular_value_distr
{
	SVD_DISTR_LEFT  = 0,
	SVD_DISTR_RIGHT = 1,
};

enum mps_orthonormalization_mode
{
	MPS_ORTHONORMAL_LEFT  = 0,
	MPS_ORTHONORMAL_RIGHT = 1,
};

/* row-major dense tensor in host memory */
struct dense_tensor
{
	void* data;
	ct_long* dim;
	enum numeric_type dtype;
	int ndim;
};

/* block-sparse tensor in host memory; 'blocks' spans the full sector grid,
 * NULL where the quantum numbers are not conserved */
struct block_sparse_tensor
{
	struct dense_tensor** blocks;
	ct_long* dim_blocks;
	ct_long* dim_logical;
	enum tensor_axis_direction* axis_dir;
	qnumber** qnums_blocks;
	qnumber** qnums_logical;
	enum numeric_type dtype;
	int ndim;
};

struct trunc_info
{
	double norm_sigma;
	double entropy;
	double tol_eff;
};

struct index_list
{
	ct_long* ind;
	ct_long num;
};

struct mps
{
	struct block_sparse_tensor* a;
	qnumber* qsite;
	ct_long d;
	int nsites;
};

struct mpo
{
	struct block_sparse_tensor* a;
	qnumber* qsite;
	ct_long d;
	int nsites;
};

/* interleaved (re, im) double pair; layout-compatible with C99 'double _Complex' */
typedef struct { double re, im; } ctb_dcomplex;

typedef void lanczos_linear_func_d(const ct_long n, const void* data, const double* v, double* ret);
typedef void lanczos_linear_func_z(const ct_long n, const void* data, const void* v, void* ret);

/* ---- MPO assembly: operator graph + look-up tables (read-only inputs of mpo_from_assembly and the coefficient gradient) ---- */

enum { OID_NOP = -1, OID_IDENTITY = 0 };     /* include/operator/local_op.h:13-17 */

struct local_op_ref
{
	int oid;   /* operator index into opmap */
	int cid;   /* coefficient index into coeffmap */
};

struct mpo_graph_vertex
{
	int* eids[2];      /* indices of left- and right-connected edges */
	int num_edges[2];
	qnumber qnum;
};

struct mpo_graph_edge
{
	int vids[2];                 /* left and right vertex */
	struct local_op_ref* opics;  /* weighted sum of local operators */
	int nopics;
};

struct mpo_graph
{
	struct mpo_graph_vertex** verts;  /* [nsites + 1][num_verts] */
	struct mpo_graph_edge** edges;    /* [nsites][num_edges] */
	int* num_verts;
	int* num_edges;
	int nsites;
};

struct mpo_assembly
{
	struct mpo_graph graph;
	struct dense_tensor* opmap;     /* local operator look-up table */
	void* coeffmap;                 /* coefficient look-up table (entries of type dtype) */
	qnumber* qsite;
	ct_long d;
	enum numeric_type dtype;
	int num_local_ops;
	int num_coeffs;
};

#ifdef __cplusplus
}
#endif

#endif
