"""ctypes view of the drop-in C boundary (include/ctb_types.h, include/chemtensor_b200.h).

The struct layouts are the reference's public C structs (reference
include/tensor/block_sparse_tensor.h:18-28, include/tensor/dense_tensor.h:17-23,
include/state/mps.h:14-20, include/operator/mpo.h:34-40), so the same Python classes can
talk to any shared library that exports the reference's symbols.  `CLibrary` binds the
signatures for one shared object; the product uses it for libchemtensor_b200.so only.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

CT_SINGLE_REAL, CT_DOUBLE_REAL, CT_SINGLE_COMPLEX, CT_DOUBLE_COMPLEX = 0, 1, 2, 3
TENSOR_AXIS_IN, TENSOR_AXIS_OUT = -1, 1
AXIS_RANGE_LEADING, AXIS_RANGE_TRAILING = 0, 1
QR_REDUCED, QR_COMPLETE = 0, 1
SVD_DISTR_LEFT, SVD_DISTR_RIGHT = 0, 1
MPS_ORTHONORMAL_LEFT, MPS_ORTHONORMAL_RIGHT = 0, 1

_NP_OF_DTYPE = {CT_SINGLE_REAL: np.float32, CT_DOUBLE_REAL: np.float64,
                CT_SINGLE_COMPLEX: np.complex64, CT_DOUBLE_COMPLEX: np.complex128}
_DTYPE_OF_NP = {np.dtype(v): k for k, v in _NP_OF_DTYPE.items()}


def np_dtype(code: int):
    return _NP_OF_DTYPE[int(code)]


def ct_dtype(npdt) -> int:
    return _DTYPE_OF_NP[np.dtype(npdt)]


class DenseTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dim", C.POINTER(C.c_int64)), ("dtype", C.c_int), ("ndim", C.c_int)]


class BlockSparseTensor(C.Structure):
    _fields_ = [("blocks", C.POINTER(C.POINTER(DenseTensor))),
                ("dim_blocks", C.POINTER(C.c_int64)),
                ("dim_logical", C.POINTER(C.c_int64)),
                ("axis_dir", C.POINTER(C.c_int)),
                ("qnums_blocks", C.POINTER(C.POINTER(C.c_int32))),
                ("qnums_logical", C.POINTER(C.POINTER(C.c_int32))),
                ("dtype", C.c_int), ("ndim", C.c_int)]


class TruncInfo(C.Structure):
    _fields_ = [("norm_sigma", C.c_double), ("entropy", C.c_double), ("tol_eff", C.c_double)]


class IndexList(C.Structure):
    _fields_ = [("ind", C.POINTER(C.c_int64)), ("num", C.c_int64)]


class MPSStruct(C.Structure):
    _fields_ = [("a", C.POINTER(BlockSparseTensor)), ("qsite", C.POINTER(C.c_int32)), ("d", C.c_int64), ("nsites", C.c_int)]


class MPOStruct(C.Structure):
    _fields_ = [("a", C.POINTER(BlockSparseTensor)), ("qsite", C.POINTER(C.c_int32)), ("d", C.c_int64), ("nsites", C.c_int)]


LANCZOS_FUNC = C.CFUNCTYPE(None, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p)
# int fn(void* ctx, const void* sendbuf, void* recvbuf, size_t bytes_per_rank, void* stream)
ALLGATHER_FUNC = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)

_P_BST = C.POINTER(BlockSparseTensor)
_P_DT = C.POINTER(DenseTensor)

# name -> (restype, argtypes); the reference-named subset every library under test exports
_SIGNATURES = {
    "allocate_block_sparse_tensor": (None, [C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_int32)), _P_BST]),
    "delete_block_sparse_tensor": (None, [_P_BST]),
    "allocate_dense_tensor": (None, [C.c_int, C.c_int, C.POINTER(C.c_int64), _P_DT]),
    "delete_dense_tensor": (None, [_P_DT]),
    "block_sparse_tensor_num_elements_blocks": (C.c_int64, [_P_BST]),
    "block_sparse_tensor_serialize_entries": (None, [_P_BST, C.c_void_p]),
    "block_sparse_tensor_deserialize_entries": (None, [_P_BST, C.c_void_p]),
    "block_sparse_tensor_transpose": (None, [C.POINTER(C.c_int), _P_BST, _P_BST]),
    "block_sparse_tensor_conjugate_transpose": (None, [C.POINTER(C.c_int), _P_BST, _P_BST]),
    "block_sparse_tensor_flatten_axes": (None, [_P_BST, C.c_int, C.c_int, _P_BST]),
    "block_sparse_tensor_split_axis": (None, [_P_BST, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_int32)), _P_BST]),
    "block_sparse_tensor_slice": (None, [_P_BST, C.c_int, C.POINTER(C.c_int64), C.c_int64, _P_BST]),
    "block_sparse_tensor_cyclic_partial_trace": (None, [_P_BST, C.c_int, _P_BST]),
    "block_sparse_tensor_multiply_pointwise_vector": (None, [_P_BST, _P_DT, C.c_int, _P_BST]),
    "block_sparse_tensor_dot": (None, [_P_BST, C.c_int, _P_BST, C.c_int, C.c_int, _P_BST]),
    "block_sparse_tensor_qr": (C.c_int, [_P_BST, C.c_int, _P_BST, _P_BST]),
    "block_sparse_tensor_rq": (C.c_int, [_P_BST, C.c_int, _P_BST, _P_BST]),
    "block_sparse_tensor_svd": (C.c_int, [_P_BST, _P_BST, _P_DT, _P_BST]),
    "von_neumann_entropy": (C.c_double, [C.POINTER(C.c_double), C.c_int64]),
    "retained_bond_indices": (None, [C.POINTER(C.c_double), C.c_int64, C.c_double, C.c_bool, C.c_int64, C.POINTER(IndexList), C.POINTER(TruncInfo)]),
    "delete_index_list": (None, [C.POINTER(IndexList)]),
    "split_block_sparse_matrix_svd": (C.c_int, [_P_BST, C.c_double, C.c_bool, C.c_int64, C.c_bool, C.c_int, _P_BST, _P_BST, C.POINTER(TruncInfo)]),
    "mps_local_orthonormalize_qr": (None, [_P_BST, _P_BST]),
    "mps_local_orthonormalize_rq": (None, [_P_BST, _P_BST]),
    "mps_orthonormalize_qr": (C.c_double, [C.POINTER(MPSStruct), C.c_int]),
    "mps_split_tensor_svd": (C.c_int, [_P_BST, C.POINTER(C.c_int64), C.POINTER(C.POINTER(C.c_int32)), C.c_double, C.c_int64, C.c_bool, C.c_int, _P_BST, _P_BST, C.POINTER(TruncInfo)]),
    "mps_merge_tensor_pair": (None, [_P_BST, _P_BST, _P_BST]),
    "mpo_merge_tensor_pair": (None, [_P_BST, _P_BST, _P_BST]),
    "create_dummy_operator_block_right": (None, [_P_BST, _P_BST, _P_BST, _P_BST]),
    "create_dummy_operator_block_left": (None, [_P_BST, _P_BST, _P_BST, _P_BST]),
    "contraction_operator_step_right": (None, [_P_BST, _P_BST, _P_BST, _P_BST, _P_BST]),
    "contraction_operator_step_left": (None, [_P_BST, _P_BST, _P_BST, _P_BST, _P_BST]),
    "compute_right_operator_blocks": (None, [C.POINTER(MPSStruct), C.POINTER(MPSStruct), C.POINTER(MPOStruct), _P_BST]),
    "apply_local_hamiltonian": (None, [_P_BST, _P_BST, _P_BST, _P_BST, _P_BST]),
    "lanczos_iteration_d": (None, [C.c_int64, LANCZOS_FUNC, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_int)]),
    "lanczos_iteration_z": (None, [C.c_int64, LANCZOS_FUNC, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_int)]),
    "eigensystem_krylov_symmetric": (C.c_int, [C.c_int64, LANCZOS_FUNC, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p]),
    "eigensystem_krylov_hermitian": (C.c_int, [C.c_int64, LANCZOS_FUNC, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p]),
    # other callers of the same primitives
    "mps_vdot": (None, [C.POINTER(MPSStruct), C.POINTER(MPSStruct), C.c_void_p]),
    "mps_norm": (C.c_double, [C.POINTER(MPSStruct)]),
    "mpo_inner_product": (None, [C.POINTER(MPSStruct), C.POINTER(MPOStruct), C.POINTER(MPSStruct), C.c_void_p]),
    "apply_mpo": (None, [C.POINTER(MPOStruct), C.POINTER(MPSStruct), C.POINTER(MPSStruct)]),
    "compute_local_hamiltonian_environment": (None, [_P_BST, _P_BST, _P_BST, _P_BST, _P_BST]),
    "split_block_sparse_matrix_svd_isometry": (C.c_int, [_P_BST, C.c_double, C.c_bool, C.c_int64, _P_BST, C.POINTER(TruncInfo)]),
    "mps_local_orthonormalize_left_svd": (C.c_int, [C.c_double, C.c_int64, C.c_bool, _P_BST, _P_BST, C.POINTER(TruncInfo)]),
    "mps_local_orthonormalize_right_svd": (C.c_int, [C.c_double, C.c_int64, C.c_bool, _P_BST, _P_BST, C.POINTER(TruncInfo)]),
    "mps_compress": (C.c_int, [C.c_double, C.c_int64, C.c_int, C.POINTER(MPSStruct), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(TruncInfo)]),
    "mpo_from_assembly": (None, [C.c_void_p, C.POINTER(MPOStruct)]),
    "operator_average_coefficient_gradient": (None, [C.c_void_p, C.POINTER(MPSStruct), C.POINTER(MPSStruct), C.c_void_p, C.c_void_p]),
    "mps_compress_rescale": (C.c_int, [C.c_double, C.c_int64, C.c_int, C.POINTER(MPSStruct), C.POINTER(C.c_double), C.POINTER(TruncInfo)]),
    "dmrg_singlesite": (C.c_int, [C.POINTER(MPOStruct), C.c_int, C.c_int, C.POINTER(MPSStruct), C.POINTER(C.c_double)]),
    "dmrg_twosite": (C.c_int, [C.POINTER(MPOStruct), C.c_int, C.c_int, C.c_double, C.c_int64, C.POINTER(MPSStruct), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
}

# engine extensions exported only by libchemtensor_b200.so (and its host-logic test build)
_EXT_SIGNATURES = {
    "ctb_init": (C.c_int, [C.c_int]),
    "ctb_backend": (C.c_int, []),
    "ctb_launch_count": (C.c_longlong, []),
    "ctb_heff_benchmark": (C.c_int, [_P_BST, _P_BST, _P_BST, _P_BST, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "ctb_dot_benchmark": (C.c_int, [_P_BST, C.c_int, _P_BST, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "ctb_get_stats": (C.c_int, [C.POINTER(C.c_double), C.c_int]),
    "ctb_heff_plan_info": (C.c_int, [_P_BST, _P_BST, _P_BST, _P_BST, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "ctb_dist_unique_id": (C.c_int, [C.c_void_p]),
    "ctb_dist_init": (C.c_int, [C.c_int, C.c_int, C.c_void_p]),
    "ctb_dist_set_allgather": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ctb_dist_finalize": (C.c_int, []),
    "ctb_dist_info": (C.c_int, [C.POINTER(C.c_longlong)]),
    "ctb_apply_local_hamiltonian_pair": (C.c_int, [_P_BST, _P_BST, _P_BST, _P_BST, _P_BST, _P_BST]),
    "ctb_dist_pull_exchanges": (C.c_longlong, []),
    "ctb_dist_push_exchanges": (C.c_longlong, []),
    "ctb_dist_multicast_exchanges": (C.c_longlong, []),
    "ctb_remap_benchmark": (C.c_int, [_P_BST, C.POINTER(C.c_double)]),
    "ctb_retained_bond_indices_device": (C.c_int, [C.POINTER(C.c_double), C.c_int64, C.c_double, C.c_bool, C.c_int64, C.POINTER(IndexList), C.POINTER(TruncInfo)]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES) + tuple(_EXT_SIGNATURES) + ("allocate_zero_dense_tensor", "allocate_block_sparse_tensor_like",
    "copy_block_sparse_tensor", "allocate_mps", "delete_mps", "allocate_mpo", "delete_mpo", "save_mps", "load_mps",
    # SU(2)-symmetric variant (include/ctb_su2.h; ctypes structures and bindings in tests/su2_helpers.py)
    "su2_dmrg_twosite", "su2_dmrg_singlesite", "su2_apply_local_hamiltonian", "su2_contraction_operator_step_left",
    "su2_contraction_operator_step_right", "su2_compute_right_operator_blocks", "su2_create_dummy_operator_block_left",
    "su2_create_dummy_operator_block_right", "su2_mpo_inner_product", "su2_mps_orthonormalize_qr", "su2_mps_local_orthonormalize_qr",
    "su2_mps_local_orthonormalize_rq", "su2_tensor_contract_simple", "su2_tensor_fmove", "su2_tensor_svd", "su2_recoupling_coefficient",
    "su2_tensor_num_elements_degensors", "su2_tensor_serialize_renormalized_entries", "su2_tensor_deserialize_renormalized_entries",
    "ctb_su2_apply_local_hamiltonian_pair", "ctb_su2_get_stats")


class CLibrary:
    """One loaded shared object exporting the reference's C API (all or part of it)."""

    def __init__(self, path: str, extensions: bool = False, mode: int = C.RTLD_LOCAL):
        if not os.path.exists(path):
            raise FileNotFoundError(f"shared library not found: {path}")
        self.path = path
        self.dll = C.CDLL(path, mode=mode)
        sigs = dict(_SIGNATURES)
        if extensions:
            sigs.update(_EXT_SIGNATURES)
        for name, (restype, argtypes) in sigs.items():
            try:
                fn = getattr(self.dll, name)
            except AttributeError:
                continue
            fn.restype = restype
            fn.argtypes = argtypes
            setattr(self, name, fn)

    def has(self, name: str) -> bool:
        return hasattr(self.dll, name)


# ------------------------------------------------------------------------------------------------
# numpy <-> struct helpers
# ------------------------------------------------------------------------------------------------

def _grid_size(t: BlockSparseTensor) -> int:
    n = 1
    for i in range(t.ndim):
        n *= t.dim_blocks[i]
    return n


class BST:
    """Owning Python handle of a host-memory `struct block_sparse_tensor` allocated by `lib`."""

    def __init__(self, lib: CLibrary, struct: BlockSparseTensor | None = None, owned: bool = True):
        self.lib = lib
        self.s = struct if struct is not None else BlockSparseTensor()
        self.owned = owned

    def __del__(self):
        try:
            if self.owned and self.s.blocks:
                self.lib.delete_block_sparse_tensor(C.byref(self.s))
                self.owned = False
        except Exception:
            pass

    @property
    def ptr(self):
        return C.byref(self.s)

    @property
    def ndim(self) -> int:
        return self.s.ndim

    @property
    def dtype(self):
        return np_dtype(self.s.dtype)

    @property
    def shape(self):
        return tuple(int(self.s.dim_logical[i]) for i in range(self.s.ndim))

    @property
    def axis_dir(self):
        return [int(self.s.axis_dir[i]) for i in range(self.s.ndim)]

    @property
    def qnums(self):
        return [np.ctypeslib.as_array(self.s.qnums_logical[i], shape=(int(self.s.dim_logical[i]),)).copy()
                for i in range(self.s.ndim)]

    @property
    def qnums_blocks(self):
        return [np.ctypeslib.as_array(self.s.qnums_blocks[i], shape=(int(self.s.dim_blocks[i]),)).copy() for i in range(self.s.ndim)]

    def blocks(self):
        """Yield (grid multi-index, numpy view of the block) for every stored block, in grid order."""
        t = self.s
        nsec = [int(t.dim_blocks[i]) for i in range(t.ndim)]
        ngrid = int(np.prod(nsec)) if t.ndim > 0 else 1
        for k in range(ngrid):
            bp = t.blocks[k]
            if not bp:
                continue
            b = bp.contents
            shape = tuple(int(b.dim[i]) for i in range(b.ndim))
            n = int(np.prod(shape)) if b.ndim > 0 else 1
            buf = (C.c_byte * (n * np.dtype(self.dtype).itemsize)).from_address(b.data)
            arr = np.frombuffer(buf, dtype=self.dtype, count=n).reshape(shape)
            idx = np.unravel_index(k, nsec) if t.ndim > 0 else ()
            yield tuple(int(i) for i in idx), arr

    def num_elements(self) -> int:
        return sum(a.size for _, a in self.blocks())

    def serialize(self) -> np.ndarray:
        parts = [a.reshape(-1) for _, a in self.blocks()]
        return np.concatenate(parts) if parts else np.zeros(0, dtype=self.dtype)

    def deserialize(self, v: np.ndarray) -> None:
        pos = 0
        for _, a in self.blocks():
            a[...] = v[pos:pos + a.size].reshape(a.shape)
            pos += a.size

    def to_dense(self) -> np.ndarray:
        out = np.zeros(self.shape, dtype=self.dtype)
        if self.ndim == 0:
            for _, a in self.blocks():
                out[...] = a
            return out
        qn = self.qnums
        qb = self.qnums_blocks
        for idx, a in self.blocks():
            sel = [np.nonzero(qn[i] == qb[i][idx[i]])[0] for i in range(self.ndim)]
            out[np.ix_(*sel)] = a
        return out

    def fill_from_dense(self, dense: np.ndarray) -> None:
        qn = self.qnums
        qb = self.qnums_blocks
        for idx, a in self.blocks():
            sel = [np.nonzero(qn[i] == qb[i][idx[i]])[0] for i in range(self.ndim)]
            a[...] = dense[np.ix_(*sel)]


def _int_array(ctype, values):
    arr = (ctype * len(values))(*[int(v) for v in values])
    return arr


def _qnum_ptrs(qnums):
    keep = [np.ascontiguousarray(q, dtype=np.int32) for q in qnums]
    ptrs = (C.POINTER(C.c_int32) * len(keep))(*[k.ctypes.data_as(C.POINTER(C.c_int32)) for k in keep])
    return ptrs, keep


def bst_allocate(lib: CLibrary, dtype, shape, axis_dir, qnums) -> BST:
    t = BST(lib)
    ndim = len(shape)
    dims = _int_array(C.c_int64, shape)
    dirs = _int_array(C.c_int, axis_dir)
    ptrs, keep = _qnum_ptrs(qnums)
    lib.allocate_block_sparse_tensor(ct_dtype(dtype), ndim, dims, dirs, ptrs, t.ptr)
    return t


def bst_from_dense(lib: CLibrary, dense: np.ndarray, axis_dir, qnums) -> BST:
    t = bst_allocate(lib, dense.dtype, dense.shape, axis_dir, qnums)
    t.fill_from_dense(dense)
    return t


def bst_clone(lib: CLibrary, src: BST) -> BST:
    """Deep copy of `src` into memory owned by `lib`."""
    t = bst_allocate(lib, src.dtype, src.shape, src.axis_dir, src.qnums)
    for (_, a), (_, b) in zip(t.blocks(), src.blocks()):
        a[...] = b
    return t


def conserving_mask(shape, axis_dir, qnums) -> np.ndarray:
    """Boolean mask of the logical entries allowed by sum_i dir_i q_i = 0."""
    tot = np.zeros(shape, dtype=np.int64)
    for i, (d, q) in enumerate(zip(axis_dir, qnums)):
        sh = [1] * len(shape)
        sh[i] = shape[i]
        tot = tot + d * np.asarray(q, dtype=np.int64).reshape(sh)
    return tot == 0


def dense_vector(lib: CLibrary, v: np.ndarray):
    """A `struct dense_tensor` view of a 1-d float64 numpy array (memory stays owned by numpy)."""
    v = np.ascontiguousarray(v, dtype=np.float64)
    dt = DenseTensor()
    dim = (C.c_int64 * 1)(v.size)
    dt.data = v.ctypes.data
    dt.dim = C.cast(dim, C.POINTER(C.c_int64))
    dt.dtype = CT_DOUBLE_REAL
    dt.ndim = 1
    return dt, (v, dim)


class Chain:
    """Owning handle of a `struct mps` or `struct mpo` whose site tensors live in memory of `lib`."""

    def __init__(self, lib: CLibrary, kind: str, qsite, tensors: list[BST]):
        self.lib = lib
        self.kind = kind
        self.qsite = np.ascontiguousarray(qsite, dtype=np.int32)
        self.nsites = len(tensors)
        self.arr = (BlockSparseTensor * self.nsites)()
        for i, t in enumerate(tensors):
            # move the payload into the array; the BST handle gives up ownership
            C.memmove(C.byref(self.arr[i]), C.byref(t.s), C.sizeof(BlockSparseTensor))
            t.owned = False
        self.s = MPSStruct() if kind == "mps" else MPOStruct()
        self.s.a = C.cast(self.arr, C.POINTER(BlockSparseTensor))
        self.s.qsite = self.qsite.ctypes.data_as(C.POINTER(C.c_int32))
        self.s.d = len(self.qsite)
        self.s.nsites = self.nsites

    def __del__(self):
        try:
            for i in range(self.nsites):
                if self.arr[i].blocks:
                    self.lib.delete_block_sparse_tensor(C.byref(self.arr[i]))
        except Exception:
            pass

    @property
    def ptr(self):
        return C.byref(self.s)

    def site(self, i: int) -> BST:
        """Non-owning view of site tensor i (valid while the chain lives and is not modified)."""
        return BST(self.lib, self.arr[i], owned=False)

    def bond_dims(self):
        last = self.nsites - 1
        ax = 2 if self.kind == "mps" else 3
        return [int(self.arr[i].dim_logical[0]) for i in range(self.nsites)] + [int(self.arr[last].dim_logical[ax])]
