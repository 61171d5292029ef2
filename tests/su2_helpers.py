"""ctypes view of the SU(2) structs (include/ctb_su2.h = the reference's include/tensor/su2_tensor.h etc.) and reference-side generators.

TEST INFRASTRUCTURE.  Inputs (Heisenberg / Fermi-Hubbard SU(2) MPOs, random SU(2) MPS) come from the unmodified reference in
oracle/_ref; the same host structs are then handed to the engine ("emu" = host logic on the CPU test double, "cuda" = the product).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

import helpers
from chemtensor_b200 import cabi


class TreeNode(C.Structure):
    pass


TreeNode._fields_ = [("i_ax", C.c_int), ("c", C.POINTER(TreeNode) * 2)]


class FuseSplitTree(C.Structure):
    _fields_ = [("tree_fuse", C.POINTER(TreeNode)), ("tree_split", C.POINTER(TreeNode)), ("ndim", C.c_int)]


class IrredList(C.Structure):
    _fields_ = [("jlist", C.POINTER(C.c_int32)), ("num", C.c_int)]


class ChargeSectors(C.Structure):
    _fields_ = [("jlists", C.POINTER(C.c_int32)), ("nsec", C.c_int64), ("ndim", C.c_int)]


class SU2Tensor(C.Structure):
    _fields_ = [("tree", FuseSplitTree), ("outer_irreps", C.POINTER(IrredList)), ("charge_sectors", ChargeSectors),
                ("degensors", C.POINTER(C.POINTER(cabi.DenseTensor))), ("dim_degen", C.POINTER(C.POINTER(C.c_int64))),
                ("dtype", C.c_int), ("ndim_logical", C.c_int), ("ndim_auxiliary", C.c_int)]


class SU2MPS(C.Structure):
    _fields_ = [("a", C.POINTER(SU2Tensor)), ("nsites", C.c_int)]


class SU2MPO(C.Structure):
    _fields_ = [("a", C.POINTER(SU2Tensor)), ("nsites", C.c_int)]


_bound: set[int] = set()


def bind(dll) -> None:
    """argument types of the SU(2) entry points (same in the reference and in the engine)"""
    if id(dll) in _bound:
        return
    _bound.add(id(dll))
    T = C.POINTER(SU2Tensor)
    ip = C.POINTER(C.c_int)
    for name, res, args in [
        ("su2_tensor_contract_simple", None, [T, ip, T, ip, C.c_int, T]),
        ("su2_tensor_fmove", None, [T, C.c_int, T]),
        ("su2_apply_local_hamiltonian", None, [T, T, T, T, T]),
        ("su2_contraction_operator_step_left", None, [T, T, T, T, T]),
        ("su2_contraction_operator_step_right", None, [T, T, T, T, T]),
        ("su2_create_dummy_operator_block_left", None, [C.c_int, T]),
        ("su2_create_dummy_operator_block_right", None, [C.c_int, C.c_int32, T]),
        ("su2_compute_right_operator_blocks", None, [C.POINTER(SU2MPS), C.POINTER(SU2MPS), C.POINTER(SU2MPO), T]),
        ("su2_mpo_inner_product", None, [C.POINTER(SU2MPS), C.POINTER(SU2MPO), C.POINTER(SU2MPS), C.c_void_p]),
        ("su2_mps_orthonormalize_qr", C.c_double, [C.POINTER(SU2MPS), C.c_int]),
        ("su2_dmrg_singlesite", C.c_int, [C.POINTER(SU2MPO), C.c_int, C.c_int, C.POINTER(SU2MPS), C.POINTER(C.c_double)]),
        ("su2_dmrg_twosite", C.c_int, [C.POINTER(SU2MPO), C.c_int, C.c_int, C.c_double, C.c_int64, C.POINTER(SU2MPS),
                                        C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        ("su2_recoupling_coefficient", C.c_double, [C.c_int32] * 6),
        ("su2_mps_local_orthonormalize_qr", None, [T, T]),
        ("su2_mps_local_orthonormalize_rq", None, [T, T]),
    ]:
        f = getattr(dll, name)
        f.restype = res
        f.argtypes = args


def bind_ref(dll) -> None:
    bind(dll)
    if getattr(dll, "_su2_ref_bound", False):
        return
    dll._su2_ref_bound = True
    T = C.POINTER(SU2Tensor)
    dll.construct_heisenberg_1d_su2_mpo.restype = None
    dll.construct_heisenberg_1d_su2_mpo.argtypes = [C.c_int, C.c_double, C.POINTER(SU2MPO)]
    dll.construct_fermi_hubbard_1d_su2_mpo.restype = None
    dll.construct_fermi_hubbard_1d_su2_mpo.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.POINTER(SU2MPO)]
    dll.construct_random_su2_mps.restype = None
    dll.construct_random_su2_mps.argtypes = [C.c_int, C.c_int, C.POINTER(IrredList), C.POINTER(C.c_int64), C.c_int32, C.c_int32, C.c_int64,
                                             C.c_void_p, C.POINTER(SU2MPS)]
    dll.construct_random_su2_mpo.restype = None
    dll.construct_random_su2_mpo.argtypes = [C.c_int, C.c_int, C.POINTER(IrredList), C.POINTER(C.c_int64), C.c_int32, C.c_int64, C.c_void_p,
                                             C.POINTER(SU2MPO)]
    dll.seed_rng_state.restype = None
    dll.seed_rng_state.argtypes = [C.c_uint64, C.c_void_p]
    dll.delete_su2_tensor.restype = None
    dll.delete_su2_tensor.argtypes = [T]
    dll.delete_su2_mps.restype = None
    dll.delete_su2_mps.argtypes = [C.POINTER(SU2MPS)]
    dll.delete_su2_mpo.restype = None
    dll.delete_su2_mpo.argtypes = [C.POINTER(SU2MPO)]
    dll.copy_su2_tensor.restype = None
    dll.copy_su2_tensor.argtypes = [T, T]
    dll.su2_tensor_allclose.restype = C.c_bool
    dll.su2_tensor_allclose.argtypes = [T, T, C.c_double]
    dll.su2_tensor_norm2.restype = C.c_double
    dll.su2_tensor_norm2.argtypes = [T]
    dll.rscale_su2_tensor.restype = None
    dll.rscale_su2_tensor.argtypes = [C.c_void_p, T]
    dll.su2_mps_is_consistent.restype = C.c_bool
    dll.su2_mps_is_consistent.argtypes = [C.POINTER(SU2MPS)]
    dll.su2_mps_to_statevector.restype = None
    dll.su2_mps_to_statevector.argtypes = [C.POINTER(SU2MPS), T]
    dll.su2_to_dense_tensor.restype = None
    dll.su2_to_dense_tensor.argtypes = [T, C.POINTER(cabi.DenseTensor)]
    dll.delete_dense_tensor.restype = None
    dll.delete_dense_tensor.argtypes = [C.POINTER(cabi.DenseTensor)]
    dll.su2_mps_contract_tensor_pair.restype = None
    dll.su2_mps_contract_tensor_pair.argtypes = [T, T, T]
    dll.su2_mps_merge_tensor_pair.restype = None
    dll.su2_mps_merge_tensor_pair.argtypes = [T, T, T]
    dll.su2_mpo_merge_tensor_pair.restype = None
    dll.su2_mpo_merge_tensor_pair.argtypes = [T, T, T]
    dll.su2_tensor_fuse_axes.restype = None
    dll.su2_tensor_fuse_axes.argtypes = [T, C.c_int, C.c_int, T]
    dll.su2_tensor_fill_random_normal.restype = None
    dll.su2_tensor_fill_random_normal.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, T]


def ref():
    lib = helpers.load("ref")
    bind_ref(lib.dll)
    return lib.dll


def engine(kind: str):
    lib = helpers.load(kind)
    bind(lib.dll)
    return lib.dll


def rng(seed: int):
    st = C.create_string_buffer(64)
    ref().seed_rng_state(seed, st)
    return st


def heisenberg_mpo(nsites: int, J: float) -> SU2MPO:
    mpo = SU2MPO()
    ref().construct_heisenberg_1d_su2_mpo(nsites, J, C.byref(mpo))
    return mpo


def fermi_hubbard_mpo(nsites: int, t: float, u: float, mu: float) -> SU2MPO:
    mpo = SU2MPO()
    ref().construct_fermi_hubbard_1d_su2_mpo(nsites, t, u, mu, C.byref(mpo))
    return mpo


def random_mps(nsites: int, site_jlist, site_dim_degen, irrep_sector: int, max_bond_irrep: int, max_bond_dim_degen: int, seed: int,
               dtype: int = 1, scale: float | None = None) -> SU2MPS:
    jl = (C.c_int32 * len(site_jlist))(*site_jlist)
    irr = IrredList(C.cast(jl, C.POINTER(C.c_int32)), len(site_jlist))
    dd = (C.c_int64 * len(site_dim_degen))(*site_dim_degen)
    st = rng(seed)
    psi = SU2MPS()
    ref().construct_random_su2_mps(dtype, nsites, C.byref(irr), dd, irrep_sector, max_bond_irrep, max_bond_dim_degen, st, C.byref(psi))
    if scale is not None:
        alpha = C.c_double(scale)
        for i in range(nsites):
            ref().rscale_su2_tensor(C.byref(alpha), C.byref(psi.a[i]))
    return psi


def copy_mps(psi: SU2MPS) -> SU2MPS:
    """deep copy through the reference (every tensor separately allocated, array of structs from the C heap)"""
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    out = SU2MPS()
    out.nsites = psi.nsites
    out.a = C.cast(libc.malloc(C.sizeof(SU2Tensor) * psi.nsites), C.POINTER(SU2Tensor))
    for i in range(psi.nsites):
        ref().copy_su2_tensor(C.byref(psi.a[i]), C.byref(out.a[i]))
    return out


def sectors(t: SU2Tensor) -> np.ndarray:
    n, d = t.charge_sectors.nsec, t.charge_sectors.ndim
    if n == 0:
        return np.zeros((0, d), dtype=np.int32)
    return np.ctypeslib.as_array(t.charge_sectors.jlists, shape=(n * d,)).reshape(n, d).copy()


def degensor(t: SU2Tensor, c: int) -> np.ndarray:
    d = t.degensors[c].contents
    shape = tuple(d.dim[i] for i in range(d.ndim))
    n = int(np.prod(shape)) if shape else 1
    dt = np.float64 if t.dtype == 1 else np.complex128
    buf = (C.c_char * (n * np.dtype(dt).itemsize)).from_address(d.data)
    return np.frombuffer(buf, dtype=dt).reshape(shape).copy()


def tree_tuple(node) -> tuple:
    n = node.contents
    if not n.c[0]:
        return (n.i_ax,)
    return (n.i_ax, tree_tuple(n.c[0]), tree_tuple(n.c[1]))


def assert_same_su2(x: SU2Tensor, y: SU2Tensor, tol: float) -> float:
    """structure bit-exact (trees, irreducible lists, degeneracy dimensions, sector table), entries within tol (relative to the largest entry)"""
    assert x.ndim_logical == y.ndim_logical and x.ndim_auxiliary == y.ndim_auxiliary and x.dtype == y.dtype
    assert tree_tuple(x.tree.tree_fuse) == tree_tuple(y.tree.tree_fuse)
    assert tree_tuple(x.tree.tree_split) == tree_tuple(y.tree.tree_split)
    for i in range(x.ndim_logical + x.ndim_auxiliary):
        a = [x.outer_irreps[i].jlist[k] for k in range(x.outer_irreps[i].num)]
        b = [y.outer_irreps[i].jlist[k] for k in range(y.outer_irreps[i].num)]
        assert a == b, (i, a, b)
        if i < x.ndim_logical:
            for j in a:
                assert x.dim_degen[i][j] == y.dim_degen[i][j]
    sx, sy = sectors(x), sectors(y)
    assert sx.shape == sy.shape and np.array_equal(sx, sy), (sx, sy)
    err, ref_mag = 0.0, 0.0
    for c in range(sx.shape[0]):
        a, b = degensor(x, c), degensor(y, c)
        assert a.shape == b.shape
        err = max(err, float(np.max(np.abs(a - b))) if a.size else 0.0)
        ref_mag = max(ref_mag, float(np.max(np.abs(b))) if b.size else 0.0)
    assert err <= tol * max(ref_mag, 1e-300), (err, ref_mag)
    return err / max(ref_mag, 1e-300)


def complexify(t: SU2Tensor) -> None:
    """in place: real double SU(2) tensor -> complex double with zero imaginary parts (the reference has no complex SU(2) Hamiltonian)"""
    libc = C.CDLL(None)
    libc.malloc.restype = C.c_void_p
    for c in range(t.charge_sectors.nsec):
        d = t.degensors[c].contents
        n = 1
        for i in range(d.ndim):
            n *= d.dim[i]
        old = np.ctypeslib.as_array(C.cast(d.data, C.POINTER(C.c_double)), shape=(n,)).copy()
        buf = libc.malloc(16 * max(n, 1))
        arr = np.ctypeslib.as_array(C.cast(buf, C.POINTER(C.c_double)), shape=(2 * n,))
        arr[0::2] = old
        arr[1::2] = 0.0
        d.data = buf
        d.dtype = 3
    t.dtype = 3
