"""Shared test infrastructure: library loading, random block-sparse inputs, reference-side generators.

Three shared objects speak the same reference C API (chemtensor_b200/cabi.py):
  * "ref"  oracle/_ref/libchemtensor_ref.so  -- the UNMODIFIED reference, compiled by oracle/Makefile (the checker)
  * "cuda" chemtensor_b200/libchemtensor_b200.so -- the product (host C + sm_100a kernels), needs a GPU
  * "emu"  tests/emu/libctb_hostlogic_emu.so -- the product's host C linked to the CPU test double of the CUDA
           layer; lets the sector bookkeeping / plan builders / sweep logic be tested without a GPU
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from chemtensor_b200 import cabi  # noqa: E402

REF_SO = os.path.join(ROOT, "oracle", "_ref", "libchemtensor_ref.so")
CUDA_SO = os.path.join(ROOT, "chemtensor_b200", "libchemtensor_b200.so")
EMU_SO = os.path.join(ROOT, "tests", "emu", "libctb_hostlogic_emu.so")

_libs: dict[str, cabi.CLibrary] = {}


def _make(target: str) -> None:
    subprocess.run(["make", "-s", "-C", ROOT, target], check=True, stdout=subprocess.DEVNULL)


def have_gpu() -> bool:
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def load(kind: str) -> cabi.CLibrary:
    if kind in _libs:
        return _libs[kind]
    if kind == "ref":
        if not os.path.exists(REF_SO):
            _make("oracle")
        lib = cabi.CLibrary(REF_SO)
        _bind_reference_generators(lib)
    elif kind == "emu":
        _make("emu")
        lib = cabi.CLibrary(EMU_SO, extensions=True)
        assert lib.ctb_backend() == 2
    elif kind == "cuda":
        if not os.path.exists(CUDA_SO):
            _make("lib")
        lib = cabi.CLibrary(CUDA_SO, extensions=True)
        assert lib.ctb_backend() == 1
        rc = lib.ctb_init(-1)
        if rc < 0:
            raise RuntimeError("libchemtensor_b200.so: no usable CUDA device (the engine has no CPU fallback)")
    else:
        raise ValueError(kind)
    _libs[kind] = lib
    return lib


# ------------------------------------------------------------------------------------------------
# reference-only generators (inputs): Hamiltonian MPOs, random MPS, RNG
# ------------------------------------------------------------------------------------------------

def _bind_reference_generators(lib: cabi.CLibrary) -> None:
    d = lib.dll
    vp = C.c_void_p
    d.seed_rng_state.restype = None
    d.seed_rng_state.argtypes = [C.c_uint64, vp]
    d.construct_random_mps.restype = None
    d.construct_random_mps.argtypes = [C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_int32), C.c_int32, C.c_int64, vp, C.POINTER(cabi.MPSStruct)]
    d.construct_heisenberg_xxz_1d_mpo_assembly.restype = None
    d.construct_heisenberg_xxz_1d_mpo_assembly.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, vp]
    d.construct_fermi_hubbard_1d_mpo_assembly.restype = None
    d.construct_fermi_hubbard_1d_mpo_assembly.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, vp]
    d.construct_spin_molecular_hamiltonian_mpo_assembly.restype = None
    d.construct_spin_molecular_hamiltonian_mpo_assembly.argtypes = [C.POINTER(cabi.DenseTensor), C.POINTER(cabi.DenseTensor), C.c_bool, vp]
    d.construct_molecular_hamiltonian_mpo_assembly.restype = None
    d.construct_molecular_hamiltonian_mpo_assembly.argtypes = [C.POINTER(cabi.DenseTensor), C.POINTER(cabi.DenseTensor), C.c_bool, vp]
    d.mpo_from_assembly.restype = None
    d.mpo_from_assembly.argtypes = [vp, C.POINTER(cabi.MPOStruct)]
    d.delete_mpo_assembly.restype = None
    d.delete_mpo_assembly.argtypes = [vp]
    d.delete_mpo.restype = None
    d.delete_mpo.argtypes = [C.POINTER(cabi.MPOStruct)]
    d.delete_mps.restype = None
    d.delete_mps.argtypes = [C.POINTER(cabi.MPSStruct)]
    d.mps_norm.restype = C.c_double
    d.mps_norm.argtypes = [C.POINTER(cabi.MPSStruct)]
    d.mps_vdot.restype = None
    d.mps_vdot.argtypes = [C.POINTER(cabi.MPSStruct), C.POINTER(cabi.MPSStruct), vp]


class RefChain:
    """A `struct mps`/`struct mpo` fully owned by the reference library."""

    def __init__(self, ref: cabi.CLibrary, kind: str):
        self.ref = ref
        self.kind = kind
        self.s = cabi.MPSStruct() if kind == "mps" else cabi.MPOStruct()
        self.alive = False

    def __del__(self):
        try:
            if self.alive:
                (self.ref.dll.delete_mps if self.kind == "mps" else self.ref.dll.delete_mpo)(C.byref(self.s))
                self.alive = False
        except Exception:
            pass

    @property
    def ptr(self):
        return C.byref(self.s)

    @property
    def nsites(self) -> int:
        return self.s.nsites

    @property
    def qsite(self) -> np.ndarray:
        return np.ctypeslib.as_array(self.s.qsite, shape=(int(self.s.d),)).copy()

    def site(self, i: int) -> cabi.BST:
        return cabi.BST(self.ref, self.s.a[i], owned=False)

    def bond_dims(self):
        ax = 2 if self.kind == "mps" else 3
        return [int(self.s.a[i].dim_logical[0]) for i in range(self.nsites)] + [int(self.s.a[self.nsites - 1].dim_logical[ax])]


def encode_qpair(qa: int, qb: int) -> int:
    """reference include/tensor/qnumber.h:47 -- (qa << 16) + qb"""
    return (qa << 16) + qb


def ref_mpo(ref: cabi.CLibrary, model: str, nsites: int, *params) -> RefChain:
    asm = C.create_string_buffer(512)
    if model == "xxz":
        ref.dll.construct_heisenberg_xxz_1d_mpo_assembly(nsites, *[float(p) for p in params], asm)
    elif model == "fermi_hubbard":
        ref.dll.construct_fermi_hubbard_1d_mpo_assembly(nsites, *[float(p) for p in params], asm)
    else:
        raise ValueError(model)
    mpo = RefChain(ref, "mpo")
    ref.dll.mpo_from_assembly(asm, mpo.ptr)
    mpo.alive = True
    ref.dll.delete_mpo_assembly(asm)
    return mpo


def perf_dmrg_coeffs(nsites: int = 9, seed: int = 42):
    """The integrals of the reference's perf/perf_dmrg_coeffs.py:6-17, in the same order of random draws (bit-identical to the
    datasets of perf/perf_dmrg_coeffs.hdf5 under numpy's default_rng; checked in tests/test_oracle_golden.py)."""
    rng = np.random.default_rng(seed)
    tkin = rng.standard_normal(2 * (nsites,))
    vint = rng.standard_normal(4 * (nsites,))
    tkin = 0.5 * (tkin + tkin.conj().T)
    vint = 0.5 * (vint + vint.conj().transpose(2, 3, 0, 1))
    return tkin, vint


def _dense_struct(arr: np.ndarray):
    arr = np.ascontiguousarray(arr)
    dt = cabi.DenseTensor()
    dim = (C.c_int64 * arr.ndim)(*arr.shape)
    dt.data = arr.ctypes.data
    dt.dim = C.cast(dim, C.POINTER(C.c_int64))
    dt.dtype = cabi.ct_dtype(arr.dtype)
    dt.ndim = arr.ndim
    return dt, (arr, dim)


def ref_molecular_mpo(ref: cabi.CLibrary, tkin: np.ndarray, vint: np.ndarray, spin: bool = True, optimize: bool = False) -> RefChain:
    asm = C.create_string_buffer(512)
    t, keep_t = _dense_struct(tkin)
    v, keep_v = _dense_struct(vint)
    fn = ref.dll.construct_spin_molecular_hamiltonian_mpo_assembly if spin else ref.dll.construct_molecular_hamiltonian_mpo_assembly
    fn(C.byref(t), C.byref(v), bool(optimize), asm)
    mpo = RefChain(ref, "mpo")
    ref.dll.mpo_from_assembly(asm, mpo.ptr)
    mpo.alive = True
    ref.dll.delete_mpo_assembly(asm)
    return mpo


def ref_random_mps(ref: cabi.CLibrary, dtype, nsites: int, qsite, qnum_sector: int, max_vdim: int, seed: int = 42) -> RefChain:
    rng = C.create_string_buffer(64)
    ref.dll.seed_rng_state(seed, rng)
    q = np.ascontiguousarray(qsite, dtype=np.int32)
    mps = RefChain(ref, "mps")
    ref.dll.construct_random_mps(cabi.ct_dtype(dtype), nsites, len(q), q.ctypes.data_as(C.POINTER(C.c_int32)), int(qnum_sector), int(max_vdim), rng, mps.ptr)
    mps.alive = True
    return mps


def clone_chain(lib: cabi.CLibrary, src) -> cabi.Chain:
    """Deep copy of a chain (RefChain or Chain) into memory owned by `lib`."""
    tensors = [cabi.bst_clone(lib, src.site(i)) for i in range(src.nsites)]
    return cabi.Chain(lib, src.kind, src.qsite, tensors)


# ------------------------------------------------------------------------------------------------
# random block-sparse inputs
# ------------------------------------------------------------------------------------------------

def random_qnums(rng: np.random.Generator, dim: int, lo: int = -2, hi: int = 3) -> np.ndarray:
    return rng.integers(lo, hi, size=dim).astype(np.int32)


def random_dense(rng: np.random.Generator, dtype, shape, axis_dir, qnums) -> np.ndarray:
    """Dense array with random entries wherever quantum numbers are conserved, zero elsewhere."""
    mask = cabi.conserving_mask(shape, axis_dir, qnums)
    a = rng.standard_normal(shape)
    if np.dtype(dtype).kind == "c":
        a = a + 1j * rng.standard_normal(shape)
    return (a * mask).astype(dtype)


def random_bst(lib: cabi.CLibrary, rng: np.random.Generator, dtype, shape, axis_dir, qnums) -> cabi.BST:
    return cabi.bst_from_dense(lib, random_dense(rng, dtype, shape, axis_dir, qnums), axis_dir, qnums)


def assert_same_structure(x: cabi.BST, y: cabi.BST) -> None:
    """Sector structure and indexing bit-exact: dims, directions, logical and block quantum numbers, block grid."""
    assert x.ndim == y.ndim
    assert x.shape == y.shape
    assert x.axis_dir == y.axis_dir
    assert x.s.dtype == y.s.dtype
    for a, b in zip(x.qnums, y.qnums):
        assert np.array_equal(a, b)
    for a, b in zip(x.qnums_blocks, y.qnums_blocks):
        assert np.array_equal(a, b)
    bx = [(idx, a.shape) for idx, a in x.blocks()]
    by = [(idx, a.shape) for idx, a in y.blocks()]
    assert bx == by


def rel_err(x: np.ndarray, y: np.ndarray) -> float:
    """Relative Frobenius-norm error of x against y."""
    ny = np.linalg.norm(y.reshape(-1))
    d = np.linalg.norm((x - y).reshape(-1))
    return float(d / ny) if ny > 0 else float(d)


def assert_bst_close(x: cabi.BST, y: cabi.BST, tol: float) -> None:
    assert_same_structure(x, y)
    err = rel_err(x.serialize(), y.serialize())
    assert err <= tol, f"relative Frobenius error {err:.3e} > {tol:.1e}"
