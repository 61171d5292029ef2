"""mpo_from_assembly (sparse-direct, no dense Dw x d x d x Dw' intermediate) and operator_average_coefficient_gradient
(reference src/operator/mpo.c:59, src/algorithm/gradient.c:15) against the compiled reference on the reference's own MPO
assemblies (XXZ, Fermi-Hubbard, real and complex molecular Hamiltonians)."""
import ctypes as C

import numpy as np
import pytest

import helpers
from chemtensor_b200 import cabi


class MpoGraph(C.Structure):
    _fields_ = [("verts", C.c_void_p), ("edges", C.c_void_p), ("num_verts", C.POINTER(C.c_int)), ("num_edges", C.POINTER(C.c_int)), ("nsites", C.c_int)]


class MpoAssembly(C.Structure):
    """reference include/operator/mpo.h:14-24"""
    _fields_ = [("graph", MpoGraph), ("opmap", C.c_void_p), ("coeffmap", C.c_void_p), ("qsite", C.POINTER(C.c_int32)), ("d", C.c_int64),
                ("dtype", C.c_int), ("num_local_ops", C.c_int), ("num_coeffs", C.c_int)]


def _assembly(ref, model, dtype=np.float64):
    asm = MpoAssembly()
    if model == "xxz":
        ref.dll.construct_heisenberg_xxz_1d_mpo_assembly(7, 1.0, 0.8, 0.1, C.byref(asm))
        sector, max_vdim = 1, 16
    elif model == "fermi_hubbard":
        ref.dll.construct_fermi_hubbard_1d_mpo_assembly(6, 1.0, 4.0, 0.3, C.byref(asm))
        sector, max_vdim = helpers.encode_qpair(6, 0), 20
    else:
        n = 4
        rng = np.random.default_rng(9)
        cplx = np.dtype(dtype).kind == "c"
        tkin = 0.5 * (rng.standard_normal((n, n)) + (1j * rng.standard_normal((n, n)) if cplx else 0))
        vint = 0.1 * (rng.standard_normal((n, n, n, n)) + (1j * rng.standard_normal((n, n, n, n)) if cplx else 0))
        tkin = 0.5 * (tkin + tkin.conj().T)
        vint = 0.5 * (vint + vint.transpose((1, 0, 3, 2)))
        vint = 0.5 * (vint + vint.transpose((2, 3, 0, 1)).conj())
        t, keep_t = helpers._dense_struct(np.ascontiguousarray(tkin.astype(dtype)))
        v, keep_v = helpers._dense_struct(np.ascontiguousarray(vint.astype(dtype)))
        ref.dll.construct_spin_molecular_hamiltonian_mpo_assembly(C.byref(t), C.byref(v), False, C.byref(asm))
        sector, max_vdim = helpers.encode_qpair(n, 0), 16
    return asm, sector, max_vdim


CASES = [("xxz", np.float64), ("fermi_hubbard", np.float64), ("molecular", np.float64), ("molecular", np.complex128)]
IDS = [f"{m}-{np.dtype(d).name}" for m, d in CASES]


@pytest.mark.parametrize("model,dtype", CASES, ids=IDS)
def test_mpo_from_assembly(eng, ref, model, dtype):
    asm, _, _ = _assembly(ref, model, dtype)
    assert asm.dtype == cabi.ct_dtype(dtype)
    mpo_r, mpo_e = helpers.RefChain(ref, "mpo"), helpers.RefChain(eng, "mpo")
    ref.mpo_from_assembly(C.byref(asm), mpo_r.ptr); mpo_r.alive = True
    eng.mpo_from_assembly(C.byref(asm), mpo_e.ptr); mpo_e.alive = True
    assert mpo_e.nsites == mpo_r.nsites and np.array_equal(mpo_e.qsite, mpo_r.qsite)
    for i in range(mpo_r.nsites):
        helpers.assert_bst_close(mpo_e.site(i), mpo_r.site(i), 1e-15)      # structure bit-exact; entries are sums of the same products
    ref.dll.delete_mpo_assembly(C.byref(asm))


@pytest.mark.parametrize("model,dtype", CASES, ids=IDS)
def test_operator_average_coefficient_gradient(eng, ref, model, dtype):
    asm, sector, max_vdim = _assembly(ref, model, dtype)
    qsite = np.ctypeslib.as_array(asm.qsite, shape=(int(asm.d),)).copy()
    L = asm.graph.nsites
    psi_r = helpers.ref_random_mps(ref, dtype, L, qsite, sector, max_vdim, seed=31)
    chi_r = helpers.ref_random_mps(ref, dtype, L, qsite, sector, max_vdim, seed=32)
    psi_e, chi_e = helpers.clone_chain(eng, psi_r), helpers.clone_chain(eng, chi_r)
    nc = asm.num_coeffs
    ref.dll.operator_average_coefficient_gradient.restype = None
    ref.dll.operator_average_coefficient_gradient.argtypes = [C.c_void_p, C.POINTER(cabi.MPSStruct), C.POINTER(cabi.MPSStruct), C.c_void_p, C.c_void_p]
    for bra in ("chi", "psi"):       # a different bra (tiny overlap) and the expectation value <psi|op|psi>
        out = []
        for lib, psi, chi in ((eng, psi_e, chi_e), (ref, psi_r, chi_r)):
            avr = np.zeros(1, dtype=dtype); dc = np.zeros(nc, dtype=dtype)
            lib.operator_average_coefficient_gradient(C.byref(asm), psi.ptr, (chi if bra == "chi" else psi).ptr, avr.ctypes.data, dc.ctypes.data)
            out.append((avr[0], dc))
        (avr_e, dc_e), (avr_r, dc_r) = out
        assert np.max(np.abs(dc_r)) > 0
        assert abs(avr_e - avr_r) <= 1e-10 * max(abs(avr_r), np.max(np.abs(dc_r)))
        assert np.max(np.abs(dc_e - dc_r)) <= 1e-10 * np.max(np.abs(dc_r))
    ref.dll.delete_mpo_assembly(C.byref(asm))
