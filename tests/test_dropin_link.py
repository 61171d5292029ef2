"""The link recipe of INTEGRATION.md section 2, executed: the reference's own objects MINUS the replaced translation units (overlapping
functions renamed away at compile time), linked against the engine (oracle/Makefile target `dropin`).  Through that ONE library the
reference's generators (construct_*_mpo_assembly, mpo_from_assembly, construct_random_mps, src/operator/hamiltonian.c, src/state/mps.c)
feed the engine's dmrg_twosite exactly as perf/perf_dmrg.c:45-96 does; energies must match the pure reference build."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers
from chemtensor_b200 import cabi

VARIANTS = [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


def load_dropin(kind):
    path = os.path.join(helpers.ROOT, "oracle", "_ref", f"libchemtensor_dropin_{kind}.so")
    if not os.path.exists(path):
        helpers._make("dropin")
    lib = cabi.CLibrary(path, extensions=True)
    helpers._bind_reference_generators(lib)
    return lib


@pytest.mark.parametrize("kind", VARIANTS)
def test_reference_callers_reach_the_engine(kind):
    lib = load_dropin(kind)
    ref = helpers.load("ref")
    # the hot-path symbols of the combined library are the engine's (backend 2 = test double, 1 = CUDA) ...
    assert lib.ctb_backend() == (2 if kind == "emu" else 1)
    dmrg_addr = C.cast(lib.dll.dmrg_twosite, C.c_void_p).value
    eng_so = helpers.EMU_SO if kind == "emu" else helpers.CUDA_SO
    eng = C.CDLL(eng_so)
    assert dmrg_addr == C.cast(eng.dmrg_twosite, C.c_void_p).value
    assert C.cast(lib.dll.block_sparse_tensor_dot, C.c_void_p).value == C.cast(eng.block_sparse_tensor_dot, C.c_void_p).value
    # ... while the generators are the reference's own code
    assert hasattr(lib.dll, "construct_fermi_hubbard_1d_mpo_assembly") and hasattr(lib.dll, "ref_block_sparse_tensor_dot")
    if kind == "cuda":
        assert lib.ctb_init(-1) == 0
    # perf_dmrg-style run: reference generators -> engine sweep, all inside the combined library
    L, D = 8, 40
    sector = helpers.encode_qpair(L, 0)
    results = {}
    for name, x in (("dropin", lib), ("reference", ref)):
        mpo = helpers.ref_mpo(x, "fermi_hubbard", L, 1.0, 4.0, 0.0)
        psi = helpers.ref_random_mps(x, np.float64, L, mpo.qsite, sector, D, seed=42)
        en = np.zeros(3); ent = np.zeros(L - 1)
        rc = x.dmrg_twosite(mpo.ptr, 3, 20, 1e-10, D, psi.ptr, en.ctypes.data_as(C.POINTER(C.c_double)), ent.ctypes.data_as(C.POINTER(C.c_double)))
        assert rc == 0
        results[name] = (en, ent, psi.bond_dims(), x.dll.mps_norm(psi.ptr))
    en_d, ent_d, bd_d, nrm_d = results["dropin"]
    en_r, ent_r, bd_r, nrm_r = results["reference"]
    assert np.max(np.abs(en_d - en_r)) <= 1e-10
    assert bd_d == bd_r
    assert np.max(np.abs(ent_d - ent_r)) <= 1e-8
    assert abs(nrm_d - 1) <= 1e-12 and abs(nrm_r - 1) <= 1e-12


@pytest.mark.parametrize("kind", VARIANTS)
def test_su2_callers_reach_the_engine(kind):
    """su2_dmrg.c / su2_chain_ops.c dropped, the overlapping functions of su2_tensor.c, su2_mps.c and
    su2_recoupling.c renamed away; the reference's SU(2) generators feed the engine's su2_dmrg_twosite inside the combined library."""
    import su2_helpers as S
    lib = load_dropin(kind)
    d = lib.dll
    S.bind_ref(d)
    eng = C.CDLL(helpers.EMU_SO if kind == "emu" else helpers.CUDA_SO)
    for name in ("su2_dmrg_twosite", "su2_dmrg_singlesite", "su2_apply_local_hamiltonian", "su2_tensor_contract_simple", "su2_tensor_fmove", "su2_tensor_svd"):
        assert C.cast(getattr(d, name), C.c_void_p).value == C.cast(getattr(eng, name), C.c_void_p).value, name
    assert hasattr(d, "ref_su2_tensor_fmove") and hasattr(d, "construct_heisenberg_1d_su2_mpo")
    if kind == "cuda":
        assert lib.ctb_init(-1) == 0
    r = S.ref()
    L, ns = 7, 2
    res = {}
    for name, x in (("dropin", d), ("reference", r)):
        mpo = S.SU2MPO()
        x.construct_heisenberg_1d_su2_mpo(L, 1.1, C.byref(mpo))
        psi = S.random_mps(L, [1], [0, 1], 1, 5, 27, 82, scale=14.0)      # same seeded input for both
        en = (C.c_double * ns)()
        ent = (C.c_double * (L - 1))()
        assert x.su2_dmrg_twosite(C.byref(mpo), ns, 5, 1e-8, 200, C.byref(psi), en, ent) == 0
        assert x.su2_mps_is_consistent(C.byref(psi))
        res[name] = (list(en), list(ent))
    assert np.allclose(res["dropin"][0], res["reference"][0], rtol=0, atol=1e-10)
    assert np.allclose(res["dropin"][1], res["reference"][1], rtol=0, atol=1e-8)
