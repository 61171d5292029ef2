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
