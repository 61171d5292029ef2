"""BASELINE.json configs[0] and configs[1] as named, on the CUDA product (-m gpu).

* configs[0] = perf/perf_dmrg.c as shipped (9 orbitals, sector (9, 1), max_vdim 512, 2 sweeps x 25 Lanczos iterations,
  tol_split 1e-8, seed 42): both sweep energies equal the reference's -51.2777797066802 (identical on
  every thread count) to 1e-10, and the unmodified reference run next to it gives the same energies and bond dimensions.
* configs[1] in small (the full L=100, D=1024 case runs in bench.py; the reference would need minutes): XXZ chain L=24,
  max bond 128, tol_split = 0, two-site AND single-site DMRG against the reference, energies to 1e-10.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import helpers
from chemtensor_b200 import cabi, workloads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not helpers.have_gpu():
        pytest.fail("GPU test selected but no CUDA device is visible: the product has no CPU fallback")
    return helpers.load("cuda")


def test_perf_dmrg_as_shipped(cuda, ref):
    ours = bench.perf_dmrg_c1(cuda)
    theirs = bench.perf_dmrg_c1(ref)
    assert ours is not None and theirs is not None
    assert abs(theirs["energies"][-1] - (-51.2777797066802)) <= 1e-10
    assert np.max(np.abs(np.array(ours["energies"]) - (-51.2777797066802))) <= 1e-10, ours["energies"]
    assert np.max(np.abs(np.array(ours["energies"]) - np.array(theirs["energies"]))) <= 1e-10
    assert ours["bond_dims"] == theirs["bond_dims"]


@pytest.mark.parametrize("single_site", [False, True])
def test_xxz_chain_sweeps(cuda, ref, single_site):
    L, D, sweeps, lanczos = 24, 128, 2, 10
    mpo_r = helpers.ref_mpo(ref, "xxz", L, 1.0, 0.8, 0.1)
    psi0 = helpers.ref_random_mps(ref, np.float64, L, mpo_r.qsite, 0, D, seed=42)
    res = []
    for lib in (cuda, ref):
        mpo, psi = helpers.clone_chain(lib, mpo_r), helpers.clone_chain(lib, psi0)
        en = np.zeros(sweeps); ent = np.zeros(L - 1)
        if single_site:
            rc = lib.dmrg_singlesite(mpo.ptr, sweeps, lanczos, psi.ptr, en.ctypes.data_as(C.POINTER(C.c_double)))
        else:
            rc = lib.dmrg_twosite(mpo.ptr, sweeps, lanczos, 0.0, D, psi.ptr, en.ctypes.data_as(C.POINTER(C.c_double)), ent.ctypes.data_as(C.POINTER(C.c_double)))
        assert rc == 0
        res.append((en, psi))
    assert np.max(np.abs(res[0][0] - res[1][0])) <= 1e-10, (res[0][0], res[1][0])
    assert res[0][1].bond_dims() == res[1][1].bond_dims()
    assert max(res[1][1].bond_dims()) == D
