"""The block index of large sector grids (host/ctb_internal.h: ctb_grid_offset): grids beyond CTB_GRID_DENSE_MAX cells keep an
open-addressing hash over their stored blocks instead of a dense cell table, and the stored blocks are enumerated over all axes but the
one with the most sectors.  The test knob CTB_GRID_DENSE_MAX=0 sends EVERY tensor of a run down that path: contractions, re-blocking
and whole sweeps must give what the dense table gives (bit-identical structure, same energies) and what the reference gives."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers
from chemtensor_b200 import cabi


@pytest.fixture
def hashed_grids():
    old = os.environ.get("CTB_GRID_DENSE_MAX")
    os.environ["CTB_GRID_DENSE_MAX"] = "0"
    yield
    if old is None:
        os.environ.pop("CTB_GRID_DENSE_MAX", None)
    else:
        os.environ["CTB_GRID_DENSE_MAX"] = old


@pytest.mark.parametrize("ndim_mult", [1, 2])
def test_dot_hashed_grid(eng, ref, rng, hashed_grids, ndim_mult):
    cdim = [7, 5][:ndim_mult]
    sshape, tshape = (6, 4) + tuple(cdim), tuple(cdim) + (5, 3, 4)
    sdirs = [1, -1] + [1] * ndim_mult
    tdirs = [-1] * ndim_mult + [1, -1, 1]
    cq = [helpers.random_qnums(rng, d) for d in cdim]
    sq = [helpers.random_qnums(rng, 6), helpers.random_qnums(rng, 4)] + cq
    tq = cq + [helpers.random_qnums(rng, 5), helpers.random_qnums(rng, 3), helpers.random_qnums(rng, 4)]
    sd = helpers.random_dense(rng, np.float64, sshape, sdirs, sq)
    td = helpers.random_dense(rng, np.float64, tshape, tdirs, tq)
    out = []
    for lib in (eng, ref):
        s = cabi.bst_from_dense(lib, sd, sdirs, sq)
        t = cabi.bst_from_dense(lib, td, tdirs, tq)
        r = cabi.BST(lib)
        lib.block_sparse_tensor_dot(s.ptr, cabi.AXIS_RANGE_TRAILING, t.ptr, cabi.AXIS_RANGE_LEADING, ndim_mult, r.ptr)
        out.append(r)
    helpers.assert_bst_close(out[0], out[1], 1e-13)


def test_twosite_sweep_hashed_grid(eng, ref, hashed_grids):
    """Fermi-Hubbard L = 6 two-site sweeps with every tensor on the hash index: energies of the reference to 1e-10."""
    L, max_vdim = 6, 32
    mpo_r = helpers.ref_mpo(ref, "fermi_hubbard", L, 1.0, 4.0, 0.0)
    psi0 = helpers.ref_random_mps(ref, np.float64, L, mpo_r.qsite, helpers.encode_qpair(L, 0), max_vdim, seed=42)
    res = []
    for lib in (eng, ref):
        mpo, psi = helpers.clone_chain(lib, mpo_r), helpers.clone_chain(lib, psi0)
        en = np.zeros(2); ent = np.zeros(L - 1)
        assert lib.dmrg_twosite(mpo.ptr, 2, 20, 1e-10, max_vdim, psi.ptr, en.ctypes.data_as(C.POINTER(C.c_double)), ent.ctypes.data_as(C.POINTER(C.c_double))) == 0
        res.append((en, psi))
    assert np.max(np.abs(res[0][0] - res[1][0])) <= 1e-10, (res[0][0], res[1][0])
    assert res[0][1].bond_dims() == res[1][1].bond_dims()


def test_pair_form_sweep_hashed_grid(eng, ref, hashed_grids):
    """complex128 molecular sweep in pair form (the 6-leg intermediates are the tensors that use the hash index in production)."""
    n = 4
    rng = np.random.default_rng(5)
    tkin = 0.5 * (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    vint = 0.1 * (rng.standard_normal((n, n, n, n)) + 1j * rng.standard_normal((n, n, n, n)))
    tkin = 0.5 * (tkin + tkin.conj().T)
    vint = 0.5 * (vint + vint.transpose((1, 0, 3, 2)))
    vint = 0.5 * (vint + vint.transpose((2, 3, 0, 1)).conj())
    mpo_r = helpers.ref_molecular_mpo(ref, tkin, vint, spin=True, optimize=False)
    L = mpo_r.nsites
    psi0 = helpers.ref_random_mps(ref, np.complex128, L, mpo_r.qsite, helpers.encode_qpair(n, 0), 24, seed=42)
    res = []
    old = os.environ.get("CTB_HEFF_PAIR")
    os.environ["CTB_HEFF_PAIR"] = "1"
    try:
        for lib in (eng, ref):
            mpo, psi = helpers.clone_chain(lib, mpo_r), helpers.clone_chain(lib, psi0)
            en = np.zeros(2); ent = np.zeros(L - 1)
            assert lib.dmrg_twosite(mpo.ptr, 2, 20, 1e-10, 24, psi.ptr, en.ctypes.data_as(C.POINTER(C.c_double)), ent.ctypes.data_as(C.POINTER(C.c_double))) == 0
            res.append(en)
    finally:
        if old is None:
            os.environ.pop("CTB_HEFF_PAIR", None)
        else:
            os.environ["CTB_HEFF_PAIR"] = old
    assert np.max(np.abs(res[0] - res[1])) <= 1e-10, (res[0], res[1])
