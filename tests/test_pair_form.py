"""The pair form of the two-site effective Hamiltonian: the two single-site MPO tensors are
applied one after the other, the merged pair tensor of the reference (mpo_merge_tensor_pair, src/operator/mpo.c:255; used
by dmrg_twosite, src/algorithm/dmrg.c:289-292) is never built.  Results must equal the reference's merged path."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers
from chemtensor_b200 import cabi


def _cast(lib, src, dtype):
    tensors = []
    for i in range(src.nsites):
        s = src.site(i)
        t = cabi.bst_allocate(lib, dtype, s.shape, s.axis_dir, s.qnums)
        for (_, a), (_, b) in zip(t.blocks(), s.blocks()):
            a[...] = b
        tensors.append(t)
    return cabi.Chain(lib, src.kind, src.qsite, tensors)


def _molecular(ref, n, seed, dtype):
    rng = np.random.default_rng(seed)
    if np.dtype(dtype).kind == "c":
        tkin = 0.5 * (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
        vint = 0.1 * (rng.standard_normal((n, n, n, n)) + 1j * rng.standard_normal((n, n, n, n)))
    else:
        tkin = 0.5 * rng.standard_normal((n, n))
        vint = 0.1 * rng.standard_normal((n, n, n, n))
    tkin = 0.5 * (tkin + tkin.conj().T)
    vint = 0.5 * (vint + vint.transpose((1, 0, 3, 2)))
    vint = 0.5 * (vint + vint.transpose((2, 3, 0, 1)).conj())
    return helpers.ref_molecular_mpo(ref, tkin, vint, spin=True, optimize=False)


CASES = [
    ("fermi_hubbard", np.float64), ("xxz", np.complex128), ("molecular", np.float64), ("molecular", np.complex128),
]


def _setup(ref, model, dtype):
    if model == "molecular":
        mpo_r = _molecular(ref, 5, 3, dtype)
        L, sector, max_vdim = mpo_r.nsites, helpers.encode_qpair(5, 1), 30
    elif model == "fermi_hubbard":
        L, sector, max_vdim = 6, helpers.encode_qpair(6, 0), 30
        mpo_r = helpers.ref_mpo(ref, model, L, 1.0, 4.0, 0.3)
    else:
        L, sector, max_vdim = 7, 1, 20
        mpo_r = helpers.ref_mpo(ref, model, L, 1.0, 0.8, 0.1)
    mpo_rc = _cast(ref, mpo_r, dtype) if mpo_r.site(0).dtype != np.dtype(dtype) else mpo_r
    psi_r = helpers.ref_random_mps(ref, dtype, L, mpo_r.qsite, sector, max_vdim, seed=21)
    return mpo_rc, psi_r, L, max_vdim


@pytest.mark.parametrize("model,dtype", CASES, ids=[f"{m}-{np.dtype(d).name}" for m, d in CASES])
def test_pair_heff_equals_reference_merged_path(eng, ref, model, dtype):
    mpo_r, psi_r, L, _ = _setup(ref, model, dtype)
    i = L // 2 - 1
    rl = (cabi.BlockSparseTensor * L)()
    ref.compute_right_operator_blocks(psi_r.ptr, psi_r.ptr, mpo_r.ptr, rl)
    l_r = cabi.BST(ref)
    ref.create_dummy_operator_block_left(psi_r.site(0).ptr, psi_r.site(0).ptr, mpo_r.site(0).ptr, l_r.ptr)
    for j in range(i):
        nxt = cabi.BST(ref)
        ref.contraction_operator_step_left(psi_r.site(j).ptr, psi_r.site(j).ptr, mpo_r.site(j).ptr, l_r.ptr, nxt.ptr)
        l_r = nxt
    r_r = cabi.BST(ref, rl[i + 1], owned=False)
    a_r, w_r, b_r = cabi.BST(ref), cabi.BST(ref), cabi.BST(ref)
    ref.mps_merge_tensor_pair(psi_r.site(i).ptr, psi_r.site(i + 1).ptr, a_r.ptr)
    ref.mpo_merge_tensor_pair(mpo_r.site(i).ptr, mpo_r.site(i + 1).ptr, w_r.ptr)
    ref.apply_local_hamiltonian(a_r.ptr, w_r.ptr, l_r.ptr, r_r.ptr, b_r.ptr)
    a_e, l_e, r_e = cabi.bst_clone(eng, a_r), cabi.bst_clone(eng, l_r), cabi.bst_clone(eng, r_r)
    w0_e, w1_e = cabi.bst_clone(eng, mpo_r.site(i)), cabi.bst_clone(eng, mpo_r.site(i + 1))
    b_e = cabi.BST(eng)
    assert eng.ctb_apply_local_hamiltonian_pair(a_e.ptr, w0_e.ptr, w1_e.ptr, l_e.ptr, r_e.ptr, b_e.ptr) == 0
    helpers.assert_bst_close(b_e, b_r, 1e-12)
    for k in range(L):
        ref.delete_block_sparse_tensor(C.byref(rl[k]))


@pytest.mark.parametrize("model,dtype", CASES, ids=[f"{m}-{np.dtype(d).name}" for m, d in CASES])
def test_dmrg_twosite_in_pair_form(eng, ref, model, dtype, monkeypatch):
    mpo_r, psi0, L, max_vdim = _setup(ref, model, dtype)
    num_sweeps, maxiter, tol = 3, 25, 1e-10
    res = []
    for lib in (eng, ref):
        monkeypatch.setenv("CTB_HEFF_PAIR", "1")      # read by the engine only
        mpo, psi = helpers.clone_chain(lib, mpo_r), helpers.clone_chain(lib, psi0)
        en = np.zeros(num_sweeps); ent = np.zeros(L - 1)
        assert lib.dmrg_twosite(mpo.ptr, num_sweeps, maxiter, tol, max_vdim, psi.ptr, en.ctypes.data_as(C.POINTER(C.c_double)), ent.ctypes.data_as(C.POINTER(C.c_double))) == 0
        res.append((en, ent, psi))
    (en_e, ent_e, psi_e), (en_r, ent_r, psi_r) = res
    assert np.max(np.abs(en_e - en_r)) <= 1e-10, (en_e, en_r)
    assert np.max(np.abs(ent_e - ent_r)) <= 1e-7
    assert psi_e.bond_dims() == psi_r.bond_dims()
    for i in range(L):
        for qa, qb in zip(psi_e.site(i).qnums, psi_r.site(i).qnums):
            assert np.array_equal(qa, qb)
    psi_e_in_ref = helpers.clone_chain(ref, psi_e)
    ov = np.zeros(1, dtype=dtype)
    ref.mps_vdot(psi_e_in_ref.ptr, psi_r.ptr, ov.ctypes.data)
    assert abs(abs(ov[0]) - 1.0) <= 1e-8
