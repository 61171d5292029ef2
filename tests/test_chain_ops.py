"""Effective-Hamiltonian application and environment updates against the compiled reference.

The reference has no direct unit test of apply_local_hamiltonian / contraction_operator_step_* in its U(1) suite
(SURVEY.md §4); here they are checked on reference-generated inputs: Hamiltonian MPOs from the reference's own
constructors (src/operator/hamiltonian.c:102, :240), a seeded random MPS (src/state/mps.c:93) and the reference's own
environments.  Tolerance: 1e-12 relative Frobenius norm per matvec (BASELINE.json north_star), structure bit-exact.
"""
import ctypes as C

import numpy as np
import pytest

import helpers
from chemtensor_b200 import cabi

CASES = [
    # model, L, params, sector, max_vdim, dtype
    ("xxz", 8, (1.0, 0.8, 0.1), 0, 24, np.float64),
    ("fermi_hubbard", 6, (1.0, 4.0, 0.3), helpers.encode_qpair(6, 0), 40, np.float64),
    ("xxz", 7, (1.0, 0.8, 0.1), 1, 17, np.complex128),
]


def _setup(ref, model, L, params, sector, max_vdim, dtype):
    mpo = helpers.ref_mpo(ref, model, L, *params)
    psi = helpers.ref_random_mps(ref, dtype, L, mpo.qsite, sector, max_vdim, seed=42)
    if np.dtype(dtype).kind == "c":
        # the reference Hamiltonians are real; give the complex case a complex MPO by cloning with a phase-free cast
        pass
    return mpo, psi


def _cast_chain(lib, src, dtype):
    """Clone a chain into `lib`, converting entries to `dtype`."""
    tensors = []
    for i in range(src.nsites):
        s = src.site(i)
        t = cabi.bst_allocate(lib, dtype, s.shape, s.axis_dir, s.qnums)
        for (_, a), (_, b) in zip(t.blocks(), s.blocks()):
            a[...] = b
        tensors.append(t)
    return cabi.Chain(lib, src.kind, src.qsite, tensors)


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"{c[0]}-L{c[1]}-{np.dtype(c[5]).name}")
def test_environments_and_heff(eng, ref, case):
    model, L, params, sector, max_vdim, dtype = case
    mpo_r, psi_r = _setup(ref, model, L, params, sector, max_vdim, dtype)
    mpo_ref = _cast_chain(ref, mpo_r, dtype)
    psi_ref = _cast_chain(ref, psi_r, dtype)
    mpo_eng = _cast_chain(eng, mpo_r, dtype)
    psi_eng = _cast_chain(eng, psi_r, dtype)

    # right environments, all at once
    rl_ref = (cabi.BlockSparseTensor * L)()
    rl_eng = (cabi.BlockSparseTensor * L)()
    ref.compute_right_operator_blocks(psi_ref.ptr, psi_ref.ptr, mpo_ref.ptr, rl_ref)
    eng.compute_right_operator_blocks(psi_eng.ptr, psi_eng.ptr, mpo_eng.ptr, rl_eng)
    r_ref = [cabi.BST(ref, rl_ref[i]) for i in range(L)]
    r_eng = [cabi.BST(eng, rl_eng[i]) for i in range(L)]
    for i in range(L):
        helpers.assert_bst_close(r_eng[i], r_ref[i], 1e-12)

    # left environments step by step
    l_ref, l_eng = [cabi.BST(ref)], [cabi.BST(eng)]
    ref.create_dummy_operator_block_left(psi_ref.site(0).ptr, psi_ref.site(0).ptr, mpo_ref.site(0).ptr, l_ref[0].ptr)
    eng.create_dummy_operator_block_left(psi_eng.site(0).ptr, psi_eng.site(0).ptr, mpo_eng.site(0).ptr, l_eng[0].ptr)
    helpers.assert_bst_close(l_eng[0], l_ref[0], 0.0)
    for i in range(L - 1):
        nr, ne = cabi.BST(ref), cabi.BST(eng)
        ref.contraction_operator_step_left(psi_ref.site(i).ptr, psi_ref.site(i).ptr, mpo_ref.site(i).ptr, l_ref[i].ptr, nr.ptr)
        eng.contraction_operator_step_left(psi_eng.site(i).ptr, psi_eng.site(i).ptr, mpo_eng.site(i).ptr, l_eng[i].ptr, ne.ptr)
        helpers.assert_bst_close(ne, nr, 1e-12)
        l_ref.append(nr); l_eng.append(ne)

    # single-site Heff at every site and two-site Heff on every pair (merged MPS / MPO tensors)
    for i in range(L):
        br, be = cabi.BST(ref), cabi.BST(eng)
        ref.apply_local_hamiltonian(psi_ref.site(i).ptr, mpo_ref.site(i).ptr, l_ref[i].ptr, r_ref[i].ptr, br.ptr)
        eng.apply_local_hamiltonian(psi_eng.site(i).ptr, mpo_eng.site(i).ptr, l_eng[i].ptr, r_eng[i].ptr, be.ptr)
        helpers.assert_bst_close(be, br, 1e-12)
    for i in range(L - 1):
        a2r, a2e, w2r, w2e = cabi.BST(ref), cabi.BST(eng), cabi.BST(ref), cabi.BST(eng)
        ref.mps_merge_tensor_pair(psi_ref.site(i).ptr, psi_ref.site(i + 1).ptr, a2r.ptr)
        eng.mps_merge_tensor_pair(psi_eng.site(i).ptr, psi_eng.site(i + 1).ptr, a2e.ptr)
        helpers.assert_bst_close(a2e, a2r, 1e-13)
        ref.mpo_merge_tensor_pair(mpo_ref.site(i).ptr, mpo_ref.site(i + 1).ptr, w2r.ptr)
        eng.mpo_merge_tensor_pair(mpo_eng.site(i).ptr, mpo_eng.site(i + 1).ptr, w2e.ptr)
        helpers.assert_bst_close(w2e, w2r, 1e-13)
        br, be = cabi.BST(ref), cabi.BST(eng)
        ref.apply_local_hamiltonian(a2r.ptr, w2r.ptr, l_ref[i].ptr, r_ref[i + 1].ptr, br.ptr)
        eng.apply_local_hamiltonian(a2e.ptr, w2e.ptr, l_eng[i].ptr, r_eng[i + 1].ptr, be.ptr)
        helpers.assert_bst_close(be, br, 1e-12)


def test_environment_step_with_distinct_bra(eng, ref):
    """<chi| op |psi> with chi != psi (complex): exercises the fused conjugation of the bra tensor."""
    L, dtype = 5, np.complex128
    mpo_r = helpers.ref_mpo(ref, "xxz", L, 1.0, 0.8, 0.1)
    psi_r = helpers.ref_random_mps(ref, dtype, L, mpo_r.qsite, 1, 13, seed=7)
    chi_r = helpers.ref_random_mps(ref, dtype, L, mpo_r.qsite, 1, 13, seed=8)
    mpo_ref, mpo_eng = _cast_chain(ref, mpo_r, dtype), _cast_chain(eng, mpo_r, dtype)
    psi_eng, chi_eng = helpers.clone_chain(eng, psi_r), helpers.clone_chain(eng, chi_r)
    rl_ref = (cabi.BlockSparseTensor * L)()
    rl_eng = (cabi.BlockSparseTensor * L)()
    ref.compute_right_operator_blocks(psi_r.ptr, chi_r.ptr, mpo_ref.ptr, rl_ref)
    eng.compute_right_operator_blocks(psi_eng.ptr, chi_eng.ptr, mpo_eng.ptr, rl_eng)
    for i in range(L):
        helpers.assert_bst_close(cabi.BST(eng, rl_eng[i]), cabi.BST(ref, rl_ref[i]), 1e-12)
