"""The other callers of the hot-path primitives against the compiled reference:
mps_vdot / mps_norm, mpo_inner_product, apply_mpo, compute_local_hamiltonian_environment,
split_block_sparse_matrix_svd_isometry, mps_local_orthonormalize_left/right_svd, mps_compress(_rescale).

Pure contractions are compared entry-wise (structure bit-exact, values to 1e-12); results with an SVD gauge freedom are
compared through gauge-free quantities (projectors, merged pairs, overlaps), as the reference's own tests do
(test/state/test_mps.c, test/algorithm/test_bond_ops.c).
"""
import ctypes as C

import numpy as np
import pytest

import helpers
from chemtensor_b200 import cabi

DTYPES = [np.float64, np.complex128]
FH = ("fermi_hubbard", 6, (1.0, 4.0, 0.3), helpers.encode_qpair(6, 0), 24)
XXZ = ("xxz", 7, (1.0, 0.8, 0.1), 1, 20)


def _cast(lib, src, dtype):
    tensors = []
    for i in range(src.nsites):
        s = src.site(i)
        t = cabi.bst_allocate(lib, dtype, s.shape, s.axis_dir, s.qnums)
        for (_, a), (_, b) in zip(t.blocks(), s.blocks()):
            a[...] = b
        tensors.append(t)
    return cabi.Chain(lib, src.kind, src.qsite, tensors)


def _inputs(ref, case, dtype, seeds=(11, 12)):
    model, L, params, sector, max_vdim = case
    mpo = helpers.ref_mpo(ref, model, L, *params)
    psi = helpers.ref_random_mps(ref, dtype, L, mpo.qsite, sector, max_vdim, seed=seeds[0])
    chi = helpers.ref_random_mps(ref, dtype, L, mpo.qsite, sector, max_vdim, seed=seeds[1])
    return mpo, psi, chi


def _scalar(dtype):
    return np.zeros(1, dtype=dtype)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", [FH, XXZ], ids=["fh", "xxz"])
def test_mps_vdot_and_norm(eng, ref, case, dtype):
    _, psi_r, chi_r = _inputs(ref, case, dtype)
    psi_e, chi_e = helpers.clone_chain(eng, psi_r), helpers.clone_chain(eng, chi_r)
    v_e, v_r = _scalar(dtype), _scalar(dtype)
    eng.mps_vdot(chi_e.ptr, psi_e.ptr, v_e.ctypes.data)
    ref.mps_vdot(chi_r.ptr, psi_r.ptr, v_r.ctypes.data)
    n_e, n_r = eng.mps_norm(psi_e.ptr), ref.mps_norm(psi_r.ptr)
    assert abs(n_e - n_r) <= 1e-12 * max(1.0, n_r)
    assert abs(v_e[0] - v_r[0]) <= 1e-12 * max(1.0, n_r * ref.mps_norm(chi_r.ptr))
    # <psi|psi> through the same path (chi aliases psi)
    eng.mps_vdot(psi_e.ptr, psi_e.ptr, v_e.ctypes.data)
    assert abs(v_e[0] - n_r ** 2) <= 1e-12 * max(1.0, n_r ** 2)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", [FH, XXZ], ids=["fh", "xxz"])
def test_mpo_inner_product(eng, ref, case, dtype):
    mpo_r, psi_r, chi_r = _inputs(ref, case, dtype)
    mpo_rc = _cast(ref, mpo_r, dtype)
    mpo_e, psi_e, chi_e = _cast(eng, mpo_r, dtype), helpers.clone_chain(eng, psi_r), helpers.clone_chain(eng, chi_r)
    v_e, v_r = _scalar(dtype), _scalar(dtype)
    eng.mpo_inner_product(chi_e.ptr, mpo_e.ptr, psi_e.ptr, v_e.ctypes.data)
    ref.mpo_inner_product(chi_r.ptr, mpo_rc.ptr, psi_r.ptr, v_r.ctypes.data)
    scale = max(1.0, abs(v_r[0]))
    assert abs(v_e[0] - v_r[0]) <= 1e-11 * scale
    # expectation value <psi|H|psi> is real for the Hermitian Hamiltonian
    eng.mpo_inner_product(psi_e.ptr, mpo_e.ptr, psi_e.ptr, v_e.ctypes.data)
    ref.mpo_inner_product(psi_r.ptr, mpo_rc.ptr, psi_r.ptr, v_r.ctypes.data)
    assert abs(v_e[0] - v_r[0]) <= 1e-11 * max(1.0, abs(v_r[0]))
    assert abs(np.imag(v_e[0])) <= 1e-11 * max(1.0, abs(v_r[0]))


def test_mpo_inner_product_golden(eng):
    """the reference's own fixture test/algorithm/data/test_mpo_inner_product.hdf5 (test_chain_ops.c: single complex inputs,
    relative tolerance 1e-6); the engine computes in complex128 on the same entries"""
    import os
    from test_golden_engine import chain_from_dense, golden
    ds, at = golden("mpo_inner_product")
    L = 5
    qsite = np.asarray(at["qsite"], dtype=np.int32)
    qb = {k: [np.asarray(at[f"qbond_{k}_{i}"], dtype=np.int32) for i in range(L + 1)] for k in ("psi", "chi", "op")}
    psi = chain_from_dense(eng, "mps", [ds[f"psi_a{i}"].astype(np.complex128) for i in range(L)], qsite, qb["psi"])
    chi = chain_from_dense(eng, "mps", [ds[f"chi_a{i}"].astype(np.complex128) for i in range(L)], qsite, qb["chi"])
    op = chain_from_dense(eng, "mpo", [ds[f"op_a{i}"].astype(np.complex128) for i in range(L)], qsite, qb["op"])
    v = _scalar(np.complex128)
    eng.mpo_inner_product(chi.ptr, op.ptr, psi.ptr, v.ctypes.data)
    s_ref = complex(np.asarray(ds["s"]).reshape(-1)[0])
    assert abs(v[0] - s_ref) / abs(s_ref) <= 1e-6


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("case", [FH, XXZ], ids=["fh", "xxz"])
def test_apply_mpo(eng, ref, case, dtype):
    mpo_r, psi_r, _ = _inputs(ref, case, dtype)
    mpo_rc = _cast(ref, mpo_r, dtype)
    mpo_e, psi_e = _cast(eng, mpo_r, dtype), helpers.clone_chain(eng, psi_r)
    out_e, out_r = helpers.RefChain(eng, "mps"), helpers.RefChain(ref, "mps")
    eng.apply_mpo(mpo_e.ptr, psi_e.ptr, out_e.ptr); out_e.alive = True
    ref.apply_mpo(mpo_rc.ptr, psi_r.ptr, out_r.ptr); out_r.alive = True
    assert out_e.nsites == out_r.nsites and np.array_equal(out_e.qsite, out_r.qsite)
    assert out_e.bond_dims() == out_r.bond_dims()
    for i in range(out_r.nsites):
        helpers.assert_bst_close(out_e.site(i), out_r.site(i), 1e-12)     # structure bit-exact, entries to 1e-12


@pytest.mark.parametrize("dtype", DTYPES)
def test_local_hamiltonian_environment(eng, ref, dtype):
    model, L, params, sector, max_vdim = FH
    mpo_r, psi_r, chi_r = _inputs(ref, FH, dtype)
    mpo_rc = _cast(ref, mpo_r, dtype)
    i = L // 2
    # environments of site i from the reference, then the MPO-tensor environment from both libraries
    rl = (cabi.BlockSparseTensor * L)()
    ref.compute_right_operator_blocks(psi_r.ptr, chi_r.ptr, mpo_rc.ptr, rl)
    l_r = cabi.BST(ref)
    ref.create_dummy_operator_block_left(psi_r.site(0).ptr, chi_r.site(0).ptr, mpo_rc.site(0).ptr, l_r.ptr)
    for j in range(i):
        nxt = cabi.BST(ref)
        ref.contraction_operator_step_left(psi_r.site(j).ptr, chi_r.site(j).ptr, mpo_rc.site(j).ptr, l_r.ptr, nxt.ptr)
        l_r = nxt
    r_r = cabi.BST(ref, rl[i], owned=False)
    dw_r = cabi.BST(ref)
    ref.compute_local_hamiltonian_environment(psi_r.site(i).ptr, chi_r.site(i).ptr, l_r.ptr, r_r.ptr, dw_r.ptr)
    a_e, b_e = cabi.bst_clone(eng, psi_r.site(i)), cabi.bst_clone(eng, chi_r.site(i))
    l_e, r_e = cabi.bst_clone(eng, l_r), cabi.bst_clone(eng, r_r)
    dw_e = cabi.BST(eng)
    eng.compute_local_hamiltonian_environment(a_e.ptr, b_e.ptr, l_e.ptr, r_e.ptr, dw_e.ptr)
    helpers.assert_bst_close(dw_e, dw_r, 1e-12)
    # consistency: <chi|op|psi> = sum_entries dw * w  at site i
    v_r = _scalar(dtype)
    ref.mpo_inner_product(chi_r.ptr, mpo_rc.ptr, psi_r.ptr, v_r.ctypes.data)
    w = mpo_rc.site(i)
    acc = np.sum(dw_e.serialize() * w.serialize())
    assert abs(acc - v_r[0]) <= 1e-10 * max(1.0, abs(v_r[0]))
    for k in range(L):
        ref.delete_block_sparse_tensor(C.byref(rl[k]))


def _dense_matrix(t: cabi.BST) -> np.ndarray:
    return t.to_dense()


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("tol,max_vdim", [(0.0, 1000), (0.05, 1000), (0.0, 7)])
def test_svd_isometry(eng, ref, rng, dtype, tol, max_vdim):
    qrow, qcol = helpers.random_qnums(rng, 23), helpers.random_qnums(rng, 31)
    a_e = helpers.random_bst(eng, rng, dtype, (23, 31), (1, -1), (qrow, qcol))
    a_r = cabi.bst_clone(ref, a_e)
    u_e, u_r = cabi.BST(eng), cabi.BST(ref)
    i_e, i_r = cabi.TruncInfo(), cabi.TruncInfo()
    assert eng.split_block_sparse_matrix_svd_isometry(a_e.ptr, tol, True, max_vdim, u_e.ptr, C.byref(i_e)) == 0
    assert ref.split_block_sparse_matrix_svd_isometry(a_r.ptr, tol, True, max_vdim, u_r.ptr, C.byref(i_r)) == 0
    helpers.assert_same_structure(u_e, u_r)
    assert abs(i_e.norm_sigma - i_r.norm_sigma) <= 1e-12 * max(1.0, i_r.norm_sigma)
    assert abs(i_e.entropy - i_r.entropy) <= 1e-10
    ue, ur = _dense_matrix(u_e), _dense_matrix(u_r)
    # isometry, and the same projector (singular vectors are fixed up to a phase / rotation inside degenerate values)
    assert np.allclose(ue.conj().T @ ue, np.eye(ue.shape[1]), atol=1e-12)
    assert np.allclose(ue @ ue.conj().T, ur @ ur.conj().T, atol=1e-9)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("side", ["left", "right"])
@pytest.mark.parametrize("tol,max_vdim", [(0.0, 1000), (1e-3, 9)])
def test_mps_local_orthonormalize_svd(eng, ref, dtype, side, tol, max_vdim):
    _, psi_r, _ = _inputs(ref, FH, dtype)
    psi_e = helpers.clone_chain(eng, psi_r)
    i = 2
    j = i + 1 if side == "left" else i - 1
    lo, hi = min(i, j), max(i, j)
    out = []
    for lib, psi in ((eng, psi_e), (ref, psi_r)):
        a, nb = cabi.bst_clone(lib, psi.site(i)), cabi.bst_clone(lib, psi.site(j))
        pair0 = cabi.BST(lib)
        first, second = (a, nb) if side == "left" else (nb, a)
        lib.mps_merge_tensor_pair(first.ptr, second.ptr, pair0.ptr)
        info = cabi.TruncInfo()
        fn = lib.mps_local_orthonormalize_left_svd if side == "left" else lib.mps_local_orthonormalize_right_svd
        assert fn(tol, max_vdim, False, a.ptr, nb.ptr, C.byref(info)) == 0
        pair1 = cabi.BST(lib)
        first, second = (a, nb) if side == "left" else (nb, a)
        lib.mps_merge_tensor_pair(first.ptr, second.ptr, pair1.ptr)
        out.append((a, nb, pair0, pair1, info))
    (a_e, n_e, p0_e, p1_e, i_e), (a_r, n_r, p0_r, p1_r, i_r) = out
    helpers.assert_same_structure(a_e, a_r)          # new bond: quantum numbers and dimension bit-exact
    helpers.assert_same_structure(n_e, n_r)
    assert abs(i_e.norm_sigma - i_r.norm_sigma) <= 1e-12 * max(1.0, i_r.norm_sigma)
    assert abs(i_e.tol_eff - i_r.tol_eff) <= 1e-12
    helpers.assert_bst_close(p1_e, p1_r, 1e-9)       # the merged pair is gauge-free
    if tol == 0.0 and max_vdim >= 1000:
        helpers.assert_bst_close(p1_e, p0_e, 1e-12)  # nothing truncated: the pair is unchanged
    # isometry of the orthonormalised site
    ad = a_e.to_dense()
    m = ad.reshape(-1, ad.shape[2]) if side == "left" else ad.reshape(ad.shape[0], -1).conj().T
    assert np.allclose(m.conj().T @ m, np.eye(m.shape[1]), atol=1e-12)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mode", [cabi.MPS_ORTHONORMAL_LEFT, cabi.MPS_ORTHONORMAL_RIGHT])
@pytest.mark.parametrize("tol,max_vdim", [(0.0, 1000), (1e-2, 1000), (0.0, 6)])
def test_mps_compress(eng, ref, dtype, mode, tol, max_vdim):
    model, L, params, sector, _ = FH
    _, psi_r, _ = _inputs(ref, FH, dtype)
    psi_e = helpers.clone_chain(eng, psi_r)
    psi0 = helpers.clone_chain(ref, psi_r)
    res = []
    for lib, psi in ((eng, psi_e), (ref, psi_r)):
        norm, scale = C.c_double(0), C.c_double(0)
        info = (cabi.TruncInfo * L)()
        assert lib.mps_compress(tol, max_vdim, mode, psi.ptr, C.byref(norm), C.byref(scale), info) == 0
        res.append((norm.value, scale.value, info))
    (n_e, s_e, i_e), (n_r, s_r, i_r) = res
    assert abs(n_e - n_r) <= 1e-12 * max(1.0, n_r)
    assert abs(s_e - s_r) <= 1e-10
    assert psi_e.bond_dims() == psi_r.bond_dims()
    for i in range(L):
        for qa, qb in zip(psi_e.site(i).qnums, psi_r.site(i).qnums):
            assert np.array_equal(qa, qb)
        assert abs(i_e[i].norm_sigma - i_r[i].norm_sigma) <= 1e-10
        assert abs(i_e[i].entropy - i_r[i].entropy) <= 1e-8
    # same (normalised) state as the reference's result, and -- without truncation -- as the input up to its norm
    psi_e_in_ref = helpers.clone_chain(ref, psi_e)
    ov = _scalar(dtype)
    ref.mps_vdot(psi_e_in_ref.ptr, psi_r.ptr, ov.ctypes.data)
    assert abs(ov[0] - 1.0) <= 1e-9
    assert abs(ref.mps_norm(psi_e_in_ref.ptr) - 1.0) <= 1e-12
    if tol == 0.0 and max_vdim >= 1000:
        ref.mps_vdot(psi0.ptr, psi_e_in_ref.ptr, ov.ctypes.data)
        assert abs(ov[0] - n_r) <= 1e-10 * max(1.0, n_r)
        assert abs(s_e - 1.0) <= 1e-12


@pytest.mark.parametrize("dtype", DTYPES)
def test_mps_compress_rescale(eng, ref, dtype):
    model, L, params, sector, _ = FH
    _, psi_r, _ = _inputs(ref, FH, dtype)
    psi_e = helpers.clone_chain(eng, psi_r)
    for lib, psi in ((eng, psi_e), (ref, psi_r)):
        scale = C.c_double(0)
        info = (cabi.TruncInfo * L)()
        assert lib.mps_compress_rescale(1e-3, 12, cabi.MPS_ORTHONORMAL_LEFT, psi.ptr, C.byref(scale), info) == 0
    psi_e_in_ref = helpers.clone_chain(ref, psi_e)
    assert abs(ref.mps_norm(psi_e_in_ref.ptr) - ref.mps_norm(psi_r.ptr)) <= 1e-11 * ref.mps_norm(psi_r.ptr)
    ov = _scalar(dtype)
    ref.mps_vdot(psi_e_in_ref.ptr, psi_r.ptr, ov.ctypes.data)
    assert abs(ov[0] - ref.mps_norm(psi_r.ptr) ** 2) <= 1e-9 * ref.mps_norm(psi_r.ptr) ** 2
