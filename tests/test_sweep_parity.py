"""Two-site DMRG sweeps at a bond dimension where the sector blocks of the split are hundreds wide (block-Jacobi SVD path,
csrc/ctbd_svd_bj.cu) against the compiled reference on identical seeded inputs: energies within 1e-10 (north_star), bond
dimensions and bond quantum numbers bit-exact.  Fermi-Hubbard chain L=14, U(1)xU(1) sector (N=14, 2Sz=0), max bond 512,
tol_split = 0 (bonds saturate), 2 sweeps x 10 Lanczos iterations; the reference needs about a minute on the box's host cores.
Also the batched SVD itself on blocks beyond shared memory, graded spectra over 14 decades, real and complex, against LAPACK
(through the reference's block_sparse_tensor_svd)."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers
from chemtensor_b200 import cabi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    if not helpers.have_gpu():
        pytest.fail("GPU test selected but no CUDA device is visible: the product has no CPU fallback")
    return helpers.load("cuda")


def test_twosite_sweep_energies_D512(cuda, ref):
    L, D, sweeps, lanczos = 14, 512, 2, 10
    os.environ.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
    mpo_r = helpers.ref_mpo(ref, "fermi_hubbard", L, 1.0, 4.0, 0.0)
    psi0 = helpers.ref_random_mps(ref, np.float64, L, mpo_r.qsite, helpers.encode_qpair(L, 0), D, seed=42)
    res = []
    for lib in (cuda, ref):
        mpo, psi = helpers.clone_chain(lib, mpo_r), helpers.clone_chain(lib, psi0)
        en = np.zeros(sweeps); ent = np.zeros(L - 1)
        rc = lib.dmrg_twosite(mpo.ptr, sweeps, lanczos, 0.0, D, psi.ptr, en.ctypes.data_as(C.POINTER(C.c_double)), ent.ctypes.data_as(C.POINTER(C.c_double)))
        assert rc == 0
        res.append((en, ent, psi))
    (en_e, ent_e, psi_e), (en_r, ent_r, psi_r) = res
    assert max(psi_r.bond_dims()) == D
    assert np.max(np.abs(en_e - en_r)) <= 1e-10, (en_e, en_r)
    assert np.max(np.abs(ent_e - ent_r)) <= 1e-7
    assert psi_e.bond_dims() == psi_r.bond_dims()
    for i in range(L):
        for qa, qb in zip(psi_e.site(i).qnums, psi_r.site(i).qnums):
            assert np.array_equal(qa, qb)


def _graded(rng, dtype, m, n, decades):
    k = min(m, n)
    cplx = (dtype == np.complex128)
    u, _ = np.linalg.qr(rng.standard_normal((m, k)) + (1j * rng.standard_normal((m, k)) if cplx else 0))
    v, _ = np.linalg.qr(rng.standard_normal((n, k)) + (1j * rng.standard_normal((n, k)) if cplx else 0))
    s = 10.0 ** (-decades * np.arange(k) / (k - 1))
    return ((u * s) @ v.conj().T).astype(dtype)


@pytest.mark.parametrize("dtype,shapes", [(np.float64, [(700, 520), (333, 470), (129, 129)]), (np.complex128, [(400, 300), (150, 410)])])
def test_svd_big_blocks_graded(cuda, ref, dtype, shapes):
    """Every block takes the QR-preconditioned block-Jacobi path; singular values against LAPACK to 1e-13 of the largest,
    reconstruction and isometry as the reference's own property test (test_block_sparse_tensor.c:1462-1475), 1e-13 per entry."""
    rng = np.random.default_rng(11)
    M, N = sum(s[0] for s in shapes), sum(s[1] for s in shapes)
    dense = np.zeros((M, N), dtype=dtype)
    qr_ = np.concatenate([np.full(m, b, dtype=np.int32) for b, (m, _) in enumerate(shapes)])
    qc_ = np.concatenate([np.full(n, b, dtype=np.int32) for b, (_, n) in enumerate(shapes)])
    r0 = c0 = 0
    for (m, n) in shapes:
        dense[r0:r0 + m, c0:c0 + n] = _graded(rng, dtype, m, n, 14)
        r0 += m; c0 += n
    a, b = cabi.bst_from_dense(cuda, dense, [1, -1], [qr_, qc_]), cabi.bst_from_dense(ref, dense, [1, -1], [qr_, qc_])
    u, vh, ur, vr = cabi.BST(cuda), cabi.BST(cuda), cabi.BST(ref), cabi.BST(ref)
    s, sr = cabi.DenseTensor(), cabi.DenseTensor()
    assert cuda.block_sparse_tensor_svd(a.ptr, u.ptr, C.byref(s), vh.ptr) == 0
    assert ref.block_sparse_tensor_svd(b.ptr, ur.ptr, C.byref(sr), vr.ptr) == 0
    helpers.assert_same_structure(u, ur)
    helpers.assert_same_structure(vh, vr)
    ns = int(s.dim[0])
    assert ns == int(sr.dim[0])
    sv = np.ctypeslib.as_array(C.cast(s.data, C.POINTER(C.c_double)), shape=(ns,)).copy()
    svr = np.ctypeslib.as_array(C.cast(sr.data, C.POINTER(C.c_double)), shape=(ns,)).copy()
    assert np.max(np.abs(sv - svr)) <= 1e-13 * np.max(svr)
    U, V = u.to_dense(), vh.to_dense()
    assert helpers.rel_err((U * sv) @ V, dense) <= 1e-13
    assert np.abs(U.conj().T @ U - np.eye(ns)).max() <= 1e-13
    assert np.abs(V @ V.conj().T - np.eye(ns)).max() <= 1e-13
    cuda.delete_dense_tensor(C.byref(s)); ref.delete_dense_tensor(C.byref(sr))
