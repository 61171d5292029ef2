"""More of the reference's own fixtures against the engine's C-ABI (tests/golden/ref_*.npz, converted from the HDF5 files of the
reference's C test-suite), mirroring the reference tests that load them:
  test_block_sparse_tensor_transpose / _reshape / _slice / _multiply_pointwise_vector   test/tensor/test_block_sparse_tensor.c:395-820
  test_eigensystem_krylov_symmetric / _hermitian                                          test/util/test_krylov.c
  test_mps_orthonormalize_qr, test_mps_split_tensor_svd                                   test/state/test_mps.c:227-350, :518-640
Single-precision fixtures are computed in double on the same entries and compared at single-precision accuracy."""
import ctypes as C

import numpy as np
import pytest

import helpers
from chemtensor_b200 import cabi
from test_golden_engine import chain_from_dense, golden, statevector


def _tensor(lib, ds_key, ds, at, ndim, dtype=None):
    dense = np.ascontiguousarray(ds[ds_key])
    if dtype is not None:
        dense = dense.astype(dtype)
    dirs = [int(x) for x in at["axis_dir"]]
    qn = [np.asarray(at[f"qnums{i}"], dtype=np.int32) for i in range(ndim)]
    return cabi.bst_from_dense(lib, dense, dirs, qn), dense, dirs, qn


def _masked(dense, dirs, qn):
    return np.where(cabi.conserving_mask(dense.shape, dirs, qn), dense, 0)


def test_transpose_golden(eng):
    ds, at = golden("block_sparse_tensor_transpose")
    t, dense, dirs, qn = _tensor(eng, "t", ds, at, 4)
    perm = [1, 3, 2, 0]
    r = cabi.BST(eng)
    eng.block_sparse_tensor_transpose((C.c_int * 4)(*perm), t.ptr, r.ptr)
    assert r.axis_dir == [dirs[p] for p in perm]
    for i, p in enumerate(perm):
        assert np.array_equal(r.qnums[i], qn[p])
    want = _masked(ds["t_tp"], [dirs[p] for p in perm], [qn[p] for p in perm])
    assert np.array_equal(r.to_dense(), want)          # pure data movement: exact


def test_reshape_golden(eng):
    ds, at = golden("block_sparse_tensor_reshape")
    t, dense, dirs, qn = _tensor(eng, "t", ds, at, 5)
    flat = cabi.BST(eng)
    eng.block_sparse_tensor_flatten_axes(t.ptr, 1, cabi.TENSOR_AXIS_OUT, flat.ptr)
    want = _masked(dense, dirs, qn)
    assert np.array_equal(flat.to_dense(), want.reshape(5, 28, 11, 3))
    # fused quantum numbers: new_dir * (dir_1 q_j + dir_2 q_k), reference block_sparse_tensor.c:972-979
    qf = (dirs[1] * qn[1][:, None] + dirs[2] * qn[2][None, :]).reshape(-1)
    assert np.array_equal(flat.qnums[1], qf)
    back = cabi.BST(eng)
    dims = (C.c_int64 * 2)(7, 4)
    d2 = (C.c_int * 2)(dirs[1], dirs[2])
    q1, q2 = np.ascontiguousarray(qn[1]), np.ascontiguousarray(qn[2])
    qp = (C.POINTER(C.c_int32) * 2)(q1.ctypes.data_as(C.POINTER(C.c_int32)), q2.ctypes.data_as(C.POINTER(C.c_int32)))
    eng.block_sparse_tensor_split_axis(flat.ptr, 1, dims, d2, qp, back.ptr)
    helpers.assert_bst_close(back, t, 0.0)


def test_slice_golden(eng):
    ds, at = golden("block_sparse_tensor_slice")
    t, dense, dirs, qn = _tensor(eng, "t", ds, at, 4, np.float64)
    ind = np.asarray(at["ind"], dtype=np.int64)
    s = cabi.BST(eng)
    eng.block_sparse_tensor_slice(t.ptr, 2, ind.ctypes.data_as(C.POINTER(C.c_int64)), len(ind), s.ptr)
    assert np.array_equal(s.qnums[2], qn[2][ind])
    qs = [qn[0], qn[1], qn[2][ind], qn[3]]
    assert np.array_equal(s.to_dense(), _masked(ds["s"].astype(np.float64), dirs, qs))


@pytest.mark.parametrize("which", [0, 1])
def test_multiply_pointwise_vector_golden(eng, which):
    ds, at = golden("block_sparse_tensor_multiply_pointwise_vector")
    s, dense, dirs, qn = _tensor(eng, "s", ds, at, 4, np.float64)
    vec = ds[f"t{which}"].astype(np.float64)
    dt, keep = cabi.dense_vector(eng, vec)
    r = cabi.BST(eng)
    eng.block_sparse_tensor_multiply_pointwise_vector(s.ptr, C.byref(dt), cabi.AXIS_RANGE_LEADING if which == 0 else cabi.AXIS_RANGE_TRAILING, r.ptr)
    want = _masked(ds[f"s_mult_t{which}"].astype(np.float64), dirs, qn)
    assert np.linalg.norm(r.to_dense() - want) <= 1e-6 * np.linalg.norm(want)
    # and exactly the product in double
    m = _masked(dense, dirs, qn)
    exact = m * (vec[:, None, None, None] if which == 0 else vec[None, None, None, :])
    assert np.array_equal(r.to_dense(), exact)


@pytest.mark.parametrize("kind", ["symmetric", "hermitian"])
def test_eigensystem_krylov_golden(eng, kind):
    ds, _ = golden(f"eigensystem_krylov_{kind}")
    a = np.ascontiguousarray(ds["a"])
    n = a.shape[0]
    dt = a.dtype
    lam_ref = np.asarray(ds["lambda"])
    numeig = len(lam_ref)

    def matvec(nn, data, v, ret):
        vin = np.ctypeslib.as_array(C.cast(v, C.POINTER(C.c_double)), shape=(nn * (2 if dt.kind == "c" else 1),)).view(dt)
        out = np.ctypeslib.as_array(C.cast(ret, C.POINTER(C.c_double)), shape=(nn * (2 if dt.kind == "c" else 1),)).view(dt)
        out[:] = a @ vin
    cb = cabi.LANCZOS_FUNC(matvec)
    v0 = np.ascontiguousarray(ds["vstart"])
    lam = np.zeros(numeig)
    u = np.zeros((n, numeig), dtype=dt)
    fn = eng.eigensystem_krylov_symmetric if kind == "symmetric" else eng.eigensystem_krylov_hermitian
    maxiter = 35 if kind == "symmetric" else 37      # constants of test/util/test_krylov.c:199, :284
    assert fn(n, cb, None, v0.ctypes.data, maxiter, numeig, lam.ctypes.data_as(C.POINTER(C.c_double)), u.ctypes.data) == 0
    assert np.max(np.abs(lam - lam_ref)) <= 1e-13      # the reference test's tolerance
    u_ref = np.asarray(ds["u_ritz"])
    for e in range(numeig):
        assert abs(abs(np.vdot(u[:, e], u_ref[:, e])) - 1.0) <= 1e-10


@pytest.mark.parametrize("mode", [cabi.MPS_ORTHONORMAL_LEFT, cabi.MPS_ORTHONORMAL_RIGHT])
def test_mps_orthonormalize_qr_golden(eng, mode):
    ds, at = golden("mps_orthonormalize_qr")
    L = 6
    qsite = np.asarray(at["qsite"], dtype=np.int32)
    qb = [np.asarray(at[f"qbond{i}"], dtype=np.int32) for i in range(L + 1)]
    psi = chain_from_dense(eng, "mps", [ds[f"a{i}"].astype(np.complex128) for i in range(L)], qsite, qb)
    v_ref = statevector(psi)
    norm = eng.mps_orthonormalize_qr(psi.ptr, mode)
    v = statevector(psi)
    assert abs(np.linalg.norm(v) - 1) <= 1e-12
    assert np.linalg.norm(norm * v - v_ref) <= 1e-12 * np.linalg.norm(v_ref)
    assert abs(norm - np.linalg.norm(v_ref)) <= 1e-12 * norm
    for i in range(L):
        a = psi.site(i).to_dense()
        m = a.reshape(-1, a.shape[2]) if mode == cabi.MPS_ORTHONORMAL_LEFT else a.reshape(a.shape[0], -1).conj().T
        assert np.allclose(m.conj().T @ m, np.eye(m.shape[1]), atol=1e-12)


@pytest.mark.parametrize("distr", [cabi.SVD_DISTR_LEFT, cabi.SVD_DISTR_RIGHT])
@pytest.mark.parametrize("truncate", [False, True])
def test_mps_split_tensor_svd_golden(eng, distr, truncate):
    ds, at = golden("mps_split_tensor_svd")
    q0, q1 = np.asarray(at["qsite0"], dtype=np.int32), np.asarray(at["qsite1"], dtype=np.int32)
    qb0, qb1 = np.asarray(at["qbonds0"], dtype=np.int32), np.asarray(at["qbonds1"], dtype=np.int32)
    q2 = (q0[:, None] + q1[None, :]).reshape(-1).astype(np.int32)
    a_pair = cabi.bst_from_dense(eng, np.ascontiguousarray(ds["a_pair"]), [1, 1, -1], [qb0, q2, qb1])
    dims = (C.c_int64 * 2)(len(q0), len(q1))
    qp = (C.POINTER(C.c_int32) * 2)(q0.ctypes.data_as(C.POINTER(C.c_int32)), q1.ctypes.data_as(C.POINTER(C.c_int32)))
    a0, a1, mrg, info = cabi.BST(eng), cabi.BST(eng), cabi.BST(eng), cabi.TruncInfo()
    tol = float(at["tol"]) if truncate else 0.0
    assert eng.mps_split_tensor_svd(a_pair.ptr, dims, qp, tol, 100, False, distr, a0.ptr, a1.ptr, C.byref(info)) == 0
    eng.mps_merge_tensor_pair(a0.ptr, a1.ptr, mrg.ptr)
    want = a_pair.to_dense() if not truncate else _masked(np.asarray(ds["a_mrg"]), [1, 1, -1], [qb0, q2, qb1])
    assert np.max(np.abs(mrg.to_dense() - want)) <= 1e-13 * max(1.0, np.max(np.abs(want)))
