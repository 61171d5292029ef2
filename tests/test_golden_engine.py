"""The engine's C-ABI entry points against the committed golden vectors (tests/golden/*.npz).

ref_*.npz are the fixtures of the reference's own C test-suite (converted by tests/golden/make_golden.py); the tests
below mirror the reference tests that load them, with the reference's tolerances:
  test_dmrg_twosite / _singlesite        reference test/algorithm/test_dmrg.c:233-460 / :9-230       (1e-12)
  test_block_sparse_tensor_dot           reference test/tensor/test_block_sparse_tensor.c:1117-1233  (1e-13)
  test_block_sparse_tensor_qr/rq/svd     reference test/tensor/test_block_sparse_tensor.c:1236-1475  (1e-13, gauge-free properties)
  test_split_block_sparse_matrix_svd     reference test/algorithm/test_bond_ops.c:7                  (#retained exact, 5e-6 / 2e-6)
  test_retained_bond_indices             reference test/algorithm/test_truncation.c                  (exact / 1e-13)
  test_lanczos_iteration_d/z             reference test/util/test_krylov.c:29, :105                  (1e-13)
refrun_known_answers.npz holds outputs of the unmodified compiled reference for apply_local_hamiltonian and the
environment steps (no fixture exists for them in the reference) and per-sweep energies of the BASELINE.json model families.
Every test runs on the host-logic test double ("emu", CPU) and on the CUDA product ("cuda", -m gpu); nothing here needs
/root/reference or oracle/_ref.
"""
import ctypes as C
import os

import numpy as np
import pytest

import helpers
from chemtensor_b200 import cabi, workloads

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    z = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    ds = {k[3:]: z[k] for k in z.files if k.startswith("ds/")}
    at = {k[3:]: z[k] for k in z.files if k.startswith("at/")}
    return ds, at


def chain_from_dense(lib, kind, tensors, qsite, qbonds):
    if kind == "mps":
        dirs = [1, 1, -1]
        sites = [cabi.bst_from_dense(lib, np.ascontiguousarray(t), dirs, [qbonds[i], qsite, qbonds[i + 1]]) for i, t in enumerate(tensors)]
    else:
        dirs = [1, 1, -1, -1]
        sites = [cabi.bst_from_dense(lib, np.ascontiguousarray(t), dirs, [qbonds[i], qsite, qsite, qbonds[i + 1]]) for i, t in enumerate(tensors)]
    return cabi.Chain(lib, kind, np.asarray(qsite, dtype=np.int32), sites)


def statevector(chain):
    v = chain.site(0).to_dense()
    for i in range(1, chain.nsites):
        v = np.tensordot(v, chain.site(i).to_dense(), axes=(v.ndim - 1, 0))
    return v.reshape(-1)


def _dmrg_fixture(name, L):
    ds, at = golden(name)
    qsite = np.asarray(at["qsite"], dtype=np.int32)
    qh = [np.asarray(at[f"h_qbond{i}"], dtype=np.int32) for i in range(L + 1)]
    qp = [np.asarray(at[f"psi_start_qbond{i}"], dtype=np.int32) for i in range(L + 1)]
    qr = [np.asarray(at[f"psi_qbond{i}"], dtype=np.int32) for i in range(L + 1)]
    return ds, at, qsite, qh, qp, qr


def test_dmrg_twosite_golden(eng):
    L = 6
    ds, at, qsite, qh, qp, qr = _dmrg_fixture("dmrg_twosite", L)
    mpo = chain_from_dense(eng, "mpo", [ds[f"h_a{i}"] for i in range(L)], qsite, qh)
    psi = chain_from_dense(eng, "mps", [ds[f"psi_start_a{i}"] for i in range(L)], qsite, qp)
    nsweeps, d = 4, len(qsite)
    en = np.zeros(nsweeps); ent = np.zeros(L - 1)
    rc = eng.dmrg_twosite(mpo.ptr, nsweeps, 25, float(at["tol_split"]), d ** (L // 2), psi.ptr,
                          en.ctypes.data_as(C.POINTER(C.c_double)), ent.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0
    assert np.max(np.abs(en - ds["en_sweeps"])) <= 1e-12
    v = statevector(psi)
    assert abs(np.linalg.norm(v) - 1) <= 1e-12
    ref_psi = chain_from_dense(eng, "mps", [ds[f"psi_a{i}"] for i in range(L)], qsite, qr)
    assert abs(abs(np.vdot(statevector(ref_psi), v)) - 1) <= 1e-12
    # sector structure of the optimised bonds is bit-exact
    for i in range(L):
        assert np.array_equal(psi.site(i).qnums[0], qr[i])


def test_dmrg_singlesite_golden(eng):
    L = 7
    ds, at, qsite, qh, qp, qr = _dmrg_fixture("dmrg_singlesite", L)
    mpo = chain_from_dense(eng, "mpo", [ds[f"h_a{i}"] for i in range(L)], qsite, qh)
    psi = chain_from_dense(eng, "mps", [ds[f"psi_start_a{i}"] for i in range(L)], qsite, qp)
    nsweeps = 6
    en = np.zeros(nsweeps)
    rc = eng.dmrg_singlesite(mpo.ptr, nsweeps, 25, psi.ptr, en.ctypes.data_as(C.POINTER(C.c_double)))
    assert rc == 0
    assert np.max(np.abs(en - ds["en_sweeps"])) <= 1e-12
    v = statevector(psi)
    ref_psi = chain_from_dense(eng, "mps", [ds[f"psi_a{i}"] for i in range(L)], qsite, qr)
    assert abs(abs(np.vdot(statevector(ref_psi), v)) - 1) <= 1e-12


def test_block_sparse_tensor_dot_golden(eng):
    ds, at = golden("block_sparse_tensor_dot")
    ndim_mult = 3
    qn = [np.asarray(at[f"qnums{i}"], dtype=np.int32) for i in range(8)]
    axis_dir = [int(x) for x in at["axis_dir"]]
    s_dense, t_dense, r_ref = ds["s"], ds["t"], ds["r"]
    # s: axes 0..4 (the last three contracted), t: contracted axes first with opposite direction, then axes 5..7
    dir_s, qn_s = axis_dir[:5], qn[:5]
    dir_t = [-d for d in axis_dir[2:5]] + axis_dir[5:8]
    qn_t = qn[2:5] + qn[5:8]
    s = cabi.bst_from_dense(eng, np.ascontiguousarray(s_dense), dir_s, qn_s)
    t = cabi.bst_from_dense(eng, np.ascontiguousarray(t_dense), dir_t, qn_t)
    # the four LEADING / TRAILING combinations, as in the reference test
    for ax_s in (cabi.AXIS_RANGE_TRAILING, cabi.AXIS_RANGE_LEADING):
        for ax_t in (cabi.AXIS_RANGE_LEADING, cabi.AXIS_RANGE_TRAILING):
            sp, tp = s, t
            if ax_s == cabi.AXIS_RANGE_LEADING:
                perm = (C.c_int * 5)(2, 3, 4, 0, 1)
                sp = cabi.BST(eng); eng.block_sparse_tensor_transpose(perm, s.ptr, sp.ptr)
            if ax_t == cabi.AXIS_RANGE_TRAILING:
                perm = (C.c_int * 6)(3, 4, 5, 0, 1, 2)
                tp = cabi.BST(eng); eng.block_sparse_tensor_transpose(perm, t.ptr, tp.ptr)
            r = cabi.BST(eng)
            eng.block_sparse_tensor_dot(sp.ptr, ax_s, tp.ptr, ax_t, ndim_mult, r.ptr)
            assert np.max(np.abs(r.to_dense() - r_ref)) <= 1e-13 * max(1.0, np.max(np.abs(r_ref)))


@pytest.mark.parametrize("which", ["qr", "rq", "svd"])
def test_factorization_golden(eng, which):
    ds, at = golden(f"block_sparse_tensor_{which}")
    for c in (0, 1):
        a_dense = ds[f"a{c}"]
        dirs = [int(x) for x in at[f"axis_dir{c}"]]
        qn = [np.asarray(at[f"qnums{c}{i}"], dtype=np.int32) for i in range(2)]
        a = cabi.bst_from_dense(eng, np.ascontiguousarray(a_dense), dirs, qn)
        a_blocked = a.to_dense()
        x, y = cabi.BST(eng), cabi.BST(eng)
        if which == "qr":
            assert eng.block_sparse_tensor_qr(a.ptr, cabi.QR_REDUCED, x.ptr, y.ptr) == 0
            q, r = x.to_dense(), y.to_dense()
            assert np.max(np.abs(q @ r - a_blocked)) <= 1e-13 * max(1.0, np.max(np.abs(a_blocked)))
            assert np.max(np.abs(q.conj().T @ q - np.eye(q.shape[1]))) <= 1e-13
        elif which == "rq":
            assert eng.block_sparse_tensor_rq(a.ptr, cabi.QR_REDUCED, x.ptr, y.ptr) == 0
            r, q = x.to_dense(), y.to_dense()
            assert np.max(np.abs(r @ q - a_blocked)) <= 1e-13 * max(1.0, np.max(np.abs(a_blocked)))
            assert np.max(np.abs(q @ q.conj().T - np.eye(q.shape[0]))) <= 1e-13
        else:
            s = cabi.DenseTensor()
            assert eng.block_sparse_tensor_svd(a.ptr, x.ptr, C.byref(s), y.ptr) == 0
            u, vh = x.to_dense(), y.to_dense()
            k = u.shape[1]
            sv = np.ctypeslib.as_array(C.cast(s.data, C.POINTER(C.c_double)), shape=(k,)).copy()
            eng.delete_dense_tensor(C.byref(s))
            assert np.all(sv >= 0)
            assert np.max(np.abs((u * sv) @ vh - a_blocked)) <= 1e-13 * max(1.0, np.max(np.abs(a_blocked)))
            assert np.max(np.abs(u.conj().T @ u - np.eye(k))) <= 1e-13
            if c == 0:     # c == 1 is the "no matching sector" case: dummy bond, vh == 0 (reference block_sparse_tensor.c:2726-2776)
                assert np.max(np.abs(vh @ vh.conj().T - np.eye(k))) <= 1e-13
            else:
                assert k == 1 and np.all(vh == 0)


def test_split_block_sparse_matrix_svd_golden(eng):
    """single-precision fixture: the engine computes in double on the up-cast input (the DMRG path asserts double
    precision, reference dmrg.c:163,271); tolerances of the reference test (5e-6 isometry, 2e-6 reconstruction)"""
    ds, at = golden("split_block_sparse_matrix_svd")
    dirs = [int(x) for x in at["axis_dir"]]
    qn = [np.asarray(at[f"qnums{i}"], dtype=np.int32) for i in range(2)]
    a = cabi.bst_from_dense(eng, np.ascontiguousarray(ds["a"].astype(np.complex128)), dirs, qn)
    for renorm, key in ((False, "a_trunc_plain"), (True, "a_trunc_renrm")):
        for distr in (cabi.SVD_DISTR_LEFT, cabi.SVD_DISTR_RIGHT):
            a0, a1 = cabi.BST(eng), cabi.BST(eng)
            info = cabi.TruncInfo()
            rc = eng.split_block_sparse_matrix_svd(a.ptr, float(at["tol"]), True, 200, renorm, distr, a0.ptr, a1.ptr, C.byref(info))
            assert rc == 0
            assert a0.shape[1] == int(at["num_retained"]) == a1.shape[0]
            m0, m1 = a0.to_dense(), a1.to_dense()
            iso = m1 if distr == cabi.SVD_DISTR_LEFT else m0
            gram = iso @ iso.conj().T if distr == cabi.SVD_DISTR_LEFT else iso.conj().T @ iso
            assert np.max(np.abs(gram - np.eye(gram.shape[0]))) <= 5e-6
            assert np.max(np.abs(m0 @ m1 - ds[key])) <= 2e-6 * max(1.0, np.max(np.abs(ds[key])))


def test_retained_bond_indices_golden(eng):
    ds, at = golden("retained_bond_indices")
    sigma = np.ascontiguousarray(ds["sigma"], dtype=np.float64)
    lst, info = cabi.IndexList(), cabi.TruncInfo()
    eng.retained_bond_indices(sigma.ctypes.data_as(C.POINTER(C.c_double)), len(sigma), float(at["tol"]), True, 2 ** 40, C.byref(lst), C.byref(info))
    ind = np.ctypeslib.as_array(lst.ind, shape=(int(lst.num),)).copy()
    eng.delete_index_list(C.byref(lst))
    assert np.array_equal(ind, ds["ind"])
    assert abs(info.norm_sigma - float(ds["norm_sigma"])) <= 1e-13
    assert abs(info.entropy - float(ds["entropy"])) <= 1e-13


@pytest.mark.parametrize("kind", ["d", "z"])
def test_lanczos_iteration_golden(eng, kind):
    ds, _ = golden(f"lanczos_iteration_{kind}")
    a = np.ascontiguousarray(ds["a"])
    n = a.shape[0]
    maxiter = len(ds["alpha"])
    dt = a.dtype

    def afunc(nn, data, v, ret):
        vin = np.ctypeslib.as_array(C.cast(v, C.POINTER(C.c_double)), shape=(nn * (2 if dt.kind == "c" else 1),)).view(dt)
        out = np.ctypeslib.as_array(C.cast(ret, C.POINTER(C.c_double)), shape=(nn * (2 if dt.kind == "c" else 1),)).view(dt)
        out[:] = a @ vin

    cb = cabi.LANCZOS_FUNC(afunc)
    vstart = np.ascontiguousarray(ds["vstart"])
    alpha = np.zeros(maxiter); beta = np.zeros(maxiter); V = np.zeros((maxiter, n), dtype=dt)
    numiter = C.c_int(0)
    fn = eng.lanczos_iteration_d if kind == "d" else eng.lanczos_iteration_z
    fn(n, cb, None, vstart.ctypes.data, maxiter, alpha.ctypes.data_as(C.POINTER(C.c_double)), beta.ctypes.data_as(C.POINTER(C.c_double)), V.ctypes.data, C.byref(numiter))
    assert numiter.value == maxiter
    assert np.max(np.abs(alpha - ds["alpha"])) <= 1e-13 * max(1.0, np.max(np.abs(ds["alpha"])))
    nb = len(ds["beta"])
    assert np.max(np.abs(beta[:nb] - ds["beta"])) <= 1e-13 * max(1.0, np.max(np.abs(ds["beta"])))
    assert np.max(np.abs(V - ds["v"])) <= 1e-10


def _bst_from_packed(lib, z, prefix):
    axis_dir = [int(x) for x in z[f"{prefix}/axis_dir"]]
    qn = [z[f"{prefix}/qnums{i}"] for i in range(len(axis_dir))]
    ent = z[f"{prefix}/entries"]
    t = cabi.bst_allocate(lib, ent.dtype, [len(q) for q in qn], axis_dir, qn)
    t.deserialize(ent)
    return t


def test_heff_and_env_known_answers(eng):
    """matvec within 1e-12 relative Frobenius norm of the reference (BASELINE.json north_star), structure bit-exact"""
    z = np.load(os.path.join(GOLDEN, "refrun_known_answers.npz"))
    a, w, l, r = (_bst_from_packed(eng, z, f"heff_d/{k}") for k in "awlr")
    b = cabi.BST(eng)
    eng.apply_local_hamiltonian(a.ptr, w.ptr, l.ptr, r.ptr, b.ptr)
    assert [int(x) for x in z["heff_d/b/axis_dir"]] == b.axis_dir
    for i, q in enumerate(b.qnums):
        assert np.array_equal(q, z[f"heff_d/b/qnums{i}"])
    assert helpers.rel_err(b.serialize(), z["heff_d/b/entries"]) <= 1e-12
    a1, w1 = _bst_from_packed(eng, z, "env_d/a_left"), _bst_from_packed(eng, z, "env_d/w_left")
    ln = cabi.BST(eng)
    eng.contraction_operator_step_left(a1.ptr, a1.ptr, w1.ptr, l.ptr, ln.ptr)
    for i, q in enumerate(ln.qnums):
        assert np.array_equal(q, z[f"env_d/l_next/qnums{i}"])
    assert helpers.rel_err(ln.serialize(), z["env_d/l_next/entries"]) <= 1e-12
    a2, w2 = _bst_from_packed(eng, z, "env_d/a_right"), _bst_from_packed(eng, z, "env_d/w_right")
    rn = cabi.BST(eng)
    eng.contraction_operator_step_right(a2.ptr, a2.ptr, w2.ptr, r.ptr, rn.ptr)
    for i, q in enumerate(rn.qnums):
        assert np.array_equal(q, z[f"env_d/r_next/qnums{i}"])
    assert helpers.rel_err(rn.serialize(), z["env_d/r_next/entries"]) <= 1e-12
