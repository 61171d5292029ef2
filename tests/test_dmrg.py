"""Krylov interface and the DMRG drivers against the compiled reference on identical seeded inputs.

Reference tests mirrored: test/util/test_krylov.c :29, :105 (alpha, beta, V of a dense matrix; Ritz values 1e-13,
vectors 1e-12 up to sign) and test/algorithm/test_dmrg.c :9, :233 (en_sweeps to 1e-12, norm 1, overlap 1).
Bars (BASELINE.json north_star): energies within 1e-10, bond quantum numbers / dimensions bit-exact.
"""
import ctypes as C

import numpy as np
import pytest

import helpers
from chemtensor_b200 import cabi


def _matvec_callback(mat, dtype):
    def cb(n, data, v, ret):
        x = np.ctypeslib.as_array(C.cast(v, C.POINTER(C.c_double)), shape=(n * (2 if dtype == np.complex128 else 1),)).view(dtype)
        y = np.ctypeslib.as_array(C.cast(ret, C.POINTER(C.c_double)), shape=(n * (2 if dtype == np.complex128 else 1),)).view(dtype)
        y[:] = mat @ x
    return cabi.LANCZOS_FUNC(cb)


@pytest.mark.parametrize("dtype,n", [(np.float64, 319), (np.complex128, 173)])
def test_lanczos_iteration(eng, ref, rng, dtype, n):
    a = rng.standard_normal((n, n))
    if dtype == np.complex128:
        a = a + 1j * rng.standard_normal((n, n))
    a = 0.5 * (a + a.conj().T)
    v0 = rng.standard_normal(n).astype(dtype)
    if dtype == np.complex128:
        v0 = v0 + 1j * rng.standard_normal(n)
    maxiter = 24
    cb = _matvec_callback(a, dtype)
    out = []
    for lib in (eng, ref):
        alpha, beta = np.zeros(maxiter), np.zeros(maxiter - 1 + 1)
        v = np.zeros((maxiter, n), dtype=dtype)
        numiter = C.c_int(0)
        fn = lib.lanczos_iteration_d if dtype == np.float64 else lib.lanczos_iteration_z
        fn(n, cb, None, v0.ctypes.data, maxiter, alpha.ctypes.data_as(C.POINTER(C.c_double)), beta.ctypes.data_as(C.POINTER(C.c_double)), v.ctypes.data, C.byref(numiter))
        out.append((alpha, beta[:maxiter - 1], v, numiter.value))
    assert out[0][3] == out[1][3] == maxiter
    # the three-term recurrence amplifies rounding differences: compare the early coefficients tightly
    assert np.max(np.abs(out[0][0][:8] - out[1][0][:8])) <= 1e-11
    assert np.max(np.abs(out[0][1][:8] - out[1][1][:8])) <= 1e-11
    assert helpers.rel_err(out[0][2][:8], out[1][2][:8]) <= 1e-10
    # orthonormality of the leading Krylov vectors
    v = out[0][2]
    assert np.linalg.norm(v[:8].conj() @ v[:8].T - np.eye(8)) <= 1e-9


@pytest.mark.parametrize("dtype,n", [(np.float64, 197), (np.complex128, 114)])
def test_eigensystem_krylov(eng, ref, rng, dtype, n):
    a = rng.standard_normal((n, n))
    if dtype == np.complex128:
        a = a + 1j * rng.standard_normal((n, n))
    a = 0.5 * (a + a.conj().T)
    v0 = rng.standard_normal(n).astype(dtype)
    maxiter, numeig = 35, 3
    cb = _matvec_callback(a, dtype)
    out = []
    for lib in (eng, ref):
        lam = np.zeros(numeig)
        u = np.zeros((n, numeig), dtype=dtype)
        fn = lib.eigensystem_krylov_symmetric if dtype == np.float64 else lib.eigensystem_krylov_hermitian
        assert fn(n, cb, None, v0.ctypes.data, maxiter, numeig, lam.ctypes.data_as(C.POINTER(C.c_double)), u.ctypes.data) == 0
        out.append((lam, u))
    assert np.max(np.abs(out[0][0] - out[1][0])) <= 1e-10
    for e in range(numeig):
        ov = abs(np.vdot(out[0][1][:, e], out[1][1][:, e]))
        assert abs(ov - 1.0) <= 1e-8


def _energy_cases():
    return [
        ("xxz", 8, (1.0, 0.8, 0.1), 0, 32, np.float64),
        ("fermi_hubbard", 6, (1.0, 4.0, 0.3), helpers.encode_qpair(6, 0), 60, np.float64),
        ("xxz", 6, (1.0, 0.8, 0.1), 0, 16, np.complex128),
    ]


@pytest.mark.parametrize("case", _energy_cases(), ids=lambda c: f"{c[0]}-L{c[1]}-{np.dtype(c[5]).name}")
def test_dmrg_twosite(eng, ref, case):
    model, L, params, sector, max_vdim, dtype = case
    mpo_r = helpers.ref_mpo(ref, model, L, *params)
    psi0 = helpers.ref_random_mps(ref, dtype, L, mpo_r.qsite, sector, max_vdim, seed=42)
    num_sweeps, maxiter, tol = 3, 25, 1e-10
    res = []
    for lib in (eng, ref):
        mpo = _cast(lib, mpo_r, dtype)
        psi = helpers.clone_chain(lib, psi0)
        en = np.zeros(num_sweeps); ent = np.zeros(L - 1)
        rc = lib.dmrg_twosite(mpo.ptr, num_sweeps, maxiter, tol, max_vdim, psi.ptr, en.ctypes.data_as(C.POINTER(C.c_double)), ent.ctypes.data_as(C.POINTER(C.c_double)))
        assert rc == 0
        res.append((en, ent, psi))
    en_e, ent_e, psi_e = res[0]
    en_r, ent_r, psi_r = res[1]
    assert np.max(np.abs(en_e - en_r)) <= 1e-10, (en_e, en_r)
    assert np.max(np.abs(ent_e - ent_r)) <= 1e-7
    assert psi_e.bond_dims() == psi_r.bond_dims()
    for i in range(L):   # bond quantum numbers bit-exact
        for qa, qb in zip(psi_e.site(i).qnums, psi_r.site(i).qnums):
            assert np.array_equal(qa, qb)
    # same state up to a phase: |<psi_e|psi_r>| = 1, norm 1
    psi_e_in_ref = helpers.clone_chain(ref, psi_e)
    ov = np.zeros(1, dtype=dtype)
    ref.dll.mps_vdot(psi_e_in_ref.ptr, psi_r.ptr, ov.ctypes.data)
    assert abs(abs(ov[0]) - 1.0) <= 1e-8
    assert abs(ref.dll.mps_norm(psi_e_in_ref.ptr) - 1.0) <= 1e-12


def _cast(lib, src, dtype):
    tensors = []
    for i in range(src.nsites):
        s = src.site(i)
        t = cabi.bst_allocate(lib, dtype, s.shape, s.axis_dir, s.qnums)
        for (_, a), (_, b) in zip(t.blocks(), s.blocks()):
            a[...] = b
        tensors.append(t)
    return cabi.Chain(lib, src.kind, src.qsite, tensors)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_dmrg_singlesite(eng, ref, dtype):
    L, max_vdim = 7, 20
    mpo_r = helpers.ref_mpo(ref, "xxz", L, 1.0, 0.8, 0.1)
    psi0 = helpers.ref_random_mps(ref, dtype, L, mpo_r.qsite, 1, max_vdim, seed=3)
    num_sweeps, maxiter = 4, 25
    res = []
    for lib in (eng, ref):
        mpo = _cast(lib, mpo_r, dtype)
        psi = helpers.clone_chain(lib, psi0)
        en = np.zeros(num_sweeps)
        assert lib.dmrg_singlesite(mpo.ptr, num_sweeps, maxiter, psi.ptr, en.ctypes.data_as(C.POINTER(C.c_double))) == 0
        res.append((en, psi))
    assert np.max(np.abs(res[0][0] - res[1][0])) <= 1e-10, (res[0][0], res[1][0])
    assert res[0][1].bond_dims() == res[1][1].bond_dims()


def test_dmrg_twosite_molecular_complex(eng, ref):
    """BASELINE.json configs[3] in small: synthetic molecular Hamiltonian from random complex integrals with the Hermitian
    symmetrisation of perf/perf_dmrg_coeffs.py:8-17 (conjugated), spatial orbitals (d = 4, packed (N, 2Sz) quantum numbers),
    complex128 two-site DMRG; energies within 1e-10 of the reference, bond structure bit-exact."""
    n = 5
    rng = np.random.default_rng(5)
    tkin = 0.5 * (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    vint = 0.1 * (rng.standard_normal((n, n, n, n)) + 1j * rng.standard_normal((n, n, n, n)))
    tkin = 0.5 * (tkin + tkin.conj().T)
    vint = 0.5 * (vint + vint.transpose((1, 0, 3, 2)))
    vint = 0.5 * (vint + vint.transpose((2, 3, 0, 1)).conj())
    mpo_r = helpers.ref_molecular_mpo(ref, tkin, vint, spin=True, optimize=False)
    L = mpo_r.nsites
    assert mpo_r.site(1).dtype == np.complex128 and len(mpo_r.qsite) == 4
    psi0 = helpers.ref_random_mps(ref, np.complex128, L, mpo_r.qsite, helpers.encode_qpair(n, 1), 40, seed=42)
    num_sweeps, maxiter, tol, max_vdim = 3, 25, 1e-10, 40
    res = []
    for lib in (eng, ref):
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
    ov = np.zeros(1, dtype=np.complex128)
    psi_e_in_ref = helpers.clone_chain(ref, psi_e)      # keep the handle alive across the call
    ref.mps_vdot(psi_e_in_ref.ptr, psi_r.ptr, ov.ctypes.data)
    assert abs(abs(ov[0]) - 1.0) <= 1e-8
