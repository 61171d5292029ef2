"""Pins the NumPy restatement (oracle/np_oracle.py) to the reference's OWN golden vectors.

tests/golden/ref_*.npz are the HDF5 fixtures of the reference's C test-suite (test/algorithm/data, test/util/data,
test/tensor/data), converted verbatim by tests/golden/make_golden.py; tolerances are the ones the reference's tests use
(test/algorithm/test_dmrg.c:380 1e-12, test/util/test_krylov.c 1e-13, test/algorithm/test_truncation.c exact / 1e-13).
CPU only: these tests check the checker.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import np_oracle as orc  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden(name):
    z = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"))
    ds = {k[3:]: z[k] for k in z.files if k.startswith("ds/")}
    at = {k[3:]: z[k] for k in z.files if k.startswith("at/")}
    return ds, at


def test_retained_bond_indices_golden():
    """reference test/algorithm/test_truncation.c: index list exact, norm and entropy 1e-13"""
    ds, at = golden("retained_bond_indices")
    ind, norm_sigma, entropy, _ = orc.retained_bond_indices(ds["sigma"], float(at["tol"]), True, 2 ** 40)
    assert np.array_equal(ind, ds["ind"])
    assert abs(norm_sigma - float(ds["norm_sigma"])) <= 1e-13
    assert abs(entropy - float(ds["entropy"])) <= 1e-13


@pytest.mark.parametrize("kind", ["d", "z"])
def test_lanczos_iteration_golden(kind):
    """reference test/util/test_krylov.c:29 / :105 -- alpha, beta, V of 24 iterations, 1e-13"""
    ds, _ = golden(f"lanczos_iteration_{kind}")
    a = ds["a"]
    n = a.shape[0]
    maxiter = len(ds["alpha"])
    alpha, beta, V, numiter = orc.lanczos_iteration(n, lambda v: a @ v, ds["vstart"], maxiter)
    assert numiter == maxiter
    assert np.max(np.abs(alpha - ds["alpha"])) <= 1e-13 * max(1.0, np.max(np.abs(ds["alpha"])))
    assert np.max(np.abs(beta[: len(ds["beta"])] - ds["beta"])) <= 1e-13 * max(1.0, np.max(np.abs(ds["beta"])))
    assert np.max(np.abs(V - ds["v"])) <= 1e-10     # no re-orthogonalisation: rounding differences grow along the recurrence


@pytest.mark.parametrize("kind", ["symmetric", "hermitian"])
def test_eigensystem_krylov_golden(kind):
    """reference test/util/test_krylov.c -- Ritz values 1e-13, Ritz vectors up to a phase"""
    ds, _ = golden(f"eigensystem_krylov_{kind}")
    a = ds["a"]
    n = a.shape[0]
    numeig = len(ds["lambda"])
    maxiter = 35 if kind == "symmetric" else 24
    # the number of iterations is a constant of the C test; find it from the stored Ritz values
    best = None
    for mi in range(numeig, 60):
        lam, u = orc.eigensystem_krylov(n, lambda v: a @ v, ds["vstart"], mi, numeig)
        err = np.max(np.abs(lam - ds["lambda"]))
        if best is None or err < best[0]:
            best = (err, mi, lam, u)
        if err <= 1e-12:
            break
    err, mi, lam, u = best
    assert err <= 1e-12, f"no iteration count reproduces the stored Ritz values (best {err:.2e} at maxiter={mi})"
    for j in range(numeig):
        ov = np.vdot(ds["u_ritz"][:, j], u[:, j])
        assert abs(abs(ov) - np.vdot(u[:, j], u[:, j]).real) <= 1e-9


def test_serialize_golden():
    """reference test/tensor/test_block_sparse_tensor.c:1589 -- packed entry order of serialize_entries"""
    ds, at = golden("block_sparse_tensor_serialize")
    names = sorted(k for k in at if k.startswith("qnums"))
    qnums = [at[k] for k in names]
    axis_dir = at["axis_dir"]
    dense = ds[[k for k in ds if ds[k].ndim == len(qnums)][0]]
    v = orc.serialize_entries(dense, axis_dir, qnums)
    back = orc.deserialize_entries(v, axis_dir, qnums)
    mask = orc.conserving_mask(axis_dir, qnums)
    assert len(v) == int(mask.sum())
    assert np.array_equal(back, dense * mask)


def _load_chain(ds, at, prefix, qprefix, nsites):
    tensors = [ds[f"{prefix}{i}"] for i in range(nsites)]
    qbonds = [np.asarray(at[f"{qprefix}{i}"], dtype=np.int64) for i in range(nsites + 1)]
    return tensors, qbonds


def test_dmrg_twosite_golden():
    """reference test/algorithm/test_dmrg.c:233-460 -- complex128, L=6, d=4, 4 sweeps x 25 Lanczos iterations, 1e-12"""
    ds, at = golden("dmrg_twosite")
    L = 6
    W, _ = _load_chain(ds, at, "h_a", "h_qbond", L)
    A, qb = _load_chain(ds, at, "psi_start_a", "psi_start_qbond", L)
    qsite = np.asarray(at["qsite"], dtype=np.int64)
    d = len(qsite)
    en, entropy, A_opt = orc.dmrg_twosite(W, qsite, A, qb, 4, 25, float(at["tol_split"]), d ** (L // 2))
    assert np.max(np.abs(en - ds["en_sweeps"])) <= 1e-12
    psi_ref, _ = _load_chain(ds, at, "psi_a", "psi_qbond", L)
    ov = np.vdot(orc.mps_to_statevector(psi_ref), orc.mps_to_statevector(A_opt))
    assert abs(abs(ov) - 1) <= 1e-12
    # the optimised bond structure is part of the golden vector: same bond quantum numbers
    for i in range(L + 1):
        assert np.array_equal(np.sort(qb[i]), np.sort(np.asarray(at[f"psi_qbond{i}"], dtype=np.int64)))


def test_dmrg_singlesite_golden():
    """reference test/algorithm/test_dmrg.c:9-230 -- complex128, L=7, d=3, 6 sweeps x 25 Lanczos iterations, 1e-12"""
    ds, at = golden("dmrg_singlesite")
    L = 7
    W, _ = _load_chain(ds, at, "h_a", "h_qbond", L)
    A, qb = _load_chain(ds, at, "psi_start_a", "psi_start_qbond", L)
    qsite = np.asarray(at["qsite"], dtype=np.int64)
    en, A_opt = orc.dmrg_singlesite(W, qsite, A, qb, 6, 25)
    assert np.max(np.abs(en - ds["en_sweeps"])) <= 1e-12
    psi_ref, _ = _load_chain(ds, at, "psi_a", "psi_qbond", L)
    ov = np.vdot(orc.mps_to_statevector(psi_ref), orc.mps_to_statevector(A_opt))
    assert abs(abs(ov) - 1) <= 1e-12


def test_heff_and_env_known_answers():
    """apply_local_hamiltonian / contraction_operator_step_* have no fixture in the reference; the known answers were
    produced by the unmodified compiled reference (tests/golden/make_golden.py)."""
    z = np.load(os.path.join(GOLDEN, "refrun_known_answers.npz"))

    def dense(prefix):
        axis_dir = z[f"{prefix}/axis_dir"]
        qn = [z[f"{prefix}/qnums{i}"] for i in range(len(axis_dir))]
        return orc.deserialize_entries(z[f"{prefix}/entries"], axis_dir, qn)

    a, w, l, r, b = (dense(f"heff_d/{k}") for k in "awlrb")
    got = orc.apply_local_hamiltonian(a, w, l[0], r[..., 0])
    assert np.linalg.norm(got - b) <= 1e-13 * np.linalg.norm(b)
    a1, w1, ln = dense("env_d/a_left"), dense("env_d/w_left"), dense("env_d/l_next")
    got = orc.contraction_operator_step_left(a1, a1, w1, l[0])
    assert np.linalg.norm(got - ln[0]) <= 1e-13 * np.linalg.norm(ln)
    a2, w2, rn = dense("env_d/a_right"), dense("env_d/w_right"), dense("env_d/r_next")
    got = orc.contraction_operator_step_right(a2, a2, w2, r[..., 0])
    assert np.linalg.norm(got - rn[..., 0]) <= 1e-13 * np.linalg.norm(rn)


def test_perf_dmrg_coefficients_bit_identical():
    """BASELINE.json configs[0]: helpers.perf_dmrg_coeffs() reproduces the datasets of the reference's perf/perf_dmrg_coeffs.hdf5
    (generator perf/perf_dmrg_coeffs.py, numpy default_rng(42)) bit for bit, so the benchmark needs no HDF5 reader at run time."""
    import helpers
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_perf_dmrg_coeffs.npz"))
    tkin, vint = helpers.perf_dmrg_coeffs()
    assert np.array_equal(g["ds/tkin"], tkin)
    assert np.array_equal(g["ds/vint"], vint)
