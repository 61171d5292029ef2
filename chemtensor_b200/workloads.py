"""Synthetic inputs of the shapes BASELINE.json names, built without the reference.

* nearest-neighbour Hamiltonian MPOs with additive quantum numbers (same operators, coefficients and
  physical quantum numbers as the reference's constructors, src/operator/hamiltonian.c:102 (XXZ) and
  :240 (Fermi-Hubbard); the virtual-bond basis is the plain finite-state-machine one, i.e. a permutation
  of the reference's graph-derived basis in the bulk and one or two extra states at the two end bonds),
* virtual-bond quantum numbers of a random MPS by the reference's rule (src/state/mps.c:93-150):
  all combinations of the previous bond with the site, sub-sampled to max_vdim,
* the four operands (a, w, l, r) of one two-site effective-Hamiltonian application at a given bond,
  with the sector structure the sweep produces there and N(0, 1) entries.

Everything returns host-memory `struct block_sparse_tensor`s owned by the given C library (cabi.BST).
"""
from __future__ import annotations

import numpy as np

from . import cabi


def encode_qpair(qa: int, qb: int) -> int:
    """reference include/tensor/qnumber.h:47"""
    return (int(qa) << 16) + int(qb)


# ------------------------------------------------------------------------------------------------
# Hamiltonians as finite-state-machine MPOs
# ------------------------------------------------------------------------------------------------

def _fsm_mpo(nsites: int, qsite, onsite: np.ndarray, terms):
    """H = sum_i [ onsite_i + sum_k A_k(i) B_k(i+1) ];  terms = [(A_k, B_k, q_k)], q_k = charge A_k adds.

    Bond states: 0 = nothing applied yet, 1..K = A_k applied, K+1 = finished.
    Returns (tensors [Dl, d, d, Dr], bond quantum numbers)."""
    d = len(qsite)
    K = len(terms)
    Dw = K + 2
    eye = np.eye(d)
    bulk = np.zeros((Dw, d, d, Dw))
    bulk[0, :, :, 0] = eye
    bulk[Dw - 1, :, :, Dw - 1] = eye
    bulk[0, :, :, Dw - 1] = onsite
    for k, (A, B, _) in enumerate(terms):
        bulk[0, :, :, 1 + k] = A
        bulk[1 + k, :, :, Dw - 1] = B
    qbulk = np.array([0] + [q for (_, _, q) in terms] + [0], dtype=np.int32)
    tensors, qbonds = [], []
    for i in range(nsites):
        t = bulk
        if i == 0:
            t = t[0:1]
        if i == nsites - 1:
            t = t[:, :, :, Dw - 1:Dw]
        tensors.append(np.ascontiguousarray(t))
    qbonds = [np.zeros(1, dtype=np.int32)] + [qbulk.copy() for _ in range(nsites - 1)] + [np.zeros(1, dtype=np.int32)]
    return tensors, qbonds


def xxz_mpo(nsites: int, J: float, D: float, h: float):
    """sum J (X X + Y Y + D Z Z) - h Z  (reference hamiltonian.c:99-160); qsite = 2 Sz = (1, -1)."""
    qsite = np.array([1, -1], dtype=np.int32)
    sup = np.array([[0., 1.], [0., 0.]])
    sdn = np.array([[0., 0.], [1., 0.]])
    sz = np.array([[0.5, 0.], [0., -0.5]])
    terms = [(0.5 * J * sup, sdn, 2), (0.5 * J * sdn, sup, -2), (J * D * sz, sz, 0)]
    tensors, qbonds = _fsm_mpo(nsites, qsite, -h * sz, terms)
    return tensors, qbonds, qsite


def fermi_hubbard_mpo(nsites: int, t: float, u: float, mu: float):
    """-t sum (a^dag_{i s} a_{i+1 s} + h.c.) + u (n_up - 1/2)(n_dn - 1/2) - mu (n_up + n_dn)
    (reference hamiltonian.c:236-370); local basis |n_up n_dn>, quantum numbers (N, 2 Sz) packed."""
    qn, qs = [0, 1, 1, 2], [0, -1, 1, 0]
    qsite = np.array([encode_qpair(a, b) for a, b in zip(qn, qs)], dtype=np.int32)
    I2 = np.eye(2)
    ad = np.array([[0., 0.], [1., 0.]])
    an = np.array([[0., 1.], [0., 0.]])
    Z = np.diag([1., -1.])
    CI, AI = np.kron(ad, I2), np.kron(an, I2)
    CZ, AZ = np.kron(ad, Z), np.kron(an, Z)
    IC, IA = np.kron(I2, ad), np.kron(I2, an)
    ZC, ZA = np.kron(Z, ad), np.kron(Z, an)
    ntot = np.diag([0., 1., 1., 2.])
    nint = np.diag([0.25, -0.25, -0.25, 0.25])
    terms = [(-t * CZ, AI, encode_qpair(1, 1)), (-t * AZ, CI, encode_qpair(-1, -1)),
             (-t * IC, ZA, encode_qpair(1, -1)), (-t * IA, ZC, encode_qpair(-1, 1))]
    tensors, qbonds = _fsm_mpo(nsites, qsite, -mu * ntot + u * nint, terms)
    return tensors, qbonds, qsite


MODELS = {"xxz": xxz_mpo, "fermi_hubbard": fermi_hubbard_mpo}


def mpo_chain(lib: cabi.CLibrary, model: str, nsites: int, params, dtype=np.float64) -> cabi.Chain:
    tensors, qbonds, qsite = MODELS[model](nsites, *params)
    dirs = [cabi.TENSOR_AXIS_OUT, cabi.TENSOR_AXIS_OUT, cabi.TENSOR_AXIS_IN, cabi.TENSOR_AXIS_IN]
    site_tensors = [cabi.bst_from_dense(lib, t.astype(dtype), dirs, [qbonds[i], qsite, qsite, qbonds[i + 1]])
                    for i, t in enumerate(tensors)]
    return cabi.Chain(lib, "mpo", qsite, site_tensors)


def mpo_to_matrix(tensors) -> np.ndarray:
    """Dense matrix of an MPO given as [Dl, d, d, Dr] arrays (small chains; test helper)."""
    m = tensors[0]
    for t in tensors[1:]:
        m = np.tensordot(m, t, axes=(m.ndim - 1, 0))
    m = m.reshape(m.shape[1:-1])
    n = m.ndim // 2
    perm = list(range(0, 2 * n, 2)) + list(range(1, 2 * n, 2))
    dims = int(np.prod([m.shape[i] for i in range(0, 2 * n, 2)]))
    return m.transpose(perm).reshape(dims, dims)


# ------------------------------------------------------------------------------------------------
# random MPS structure and tensors
# ------------------------------------------------------------------------------------------------

def random_bond_qnums(nsites: int, qsite, qnum_sector: int, max_vdim: int, rng: np.random.Generator):
    """Virtual-bond quantum numbers by the rule of the reference's construct_random_mps (mps.c:99-150)."""
    qsite = np.asarray(qsite, dtype=np.int64)
    qb = [None] * (nsites + 1)
    qb[0] = np.zeros(1, dtype=np.int64)
    qb[nsites] = np.array([qnum_sector], dtype=np.int64)
    for l in range(1, (nsites + 1) // 2):
        full = (qb[l - 1][:, None] + qsite[None, :]).reshape(-1)
        qb[l] = full if len(full) <= max_vdim else full[rng.choice(len(full), size=max_vdim, replace=False)]
    for l in range(nsites - 1, (nsites + 1) // 2 - 1, -1):
        full = (qb[l + 1][:, None] - qsite[None, :]).reshape(-1)
        qb[l] = full if len(full) <= max_vdim else full[rng.choice(len(full), size=max_vdim, replace=False)]
    return [q.astype(np.int32) for q in qb]


def fill_random(t: cabi.BST, rng: np.random.Generator, scale: float = 1.0) -> None:
    for _, a in t.blocks():
        a[...] = scale * rng.standard_normal(a.shape)
        if np.dtype(t.dtype).kind == "c":
            a[...] += 1j * scale * rng.standard_normal(a.shape)


def random_mps(lib: cabi.CLibrary, dtype, nsites: int, qsite, qnum_sector: int, max_vdim: int, seed: int = 42) -> cabi.Chain:
    rng = np.random.default_rng(seed)
    qb = random_bond_qnums(nsites, qsite, qnum_sector, max_vdim, rng)
    dirs = [cabi.TENSOR_AXIS_OUT, cabi.TENSOR_AXIS_OUT, cabi.TENSOR_AXIS_IN]
    tensors = []
    for i in range(nsites):
        shape = (len(qb[i]), len(qsite), len(qb[i + 1]))
        t = cabi.bst_allocate(lib, dtype, shape, dirs, [qb[i], qsite, qb[i + 1]])
        fill_random(t, rng, 1.0 / np.sqrt(float(np.prod(shape))))
        tensors.append(t)
    return cabi.Chain(lib, "mps", np.asarray(qsite, dtype=np.int32), tensors)


# ------------------------------------------------------------------------------------------------
# operands of one two-site effective-Hamiltonian application
# ------------------------------------------------------------------------------------------------

def measured_bond_qnums(data_file: str, bond: int, scale: int = 1, shift_q: int = 0, max_vdim: int | None = None):
    """Bond quantum numbers with the sector histogram a converged two-site sweep produced (chemtensor_b200/data/bonds_*.json,
    recorded by tools/measure_bond_structure.py on the GPU), every multiplicity multiplied by `scale` and every quantum
    number shifted by `shift_q` (a longer chain at the same filling).  Ordering as the SVD split leaves it: sectors
    ascending, all entries of a sector contiguous (reference block_sparse_tensor.c:2694-2724)."""
    import json
    import os
    path = data_file if os.path.isabs(data_file) else os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", data_file)
    with open(path) as f:
        rec = json.load(f)
    hist = {int(q) + shift_q: int(m) * scale for q, m in rec["bonds"][bond]["sectors"].items()}
    if max_vdim is not None:
        # trim the largest sectors proportionally if integer scaling overshoots the bond dimension
        tot = sum(hist.values())
        while tot > max_vdim:
            q = max(hist, key=hist.get)
            hist[q] -= 1
            tot -= 1
    qs = sorted(hist)
    return np.concatenate([np.full(hist[q], q, dtype=np.int32) for q in qs if hist[q] > 0])


def heff_operands(lib: cabi.CLibrary, model: str, nsites: int, params, qnum_sector: int, max_vdim: int, site: int | None = None,
                  dtype=np.float64, seed: int = 42, bonds=None, with_sites: bool = False):
    """(a, w, l, r) for the pair (site, site+1): the merged two-site MPS tensor a[Dl, d^2, Dr], the merged MPO tensor
    w[Dw, d^2, d^2, Dw'] (real Hamiltonian entries), and environments l[1, Dl, Dw, Dl], r[Dr, Dw', Dr, 1] with the
    structure of the reference's contraction_operator_step_left/right outputs (chain_ops.c:116, :196) and random entries."""
    rng = np.random.default_rng(seed)
    tensors, qwb, qsite = MODELS[model](nsites, *params)
    if site is None:
        site = nsites // 2 - 1
    if bonds is None:
        qb = random_bond_qnums(nsites, qsite, qnum_sector, max_vdim, rng)
    else:
        # measured (converged) sector structure of the two bonds around the pair: (q_left, q_right)
        qb = {site: np.asarray(bonds[0], dtype=np.int32), site + 2: np.asarray(bonds[1], dtype=np.int32)}
    OUT, IN = cabi.TENSOR_AXIS_OUT, cabi.TENSOR_AXIS_IN
    # merged physical leg: logical index j*d + k, quantum number q_j + q_k (flatten_axes with both legs OUT)
    q2 = (np.asarray(qsite, dtype=np.int64)[:, None] + np.asarray(qsite, dtype=np.int64)[None, :]).reshape(-1).astype(np.int32)
    ql, qr = qb[site], qb[site + 2]
    a = cabi.bst_allocate(lib, dtype, (len(ql), len(q2), len(qr)), [OUT, OUT, IN], [ql, q2, qr])
    fill_random(a, rng, 1.0)
    # two-site MPO tensor from the real site tensors: w[l, (s0 s1), (t0 t1), r]
    w2 = np.tensordot(tensors[site], tensors[site + 1], axes=(3, 0)).transpose(0, 1, 3, 2, 4, 5)
    d = len(qsite)
    w2 = w2.reshape(w2.shape[0], d * d, d * d, w2.shape[5]).astype(dtype)
    w = cabi.bst_from_dense(lib, w2, [OUT, OUT, IN, IN], [qwb[site], q2, q2, qwb[site + 2]])
    # environments; the outer dummy leg carries q = q_a + q_w - q_b of the chain end, which is 0 on both sides
    q0 = np.zeros(1, dtype=np.int32)
    l = cabi.bst_allocate(lib, dtype, (1, len(ql), len(qwb[site]), len(ql)), [OUT, IN, IN, OUT], [q0, ql, qwb[site], ql])
    r = cabi.bst_allocate(lib, dtype, (len(qr), len(qwb[site + 2]), len(qr), 1), [OUT, OUT, IN, IN], [qr, qwb[site + 2], qr, q0])
    fill_random(l, rng, 1.0 / np.sqrt(len(ql)))
    fill_random(r, rng, 1.0 / np.sqrt(len(qr)))
    if with_sites:
        # the two single-site MPO tensors the merged one was built from (pair form of the effective Hamiltonian)
        qs = np.asarray(qsite, dtype=np.int32)
        w0 = cabi.bst_from_dense(lib, np.ascontiguousarray(tensors[site]).astype(dtype), [OUT, OUT, IN, IN], [qwb[site], qs, qs, qwb[site + 1]])
        w1 = cabi.bst_from_dense(lib, np.ascontiguousarray(tensors[site + 1]).astype(dtype), [OUT, OUT, IN, IN], [qwb[site + 1], qs, qs, qwb[site + 2]])
        return a, w, l, r, w0, w1
    return a, w, l, r
