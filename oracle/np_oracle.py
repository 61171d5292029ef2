"""TEST INFRASTRUCTURE -- dense NumPy restatement of the reference's DMRG hot path.  Never imported by the product.

Every function restates the ALGORITHM of one reference routine (cited file:line under the reference tree) on plain
dense arrays plus per-axis quantum-number lists: quantum-number conservation appears as zero patterns, not as block
storage, so sizes are limited to bond dimensions of ~10^2.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import this module, and only as the checker.

Pinned against the reference's own golden vectors (tests/golden/ref_*.npz, converted from the HDF5 fixtures of the
reference's C test-suite) in tests/test_oracle_golden.py: truncation rule (exact index list), Lanczos (alpha, beta, V),
Krylov eigen-solvers, sector/serialisation order, and the per-sweep energies of test_dmrg_twosite / test_dmrg_singlesite
(1e-12).  The second, stronger oracle is the unmodified reference itself compiled into oracle/_ref (oracle/Makefile).

Conventions (reference src/state/mps.c:41, src/operator/mpo.c:48):
  MPS site tensor a[Dl, d, Dr]      directions (OUT, OUT, IN):      q_l + q_s - q_r = 0
  MPO site tensor w[Dw, d, d, Dw']  directions (OUT, OUT, IN, IN):  q_w + q_s' - q_s - q_w' = 0
  environments    l[Dl, Dw, Dl'], r[Dr, Dw', Dr'] (the reference's dummy outer legs of dimension 1 are dropped)
"""
from __future__ import annotations

import numpy as np

OUT, IN = 1, -1


# ------------------------------------------------------------------------------------------------------------------
# sector structure (reference src/tensor/block_sparse_tensor.c:37-141)
# ------------------------------------------------------------------------------------------------------------------

def conserving_mask(axis_dir, qnums) -> np.ndarray:
    """True where sum_i dir_i q_i == 0 (block_sparse_tensor.c:126-133)."""
    shape = tuple(len(q) for q in qnums)
    tot = np.zeros(shape, dtype=np.int64)
    for i, (d, q) in enumerate(zip(axis_dir, qnums)):
        sh = [1] * len(shape)
        sh[i] = shape[i]
        tot = tot + int(d) * np.asarray(q, dtype=np.int64).reshape(sh)
    return tot == 0


def sector_lists(qnums):
    """Per axis: sorted distinct quantum numbers (block_sparse_tensor.c:79-117)."""
    return [np.unique(np.asarray(q, dtype=np.int64)) for q in qnums]


def serialize_entries(dense: np.ndarray, axis_dir, qnums) -> np.ndarray:
    """Packed vector of the stored blocks: conserving cells of the sector grid in row-major order, each block row-major
    with rows/columns in order of appearance (block_sparse_tensor_serialize_entries, block_sparse_tensor.c:3131-3150)."""
    secs = sector_lists(qnums)
    ndim = dense.ndim
    parts = []
    for cell in np.ndindex(*[len(s) for s in secs]):
        if sum(int(axis_dir[i]) * int(secs[i][cell[i]]) for i in range(ndim)) != 0:
            continue
        sel = [np.nonzero(np.asarray(qnums[i]) == secs[i][cell[i]])[0] for i in range(ndim)]
        parts.append(dense[np.ix_(*sel)].reshape(-1))
    return np.concatenate(parts) if parts else np.zeros(0, dtype=dense.dtype)


def deserialize_entries(entries: np.ndarray, axis_dir, qnums) -> np.ndarray:
    """Inverse of serialize_entries (block_sparse_tensor.c:3157-3176)."""
    secs = sector_lists(qnums)
    ndim = len(qnums)
    dense = np.zeros(tuple(len(q) for q in qnums), dtype=entries.dtype)
    pos = 0
    for cell in np.ndindex(*[len(s) for s in secs]):
        if sum(int(axis_dir[i]) * int(secs[i][cell[i]]) for i in range(ndim)) != 0:
            continue
        sel = [np.nonzero(np.asarray(qnums[i]) == secs[i][cell[i]])[0] for i in range(ndim)]
        shape = tuple(len(s) for s in sel)
        n = int(np.prod(shape))
        dense[np.ix_(*sel)] = entries[pos:pos + n].reshape(shape)
        pos += n
    assert pos == len(entries)
    return dense


def flatten_qnums(q0, dir0, q1, dir1, new_dir):
    """Quantum numbers of two fused neighbouring legs (block_sparse_tensor_flatten_axes, block_sparse_tensor.c:972-979)."""
    q0 = np.asarray(q0, dtype=np.int64)
    q1 = np.asarray(q1, dtype=np.int64)
    return (new_dir * (dir0 * q0[:, None] + dir1 * q1[None, :])).reshape(-1)


# ------------------------------------------------------------------------------------------------------------------
# chain operations (reference src/algorithm/chain_ops.c)
# ------------------------------------------------------------------------------------------------------------------

def apply_local_hamiltonian(a, w, l, r):
    """b = L . W . A . R in the reference's contraction order (apply_local_hamiltonian, chain_ops.c:353-390):
    step 1 (:363) s = a . r; step 2 (:367-372) s = w . s over (d_in, Dw'); step 4 (:380-385) b = l . s over (Dl, Dw)."""
    s = np.tensordot(a, r, axes=(2, 0))                       # [Dl, d_in, Dw', Dr']
    s = np.tensordot(w, s, axes=((2, 3), (1, 2)))             # [Dw, d_out, Dl, Dr']
    b = np.tensordot(l, s, axes=((0, 1), (2, 0)))             # [Dl', d_out, Dr']
    return b


def contraction_operator_step_right(a, b, w, r):
    """r_next[Dl, Dw, Dl'] from r[Dr, Dw', Dr'] (contraction_operator_step_right, chain_ops.c:116-165):
    a . r, then w over (d_in, Dw'), then conj(b) over (d_out, Dr')."""
    s = np.tensordot(a, r, axes=(2, 0))                       # [Dl, d, Dw', Dr']
    s = np.tensordot(w, s, axes=((2, 3), (1, 2)))             # [Dw, d', Dl, Dr']
    t = np.tensordot(s, b.conj(), axes=((1, 3), (1, 2)))      # [Dw, Dl, Dl']
    return t.transpose(1, 0, 2)


def contraction_operator_step_left(a, b, w, l):
    """l_next[Dr, Dw', Dr'] from l[Dl, Dw, Dl'] (contraction_operator_step_left, chain_ops.c:196-245):
    l . conj(b), then w over (Dw, d_out), then a over (Dl, d_in)."""
    s = np.tensordot(l, b.conj(), axes=(2, 0))                # [Dl, Dw, d', Dr']
    s = np.tensordot(s, w, axes=((1, 2), (0, 1)))             # [Dl, Dr', d, Dw']
    t = np.tensordot(a, s, axes=((0, 1), (0, 2)))             # [Dr, Dr', Dw']
    return t.transpose(0, 2, 1)


def mpo_merge_tensor_pair(w0, w1):
    """Two-site MPO tensor [Dw, d0 d1, d0 d1, Dw''] (mpo_merge_tensor_pair, src/operator/mpo.c:255-277)."""
    t = np.tensordot(w0, w1, axes=(3, 0)).transpose(0, 1, 3, 2, 4, 5)
    return t.reshape(t.shape[0], t.shape[1] * t.shape[2], t.shape[3] * t.shape[4], t.shape[5])


def mps_merge_tensor_pair(a0, a1):
    """[Dl, d0 d1, Dr] (mps_merge_tensor_pair, src/state/mps.c:1166-1178)."""
    t = np.tensordot(a0, a1, axes=(2, 0))
    return t.reshape(t.shape[0], t.shape[1] * t.shape[2], t.shape[3])


# ------------------------------------------------------------------------------------------------------------------
# truncation (reference src/algorithm/truncation.c)
# ------------------------------------------------------------------------------------------------------------------

def von_neumann_entropy(sigma) -> float:
    """-sum p log p with p = sigma^2 (von_neumann_entropy, truncation.c:13-27)."""
    s = 0.0
    for x in sigma:
        if x > 0:
            p = float(x) * float(x)
            s -= p * np.log(p)
    return s


def retained_bond_indices(sigma, tol: float, relative_thresh: bool, max_vdim: int):
    """(index list ascending, norm_sigma, entropy, tol_eff)  (retained_bond_indices, truncation.c:110-223):
    sort ascending, square, optionally normalise by the sum, running sum from the smallest, zero the sums cut by
    max_vdim, keep indices whose running sum exceeds tol."""
    sigma = np.asarray(sigma, dtype=np.float64)
    n = len(sigma)
    tol_eff = tol
    order = np.argsort(sigma, kind="stable")
    v = sigma[order] ** 2
    sqsum = 0.0
    for x in v:
        sqsum += x
    if sqsum == 0:
        return np.zeros(0, dtype=np.int64), 0.0, 0.0, tol_eff
    if relative_thresh:
        v = v / sqsum
    acc = np.empty(n)
    run = 0.0
    for i in range(n):
        run = v[i] if i == 0 else acc[i - 1] + v[i]
        acc[i] = run
    if max_vdim < n:
        tol_eff = max(tol, acc[n - max_vdim - 1])
        acc[: n - max_vdim] = 0
    accum = np.empty(n)
    accum[order] = acc
    ind = np.nonzero(accum > tol)[0].astype(np.int64)
    if len(ind) == 0:
        return ind, 0.0, 0.0, tol_eff
    retained = sigma[ind]
    norm_sigma = float(np.linalg.norm(retained))
    return ind, norm_sigma, von_neumann_entropy(retained / norm_sigma), tol_eff


# ------------------------------------------------------------------------------------------------------------------
# Krylov (reference src/util/krylov.c)
# ------------------------------------------------------------------------------------------------------------------

def lanczos_iteration(n: int, afunc, vstart, maxiter: int):
    """(alpha, beta, V[numiter, n], numiter)  (lanczos_iteration_d/_z, krylov.c:24-89 / :96-167): three-term recurrence
    without re-orthogonalisation, alpha_j = Re <v_j, w>, breakdown beta_j < 100 n eps ends with numiter = j + 1, the last
    iteration computes alpha only."""
    vstart = np.asarray(vstart)
    dtype = np.result_type(vstart.dtype, np.float64)
    V = np.zeros((maxiter, n), dtype=dtype)
    alpha = np.zeros(maxiter)
    beta = np.zeros(max(maxiter - 1, 0))
    V[0] = vstart / np.linalg.norm(vstart)
    for j in range(maxiter - 1):
        w = afunc(V[j])
        alpha[j] = np.vdot(V[j], w).real
        w = w - alpha[j] * V[j] - (beta[j - 1] * V[j - 1] if j > 0 else 0)
        beta[j] = np.linalg.norm(w)
        if beta[j] < 100 * n * np.finfo(np.float64).eps:
            return alpha[: j + 1], beta[:j], V[: j + 1], j + 1
        V[j + 1] = w / beta[j]
    j = maxiter - 1
    w = afunc(V[j])
    alpha[j] = np.vdot(V[j], w).real
    return alpha, beta, V, maxiter


def eigensystem_krylov(n: int, afunc, vstart, maxiter: int, numeig: int):
    """(lambda[numeig], u_ritz[n, numeig])  (eigensystem_krylov_symmetric/_hermitian, krylov.c:172-251 / :258-345):
    Lanczos, eigen-decomposition of the tridiagonal matrix, Ritz vectors V^T U[:, :numeig] (not re-normalised)."""
    alpha, beta, V, numiter = lanczos_iteration(n, afunc, vstart, maxiter)
    if numiter < numeig:
        raise RuntimeError("Lanczos breakdown before 'numeig' iterations")
    T = np.diag(alpha[:numiter]) + np.diag(beta[: numiter - 1], 1) + np.diag(beta[: numiter - 1], -1)
    lam, U = np.linalg.eigh(T)
    return lam[:numeig], V[:numiter].T @ U[:, :numeig]


# ------------------------------------------------------------------------------------------------------------------
# block-wise factorisations of a matrix with row / column quantum numbers
# (reference src/tensor/block_sparse_tensor.c: _qr :2402, _rq :2544, _svd :2686)
# ------------------------------------------------------------------------------------------------------------------

def _matrix_sectors(qrow, qcol, by_rows: bool):
    """Charges q that own a stored block (rows with qrow == q and columns with qcol == q), ascending in q.
    The matrix has directions (OUT, IN): conservation reads qrow - qcol = 0."""
    qs = np.unique(np.asarray(qrow if by_rows else qcol, dtype=np.int64))
    out = []
    for q in qs:
        rows = np.nonzero(np.asarray(qrow) == q)[0]
        cols = np.nonzero(np.asarray(qcol) == q)[0]
        if len(rows) > 0 and len(cols) > 0:
            out.append((int(q), rows, cols))
    return out


def block_svd(mat, qrow, qcol):
    """(u, s, vh, qbond): per-sector economy SVD; the new bond lists, for column sectors ascending, min(m, n) entries
    each with singular values descending (block_sparse_tensor_svd, block_sparse_tensor.c:2694-2724, :2824-2837)."""
    secs = _matrix_sectors(qrow, qcol, by_rows=False)
    k_tot = sum(min(len(r), len(c)) for _, r, c in secs)
    u = np.zeros((mat.shape[0], k_tot), dtype=mat.dtype)
    vh = np.zeros((k_tot, mat.shape[1]), dtype=mat.dtype)
    s = np.zeros(k_tot)
    qbond = np.zeros(k_tot, dtype=np.int64)
    pos = 0
    for q, rows, cols in secs:
        ub, sb, vb = np.linalg.svd(mat[np.ix_(rows, cols)], full_matrices=False)
        k = len(sb)
        u[np.ix_(rows, np.arange(pos, pos + k))] = ub
        vh[np.ix_(np.arange(pos, pos + k), cols)] = vb
        s[pos:pos + k] = sb
        qbond[pos:pos + k] = q
        pos += k
    return u, s, vh, qbond


def block_rq(mat, qrow, qcol):
    """(r, q, qbond): per-sector reduced RQ, new bond ordered by row sectors ascending
    (block_sparse_tensor_rq, block_sparse_tensor.c:2552-2581)."""
    secs = _matrix_sectors(qrow, qcol, by_rows=True)
    k_tot = sum(min(len(r), len(c)) for _, r, c in secs)
    rr = np.zeros((mat.shape[0], k_tot), dtype=mat.dtype)
    qq = np.zeros((k_tot, mat.shape[1]), dtype=mat.dtype)
    qbond = np.zeros(k_tot, dtype=np.int64)
    pos = 0
    for q, rows, cols in secs:
        blk = mat[np.ix_(rows, cols)]
        # A = R Q  <=>  A^H = Q^H R^H: QR of the adjoint
        qh, rh = np.linalg.qr(blk.conj().T, mode="reduced")
        k = qh.shape[1]
        rr[np.ix_(rows, np.arange(pos, pos + k))] = rh.conj().T
        qq[np.ix_(np.arange(pos, pos + k), cols)] = qh.conj().T
        qbond[pos:pos + k] = q
        pos += k
    return rr, qq, qbond


def block_qr(mat, qrow, qcol):
    """(q, r, qbond): per-sector reduced QR, new bond ordered by column sectors ascending
    (block_sparse_tensor_qr, block_sparse_tensor.c:2410-2439)."""
    secs = _matrix_sectors(qrow, qcol, by_rows=False)
    k_tot = sum(min(len(r), len(c)) for _, r, c in secs)
    qq = np.zeros((mat.shape[0], k_tot), dtype=mat.dtype)
    rr = np.zeros((k_tot, mat.shape[1]), dtype=mat.dtype)
    qbond = np.zeros(k_tot, dtype=np.int64)
    pos = 0
    for q, rows, cols in secs:
        qb, rb = np.linalg.qr(mat[np.ix_(rows, cols)], mode="reduced")
        k = qb.shape[1]
        qq[np.ix_(rows, np.arange(pos, pos + k))] = qb
        rr[np.ix_(np.arange(pos, pos + k), cols)] = rb
        qbond[pos:pos + k] = q
        pos += k
    return qq, rr, qbond


def split_matrix_svd(mat, qrow, qcol, tol, relative_thresh, max_vdim, renormalize, distr_right: bool):
    """(a0, a1, qbond, info) (split_block_sparse_matrix_svd, src/algorithm/bond_ops.c:15-138)."""
    u, s, vh, qbond = block_svd(mat, qrow, qcol)
    ind, norm_sigma, entropy, tol_eff = retained_bond_indices(s, tol, relative_thresh, max_vdim)
    info = {"norm_sigma": norm_sigma, "entropy": entropy, "tol_eff": tol_eff}
    if len(ind) == 0:
        # dummy bond of dimension 1 (bond_ops.c:50-84)
        q0 = int(np.asarray(qcol)[0]) if len(qbond) == 0 else int(qbond[0])
        a0 = np.zeros((mat.shape[0], 1), dtype=mat.dtype)
        a1 = np.zeros((1, mat.shape[1]), dtype=mat.dtype)
        return a0, a1, np.array([q0], dtype=np.int64), info
    sr = s[ind]
    if renormalize:
        sr = sr * (np.linalg.norm(s) / norm_sigma)
    u, vh, qbond = u[:, ind], vh[ind, :], qbond[ind]
    if distr_right:
        return u, sr[:, None] * vh, qbond, info
    return u * sr[None, :], vh, qbond, info


# ------------------------------------------------------------------------------------------------------------------
# MPS pieces and the DMRG sweeps (reference src/state/mps.c, src/algorithm/dmrg.c)
# ------------------------------------------------------------------------------------------------------------------

def mps_local_orthonormalize_rq(a, qbonds_lr, qsite, a_prev):
    """Right-orthonormalise a[Dl, d, Dr] and absorb R into a_prev (mps_local_orthonormalize_rq, mps.c:560-602).
    Returns (a_new, a_prev_new, new left bond quantum numbers)."""
    ql, qr = qbonds_lr
    dl, d, dr = a.shape
    qcol = flatten_qnums(qsite, OUT, qr, IN, IN)          # fused (d, Dr) leg with direction IN: -(q_s - q_r)
    r, q, qb = block_rq(a.reshape(dl, d * dr), ql, qcol)
    return q.reshape(len(qb), d, dr), np.tensordot(a_prev, r, axes=(2, 0)), qb


def mps_local_orthonormalize_qr(a, qbonds_lr, qsite, a_next):
    """Left-orthonormalise a and absorb R into a_next (mps_local_orthonormalize_qr, mps.c:513-555)."""
    ql, qr = qbonds_lr
    dl, d, dr = a.shape
    qrow = flatten_qnums(ql, OUT, qsite, OUT, OUT)
    q, r, qb = block_qr(a.reshape(dl * d, dr), qrow, qr)
    return q.reshape(dl, d, len(qb)), np.tensordot(r, a_next, axes=(1, 0)), qb


def mps_split_tensor_svd(a, ql, qr, qsite0, qsite1, tol, max_vdim, distr_right: bool):
    """Split a[Dl, d0 d1, Dr] into two site tensors (mps_split_tensor_svd, mps.c:1119-1159; relative threshold)."""
    dl, dd, dr = a.shape
    d0, d1 = len(qsite0), len(qsite1)
    qrow = flatten_qnums(ql, OUT, qsite0, OUT, OUT)
    qcol = flatten_qnums(qsite1, OUT, qr, IN, IN)
    m0, m1, qb, info = split_matrix_svd(a.reshape(dl * d0, d1 * dr), qrow, qcol, tol, True, max_vdim, False, distr_right)
    return m0.reshape(dl, d0, len(qb)), m1.reshape(len(qb), d1, dr), qb, info


def _minimize_local_energy(w, l, r, a_start, mask, maxiter):
    """Lanczos ground state of the local effective Hamiltonian (minimize_local_energy, dmrg.c:86-149)."""
    n = int(mask.sum())      # length of the packed vector of stored entries (dmrg.c:93)
    shape = a_start.shape

    def unpack(v):
        t = np.zeros(shape, dtype=v.dtype)
        t[mask] = v
        return t

    def afunc(v):
        return apply_local_hamiltonian(unpack(v), w, l, r)[mask]

    lam, u = eigensystem_krylov(n, afunc, a_start[mask], maxiter, 1)
    return float(lam[0]), unpack(u[:, 0])


def _right_environments(A, W):
    L = len(A)
    R = [None] * L
    R[L - 1] = np.ones((1, 1, 1), dtype=A[0].dtype)
    for i in range(L - 1, 0, -1):
        R[i - 1] = contraction_operator_step_right(A[i], A[i], W[i], R[i])
    return R


def _orthonormalize_right(A, qbonds, qsite):
    """mps_orthonormalize_qr(RIGHT) (mps.c:609-757); returns the norm, sign folded into the first tensor."""
    L = len(A)
    for i in range(L - 1, 0, -1):
        A[i], A[i - 1], qbonds[i] = mps_local_orthonormalize_rq(A[i], (qbonds[i], qbonds[i + 1]), qsite, A[i - 1])
    head = np.ones((1, 1, 1), dtype=A[0].dtype)
    A[0], head, qbonds[0] = mps_local_orthonormalize_rq(A[0], (qbonds[0], qbonds[1]), qsite, head)
    nrm = head.reshape(-1)[0]
    if nrm.real < 0:
        A[0] = -A[0]
        nrm = -nrm
    return float(np.real(nrm))


def dmrg_twosite(W, qsite, A, qbonds, num_sweeps, maxiter_lanczos, tol_split, max_vdim):
    """Two-site DMRG (dmrg_twosite, src/algorithm/dmrg.c:262-399).  W, A: lists of dense site tensors; qbonds: list of
    nsites+1 bond quantum-number arrays of the MPS (updated in place).  Returns (en_sweeps, entropy)."""
    L = len(A)
    A = list(A)
    qsite = np.asarray(qsite, dtype=np.int64)
    _orthonormalize_right(A, qbonds, qsite)
    R = _right_environments(A, W)
    Lb = [None] * L
    Lb[0] = np.ones((1, 1, 1), dtype=A[0].dtype)
    h2 = [mpo_merge_tensor_pair(W[i], W[i + 1]) for i in range(L - 1)]
    q2 = flatten_qnums(qsite, OUT, qsite, OUT, OUT)
    en_sweeps = np.zeros(num_sweeps)
    entropy = np.zeros(L - 1)
    for n in range(num_sweeps):
        en = 0.0
        for i in range(0, L - 2):
            a = mps_merge_tensor_pair(A[i], A[i + 1])
            mask = conserving_mask([OUT, OUT, IN], [qbonds[i], q2, qbonds[i + 2]])
            en, a = _minimize_local_energy(h2[i], Lb[i], R[i + 1], a, mask, maxiter_lanczos)
            A[i], A[i + 1], qbonds[i + 1], _ = mps_split_tensor_svd(a, qbonds[i], qbonds[i + 2], qsite, qsite, tol_split, max_vdim, True)
            Lb[i + 1] = contraction_operator_step_left(A[i], A[i], W[i], Lb[i])
        for i in range(L - 2, -1, -1):
            a = mps_merge_tensor_pair(A[i], A[i + 1])
            mask = conserving_mask([OUT, OUT, IN], [qbonds[i], q2, qbonds[i + 2]])
            en, a = _minimize_local_energy(h2[i], Lb[i], R[i + 1], a, mask, maxiter_lanczos)
            A[i], A[i + 1], qbonds[i + 1], info = mps_split_tensor_svd(a, qbonds[i], qbonds[i + 2], qsite, qsite, tol_split, max_vdim, False)
            entropy[i] = info["entropy"]
            R[i] = contraction_operator_step_right(A[i + 1], A[i + 1], W[i + 1], R[i + 1])
        A[0] = A[0] / np.linalg.norm(A[0].reshape(-1))      # dmrg.c:366-378 (RQ against a dummy tensor)
        en_sweeps[n] = en
    return en_sweeps, entropy, A


def dmrg_singlesite(W, qsite, A, qbonds, num_sweeps, maxiter_lanczos):
    """Single-site DMRG (dmrg_singlesite, src/algorithm/dmrg.c:155-258)."""
    L = len(A)
    A = list(A)
    qsite = np.asarray(qsite, dtype=np.int64)
    _orthonormalize_right(A, qbonds, qsite)
    R = _right_environments(A, W)
    Lb = [None] * L
    Lb[0] = np.ones((1, 1, 1), dtype=A[0].dtype)
    en_sweeps = np.zeros(num_sweeps)
    for n in range(num_sweeps):
        en = 0.0
        for i in range(0, L - 1):
            mask = conserving_mask([OUT, OUT, IN], [qbonds[i], qsite, qbonds[i + 1]])
            en, A[i] = _minimize_local_energy(W[i], Lb[i], R[i], A[i], mask, maxiter_lanczos)
            A[i], A[i + 1], qbonds[i + 1] = mps_local_orthonormalize_qr(A[i], (qbonds[i], qbonds[i + 1]), qsite, A[i + 1])
            Lb[i + 1] = contraction_operator_step_left(A[i], A[i], W[i], Lb[i])
        for i in range(L - 1, 0, -1):
            mask = conserving_mask([OUT, OUT, IN], [qbonds[i], qsite, qbonds[i + 1]])
            en, A[i] = _minimize_local_energy(W[i], Lb[i], R[i], A[i], mask, maxiter_lanczos)
            A[i], A[i - 1], qbonds[i] = mps_local_orthonormalize_rq(A[i], (qbonds[i], qbonds[i + 1]), qsite, A[i - 1])
            R[i - 1] = contraction_operator_step_right(A[i], A[i], W[i], R[i])
        A[0] = A[0] / np.linalg.norm(A[0].reshape(-1))
        en_sweeps[n] = en
    return en_sweeps, A


def mps_to_statevector(A) -> np.ndarray:
    v = A[0]
    for t in A[1:]:
        v = np.tensordot(v, t, axes=(v.ndim - 1, 0))
    return v.reshape(-1)
