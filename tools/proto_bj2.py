"""NumPy prototype of the round-2 block-Jacobi SVD (csrc/ctbd_svd_bj.cu): cyclic one-sided block Jacobi on the ROWS of [G | W],
pair Gram matrix P = X X^T by GEMM, P diagonalised by two-sided cyclic Jacobi ROTATIONS (relative accuracy on graded Gram matrices,
Demmel-Veselic), rows updated by GEMM.  Variants: plain / sorted (de Rijk) / QR-preconditioned.  Prints relative row overlap per sweep.

usage: proto_bj2.py R C b variant[,variant..] [decays]"""
import sys
import numpy as np
import scipy.linalg

rng = np.random.default_rng(0)


def make(R, C, decay):
    u, _ = np.linalg.qr(rng.standard_normal((R, R)))
    v, _ = np.linalg.qr(rng.standard_normal((C, R)))
    s = 10.0 ** (-decay * np.arange(R) / R)
    return (u * s) @ v.T, s


def rel_overlap(X):
    G = X @ X.T
    d = np.sqrt(np.diag(G))
    M = np.abs(G) / np.maximum(np.outer(d, d), 1e-300)
    np.fill_diagonal(M, 0)
    return M


def tournament(N):
    for r in range(N - 1):
        prs = [(N - 1, r)]
        for i in range(1, N // 2):
            prs.append(((r + i) % (N - 1), (r - i + (N - 1)) % (N - 1)))
        yield prs


_TOUR = {}


def tour_arrays(n):
    if n not in _TOUR:
        N = n + (n & 1)
        out = []
        for prs in tournament(N):
            pq = [(min(p, q), max(p, q)) for p, q in prs if p < n and q < n]
            out.append((np.array([p for p, q in pq]), np.array([q for p, q in pq])))
        _TOUR[n] = out
    return _TOUR[n]


INNER_CAP = 30
FLOAT_T = False


def cross_arrays(n):
    """bipartite round robin between the first and second half of the indices: n/2 steps of n/2 disjoint cross pairs"""
    key = ("cross", n)
    if key not in _TOUR:
        h = n // 2
        _TOUR[key] = [(np.arange(h), h + (np.arange(h) + s) % h) for s in range(h)]
    return _TOUR[key]


CROSS_ONLY = False


def jacobi_eig(P, tol, sort=False, max_sweeps=None, cross=False):
    """two-sided cyclic Jacobi (parallel ordering) on symmetric P; returns Q with Q^T P Q ~ diagonal"""
    n = P.shape[0]
    P = P.copy()
    Q = np.eye(n)
    if max_sweeps is None:
        max_sweeps = INNER_CAP
    for sweep in range(max_sweeps):
        nrot = 0
        for ps, qs in (cross_arrays(n) if cross else tour_arrays(n)):
            g = P[ps, qs]
            a = P[ps, ps]
            b = P[qs, qs]
            act = (g != 0) & (np.abs(g) > tol * np.sqrt(np.abs(a * b)))
            swap = np.zeros(len(ps), bool)
            if sort:
                swap = (b > a)      # de Rijk: keep the larger diagonal entry at the smaller index
                act = act | swap
            if not act.any():
                continue
            nrot += int((act & ~swap).sum()) + int(((np.abs(g) > tol * np.sqrt(np.abs(a * b))) & swap).sum())
            zeta = (b - a) / (2 * np.where(g == 0, 1.0, g))
            t = np.where(zeta == 0, 1.0, np.sign(zeta) / (np.abs(zeta) + np.sqrt(1 + zeta * zeta)))
            t = np.where(g == 0, 0.0, t)
            if FLOAT_T:
                t = t.astype(np.float32).astype(np.float64)
            c = 1 / np.sqrt(1 + t * t)
            s = c * t
            c = np.where(act, c, 1.0)
            s = np.where(act, s, 0.0)
            if sort:
                # rotation followed by a swap of p and q when the larger eigenvalue would land at q
                ap = a - t * g
                bq = b + t * g
                sw = act & (bq > ap)
                # new p column = old q-combination: (c,s) -> swap columns of J
                cp = np.where(sw, s, c); sp = np.where(sw, -c, s)      # J[:, p] = [cp; -sp'] ...
                # build J explicitly for clarity (n <= 128)
                J = np.eye(n)
                J[ps, ps] = c; J[qs, qs] = c; J[ps, qs] = s; J[qs, ps] = -s
                for k in np.nonzero(sw)[0]:
                    J[:, [ps[k], qs[k]]] = J[:, [qs[k], ps[k]]]
                P = J.T @ P @ J
                Q = Q @ J
            else:
                # columns
                Pp = P[:, ps].copy(); Pq = P[:, qs].copy()
                P[:, ps] = c * Pp - s * Pq
                P[:, qs] = s * Pp + c * Pq
                Pp = P[ps, :].copy(); Pq = P[qs, :].copy()
                P[ps, :] = c[:, None] * Pp - s[:, None] * Pq
                P[qs, :] = s[:, None] * Pp + c[:, None] * Pq
                Qp = Q[:, ps].copy(); Qq = Q[:, qs].copy()
                Q[:, ps] = c * Qp - s * Qq
                Q[:, qs] = s * Qp + c * Qq
            P = 0.5 * (P + P.T)
        if nrot == 0:
            break
    return Q, sweep + 1


def block_jacobi(A, b, tol, sort=False, max_sweeps=40, verbose=True):
    R, C = A.shape
    X = np.hstack([A, np.eye(R)])
    nb = (R + b - 1) // b
    N = nb + (nb & 1)
    hist = []
    for sweep in range(max_sweeps):
        active = 0
        inner = 0
        for rnd, prs in enumerate(tournament(N) if nb > 1 else [[(0, 1)]]):
            for p, q in prs:
                if p > q:
                    p, q = q, p
                idx = np.r_[p * b:min((p + 1) * b, R), (q * b if q < nb else R):(min((q + 1) * b, R) if q < nb else R)]
                if len(idx) < 2:
                    continue
                XP = X[idx]
                P = XP[:, :C] @ XP[:, :C].T
                d = np.sqrt(np.diag(P))
                M = np.abs(P) / np.maximum(np.outer(d, d), 1e-300)
                np.fill_diagonal(M, 0)
                if M.max() <= tol:
                    continue
                active += 1
                full_pair = (len(idx) == 2 * b)
                Q, ns = jacobi_eig(P, tol, sort=sort, cross=(CROSS_ONLY and rnd > 0 and full_pair))
                inner += ns
                X[idx] = Q.T @ XP
        ro = rel_overlap(X[:, :C]).max()
        hist.append(ro)
        if verbose:
            print(f"  sweep {sweep + 1}: active pairs {active}, mean inner sweeps {inner / max(active, 1):.1f}, max rel overlap {ro:.2e}", flush=True)
        if active == 0:
            break
    return X, hist


def report(A, G, W, left=None):
    R = A.shape[0]
    sig = np.linalg.norm(G, axis=1)
    order = np.argsort(-sig)
    sv = np.linalg.svd(A, compute_uv=False)
    Vh = G / sig[:, None]
    U = W.T if left is None else left @ W.T
    print("  sigma rel err max %.2e" % np.max(np.abs(sig[order] - sv) / sv), " U orth %.2e" % np.abs(U.T @ U - np.eye(R)).max(),
          " Vh orth %.2e" % np.abs(Vh @ Vh.T - np.eye(R)).max(), " recon %.2e" % (np.abs((U * sig) @ Vh - A).max() / np.abs(A).max()))


if __name__ == "__main__":
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    b = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    variants = sys.argv[4].split(",") if len(sys.argv) > 4 else ["plain"]
    decays = [float(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [2, 8, 16]
    if len(sys.argv) > 6:
        INNER_CAP = int(sys.argv[6])
    if len(sys.argv) > 7:
        FLOAT_T = bool(int(sys.argv[7]))
    if len(sys.argv) > 8:
        CROSS_ONLY = bool(int(sys.argv[8]))
    for decay in decays:
        A, s = make(R, C, decay)
        tol = np.finfo(float).eps * np.sqrt(C)
        for var in variants:
            print(f"R={R} C={C} b={b} decay={decay} variant={var}")
            if var == "plain":
                X, hist = block_jacobi(A, b, tol)
                report(A, X[:, :C], X[:, C:])
            elif var == "sorted":
                X, hist = block_jacobi(A, b, tol, sort=True)
                report(A, X[:, :C], X[:, C:])
            elif var in ("qr", "qrs"):
                # A^T = Q R (C x R . R x R);  A = R^T Q^T;  Jacobi on the rows of R: J^T R = S Vh_R  ->  R = J S Vh_R
                # A = R^T Q^T = Vh_R^T S J^T Q^T: left vectors Vh_R^T, right vectors (Q J)^T
                Qf, Rf = np.linalg.qr(A.T)
                X, hist = block_jacobi(Rf, b, tol, sort=(var == "qrs"))
                G, W = X[:, :R], X[:, R:]
                sig = np.linalg.norm(G, axis=1)
                sv = np.linalg.svd(A, compute_uv=False)
                Uo = (G / sig[:, None]).T
                Vho = (Qf @ W.T).T
                print("  sigma rel err max %.2e" % np.max(np.abs(np.sort(sig)[::-1] - sv) / sv), " U orth %.2e" % np.abs(Uo.T @ Uo - np.eye(R)).max(),
                      " Vh orth %.2e" % np.abs(Vho @ Vho.T - np.eye(R)).max(), " recon %.2e" % (np.abs((Uo * sig) @ Vho - A).max() / np.abs(A).max()))
            elif var in ("qrlq", "qrlqqr"):
                # two (three) preconditioning steps: A^T = Q1 R1, R1^T = Q2 R2, (R2^T = Q3 R3); Jacobi on the rows of the last factor
                Q1, R1 = np.linalg.qr(A.T)
                Q2, R2 = np.linalg.qr(R1.T)
                Rl = R2
                if var == "qrlqqr":
                    Q3, Rl = np.linalg.qr(R2.T)
                X, hist = block_jacobi(Rl, b, tol)
            elif var == "qrp":
                Qf, Rf, piv = scipy.linalg.qr(A.T, mode="economic", pivoting=True)
                X, hist = block_jacobi(Rf, b, tol)
            elif var == "lq":
                # A = L Q (R x R . R x C); Jacobi on the rows of L^T?  rows of L: J^T L = S Vh_L
                Qf, Rf = np.linalg.qr(A.T)
                L = Rf.T
                X, hist = block_jacobi(L, b, tol)
