import numpy as np
from scipy.optimize import linear_sum_assignment
"""NumPy prototype behind DESIGN.md's note on the GEMM-driven block-Jacobi SVD stage (round-2 plan): cyclic block one-sided Jacobi on
the rows of a graded matrix, pair Gram matrices diagonalised by a small SVD.  Findings (numbers in DESIGN.md section 6):
  * eigenvector matrices in SORTED order (what an SVD routine returns) make the cyclic scheme stagnate -- the transformation must be
    the one closest to the identity (column assignment maximising the diagonal);
  * with that, a mildly graded spectrum (1e-2) converges quadratically in ~10 sweeps, but a DMRG-like spectrum (16 decades) does
    NOT reach relative orthogonality of the small rows: the Gram matrix squares the condition inside a pair."""
rng = np.random.default_rng(0)


def make(R, C, decay):
    u, _ = np.linalg.qr(rng.standard_normal((R, R))); v, _ = np.linalg.qr(rng.standard_normal((C, R)))
    s = 10.0 ** (-decay * np.arange(R) / R)
    return (u * s) @ v.T, s


def rel_overlap(X):
    G = X @ X.T; d = np.sqrt(np.diag(G)); M = np.abs(G) / np.outer(d, d); np.fill_diagonal(M, 0); return M


def rounds(nb):
    n = nb + (nb & 1); pl = list(range(n))
    for r in range(n - 1):
        pairs = [(pl[i], pl[n - 1 - i]) for i in range(n // 2)]
        yield [(p, q) for p, q in pairs if p < nb and q < nb], []
        pl = [pl[0]] + [pl[-1]] + pl[1:-1]

def fix(U):
    r,c=linear_sum_assignment(-np.abs(U))   # maximise sum |U[i, perm(i)]|
    Up=U[:,c]                                # column c[i] goes to position i
    return Up
for decay in (2,16):
    A,s=make(128,256,decay)
    for label,fixer in (("sorted",lambda U:U),("near-identity",fix)):
        X=A.copy(); w=16; hist=[]
        for sweep in range(14):
            off=0
            for pairs,byes in rounds(8):
                for p,q in pairs:
                    idx=np.r_[p*w:(p+1)*w, q*w:(q+1)*w]
                    XP=X[idx]; G=XP@XP.T
                    U,_,_=np.linalg.svd(G); U=fixer(U)
                    X[idx]=U.T@XP
            hist.append(rel_overlap(X).max())
        print(decay,label,['%.1e'%h for h in hist])
