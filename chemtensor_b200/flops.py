"""Algorithmic flop count of the hot path from sector metadata alone (integer exact).

F_alg = sum over the contractions of the path, over output blocks, over contracted sector tuples allowed by
conservation, of 2 m n k (x4 for complex128) -- exactly the dense_tensor_dot_update calls the reference issues at
src/tensor/block_sparse_tensor.c:1953-1994.  Used by bench.py for the arm that runs without the
engine, and cross-checked in the tests against the engine's own plan-time count.
"""
from __future__ import annotations

import itertools
from collections import defaultdict

import numpy as np


class SectorTensor:
    """Block structure of a block-sparse tensor: per axis a direction and {quantum number: multiplicity}."""

    def __init__(self, dirs, sectors):
        self.dirs = list(dirs)
        self.sectors = [dict(s) for s in sectors]

    @classmethod
    def from_bst(cls, t):
        secs = []
        for q in t.qnums:
            vals, counts = np.unique(np.asarray(q), return_counts=True)
            secs.append({int(v): int(c) for v, c in zip(vals, counts)})
        return cls(t.axis_dir, secs)

    @property
    def ndim(self):
        return len(self.dirs)

    def blocks(self):
        """Yield tuples of quantum numbers of the conserving blocks."""
        keys = [sorted(s) for s in self.sectors]
        if self.ndim == 0:
            return
        # enumerate all but the last axis, solve for the last
        last = self.ndim - 1
        for combo in itertools.product(*keys[:last]):
            tot = sum(d * q for d, q in zip(self.dirs[:last], combo))
            qlast = -tot * self.dirs[last]      # dir_last * q_last = -tot, dir = +-1
            if qlast in self.sectors[last]:
                yield combo + (qlast,)

    def transpose(self, perm):
        return SectorTensor([self.dirs[p] for p in perm], [self.sectors[p] for p in perm])


def dot_flops(s: SectorTensor, s_leading: bool, t: SectorTensor, t_leading: bool, ndim_mult: int):
    """(flops with factor 2, result structure) of block_sparse_tensor_dot(s, axrange_s, t, axrange_t, ndim_mult)."""
    cs = list(range(ndim_mult)) if s_leading else list(range(s.ndim - ndim_mult, s.ndim))
    ct = list(range(ndim_mult)) if t_leading else list(range(t.ndim - ndim_mult, t.ndim))
    fs = [i for i in range(s.ndim) if i not in cs]
    ft = [i for i in range(t.ndim) if i not in ct]
    msum = defaultdict(int)
    for b in s.blocks():
        kap = tuple(b[i] for i in cs)
        msum[kap] += int(np.prod([s.sectors[i][b[i]] for i in fs], dtype=object)) if fs else 1
    nsum = defaultdict(int)
    for b in t.blocks():
        kap = tuple(b[i] for i in ct)
        nsum[kap] += int(np.prod([t.sectors[i][b[i]] for i in ft], dtype=object)) if ft else 1
    flops = 0
    for kap, m in msum.items():
        n = nsum.get(kap, 0)
        if n:
            k = 1
            for i, q in zip(cs, kap):
                k *= s.sectors[i][q]
            flops += 2 * m * n * k
    res = SectorTensor([s.dirs[i] for i in fs] + [t.dirs[i] for i in ft], [s.sectors[i] for i in fs] + [t.sectors[i] for i in ft])
    return flops, res


def heff_flops(a, w, l, r) -> float:
    """One apply_local_hamiltonian (reference chain_ops.c:353-390): a.r, w.(ar), l.(war)."""
    A, W, Lt, R = (SectorTensor.from_bst(x) for x in (a, w, l, r))
    f1, s = dot_flops(A, False, R, True, 1)                 # [Dl, dd, Dw', Dr', x]
    t = s.transpose([1, 2, 0, 3, 4])
    f2, s = dot_flops(W, False, t, True, 2)                 # [Dw, dd, Dl, Dr', x]
    t = s.transpose([2, 0, 1, 3, 4])
    k = Lt.transpose([0, 3, 1, 2])
    f3, _ = dot_flops(k, False, t, True, 2)
    total = f1 + f2 + f3
    if np.dtype(a.dtype).kind == "c":
        total *= 4
    return float(total)
