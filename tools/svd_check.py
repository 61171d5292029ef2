"""GPU check + timing of the big-block SVD path (csrc/ctbd_svd_bj.cu) through block_sparse_tensor_svd on the C-ABI.

    python tools/svd_check.py [real|complex|both] [sizes "m,n;m,n;..."] [decades] [--old]

For each case: one block-sparse matrix whose sectors give a few blocks of about the given size, graded singular values over
`decades` decades; prints reconstruction error, orthogonality of U / Vh, singular values against numpy and the wall time of the call
(host structs in and out) for the new path and, with --old, the scalar tournament (CTB_SVD_BJ=0)."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from chemtensor_b200 import cabi  # noqa: E402


def graded(rng, dtype, m, n, decades):
    k = min(m, n)
    a = rng.standard_normal((m, k)) + (1j * rng.standard_normal((m, k)) if dtype == np.complex128 else 0)
    b = rng.standard_normal((n, k)) + (1j * rng.standard_normal((n, k)) if dtype == np.complex128 else 0)
    u, _ = np.linalg.qr(a)
    v, _ = np.linalg.qr(b)
    s = 10.0 ** (-decades * np.arange(k) / max(k - 1, 1))
    return ((u * s) @ v.conj().T).astype(dtype), s


def run_case(eng, rng, dtype, shapes, decades, label):
    # block-diagonal matrix: block b carries quantum number b on both axes
    ms = [s[0] for s in shapes]
    ns = [s[1] for s in shapes]
    M, N = sum(ms), sum(ns)
    dense = np.zeros((M, N), dtype=dtype)
    qr_ = np.concatenate([np.full(m, b, dtype=np.int32) for b, m in enumerate(ms)])
    qc_ = np.concatenate([np.full(n, b, dtype=np.int32) for b, n in enumerate(ns)])
    r0 = c0 = 0
    svs = []
    for (m, n) in shapes:
        blk, s = graded(rng, dtype, m, n, decades)
        dense[r0:r0 + m, c0:c0 + n] = blk
        svs.append(s)
        r0 += m
        c0 += n
    a = cabi.bst_from_dense(eng, dense, [1, -1], [qr_, qc_])
    best = None
    for rep in range(2):
        u, vh, s = cabi.BST(eng), cabi.BST(eng), cabi.DenseTensor()
        t0 = time.perf_counter()
        rc = eng.block_sparse_tensor_svd(a.ptr, u.ptr, C.byref(s), vh.ptr)
        dt = time.perf_counter() - t0
        assert rc == 0, rc
        best = dt if best is None else min(best, dt)
        if rep == 0:
            U, Vh = u.to_dense(), vh.to_dense()
            k = U.shape[1]
            sv = np.ctypeslib.as_array(C.cast(s.data, C.POINTER(C.c_double)), shape=(k,)).copy()
        eng.delete_dense_tensor(C.byref(s))
    rec = np.linalg.norm((U * sv) @ Vh - dense) / np.linalg.norm(dense)
    ou = np.abs(U.conj().T @ U - np.eye(k)).max()
    ov = np.abs(Vh @ Vh.conj().T - np.eye(k)).max()
    sv_ref = np.concatenate(svs)
    # singular values come per block, descending
    serr = np.max(np.abs(np.sort(sv)[::-1] - np.sort(sv_ref)[::-1])) / sv_ref.max()
    print(f"{label:10s} {np.dtype(dtype).name:10s} blocks {shapes}  decades {decades}: recon {rec:.2e}  U orth {ou:.2e}  Vh orth {ov:.2e}  "
          f"sigma abs err/max {serr:.2e}  time {best * 1e3:9.2f} ms", flush=True)
    return rec, ou, ov, serr


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "real"
    sizes = sys.argv[2] if len(sys.argv) > 2 else "200,200;300,260;150,400;97,33"
    decades = float(sys.argv[3]) if len(sys.argv) > 3 else 12
    old = "--old" in sys.argv
    shapes = [tuple(int(x) for x in s.split(",")) for s in sizes.split(";")]
    eng = helpers.load("cuda")
    dts = {"real": [np.float64], "complex": [np.complex128], "both": [np.float64, np.complex128]}[which]
    bad = 0
    for dt in dts:
        for label, env in ([("block-jacobi", "1")] + ([("tournament", "0")] if old else [])):
            os.environ["CTB_SVD_BJ"] = env
            rng = np.random.default_rng(7)
            rec, ou, ov, serr = run_case(eng, rng, dt, shapes, decades, label)
            if not (rec < 1e-13 and ou < 1e-12 and ov < 1e-12 and serr < 1e-13):
                bad += 1
    os.environ.pop("CTB_SVD_BJ", None)
    print("FAILED" if bad else "OK")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
