#!/usr/bin/env python
"""Kernel-quality sweep of the grouped FP64 GEMM: block-diagonal contractions of nb equal sector blocks (m x k) . (k x n)
through ctb_dot_benchmark, per tile class (CTB_GEMM_CLASS), next to cuBLAS DGEMM (torch.matmul / bmm) on the same shapes.
Run on the GPU box; prints one JSON line per case."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chemtensor_b200 import cabi  # noqa: E402


def block_diag_pair(lib, nb, m, n, k, dtype, axr_s, axr_t):
    """s, t two-leg tensors whose stored blocks are nb diagonal sector blocks; contraction over one axis."""
    rng = np.random.default_rng(1)
    qm = np.repeat(np.arange(nb, dtype=np.int32), m)
    qn = np.repeat(np.arange(nb, dtype=np.int32), n)
    qk = np.repeat(np.arange(nb, dtype=np.int32), k)
    # s = [M, K] (TRAILING contracted) or [K, M] (LEADING); t = [K, N] (LEADING) or [N, K] (TRAILING)
    if axr_s == cabi.AXIS_RANGE_TRAILING:
        s = cabi.bst_allocate(lib, dtype, (nb * m, nb * k), [1, -1], [qm, qk])
    else:
        s = cabi.bst_allocate(lib, dtype, (nb * k, nb * m), [-1, 1], [qk, qm])
    if axr_t == cabi.AXIS_RANGE_LEADING:
        t = cabi.bst_allocate(lib, dtype, (nb * k, nb * n), [1, -1], [qk, qn])
    else:
        t = cabi.bst_allocate(lib, dtype, (nb * n, nb * k), [-1, 1], [qn, qk])
    for x in (s, t):
        for _, a in x.blocks():
            a[...] = rng.standard_normal(a.shape)
    return s, t


def main():
    import torch
    lib = cabi.CLibrary(os.path.join(ROOT, "chemtensor_b200", "libchemtensor_b200.so"), extensions=True)
    assert lib.ctb_init(-1) == 0
    cases = [(1, 4096, 4096, 4096), (1, 2048, 2048, 2048), (8, 1024, 1024, 1024), (64, 512, 512, 512), (256, 256, 256, 256),
             (1024, 128, 128, 128), (4096, 64, 64, 64), (8192, 32, 32, 32), (64, 256, 1024, 256), (64, 1024, 256, 256), (64, 96, 1536, 384)]
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        cases = cases[:3]
    layouts = [(cabi.AXIS_RANGE_TRAILING, cabi.AXIS_RANGE_LEADING, "NN")]
    if "--layouts" in sys.argv:
        layouts += [(cabi.AXIS_RANGE_LEADING, cabi.AXIS_RANGE_LEADING, "TN"), (cabi.AXIS_RANGE_TRAILING, cabi.AXIS_RANGE_TRAILING, "NT")]
    for dtype, ncls in ((np.float64, 7), (np.complex128, 3)):
        for (nb, m, n, k) in cases:
            if dtype == np.complex128 and nb * m * k > 2 ** 25:
                continue
            # cuBLAS
            tdt = torch.float64 if dtype == np.float64 else torch.complex128
            a = torch.randn(nb, m, k, dtype=tdt, device="cuda"); b = torch.randn(nb, k, n, dtype=tdt, device="cuda")
            for _ in range(2):
                torch.bmm(a, b)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                torch.bmm(a, b)
            e1.record(); torch.cuda.synchronize()
            fl = (2.0 if dtype == np.float64 else 8.0) * nb * m * n * k
            cublas_tf = fl / (e0.elapsed_time(e1) / 5 * 1e-3) / 1e12
            del a, b
            torch.cuda.empty_cache()
            for (axs, axt, lname) in layouts:
                s, t = block_diag_pair(lib, nb, m, n, k, dtype, axs, axt)
                res = {}
                for cls in [-1] + list(range(ncls)):
                    if cls >= 0:
                        os.environ["CTB_GEMM_CLASS"] = str(cls)
                    else:
                        os.environ.pop("CTB_GEMM_CLASS", None)
                    ms = C.c_double(); flops = C.c_double()
                    rc = lib.ctb_dot_benchmark(s.ptr, axs, t.ptr, axt, 1, 2, 5, 0, C.byref(ms), C.byref(flops))
                    assert rc == 0 and abs(flops.value - fl) < 1e-6 * fl
                    res["auto" if cls < 0 else f"c{cls}"] = round(flops.value / (ms.value * 1e-3) / 1e12, 3)
                os.environ.pop("CTB_GEMM_CLASS", None)
                print(json.dumps({"dtype": np.dtype(dtype).name, "nb": nb, "m": m, "n": n, "k": k, "layout": lname, "cublas_tflops": round(cublas_tf, 3), "ours_tflops": res}), flush=True)
                del s, t


if __name__ == "__main__":
    main()
