"""Time the split SVD of a two-site tensor of a bench workload on the GPU, for the knobs of the SVD path:
    python tools/svd_bench.py [workload] -- prints ms per split for the direct path, CTB_SVD_BLOCK_ROWS=8 and CTB_SVD_PRECONDITION=1.
The matrix is the two-site tensor a[Dl, d^2, Dr] of bench.build_operands() regrouped as (Dl d) x (d Dr), as mps_split_tensor_svd does."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from chemtensor_b200 import cabi, workloads  # noqa: E402


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "fh_L32_D1024"
    lib = cabi.CLibrary(bench.CUDA_SO, extensions=True)
    assert lib.ctb_init(-1) == 0
    model, L, params, sector, D, dtype, _ = bench.WORKLOADS[wl]
    a, w, l, r = bench.build_operands(lib, wl)
    qsite = workloads.MODELS[model](L, *params)[2]
    d = len(qsite)
    dims = (C.c_int64 * 2)(d, d)
    q = np.ascontiguousarray(qsite, dtype=np.int32)
    qptr = (C.POINTER(C.c_int32) * 2)(q.ctypes.data_as(C.POINTER(C.c_int32)), q.ctypes.data_as(C.POINTER(C.c_int32)))
    for label, env in (("direct", {}), ("block rows 8", {"CTB_SVD_BLOCK_ROWS": "8"}), ("qr-preconditioned", {"CTB_SVD_PRECONDITION": "1"}),
                       ("qr-preconditioned, block rows 8", {"CTB_SVD_PRECONDITION": "1", "CTB_SVD_BLOCK_ROWS": "8"})):
        for k in ("CTB_SVD_BLOCK_ROWS", "CTB_SVD_PRECONDITION"):
            os.environ.pop(k, None)
        os.environ.update(env)
        ts, sig = [], None
        for rep in range(3):
            a0, a1, info = cabi.BST(lib), cabi.BST(lib), cabi.TruncInfo()
            t0 = time.perf_counter()
            rc = lib.mps_split_tensor_svd(a.ptr, dims, qptr, 0.0, D, False, cabi.SVD_DISTR_RIGHT, a0.ptr, a1.ptr, C.byref(info))
            ts.append(time.perf_counter() - t0)
            assert rc == 0
            sig = info.norm_sigma
        print(f"{wl}: {label:34s} {min(ts) * 1e3:9.2f} ms per split (incl. host<->device copies), norm_sigma {sig:.15g}", flush=True)


if __name__ == "__main__":
    main()
