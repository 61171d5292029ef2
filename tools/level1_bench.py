"""HBM-roofline numbers of the HBM-bound kernels of the path (Lanczos level-1 work and the re-blocking / permutation
kernels are reported against the measured copy bandwidth, not the tensor pipe):

    python tools/level1_bench.py [n_doubles]       (default: the Lanczos vector of the D=4096 Fermi-Hubbard centre bond, 17.1 M doubles)

Times every level-1 kernel of csrc/ctbd_level1.cu through the thin C-ABI (ctbd_*, device pointers, CUDA events on the layer's stream,
192 MiB written between repetitions so that nothing is served from L2) and the remap kernel on the real two-site tensor (transpose of
a[Dl, dd, Dr] and the flatten / split pair of mps_split_tensor_svd).  Prints one JSON line per kernel: algorithmic bytes, time, GB/s and
the fraction of MEASURED_PEAKS.json's copy bandwidth."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from chemtensor_b200 import cabi  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 17119384
    lib = cabi.CLibrary(bench.CUDA_SO, extensions=True)
    assert lib.ctb_init(-1) == 0
    d = lib.dll
    vp = C.c_void_p
    for name, args in (("ctbd_malloc", [C.POINTER(vp), C.c_size_t]), ("ctbd_free", [vp]), ("ctbd_memset_zero", [vp, C.c_size_t]), ("ctbd_h2d", [vp, vp, C.c_size_t]),
                       ("ctbd_event_create", [C.POINTER(vp)]), ("ctbd_event_record", [vp]), ("ctbd_event_elapsed_ms", [vp, vp, C.POINTER(C.c_float)]),
                       ("ctbd_dotc", [C.c_int, C.c_int64, vp, vp, vp]), ("ctbd_nrm2", [C.c_int, C.c_int64, vp, vp]), ("ctbd_rscale", [C.c_int, C.c_int64, vp, vp, C.c_int, vp]),
                       ("ctbd_lanczos_update", [C.c_int, C.c_int64, vp, vp, vp, vp, vp, vp]), ("ctbd_lincomb", [C.c_int, C.c_int64, vp, C.c_int64, C.c_int, vp, vp])):
        getattr(d, name).restype = C.c_int
        getattr(d, name).argtypes = args
    peaks, kind = bench.load_measured_peaks()
    peak = float(peaks["hbm_gbs"])

    def dmalloc(nbytes):
        p = vp()
        assert d.ctbd_malloc(C.byref(p), nbytes) == 0
        return p

    m = 10
    V = dmalloc(m * n * 8)
    w = dmalloc(n * 8)
    out = dmalloc(n * 8)
    scal = dmalloc(64 * 8)
    flush = dmalloc(192 << 20)
    host = np.random.default_rng(0).standard_normal(n)
    for j in range(m):
        d.ctbd_h2d(C.c_void_p(V.value + j * n * 8), host.ctypes.data, n * 8)
    d.ctbd_h2d(w, host.ctypes.data, n * 8)
    one = np.array([1.0, 0.5, 0.25, 2.0])
    d.ctbd_h2d(scal, one.ctypes.data, 32)
    e0, e1 = vp(), vp()
    d.ctbd_event_create(C.byref(e0)); d.ctbd_event_create(C.byref(e1))
    coef = np.ones(m)

    def vj(j):
        return C.c_void_p(V.value + j * n * 8)

    def sc(i):
        return C.c_void_p(scal.value + 8 * i)

    cases = [
        ("dotc (alpha_j = <w, v_j>)", 2 * n * 8, lambda: d.ctbd_dotc(1, n, w, vj(0), sc(8))),
        ("nrm2", n * 8, lambda: d.ctbd_nrm2(1, n, w, sc(8))),
        ("rscale (v_{j+1} = w / beta_j)", 2 * n * 8, lambda: d.ctbd_rscale(1, n, w, sc(3), 1, vj(1))),
        ("lanczos_update (w -= alpha v_j + beta v_{j-1}; ||w||)", 4 * n * 8, lambda: d.ctbd_lanczos_update(1, n, w, vj(0), vj(1), sc(1), sc(2), sc(9))),
        (f"lincomb (Ritz vector from {m} Krylov vectors)", (m + 1) * n * 8, lambda: d.ctbd_lincomb(1, n, V, n, m, coef.ctypes.data, out)),
    ]
    for name, nbytes, fn in cases:
        ts = []
        for rep in range(6):
            d.ctbd_memset_zero(flush, 192 << 20)
            d.ctbd_event_record(e0)
            assert fn() == 0
            d.ctbd_event_record(e1)
            ms = C.c_float(0)
            d.ctbd_event_elapsed_ms(e0, e1, C.byref(ms))
            ts.append(ms.value)
        t = min(ts[1:])
        print(json.dumps({"kernel": name, "n": n, "algorithmic_bytes": nbytes, "ms": t, "GB_per_s": nbytes / t / 1e6, "frac_of_measured_copy_peak": nbytes / t / 1e6 / peak, "peak_GB_per_s": peak, "peak_source": kind}), flush=True)

    # the re-blocking kernel on the two-site tensor of the bench workload: transpose [0,2,1] and the flatten pair of the split
    a, wt, l, r = bench.build_operands(lib, "fh_L64_D4096")
    nel = a.num_elements()
    import time
    for name, call in (("block_sparse_tensor_transpose of a[Dl, dd, Dr] (host structs in and out)", None),):
        pass
    if lib.has("ctb_remap_benchmark"):
        res = (C.c_double * 6)()
        if lib.ctb_remap_benchmark(a.ptr, res) == 0:
            for k, label in enumerate(("transpose [2, 1, 0]", "flatten axes (0, 1)", "flatten axes (1, 2)")):
                t = res[2 * k]; nbytes = res[2 * k + 1]
                print(json.dumps({"kernel": f"remap_kernel: {label} of the two-site tensor", "stored_entries": int(nel), "algorithmic_bytes": nbytes, "ms": t, "GB_per_s": nbytes / t / 1e6,
                                  "frac_of_measured_copy_peak": nbytes / t / 1e6 / peak, "peak_GB_per_s": peak, "peak_source": kind}), flush=True)


if __name__ == "__main__":
    main()
