"""HBM roofline of the block linear combination kernel (csrc/ctbd_blocklc.cu): F-move shaped work lists, kernel time alone (CUDA events, L2 flushed)."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = C.CDLL(os.path.join(ROOT, "chemtensor_b200", "libchemtensor_b200.so"))
lib.ctb_su2_lc_benchmark.restype = C.c_int
lib.ctb_su2_lc_benchmark.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
peak = 6549.4
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
for nelem, nblk, nterm, cplx in [(1 << 20, 64, 1, 0), (1 << 20, 32, 2, 0), (1 << 20, 32, 3, 0), (1 << 14, 2048, 2, 0), (256, 1 << 17, 2, 0), (1 << 20, 16, 2, 1)]:
    out = (C.c_double * 3)()
    rc = lib.ctb_su2_lc_benchmark(nelem, nblk, nterm, cplx, out)
    print(json.dumps({"kernel": "lc_kernel", "dtype": "c128" if cplx else "f64", "block_entries": nelem, "blocks": nblk, "terms_per_block": nterm, "rc": rc,
                      "ms": out[0], "GBps": out[1], "algorithmic_bytes": out[2], "peak_GBps": peak, "frac": out[1] / peak}), flush=True)
