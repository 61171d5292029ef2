import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CTB_TRACE"] = "1"
import bench
from chemtensor_b200 import cabi
lib = cabi.CLibrary(bench.CUDA_SO, extensions=True)
assert lib.ctb_init(-1) == 0
for wl in sys.argv[1:]:
    a, w, l, r = bench.build_operands(lib, wl)
    for _ in range(3):
        t0 = time.perf_counter(); b = cabi.BST(lib); lib.apply_local_hamiltonian(a.ptr, w.ptr, l.ptr, r.ptr, b.ptr); t1 = time.perf_counter(); del b
        print(wl, "total ms", (t1 - t0) * 1e3, "free ms", (time.perf_counter() - t1) * 1e3, flush=True)
