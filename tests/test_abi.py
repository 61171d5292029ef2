"""The C-ABI boundary: libchemtensor_b200.so loads (all symbols resolved, RTLD_NOW) and exports every function declared in
include/chemtensor_b200.h and include/ctb_device.h; the CPU test double exports the same ctbd_* layer.  No compute calls."""
import ctypes as C
import os
import re

import helpers

INCLUDE = os.path.join(helpers.ROOT, "include")


def declared_functions(header):
    src = open(os.path.join(INCLUDE, header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set()
    for m in re.finditer(r"^\s*(?:const\s+)?[A-Za-z_][A-Za-z0-9_ \*]*?\b([a-z_][a-z0-9_]*)\s*\([^;{]*\)\s*;", src, flags=re.M):
        if "typedef" in m.group(0) or "(*" in m.group(0).split("(")[0]:
            continue
        names.add(m.group(1))
    return names - {"defined", "sizeof"}


def test_product_library_loads_and_exports_the_declared_abi():
    helpers._make("lib")
    lib = C.CDLL(helpers.CUDA_SO, mode=os.RTLD_NOW)      # every undefined symbol must resolve at load time
    api = declared_functions("chemtensor_b200.h") | declared_functions("ctb_su2.h")
    dev = declared_functions("ctb_device.h")
    assert len(api) > 50 and len(dev) > 40
    missing = sorted(n for n in api | dev if not hasattr(lib, n))
    assert not missing, f"libchemtensor_b200.so lacks: {missing}"
    assert lib.ctbd_backend() == 1


def test_test_double_exports_the_device_layer():
    helpers._make("emu")
    lib = C.CDLL(helpers.EMU_SO, mode=os.RTLD_NOW)
    missing = sorted(n for n in declared_functions("ctb_device.h") | declared_functions("chemtensor_b200.h") | declared_functions("ctb_su2.h") if not hasattr(lib, n))
    assert not missing, f"test double lacks: {missing}"
    assert lib.ctbd_backend() == 2


def test_product_library_does_not_contain_the_test_double_or_an_oracle():
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", helpers.CUDA_SO], capture_output=True, text=True).stdout
    assert "cblas_" not in out and "dgesvd" not in out
    deps = subprocess.run(["ldd", helpers.CUDA_SO], capture_output=True, text=True).stdout
    for forbidden in ("openblas", "chemtensor_ref", "hostlogic_emu", "lapack"):
        assert forbidden not in deps
