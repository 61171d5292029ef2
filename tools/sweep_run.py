"""Run two-site DMRG sweeps of a named chain on the GPU engine (optionally on the compiled reference too) and print seconds per sweep,
energies and the per-phase split:   python tools/sweep_run.py fermi_hubbard|xxz L D [sweeps] [lanczos] [--ref] [--single]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from chemtensor_b200 import cabi, workloads  # noqa: E402


def main():
    model, L, D = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    pos = [a for a in sys.argv[4:] if not a.startswith("--")]
    sweeps = int(pos[0]) if len(pos) > 0 else 2
    lanczos = int(pos[1]) if len(pos) > 1 else 10
    params = {"fermi_hubbard": (1.0, 4.0, 0.0), "xxz": (1.0, 0.8, 0.1)}[model]
    sector = workloads.encode_qpair(L, 0) if model == "fermi_hubbard" else 0
    which = bench.REF_SO if "--ref" in sys.argv else bench.CUDA_SO
    lib = cabi.CLibrary(which, extensions="--ref" not in sys.argv)
    if "--ref" not in sys.argv:
        assert lib.ctb_init(-1) == 0
    out = bench.sweep_seconds(lib, model, L, params, sector, D, sweeps=sweeps, lanczos=lanczos)
    print(json.dumps({"model": model, "L": L, "D": D, "sweeps": sweeps, "lanczos": lanczos, "impl": "reference" if "--ref" in sys.argv else "b200", **(out or {"failed": True})}), flush=True)


if __name__ == "__main__":
    main()
