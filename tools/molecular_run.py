"""BASELINE.json configs[3] at growing size: complex128 two-site DMRG on a synthetic molecular Hamiltonian with n spatial orbitals (d = 4,
MPO bond 2 n^2 + 3 n + 2), bond dimension D, through the public dmrg_twosite() of the engine:

    python tools/molecular_run.py n D [sweeps] [lanczos] [--merged]

Prints one JSON line: seconds per sweep, energies, MPO / MPS bond dimensions, the per-phase split and the device memory high-water mark.
The MPO comes from the reference's generator (oracle/_ref, input generator only); the engine chooses the pair form of the effective
Hamiltonian by itself once the merged pair tensor would exceed 2^28 dense entries (--merged forces the merged form)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from chemtensor_b200 import cabi  # noqa: E402


def main():
    n, D = int(sys.argv[1]), int(sys.argv[2])
    pos = [a for a in sys.argv[3:] if not a.startswith("--")]
    sweeps = int(pos[0]) if len(pos) > 0 else 1
    lanczos = int(pos[1]) if len(pos) > 1 else 10
    lib = cabi.CLibrary(bench.CUDA_SO, extensions=True)
    assert lib.ctb_init(-1) == 0
    out = bench.molecular_sweep_seconds(lib, n, D, False if "--merged" in sys.argv else None, sweeps=sweeps, lanczos=lanczos)
    rec = {"config": f"molecular_c128_n{n}_D{D}", "orbitals": n, "D": D, "sweeps": sweeps, "lanczos": lanczos, "dtype": "c128", **(out or {"failed": True})}
    ph = bench.phases(lib)
    if ph is not None:
        rec["phases_s"] = ph
    print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()
