#!/usr/bin/env python
"""Run the engine's two-site DMRG on a named Hamiltonian from a seeded random MPS and record, per virtual bond, the
sector histogram {quantum number: multiplicity} the sweep produces, plus timing statistics.

The histograms are committed under chemtensor_b200/data/ and are what bench.py's synthetic Heff operands are built
from ("(a, w, l, r) ... built synthetically with the measured sector histogram and randn entries").
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from chemtensor_b200 import cabi, workloads  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="fermi_hubbard")
    ap.add_argument("--nsites", type=int, default=16)
    ap.add_argument("--max-vdim", type=int, default=256)
    ap.add_argument("--sweeps", type=int, default=1)
    ap.add_argument("--lanczos", type=int, default=10)
    ap.add_argument("--tol", type=float, default=0.0)
    ap.add_argument("--out", default=None)
    ap.add_argument("--lib", default=os.path.join(ROOT, "chemtensor_b200", "libchemtensor_b200.so"))
    args = ap.parse_args()

    lib = cabi.CLibrary(args.lib, extensions=True)
    assert lib.ctb_init(-1) == 0
    if args.model == "xxz":
        params, sector = (1.0, 0.8, 0.1), 0
    else:
        params, sector = (1.0, 4.0, 0.0), workloads.encode_qpair(args.nsites, 0)
    mpo = workloads.mpo_chain(lib, args.model, args.nsites, params)
    psi = workloads.random_mps(lib, np.float64, args.nsites, mpo.qsite, sector, args.max_vdim, seed=42)
    en = np.zeros(args.sweeps); ent = np.zeros(args.nsites - 1)
    t0 = time.perf_counter()
    rc = lib.dmrg_twosite(mpo.ptr, args.sweeps, args.lanczos, args.tol, args.max_vdim, psi.ptr,
                          en.ctypes.data_as(C.POINTER(C.c_double)), ent.ctypes.data_as(C.POINTER(C.c_double)))
    dt = time.perf_counter() - t0
    assert rc == 0
    st = (C.c_double * 9)()
    lib.ctb_get_stats(st, 9)
    stats = dict(zip(["heff_flops", "heff_calls", "env_flops", "lanczos_ms", "svd_ms", "env_ms", "total_ms", "max_vector_len", "max_bond_dim"], list(st)))
    bonds = []
    for i in range(args.nsites):
        q = psi.site(i).qnums[0]
        v, c = np.unique(q, return_counts=True)
        bonds.append({"bond": i, "dim": int(len(q)), "sectors": {str(int(a)): int(b) for a, b in zip(v, c)}})
    res = {"model": args.model, "nsites": args.nsites, "params": params, "sector": sector, "max_vdim": args.max_vdim, "sweeps": args.sweeps,
           "lanczos": args.lanczos, "tol_split": args.tol, "energies": list(en), "wall_s": dt, "stats": stats, "bonds": bonds}
    c = bonds[args.nsites // 2]
    print(json.dumps({k: res[k] for k in ("model", "nsites", "max_vdim", "energies", "wall_s", "stats")}))
    print("centre bond: dim", c["dim"], "nsectors", len(c["sectors"]), "largest", sorted(c["sectors"].values())[-6:])
    if args.out:
        with open(args.out, "w") as f:
            json.dump(res, f)


if __name__ == "__main__":
    main()
