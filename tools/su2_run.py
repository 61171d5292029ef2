"""SU(2)-symmetric Heisenberg chain (BASELINE configs[4]): two-site / single-site DMRG sweeps on the engine, the reference timed beside it.

usage: python tools/su2_run.py <engine: cuda|emu> <L> <max_vdim> [--sweeps N] [--lanczos K] [--degen D] [--ref] [--single] [--out file.json]

Inputs come from the reference's generators in oracle/_ref (construct_heisenberg_1d_su2_mpo, construct_random_su2_mps with bond irreducible
representations up to 2j = 5 and 'degen' multiplets each); max_vdim is the LOGICAL bond dimension (multiplet dimensions included), as in the reference.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import su2_helpers as S  # noqa: E402


def bond_summary(psi):
    out = []
    for i in range(psi.nsites):
        t = psi.a[i]
        js = [t.outer_irreps[2].jlist[k] for k in range(t.outer_irreps[2].num)]
        out.append({"j2": js, "multiplets": [int(t.dim_degen[2][j]) for j in js]})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("engine")
    ap.add_argument("L", type=int)
    ap.add_argument("max_vdim", type=int)
    ap.add_argument("--sweeps", type=int, default=2)
    ap.add_argument("--lanczos", type=int, default=10)
    ap.add_argument("--degen", type=int, default=8)
    ap.add_argument("--tol", type=float, default=0.0)
    ap.add_argument("--max-irrep", type=int, default=5)
    ap.add_argument("--ref", action="store_true")
    ap.add_argument("--single", action="store_true")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()

    r = S.ref()
    e = S.engine(a.engine)
    e.ctb_su2_get_stats.restype = None
    e.ctb_su2_get_stats.argtypes = [C.POINTER(C.c_double)]
    mpo = S.heisenberg_mpo(a.L, 1.0)
    psi = S.random_mps(a.L, [1], [0, 1], a.L % 2, a.max_irrep, a.degen, 42, scale=1.0)
    res = {"config": {"model": "heisenberg_su2", "L": a.L, "max_vdim_logical": a.max_vdim, "sweeps": a.sweeps, "lanczos": a.lanczos,
                      "start_degen": a.degen, "tol_split": a.tol, "single_site": a.single}, "engine": a.engine}
    runs = [("engine", e)] + ([("reference", r)] if a.ref else [])
    for name, lib in runs:
        p = S.copy_mps(psi)
        en = (C.c_double * a.sweeps)()
        ent = (C.c_double * max(a.L - 1, 1))()
        per_sweep = []
        t0 = time.perf_counter()
        # ONE call for all sweeps, as a user makes it (orthonormalisation and the right environments are built once)
        if a.single:
            rc = lib.su2_dmrg_singlesite(C.byref(mpo), a.sweeps, a.lanczos, C.byref(p), en)
        else:
            rc = lib.su2_dmrg_twosite(C.byref(mpo), a.sweeps, a.lanczos, a.tol, a.max_vdim, C.byref(p), en, ent)
        wall = time.perf_counter() - t0
        rec = {"rc": rc, "wall_s": wall, "s_per_sweep": wall / a.sweeps, "energies": list(en)}
        if name == "engine":
            st = (C.c_double * 16)()
            e.ctb_su2_get_stats(st)
            nsw = int(st[6])
            rec["sweep_s"] = [st[7 + k] for k in range(min(nsw, 9))]
            rec["sweeps_completed"] = nsw
            rec["stats"] = {"launches": st[0], "heff_applications": st[1], "local_solve_s": st[2], "split_qr_s": st[3], "environment_s": st[4],
                            "orthonormalise_and_right_environments_s": st[5]}
        b = bond_summary(p)
        rec["centre_bond"] = b[a.L // 2 - 1]
        rec["max_multiplets"] = max(sum(x["multiplets"]) for x in b)
        rec["max_logical_bond"] = max(sum(m * (j + 1) for j, m in zip(x["j2"], x["multiplets"])) for x in b)
        res[name] = rec
        print(name, json.dumps(rec), flush=True)
        if a.out:
            with open(a.out + ".partial", "w") as f:
                json.dump(res, f)
    if a.ref:
        res["energy_diff"] = max(abs(x - y) for x, y in zip(res["engine"]["energies"], res["reference"]["energies"]))
        print("energy_diff", res["energy_diff"])
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
