"""Randomised parity sweep of the SU(2) path on the CPU test double against the unmodified reference: Heisenberg and Fermi-Hubbard SU(2) chains,
both DMRG variants, with and without truncation, inside the range the reference's recoupling tables cover (start bonds 2j <= 2, L <= 8).
usage: python tools/su2_fuzz.py   (216 cases, about 20 s; prints every mismatch above 1e-9 and the count)"""
import ctypes as C, numpy as np, sys, itertools
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import su2_helpers as S
r = S.ref(); e = S.engine("emu")
bad = 0; n = 0
for seed, L, sector, mirr, deg, model, (tol, mv) in itertools.product([1, 2, 3], [4, 7, 8], [0, 1], [1, 2], [1, 5], ["h", "fh"], [(1e-6, 40), (0.0, 9)]):
    if model == "h":
        if (L + sector) % 2: continue
        mpo = S.heisenberg_mpo(L, 0.9); site = ([1], [0, 1])
    else:
        mpo = S.fermi_hubbard_mpo(L, 1.0, 3.0, 0.7); site = ([0, 1], [2, 1])
    psi = S.random_mps(L, site[0], site[1], sector, mirr, deg, seed, scale=2.0)
    if not r.su2_mps_is_consistent(C.byref(psi)): continue
    if any(psi.a[i].charge_sectors.nsec == 0 for i in range(L)): continue
    print('cfg', seed, L, sector, mirr, deg, model, tol, mv, flush=True)
    p1 = S.copy_mps(psi); p2 = S.copy_mps(psi)
    e1 = (C.c_double*2)(); e2 = (C.c_double*2)(); s1=(C.c_double*L)(); s2=(C.c_double*L)()
    rc1 = r.su2_dmrg_twosite(C.byref(mpo), 2, 4, tol, mv, C.byref(p1), e1, s1)
    rc2 = e.su2_dmrg_twosite(C.byref(mpo), 2, 4, tol, mv, C.byref(p2), e2, s2)
    d = max(abs(a-b) for a, b in zip(e1, e2))
    n += 1
    if rc1 != rc2 or d > 1e-9 or not r.su2_mps_is_consistent(C.byref(p2)):
        bad += 1; print("MISMATCH", rc1, rc2, list(e1), list(e2), flush=True)
    p1 = S.copy_mps(psi); p2 = S.copy_mps(psi)
    rc1 = r.su2_dmrg_singlesite(C.byref(mpo), 2, 4, C.byref(p1), e1)
    rc2 = e.su2_dmrg_singlesite(C.byref(mpo), 2, 4, C.byref(p2), e2)
    d = max(abs(a-b) for a, b in zip(e1, e2))
    if rc1 != rc2 or d > 1e-9:
        bad += 1; print("MISMATCH single", rc1, rc2, list(e1), list(e2), flush=True)
print("cases", n, "bad", bad)
