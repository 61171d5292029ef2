"""Worker of tests/test_pymodule.py: imports ONE build of the reference's CPython extension (python/pymodule.c, unmodified) --
`py_ref` = linked against the pure reference, `py_emu` / `py_cuda` = relinked against the engine (oracle/Makefile `pymodule`) --
and runs chemtensor.dmrg() as a user of the reference's Python package would (reference python/pymodule.c:3238-3359).
usage: python pymodule_worker.py <py_ref|py_emu|py_cuda> <out.json>"""
import json
import os
import sys

variant, out = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", variant))
import chemtensor_pymodule as ct  # noqa: E402

res = {}
# Fermi-Hubbard chain, U(1) x U(1) sector (N = L, 2 Sz = 0): the model of BASELINE.json configs[2] in small
L = 8
mpo = ct.construct_fermi_hubbard_1d_mpo(L, 1.0, 4.0, 0.0)
psi, en, ent = ct.dmrg(mpo, num_sweeps=3, maxiter_lanczos=20, tol_split=1e-10, max_vdim=64, qnum_sector=ct.encode_quantum_number_pair(L, 0), rng_seed=42)
res["fh_energies"] = [float(x) for x in en]
res["fh_entropy"] = [float(x) for x in ent]
res["fh_bond_dims"] = [int(x) for x in psi.bond_dims]
# XXZ chain (configs[1] in small)
mpo = ct.construct_heisenberg_xxz_1d_mpo(12, 1.0, 0.8, 0.1)
psi, en, ent = ct.dmrg(mpo, num_sweeps=3, maxiter_lanczos=20, tol_split=1e-10, max_vdim=48, qnum_sector=0, rng_seed=42)
res["xxz_energies"] = [float(x) for x in en]
res["xxz_bond_dims"] = [int(x) for x in psi.bond_dims]
with open(out, "w") as f:
    json.dump(res, f)
