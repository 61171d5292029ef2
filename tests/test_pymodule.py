"""The reference's Python binding on the engine (SURVEY.md 8(b) "Python", north_star "the Python pymodule bindings").

python/pymodule.c of the reference is compiled UNMODIFIED (oracle/Makefile target `pymodule`) three times: against the pure reference
(py_ref, the checker), against the reference relinked to the engine's CPU test double (py_emu) and against the reference relinked to
the CUDA product (py_cuda).  A user's `chemtensor.dmrg(mpo, num_sweeps, maxiter_lanczos, tol_split, max_vdim, qnum_sector, rng_seed)`
(pymodule.c:3238-3359 -> construct_random_mps :3309 -> dmrg_twosite :3317) must give the same energies (1e-10), entropies and bond
dimensions whichever library sits underneath.  One process per variant: the three builds share the module name."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import helpers

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _run(variant, tmp_path):
    so = os.path.join(ROOT, "oracle", "_ref", variant, "chemtensor_pymodule.so")
    if not os.path.exists(so):
        subprocess.run(["make", "-s", "-C", ROOT, "dropin"], check=True, stdout=subprocess.DEVNULL)
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "pymodule"], check=True, stdout=subprocess.DEVNULL)
    out = str(tmp_path / f"{variant}.json")
    r = subprocess.run([sys.executable, os.path.join(HERE, "pymodule_worker.py"), variant, out], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    with open(out) as f:
        return json.load(f)


def _compare(a, b):
    for key in ("fh_energies", "xxz_energies"):
        assert np.max(np.abs(np.array(a[key]) - np.array(b[key]))) <= 1e-10, (key, a[key], b[key])
    assert np.max(np.abs(np.array(a["fh_entropy"]) - np.array(b["fh_entropy"]))) <= 1e-7
    assert a["fh_bond_dims"] == b["fh_bond_dims"] and a["xxz_bond_dims"] == b["xxz_bond_dims"]


def test_pymodule_dmrg_on_host_logic(tmp_path):
    _compare(_run("py_emu", tmp_path), _run("py_ref", tmp_path))


@pytest.mark.gpu
def test_pymodule_dmrg_on_cuda(tmp_path):
    if not helpers.have_gpu():
        pytest.fail("GPU test selected but no CUDA device is visible: the product has no CPU fallback")
    _compare(_run("py_cuda", tmp_path), _run("py_ref", tmp_path))
