"""Parity at BASELINE.json's full size: the bench workload itself (Fermi-Hubbard L=64, U(1)xU(1), bond dimension 4096,
two-site effective Hamiltonian at the centre bond, converged sector structure), through the reference-named C-ABI call.

  * against the compiled reference (oracle/_ref, ~1.5 s per application on the box's host cores) on the same seeded
    operands: sector structure bit-exact, relative Frobenius error <= 1e-12 (north_star's matvec tolerance);
  * size-independent properties of the engine alone: linearity in the two-site tensor, linearity in the right
    environment, and run-to-run bit-identity (the contraction order is fixed by the plans).

The CPU test double would need minutes for 1.4e11 flops, so this file runs on the CUDA product only (-m gpu).
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import helpers
from chemtensor_b200 import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

pytestmark = pytest.mark.gpu

WORKLOAD = "fh_L64_D4096"


@pytest.fixture(scope="module")
def cuda():
    if not helpers.have_gpu():
        pytest.fail("GPU test selected but no CUDA device is visible: the product has no CPU fallback")
    return helpers.load("cuda")


@pytest.fixture(scope="module")
def operands(cuda):
    return bench.build_operands(cuda, WORKLOAD, seed=42)


def _apply(lib, a, w, l, r):
    b = cabi.BST(lib)
    lib.apply_local_hamiltonian(a.ptr, w.ptr, l.ptr, r.ptr, b.ptr)
    return b


def _like(lib, t, values=None, rng=None):
    """A tensor with the structure of t, entries from 'values' (packed order) or N(0,1)."""
    x = cabi.bst_allocate(lib, t.dtype, t.shape, t.axis_dir, t.qnums)
    x.deserialize(values if values is not None else rng.standard_normal(t.num_elements()))
    return x


def test_full_size_matvec_matches_the_reference(cuda, operands):
    ref = helpers.load("ref")
    a, w, l, r = operands
    ar, wr, lr, rr = bench.build_operands(ref, WORKLOAD, seed=42)      # same seed, same generator: identical entries
    assert np.array_equal(a.serialize(), ar.serialize())
    b = _apply(cuda, a, w, l, r)
    b_ref = _apply(ref, ar, wr, lr, rr)
    helpers.assert_same_structure(b, b_ref)
    helpers.assert_same_structure(b, a)
    err = helpers.rel_err(b.serialize(), b_ref.serialize())
    assert err <= 1e-12, f"relative Frobenius error {err:.3e} at D=4096"
    # vector length of the bench line
    assert b.num_elements() == a.num_elements() == 17119384


def test_full_size_linearity_and_determinism(cuda, operands):
    a, w, l, r = operands
    rng = np.random.default_rng(7)
    x = _like(cuda, a, rng=rng)
    y = _like(cuda, a, rng=rng)
    alpha, beta = 0.75, -1.25
    z = _like(cuda, a, values=alpha * x.serialize() + beta * y.serialize())
    hx, hy, hz = (_apply(cuda, t, w, l, r).serialize() for t in (x, y, z))
    assert helpers.rel_err(hz, alpha * hx + beta * hy) <= 1e-12
    # same plans, same order of accumulation: bit-identical from run to run
    assert np.array_equal(_apply(cuda, x, w, l, r).serialize(), hx)
    # linear in the right environment as well
    r2 = _like(cuda, r, values=2.0 * r.serialize())
    assert helpers.rel_err(_apply(cuda, x, w, l, r2).serialize(), 2.0 * hx) <= 1e-13
