"""Re-blocking primitives on tensors large enough for the run form of the device kernel (csrc/ctbd_remap.cu, remap_run_kernel: probe
decode per 64 entries, equally spaced copies, entry-by-entry decode where a 64-entry piece crosses a row or sector boundary):
long and short innermost runs, runs that end inside a piece, slices of the innermost axis with ascending, repeated and descending index
lists, scaling along the innermost and an outer axis.  Checked entry for entry (exactly) against the compiled reference
(src/tensor/block_sparse_tensor.c:785 transpose, :950 flatten, :1123 split, :1446 slice, :1654 multiply_pointwise_vector)."""
import ctypes as C

import numpy as np
import pytest

import helpers
from chemtensor_b200 import cabi

DTYPES = [np.float64, np.complex128]


def _qn(rng, dim, nsec):
    """few sectors -> long runs; the logical order interleaves the sectors"""
    return rng.integers(0, nsec, size=dim).astype(np.int32)


def _pair(eng, ref, rng, dtype, shape, dirs, qnums):
    dense = helpers.random_dense(rng, dtype, shape, dirs, qnums)
    return cabi.bst_from_dense(eng, dense, dirs, qnums), cabi.bst_from_dense(ref, dense, dirs, qnums), dense


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("perm", [(0, 1, 2), (1, 0, 2), (2, 1, 0), (1, 2, 0)])
def test_transpose_runs(eng, ref, rng, dtype, perm):
    shape, dirs = (23, 6, 700), [1, -1, 1]
    qn = [_qn(rng, 23, 3), _qn(rng, 6, 2), np.sort(_qn(rng, 700, 3))]     # sorted innermost axis: runs of ~230 entries
    a, b, _ = _pair(eng, ref, rng, dtype, shape, dirs, qn)
    p = (C.c_int * 3)(*perm)
    ra, rb = cabi.BST(eng), cabi.BST(ref)
    eng.block_sparse_tensor_transpose(p, a.ptr, ra.ptr)
    ref.block_sparse_tensor_transpose(p, b.ptr, rb.ptr)
    helpers.assert_bst_close(ra, rb, 0.0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("i_ax", [0, 1])
@pytest.mark.parametrize("sorted_inner", [True, False])
def test_flatten_split_runs(eng, ref, rng, dtype, i_ax, sorted_inner):
    shape, dirs = (19, 4, 450), [1, 1, -1]
    inner = _qn(rng, 450, 4)
    qn = [_qn(rng, 19, 3), _qn(rng, 4, 2), np.sort(inner) if sorted_inner else inner]
    a, b, dense = _pair(eng, ref, rng, dtype, shape, dirs, qn)
    fa, fb = cabi.BST(eng), cabi.BST(ref)
    eng.block_sparse_tensor_flatten_axes(a.ptr, i_ax, 1, fa.ptr)
    ref.block_sparse_tensor_flatten_axes(b.ptr, i_ax, 1, fb.ptr)
    helpers.assert_bst_close(fa, fb, 0.0)
    new_dim = (C.c_int64 * 2)(shape[i_ax], shape[i_ax + 1])
    new_dirs = (C.c_int * 2)(dirs[i_ax], dirs[i_ax + 1])
    ptrs, keep = cabi._qnum_ptrs([qn[i_ax], qn[i_ax + 1]])
    sa, sb = cabi.BST(eng), cabi.BST(ref)
    eng.block_sparse_tensor_split_axis(fa.ptr, i_ax, new_dim, new_dirs, ptrs, sa.ptr)
    ref.block_sparse_tensor_split_axis(fb.ptr, i_ax, new_dim, new_dirs, ptrs, sb.ptr)
    helpers.assert_bst_close(sa, sb, 0.0)
    assert np.array_equal(sa.to_dense(), dense)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("kind", ["ascending", "repeated", "descending", "outer"])
def test_slice_runs(eng, ref, rng, dtype, kind):
    shape, dirs = (17, 5, 640), [1, -1, 1]
    qn = [_qn(rng, 17, 3), _qn(rng, 5, 2), np.sort(_qn(rng, 640, 3))]
    a, b, _ = _pair(eng, ref, rng, dtype, shape, dirs, qn)
    if kind == "outer":
        i_ax, ind = 0, np.array([16, 3, 3, 0, 9, 10, 11], dtype=np.int64)
    else:
        i_ax = 2
        ind = np.sort(rng.choice(640, size=400, replace=False)).astype(np.int64)
        if kind == "repeated":
            ind = np.sort(np.concatenate([ind, ind[:150]])).astype(np.int64)
        elif kind == "descending":
            ind = ind[::-1].copy()
    ra, rb = cabi.BST(eng), cabi.BST(ref)
    p = ind.ctypes.data_as(C.POINTER(C.c_int64))
    eng.block_sparse_tensor_slice(a.ptr, i_ax, p, len(ind), ra.ptr)
    ref.block_sparse_tensor_slice(b.ptr, i_ax, p, len(ind), rb.ptr)
    helpers.assert_bst_close(ra, rb, 0.0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("axrange", [cabi.AXIS_RANGE_LEADING, cabi.AXIS_RANGE_TRAILING])
def test_scale_runs(eng, ref, rng, dtype, axrange):
    shape, dirs = (21, 6, 520), [1, 1, -1]
    qn = [_qn(rng, 21, 3), _qn(rng, 6, 2), np.sort(_qn(rng, 520, 4))]
    a, b, _ = _pair(eng, ref, rng, dtype, shape, dirs, qn)
    v = rng.standard_normal(shape[0] if axrange == cabi.AXIS_RANGE_LEADING else shape[-1])
    dt, keep = cabi.dense_vector(eng, v)
    ra, rb = cabi.BST(eng), cabi.BST(ref)
    eng.block_sparse_tensor_multiply_pointwise_vector(a.ptr, C.byref(dt), axrange, ra.ptr)
    ref.block_sparse_tensor_multiply_pointwise_vector(b.ptr, C.byref(dt), axrange, rb.ptr)
    helpers.assert_bst_close(ra, rb, 1e-15)
