"""Block-sparse primitives of the hot path against the compiled reference (oracle/_ref).

Reference tests mirrored: test/tensor/test_block_sparse_tensor.c :385 (transpose), :456/:546 (flatten/split),
:659 (slice), :740 (multiply_pointwise_vector), :1117 (dot, all four LEADING/TRAILING combinations),
:1589 (serialize).  Structure must be bit-exact, entries agree to 1e-13 (contractions) or exactly (moves).
"""
import ctypes as C

import numpy as np
import pytest

import helpers
from chemtensor_b200 import cabi

DTYPES = [np.float64, np.complex128]


def _pair(eng, ref, rng, dtype, shape, dirs, qnums):
    dense = helpers.random_dense(rng, dtype, shape, dirs, qnums)
    return cabi.bst_from_dense(eng, dense, dirs, qnums), cabi.bst_from_dense(ref, dense, dirs, qnums), dense


@pytest.mark.parametrize("dtype", DTYPES)
def test_serialize_roundtrip(eng, ref, rng, dtype):
    shape, dirs = (5, 7, 4), [1, -1, 1]
    qn = [helpers.random_qnums(rng, d) for d in shape]
    a, b, dense = _pair(eng, ref, rng, dtype, shape, dirs, qn)
    n = eng.block_sparse_tensor_num_elements_blocks(a.ptr)
    assert n == ref.block_sparse_tensor_num_elements_blocks(b.ptr)
    va = np.zeros(n, dtype=dtype); vb = np.zeros(n, dtype=dtype)
    eng.block_sparse_tensor_serialize_entries(a.ptr, va.ctypes.data)
    ref.block_sparse_tensor_serialize_entries(b.ptr, vb.ctypes.data)
    assert np.array_equal(va, vb)
    z = cabi.bst_allocate(eng, dtype, shape, dirs, qn)
    eng.block_sparse_tensor_deserialize_entries(z.ptr, va.ctypes.data)
    assert np.array_equal(z.to_dense(), dense)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("conj", [False, True])
def test_transpose(eng, ref, rng, dtype, conj):
    shape, dirs = (4, 6, 3, 5), [1, -1, 1, -1]
    qn = [helpers.random_qnums(rng, d) for d in shape]
    a, b, _ = _pair(eng, ref, rng, dtype, shape, dirs, qn)
    perm = (C.c_int * 4)(2, 0, 3, 1)
    ra, rb = cabi.BST(eng), cabi.BST(ref)
    if conj:
        eng.block_sparse_tensor_conjugate_transpose(perm, a.ptr, ra.ptr)
        ref.block_sparse_tensor_conjugate_transpose(perm, b.ptr, rb.ptr)
    else:
        eng.block_sparse_tensor_transpose(perm, a.ptr, ra.ptr)
        ref.block_sparse_tensor_transpose(perm, b.ptr, rb.ptr)
    helpers.assert_bst_close(ra, rb, 0.0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("i_ax,new_dir", [(0, 1), (1, -1), (2, 1)])
def test_flatten_and_split(eng, ref, rng, dtype, i_ax, new_dir):
    shape, dirs = (3, 5, 4, 6), [1, 1, -1, -1]
    qn = [helpers.random_qnums(rng, d) for d in shape]
    a, b, dense = _pair(eng, ref, rng, dtype, shape, dirs, qn)
    fa, fb = cabi.BST(eng), cabi.BST(ref)
    eng.block_sparse_tensor_flatten_axes(a.ptr, i_ax, new_dir, fa.ptr)
    ref.block_sparse_tensor_flatten_axes(b.ptr, i_ax, new_dir, fb.ptr)
    helpers.assert_bst_close(fa, fb, 0.0)
    # split back
    new_dim = (C.c_int64 * 2)(shape[i_ax], shape[i_ax + 1])
    new_dirs = (C.c_int * 2)(dirs[i_ax], dirs[i_ax + 1])
    ptrs, keep = cabi._qnum_ptrs([qn[i_ax], qn[i_ax + 1]])
    sa, sb = cabi.BST(eng), cabi.BST(ref)
    eng.block_sparse_tensor_split_axis(fa.ptr, i_ax, new_dim, new_dirs, ptrs, sa.ptr)
    ref.block_sparse_tensor_split_axis(fb.ptr, i_ax, new_dim, new_dirs, ptrs, sb.ptr)
    helpers.assert_bst_close(sa, sb, 0.0)
    assert np.array_equal(sa.to_dense(), dense)


@pytest.mark.parametrize("dtype", DTYPES)
def test_slice(eng, ref, rng, dtype):
    shape, dirs = (6, 9, 5), [1, -1, 1]
    qn = [helpers.random_qnums(rng, d) for d in shape]
    a, b, _ = _pair(eng, ref, rng, dtype, shape, dirs, qn)
    ind = np.array([7, 2, 2, 8, 0, 5], dtype=np.int64)
    ra, rb = cabi.BST(eng), cabi.BST(ref)
    p = ind.ctypes.data_as(C.POINTER(C.c_int64))
    eng.block_sparse_tensor_slice(a.ptr, 1, p, len(ind), ra.ptr)
    ref.block_sparse_tensor_slice(b.ptr, 1, p, len(ind), rb.ptr)
    helpers.assert_bst_close(ra, rb, 0.0)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("axrange", [cabi.AXIS_RANGE_LEADING, cabi.AXIS_RANGE_TRAILING])
def test_multiply_pointwise_vector(eng, ref, rng, dtype, axrange):
    shape, dirs = (6, 4, 7), [1, 1, -1]
    qn = [helpers.random_qnums(rng, d) for d in shape]
    a, b, _ = _pair(eng, ref, rng, dtype, shape, dirs, qn)
    v = rng.standard_normal(shape[0] if axrange == cabi.AXIS_RANGE_LEADING else shape[-1])
    dt, keep = cabi.dense_vector(eng, v)
    ra, rb = cabi.BST(eng), cabi.BST(ref)
    eng.block_sparse_tensor_multiply_pointwise_vector(a.ptr, C.byref(dt), axrange, ra.ptr)
    ref.block_sparse_tensor_multiply_pointwise_vector(b.ptr, C.byref(dt), axrange, rb.ptr)
    helpers.assert_bst_close(ra, rb, 1e-15)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("ar_s", [cabi.AXIS_RANGE_LEADING, cabi.AXIS_RANGE_TRAILING])
@pytest.mark.parametrize("ar_t", [cabi.AXIS_RANGE_LEADING, cabi.AXIS_RANGE_TRAILING])
@pytest.mark.parametrize("ndim_mult", [1, 2, 3])
def test_dot(eng, ref, rng, dtype, ar_s, ar_t, ndim_mult):
    """5-leg . 6-leg over up to three axes, all four axis-range combinations (reference test :1117-1233)."""
    cdim = [7, 4, 11][:ndim_mult]
    cq = [helpers.random_qnums(rng, d) for d in cdim]
    cdir = [1, -1, 1][:ndim_mult]
    fs_dim, ft_dim = [5, 6, 3][: 5 - ndim_mult], [4, 2, 8, 3, 5][: 6 - ndim_mult]
    fs_q = [helpers.random_qnums(rng, d) for d in fs_dim]
    ft_q = [helpers.random_qnums(rng, d) for d in ft_dim]
    fs_dir = [1, -1, -1][: len(fs_dim)]
    ft_dir = [-1, 1, 1, -1, 1][: len(ft_dim)]
    if ar_s == cabi.AXIS_RANGE_LEADING:
        s_shape, s_dirs, s_q = cdim + fs_dim, cdir + fs_dir, cq + fs_q
    else:
        s_shape, s_dirs, s_q = fs_dim + cdim, fs_dir + cdir, fs_q + cq
    ncd = [-d for d in cdir]
    if ar_t == cabi.AXIS_RANGE_LEADING:
        t_shape, t_dirs, t_q = cdim + ft_dim, ncd + ft_dir, cq + ft_q
    else:
        t_shape, t_dirs, t_q = ft_dim + cdim, ft_dir + ncd, ft_q + cq
    sa, sb, sd = _pair(eng, ref, rng, dtype, tuple(s_shape), s_dirs, s_q)
    ta, tb, td = _pair(eng, ref, rng, dtype, tuple(t_shape), t_dirs, t_q)
    ra, rb = cabi.BST(eng), cabi.BST(ref)
    eng.block_sparse_tensor_dot(sa.ptr, ar_s, ta.ptr, ar_t, ndim_mult, ra.ptr)
    ref.block_sparse_tensor_dot(sb.ptr, ar_s, tb.ptr, ar_t, ndim_mult, rb.ptr)
    helpers.assert_bst_close(ra, rb, 1e-13)
    # and against a dense contraction
    ax_s = list(range(ndim_mult)) if ar_s == cabi.AXIS_RANGE_LEADING else list(range(len(s_shape) - ndim_mult, len(s_shape)))
    ax_t = list(range(ndim_mult)) if ar_t == cabi.AXIS_RANGE_LEADING else list(range(len(t_shape) - ndim_mult, len(t_shape)))
    dense = np.tensordot(sd, td, axes=(ax_s, ax_t))
    assert helpers.rel_err(ra.to_dense(), dense) <= 1e-13


@pytest.mark.parametrize("dtype", DTYPES)
def test_dot_larger_blocks(eng, ref, rng, dtype):
    """Blocks wider than one GEMM tile with ragged edges (sector multiplicities 70-150)."""
    def qn(dim, nsec):
        return np.sort(rng.integers(0, nsec, size=dim)).astype(np.int32)
    s_shape, t_shape = (230, 3, 190), (190, 4, 170)
    s_q = [qn(230, 3), np.array([0, 1, -1], dtype=np.int32), qn(190, 2)]
    t_q = [s_q[2], np.array([0, 1, 0, -1], dtype=np.int32), qn(170, 3)]
    s_dirs, t_dirs = [1, 1, -1], [1, 1, -1]
    sa, sb, _ = _pair(eng, ref, rng, dtype, s_shape, s_dirs, s_q)
    ta, tb, _ = _pair(eng, ref, rng, dtype, t_shape, t_dirs, t_q)
    ra, rb = cabi.BST(eng), cabi.BST(ref)
    eng.block_sparse_tensor_dot(sa.ptr, cabi.AXIS_RANGE_TRAILING, ta.ptr, cabi.AXIS_RANGE_LEADING, 1, ra.ptr)
    ref.block_sparse_tensor_dot(sb.ptr, cabi.AXIS_RANGE_TRAILING, tb.ptr, cabi.AXIS_RANGE_LEADING, 1, rb.ptr)
    helpers.assert_bst_close(ra, rb, 1e-13)
