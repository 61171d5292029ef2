"""Edge cases of the primitives the reference handles explicitly: contraction of ALL axes to a 0-dimensional tensor
(block_sparse_tensor.c:1848-1885), operands without a single conserving block, 1 x 1 x 1 tensors."""
import ctypes as C

import numpy as np
import pytest

import helpers
from chemtensor_b200 import cabi

DTYPES = [np.float64, np.complex128]


@pytest.mark.parametrize("dtype", DTYPES)
def test_full_contraction_gives_a_scalar_tensor(eng, ref, rng, dtype):
    q0, q1 = helpers.random_qnums(rng, 5), helpers.random_qnums(rng, 7)
    s_e = helpers.random_bst(eng, rng, dtype, (5, 7), (1, -1), (q0, q1))
    t_e = helpers.random_bst(eng, rng, dtype, (5, 7), (-1, 1), (q0, q1))
    s_r, t_r = cabi.bst_clone(ref, s_e), cabi.bst_clone(ref, t_e)
    out = []
    for lib, s, t in ((eng, s_e, t_e), (ref, s_r, t_r)):
        r = cabi.BST(lib)
        lib.block_sparse_tensor_dot(s.ptr, cabi.AXIS_RANGE_TRAILING, t.ptr, cabi.AXIS_RANGE_LEADING, 2, r.ptr)
        out.append(r)
    assert out[0].ndim == 0 and out[1].ndim == 0
    assert abs(out[0].serialize()[0] - out[1].serialize()[0]) <= 1e-13 * max(1.0, abs(out[1].serialize()[0]))
    assert abs(out[0].serialize()[0] - np.sum(s_e.to_dense() * t_e.to_dense())) <= 1e-12


@pytest.mark.parametrize("dtype", DTYPES)
def test_operands_without_any_conserving_block(eng, ref, rng, dtype):
    q0, q1, q2 = np.full(4, 5, dtype=np.int32), np.zeros(6, dtype=np.int32), helpers.random_qnums(rng, 3)
    s_e = cabi.bst_allocate(eng, dtype, (4, 6), (1, -1), (q0, q1))
    assert len(list(s_e.blocks())) == 0
    t_e = helpers.random_bst(eng, rng, dtype, (6, 3), (1, -1), (q1, q2))
    s_r, t_r = cabi.bst_clone(ref, s_e), cabi.bst_clone(ref, t_e)
    res = []
    for lib, s, t in ((eng, s_e, t_e), (ref, s_r, t_r)):
        r = cabi.BST(lib)
        lib.block_sparse_tensor_dot(s.ptr, cabi.AXIS_RANGE_TRAILING, t.ptr, cabi.AXIS_RANGE_LEADING, 1, r.ptr)
        res.append(r)
    helpers.assert_same_structure(res[0], res[1])
    assert res[0].num_elements() == 0
    perm = (C.c_int * 2)(1, 0)
    tr = []
    for lib, s in ((eng, s_e), (ref, s_r)):
        r = cabi.BST(lib)
        lib.block_sparse_tensor_transpose(perm, s.ptr, r.ptr)
        tr.append(r)
    helpers.assert_same_structure(tr[0], tr[1])
    fl = []
    for lib, t in ((eng, t_e), (ref, t_r)):
        r = cabi.BST(lib)
        lib.block_sparse_tensor_flatten_axes(t.ptr, 0, 1, r.ptr)
        fl.append(r)
    helpers.assert_bst_close(fl[0], fl[1], 0.0)


@pytest.mark.parametrize("dtype", DTYPES)
def test_single_entry_tensor_through_the_split(eng, ref, rng, dtype):
    z = np.zeros(1, dtype=np.int32)
    a_e = helpers.random_bst(eng, rng, dtype, (1, 1, 1), (1, 1, -1), (z, z, z))
    a_r = cabi.bst_clone(ref, a_e)
    out = []
    for lib, a in ((eng, a_e), (ref, a_r)):
        m, u, vh = cabi.BST(lib), cabi.BST(lib), cabi.BST(lib)
        lib.block_sparse_tensor_flatten_axes(a.ptr, 0, 1, m.ptr)
        info = cabi.TruncInfo()
        assert lib.split_block_sparse_matrix_svd(m.ptr, 0.0, True, 10, False, cabi.SVD_DISTR_LEFT, u.ptr, vh.ptr, C.byref(info)) == 0
        out.append((u, vh, info.norm_sigma))
    helpers.assert_same_structure(out[0][0], out[1][0])
    helpers.assert_same_structure(out[0][1], out[1][1])
    assert abs(out[0][2] - out[1][2]) <= 1e-15
    # u s vh reproduces the entry (the phase may sit in either factor)
    assert abs(out[0][0].serialize()[0] * out[0][1].serialize()[0] - a_e.serialize()[0]) <= 1e-15


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("ntrace", [1, 2])
def test_cyclic_partial_trace_over_real_bonds(eng, ref, rng, dtype, ntrace):
    """general form of block_sparse_tensor_cyclic_partial_trace (reference block_sparse_tensor.c:1560): traced legs of dimension > 1"""
    lead_dims = [5, 3][:ntrace]
    lead_q = [helpers.random_qnums(rng, d) for d in lead_dims]
    lead_dir = [1, -1][:ntrace]
    free_dims, free_dir = [4, 6], [1, -1]
    free_q = [helpers.random_qnums(rng, d) for d in free_dims]
    shape = tuple(lead_dims + free_dims + lead_dims)
    dirs = tuple(lead_dir + free_dir + [-x for x in lead_dir])
    qn = tuple(lead_q + free_q + lead_q)
    t_e = helpers.random_bst(eng, rng, dtype, shape, dirs, qn)
    t_r = cabi.bst_clone(ref, t_e)
    r_e, r_r = cabi.BST(eng), cabi.BST(ref)
    eng.block_sparse_tensor_cyclic_partial_trace(t_e.ptr, ntrace, r_e.ptr)
    ref.block_sparse_tensor_cyclic_partial_trace(t_r.ptr, ntrace, r_r.ptr)
    helpers.assert_bst_close(r_e, r_r, 1e-13)
    dense = t_e.to_dense()
    want = np.einsum("iabi->ab", dense) if ntrace == 1 else np.einsum("ijabij->ab", dense)
    assert np.allclose(r_e.to_dense(), want, atol=1e-13)


def test_cyclic_partial_trace_golden(eng):
    """the reference's fixture test/tensor/data/test_block_sparse_tensor_cyclic_partial_trace.hdf5 (single complex data, computed here in
    complex128; the reference test compares at 1e-6... of float32 accuracy)"""
    from test_golden_engine import golden
    ds, at = golden("block_sparse_tensor_cyclic_partial_trace")
    dirs = [int(x) for x in at["axis_dir"]]
    qn = [np.asarray(at[f"qnums{i}"], dtype=np.int32) for i in range(7)]
    t = cabi.bst_from_dense(eng, np.ascontiguousarray(ds["t"].astype(np.complex128)), dirs, qn)
    r = cabi.BST(eng)
    eng.block_sparse_tensor_cyclic_partial_trace(t.ptr, 2, r.ptr)
    want = ds["t_tr"].astype(np.complex128)
    assert r.shape == want.shape
    assert np.linalg.norm(r.to_dense() - want) <= 2e-6 * np.linalg.norm(want)
