"""The big-block SVD in pieces (include/ctb_device.h: ctbd_svdws_create / ctbd_svdws_finish / ctbd_gram_offdiag), the groundwork of a
GEMM-driven block-Jacobi stage: work matrices [G | W] exposed to the host, unitary row operations applied from outside, polish +
finish inside the device layer.  Checked here on the CPU test double of the device layer (device pointers are host pointers there)."""
import ctypes as C

import numpy as np
import pytest

import helpers
from chemtensor_b200 import cabi


class MatDesc(C.Structure):
    _fields_ = [("a_off", C.c_int64), ("m", C.c_int32), ("n", C.c_int32), ("o0_off", C.c_int64), ("o1_off", C.c_int64), ("s_off", C.c_int64)]


def _setup(rng, dtype, shapes):
    descs = (MatDesc * len(shapes))()
    a_off = u_off = v_off = s_off = 0
    mats = []
    for b, (m, n) in enumerate(shapes):
        k = min(m, n)
        descs[b] = MatDesc(a_off, m, n, u_off, v_off, s_off)
        a = rng.standard_normal((m, n)) + (1j * rng.standard_normal((m, n)) if np.dtype(dtype).kind == "c" else 0)
        mats.append(a.astype(dtype))
        a_off += m * n; u_off += m * k; v_off += k * n; s_off += k
    A = np.concatenate([a.reshape(-1) for a in mats])
    return descs, mats, A, np.zeros(u_off, dtype=dtype), np.zeros(v_off, dtype=dtype), np.zeros(s_off)


def _check(descs, mats, U, Vh, S):
    for b, a in enumerate(mats):
        m, n = a.shape; k = min(m, n)
        u = U[descs[b].o0_off:descs[b].o0_off + m * k].reshape(m, k)
        vh = Vh[descs[b].o1_off:descs[b].o1_off + k * n].reshape(k, n)
        s = S[descs[b].s_off:descs[b].s_off + k]
        assert np.all(np.diff(s) <= 1e-13 * s[0])                      # descending
        assert np.allclose(u.conj().T @ u, np.eye(k), atol=1e-12)
        assert np.allclose(vh @ vh.conj().T, np.eye(k), atol=1e-12)
        assert np.linalg.norm(u @ np.diag(s) @ vh - a) <= 1e-12 * np.linalg.norm(a)
        assert np.allclose(s, np.linalg.svd(a, compute_uv=False), rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_workspace_pieces_equal_the_batched_svd(rng, dtype):
    emu = helpers.load("emu").dll
    descs, mats, A, U, Vh, S = _setup(rng, dtype, [(9, 14), (13, 6), (8, 8)])
    dt = cabi.ct_dtype(dtype)
    ws, G, gtot = C.c_void_p(), C.c_void_p(), C.c_int64(0)
    assert emu.ctbd_svdws_create(dt, len(mats), descs, A.ctypes.data_as(C.c_void_p), C.byref(ws), C.byref(G), C.byref(gtot)) == 0
    assert gtot.value == sum(min(a.shape) * (max(a.shape) + min(a.shape)) for a in mats)
    # a unitary row operation applied from outside to every work matrix (same on G and W): the factorisation must not change
    g = np.ctypeslib.as_array(C.cast(G, C.POINTER(C.c_double)), shape=(gtot.value * (2 if np.dtype(dtype).kind == "c" else 1),)).view(dtype)
    pos = 0
    for a in mats:
        R, Cc = min(a.shape), max(a.shape)
        q, _ = np.linalg.qr(rng.standard_normal((R, R)) + (1j * rng.standard_normal((R, R)) if np.dtype(dtype).kind == "c" else 0))
        blk = g[pos:pos + R * (Cc + R)].reshape(R, Cc + R)
        blk[...] = q.astype(dtype) @ blk
        pos += R * (Cc + R)
    fin = emu.ctbd_svdws_finish
    fin.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    assert fin(ws, G, 1, U.ctypes.data_as(C.c_void_p), Vh.ctypes.data_as(C.c_void_p), S.ctypes.data_as(C.c_void_p)) == 0
    _check(descs, mats, U, Vh, S)


def test_gram_offdiag_measure(rng):
    emu = helpers.load("emu").dll
    n = 5
    x = rng.standard_normal((n, 40))
    x[3] *= 1e-6
    g = x @ x.T
    off = np.array([0], dtype=np.int64); dim = np.array([n], dtype=np.int32)
    out = np.zeros(2)
    fn = emu.ctbd_gram_offdiag
    fn.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]
    assert fn(cabi.ct_dtype(np.float64), 1, off.ctypes.data, dim.ctypes.data, g.ctypes.data, 1e-18, out.ctypes.data) == 0
    d = np.sqrt(np.outer(np.diag(g), np.diag(g)))
    rel = (g / d) ** 2
    np.fill_diagonal(rel, 0)
    assert abs(out[1] - np.max(np.diag(g))) <= 1e-15 * out[1]
    assert abs(out[0] - rel.max()) <= 1e-12 * rel.max()
    # an orthogonal set measures zero
    q, _ = np.linalg.qr(rng.standard_normal((40, n)))
    g2 = np.ascontiguousarray(q.T @ q)
    out[:] = 0
    assert fn(cabi.ct_dtype(np.float64), 1, off.ctypes.data, dim.ctypes.data, g2.ctypes.data, 1e-18, out.ctypes.data) == 0
    assert out[0] <= 1e-28
