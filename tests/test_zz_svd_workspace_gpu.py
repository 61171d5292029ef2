"""The SVD work-matrix API of the device layer (ctbd_svdws_create / ctbd_svdws_finish / ctbd_gram_offdiag) on the CUDA product:
same checks as tests/test_svd_workspace.py, with real device buffers."""
import ctypes as C

import numpy as np
import pytest

import helpers
from chemtensor_b200 import cabi
from test_svd_workspace import _check, _setup

pytestmark = pytest.mark.gpu


def _dev(dll, arr):
    p = C.c_void_p()
    assert dll.ctbd_malloc(C.byref(p), C.c_size_t(max(arr.nbytes, 16))) == 0
    if arr.nbytes:
        assert dll.ctbd_h2d(p, arr.ctypes.data_as(C.c_void_p), C.c_size_t(arr.nbytes)) == 0
    return p


def _bind(dll):
    dll.ctbd_malloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    dll.ctbd_free.argtypes = [C.c_void_p]
    dll.ctbd_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    dll.ctbd_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    dll.ctbd_svdws_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    dll.ctbd_svdws_finish.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    dll.ctbd_gram_offdiag.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p]


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
def test_workspace_pieces_on_the_device(rng, dtype):
    if not helpers.have_gpu():
        pytest.fail("GPU test selected but no CUDA device is visible")
    dll = helpers.load("cuda").dll
    _bind(dll)
    descs, mats, A, U, Vh, S = _setup(rng, dtype, [(40, 70), (90, 33), (24, 24)])
    dA, dU, dVh, dS = _dev(dll, A), _dev(dll, U), _dev(dll, Vh), _dev(dll, S)
    ws, G, gtot = C.c_void_p(), C.c_void_p(), C.c_int64(0)
    assert dll.ctbd_svdws_create(cabi.ct_dtype(dtype), len(mats), descs, dA, C.byref(ws), C.byref(G), C.byref(gtot)) == 0
    g = np.zeros(gtot.value, dtype=dtype)
    assert dll.ctbd_d2h(g.ctypes.data_as(C.c_void_p), G, C.c_size_t(g.nbytes)) == 0
    pos = 0
    for a in mats:      # a unitary row operation applied from outside
        R, Cc = min(a.shape), max(a.shape)
        q, _ = np.linalg.qr(rng.standard_normal((R, R)) + (1j * rng.standard_normal((R, R)) if np.dtype(dtype).kind == "c" else 0))
        blk = g[pos:pos + R * (Cc + R)].reshape(R, Cc + R)
        blk[...] = q.astype(dtype) @ blk
        pos += R * (Cc + R)
    assert dll.ctbd_h2d(G, g.ctypes.data_as(C.c_void_p), C.c_size_t(g.nbytes)) == 0
    assert dll.ctbd_svdws_finish(ws, G, 1, dU, dVh, dS) == 0
    for host, dev in ((U, dU), (Vh, dVh), (S, dS)):
        assert dll.ctbd_d2h(host.ctypes.data_as(C.c_void_p), dev, C.c_size_t(host.nbytes)) == 0
    _check(descs, mats, U, Vh, S)
    for p in (dA, dU, dVh, dS):
        dll.ctbd_free(p)


def test_gram_offdiag_on_the_device(rng):
    if not helpers.have_gpu():
        pytest.fail("GPU test selected but no CUDA device is visible")
    dll = helpers.load("cuda").dll
    _bind(dll)
    n = 37
    x = rng.standard_normal((n, 80)); x[5] *= 1e-6
    g = np.ascontiguousarray(x @ x.T)
    off = np.array([0], dtype=np.int64); dim = np.array([n], dtype=np.int32); out = np.zeros(2)
    dg, doff, ddim, dout = _dev(dll, g), _dev(dll, off), _dev(dll, dim), _dev(dll, out)
    assert dll.ctbd_gram_offdiag(cabi.ct_dtype(np.float64), 1, doff, ddim, dg, 1e-18, dout) == 0
    assert dll.ctbd_d2h(out.ctypes.data_as(C.c_void_p), dout, C.c_size_t(out.nbytes)) == 0
    d = np.sqrt(np.outer(np.diag(g), np.diag(g)))
    rel = (g / d) ** 2
    np.fill_diagonal(rel, 0)
    assert abs(out[1] - np.max(np.diag(g))) <= 1e-14 * out[1]
    assert abs(out[0] - rel.max()) <= 1e-10 * rel.max()
    for p in (dg, doff, ddim, dout):
        dll.ctbd_free(p)
