"""The block linear combination primitive (ctbd_lc_*, include/ctb_device.h; csrc/ctbd_blocklc.cu and its CPU twin) against NumPy.

Bit-exact for single-term copies / permutations (pure data movement), 1e-15 relative for weighted sums.  Covers what the SU(2) layer asks of
it: contiguous blocks of any length and offset parity (16-byte fast path and its tails), strided source and destination maps (transposes,
stacking into matrices), several terms per block, blocks without terms (zero fill), in-place scaling, complex conjugation.
"""
import ctypes as C

import numpy as np
import pytest

import helpers

MAXD = 8


class LcTerm(C.Structure):
    _fields_ = [("src_off", C.c_int64), ("coef", C.c_double)]


class LcBlock(C.Structure):
    _fields_ = [("dst_off", C.c_int64), ("term_begin", C.c_int32), ("term_end", C.c_int32), ("ndim", C.c_int32), ("pad_", C.c_int32),
                ("dim", C.c_int32 * MAXD), ("dstride", C.c_int64 * MAXD), ("sstride", C.c_int64 * MAXD)]


def _dev(lib):
    d = lib.dll
    d.ctbd_malloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    d.ctbd_free.argtypes = [C.c_void_p]
    d.ctbd_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    d.ctbd_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    d.ctbd_lc_plan_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(LcBlock), C.c_int, C.POINTER(LcTerm), C.POINTER(C.c_void_p)]
    d.ctbd_lc_plan_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    d.ctbd_lc_plan_destroy.argtypes = [C.c_void_p]
    return d


def _run(lib, cplx, conj, blocks, terms, src, ndst, dst_init=None, inplace=False):
    d = _dev(lib)
    es = 16 if cplx else 8
    ps, pd, plan = C.c_void_p(), C.c_void_p(), C.c_void_p()
    assert d.ctbd_malloc(C.byref(ps), max(src.size, 1) * es) == 0
    assert d.ctbd_h2d(ps, src.ctypes.data, src.size * es) == 0
    if inplace:
        pd = ps
    else:
        assert d.ctbd_malloc(C.byref(pd), max(ndst, 1) * es) == 0
        init = dst_init if dst_init is not None else np.full(ndst, 7.25, dtype=src.dtype)
        assert d.ctbd_h2d(pd, init.ctypes.data, ndst * es) == 0
    ba = (LcBlock * len(blocks))(*blocks)
    ta = (LcTerm * max(len(terms), 1))(*terms)
    assert d.ctbd_lc_plan_create(3 if cplx else 1, conj, len(blocks), ba, len(terms), ta, C.byref(plan)) == 0
    assert d.ctbd_lc_plan_run(plan, ps, pd) == 0
    out = np.empty(ndst, dtype=src.dtype)
    assert d.ctbd_d2h(out.ctypes.data, pd, ndst * es) == 0
    d.ctbd_lc_plan_destroy(plan)
    d.ctbd_free(ps)
    if not inplace:
        d.ctbd_free(pd)
    return out


def _block(dst_off, tb, te, dims, dstr, sstr):
    b = LcBlock()
    b.dst_off, b.term_begin, b.term_end, b.ndim = dst_off, tb, te, len(dims)
    for i in range(MAXD):
        b.dim[i] = dims[i] if i < len(dims) else 1
        b.dstride[i] = dstr[i] if i < len(dims) else 0
        b.sstride[i] = sstr[i] if i < len(dims) else 0
    return b


@pytest.mark.parametrize("cplx", [False, True])
def test_contiguous_blocks_all_lengths_and_parities(eng, cplx):
    """weighted sums of 0-3 contiguous source blocks; lengths around the chunk (1024) and vector (2) granules, odd and even offsets"""
    rng = np.random.default_rng(5)
    dt = np.complex128 if cplx else np.float64
    lengths = [1, 2, 3, 31, 32, 33, 255, 1023, 1024, 1025, 2047, 4099, 70001]
    src = rng.standard_normal(400000).astype(dt)
    if cplx:
        src = src + 1j * rng.standard_normal(src.size)
    blocks, terms, expect = [], [], []
    doff = 0
    for k, n in enumerate(lengths * 2):
        nterm = k % 4
        doff += k % 2                 # alternate the parity of the destination offset
        tb = len(terms)
        acc = np.zeros(n, dtype=dt)
        for t in range(nterm):
            so = int(rng.integers(0, src.size - n))
            co = float(rng.standard_normal())
            terms.append(LcTerm(so, co))
            acc += co * src[so:so + n]
        blocks.append(_block(doff, tb, len(terms), [n], [1], [1]))
        expect.append((doff, acc))
        doff += n
    for conj in ((0, 1) if cplx else (0,)):
        out = _run(eng, cplx, conj, blocks, terms, src, doff)
        for off, acc in expect:
            ref = np.conj(acc) if conj else acc
            assert np.allclose(out[off:off + acc.size], ref, rtol=0, atol=1e-14 * max(1.0, np.abs(ref).max()))


def test_permutations_and_stacking_are_exact(eng):
    """single-term blocks with strided maps on either side: axis permutations of 2- to 6-dimensional blocks and rows / columns placed into a
    larger matrix -- pure data movement, compared bit for bit"""
    rng = np.random.default_rng(7)
    src_parts, blocks, terms, checks = [], [], [], []
    soff = doff = 0
    for nd in (2, 3, 4, 5, 6):
        for _ in range(3):
            dims = [int(x) for x in rng.integers(1, 6 if nd > 4 else 9, nd)]
            perm = [int(x) for x in rng.permutation(nd)]
            x = rng.standard_normal(dims)
            y = np.transpose(x, perm)                       # destination axis i = source axis perm[i]
            sstr_src = [int(s // 8) for s in x.strides]
            ddims = [dims[p] for p in perm]
            dstr = [int(np.prod(ddims[i + 1:])) for i in range(nd)]
            sstr = [sstr_src[p] for p in perm]
            terms.append(LcTerm(soff, 1.0))
            blocks.append(_block(doff, len(terms) - 1, len(terms), ddims, dstr, sstr))
            checks.append((doff, np.ascontiguousarray(y).ravel()))
            src_parts.append(x.ravel())
            soff += x.size
            doff += x.size
    # a 7 x 5 block into rows 3.. and columns 2.. of a 12 x 11 row-major matrix (strided destination)
    x = rng.standard_normal((7, 5))
    terms.append(LcTerm(soff, 1.0))
    mat_off = doff
    blocks.append(_block(mat_off + 3 * 11 + 2, len(terms) - 1, len(terms), [7, 5], [11, 1], [5, 1]))
    src_parts.append(x.ravel())
    soff += x.size
    doff += 12 * 11
    src = np.concatenate(src_parts)
    out = _run(eng, False, 0, blocks, terms, src, doff)
    for off, y in checks:
        assert np.array_equal(out[off:off + y.size], y)
    m = out[mat_off:mat_off + 132].reshape(12, 11)
    assert np.array_equal(m[3:10, 2:7], x)
    assert np.all(m[:3] == 7.25) and np.all(m[:, :2] == 7.25)        # untouched entries keep their content


def test_in_place_scaling(eng):
    rng = np.random.default_rng(11)
    src = rng.standard_normal(5000)
    blocks, terms = [], []
    off = 0
    for n, f in ((1, 2.0), (777, -1.0), (1024, 0.5), (3198, 3.0)):
        terms.append(LcTerm(off, f))
        blocks.append(_block(off, len(terms) - 1, len(terms), [n], [1], [1]))
        off += n
    out = _run(eng, False, 0, blocks, terms, src.copy(), 5000, inplace=True)
    off = 0
    for n, f in ((1, 2.0), (777, -1.0), (1024, 0.5), (3198, 3.0)):
        assert np.array_equal(out[off:off + n], f * src[off:off + n])
        off += n
