"""save_mps / load_mps in the reference's HDF5 layout without libhdf5 (chemtensor_b200/host/mps_io.c; reference src/state/mps.c:1219-1460,
test/state/test_mps.c:743 test_save_mps: save, load, compare).

Checks: (i) the round trip through the engine's own writer and reader reproduces every block, quantum number and dimension;
(ii) the file is valid for an INDEPENDENT reader of the dialect libhdf5 writes by default -- oracle/hdf5_v0.py, which was written
against the reference's 86 fixture files (h5py / libhdf5 output) -- and holds exactly what src/state/mps.c:1232-1268 stores: attributes
nsites / qsite / qbond_<i> as int32, datasets tensor_<i> as the dense site tensors (compound {r, i} for complex); (iii) broken input is
rejected with a negative return value.  Host code only: runs identically on the CUDA product and on the host-logic test build."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

import helpers
from chemtensor_b200 import cabi

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import hdf5_v0  # noqa: E402


def _bind(lib):
    lib.dll.save_mps.restype = C.c_int
    lib.dll.save_mps.argtypes = [C.c_char_p, C.POINTER(cabi.MPSStruct)]
    lib.dll.load_mps.restype = C.c_int
    lib.dll.load_mps.argtypes = [C.c_char_p, C.POINTER(cabi.MPSStruct)]


@pytest.fixture(scope="module")
def io_lib():
    """the host-logic build exports the same functions as the product; no device is touched by mps_io.c"""
    lib = helpers.load("emu")
    _bind(lib)
    return lib


@pytest.mark.parametrize("dtype", [np.float64, np.complex128])
@pytest.mark.parametrize("nsites,max_vdim", [(1, 4), (5, 13), (11, 30)])
def test_save_load_roundtrip(io_lib, ref, tmp_path, dtype, nsites, max_vdim):
    qsite = np.array([0, 1, -1, 2, 0], dtype=np.int32)[: 3 + (nsites % 3)]
    psi_r = helpers.ref_random_mps(ref, dtype, nsites, qsite, 1 if nsites > 1 else 0, max_vdim, seed=7)
    psi = helpers.clone_chain(io_lib, psi_r)
    path = str(tmp_path / f"mps_{nsites}.hdf5").encode()
    assert io_lib.dll.save_mps(path, psi.ptr) == 0

    # (ii) independent reader
    dsets, attrs = hdf5_v0.load(path.decode())
    assert attrs["nsites"].dtype == np.int32 and int(attrs["nsites"]) == nsites
    assert attrs["qsite"].dtype == np.int32 and np.array_equal(attrs["qsite"], qsite)
    for i in range(nsites):
        t = psi.site(i)
        assert np.array_equal(attrs[f"qbond_{i}"], t.qnums[0])
        dense = dsets[f"tensor_{i}"]
        assert dense.dtype == dtype and dense.shape == tuple(t.shape)
        assert np.array_equal(dense, t.to_dense())
    assert np.array_equal(attrs[f"qbond_{nsites}"], psi.site(nsites - 1).qnums[2])
    assert sorted(dsets) == sorted(f"tensor_{i}" for i in range(nsites))

    # (i) round trip through load_mps
    loaded = cabi.MPSStruct()
    assert io_lib.dll.load_mps(path, C.byref(loaded)) == 0
    try:
        assert loaded.nsites == nsites and loaded.d == len(qsite)
        assert np.array_equal(np.ctypeslib.as_array(loaded.qsite, shape=(len(qsite),)), qsite)
        for i in range(nsites):
            a = cabi.BST(io_lib, loaded.a[i], owned=False)
            b = psi.site(i)
            helpers.assert_bst_close(a, b, 0.0)
    finally:
        io_lib.dll.delete_mps(C.byref(loaded))


def test_load_rejects_garbage(io_lib, tmp_path):
    bad = tmp_path / "bad.hdf5"
    bad.write_bytes(b"\x89HDF\r\n\x1a\n" + b"\x02" + b"\x00" * 200)      # superblock version 2
    loaded = cabi.MPSStruct()
    assert io_lib.dll.load_mps(str(bad).encode(), C.byref(loaded)) < 0
    assert io_lib.dll.load_mps(str(tmp_path / "missing.hdf5").encode(), C.byref(loaded)) < 0


def test_load_truncated_file(io_lib, ref, tmp_path):
    psi_r = helpers.ref_random_mps(ref, np.float64, 4, np.array([0, 1], dtype=np.int32), 1, 8, seed=3)
    psi = helpers.clone_chain(io_lib, psi_r)
    path = tmp_path / "t.hdf5"
    assert io_lib.dll.save_mps(str(path).encode(), psi.ptr) == 0
    raw = path.read_bytes()
    (tmp_path / "cut.hdf5").write_bytes(raw[: len(raw) - 40])
    loaded = cabi.MPSStruct()
    assert io_lib.dll.load_mps(str(tmp_path / "cut.hdf5").encode(), C.byref(loaded)) < 0


def test_product_library_exports_io():
    """the product library carries the same two symbols (checked without touching a device)"""
    lib = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "chemtensor_b200", "libchemtensor_b200.so"))
    assert hasattr(lib, "save_mps") and hasattr(lib, "load_mps")


# message bodies as libhdf5 (through h5py) wrote them into the reference's fixture files, e.g. test/util/data/test_lanczos_iteration_z.hdf5
# (dataset 'a': compound {r, i} of IEEE doubles) and test/algorithm/data/test_dmrg_twosite.hdf5 -- the engine's writer must emit the same bytes
LIBHDF5_F64 = "11203f000800000000004000340b0034ff030000"
LIBHDF5_C128 = ("1602000010000000" "7200000000000000" "00000000" "00000000" "00000000" "00000000" + "00" * 16 + LIBHDF5_F64
                + "6900000000000000" "08000000" "00000000" "00000000" "00000000" + "00" * 16 + LIBHDF5_F64)
LIBHDF5_FILL_V2 = "0202020100000000"


@pytest.mark.parametrize("dtype,expected", [(np.float64, LIBHDF5_F64), (np.complex128, LIBHDF5_C128)])
def test_message_bytes_match_libhdf5(io_lib, ref, tmp_path, dtype, expected):
    psi_r = helpers.ref_random_mps(ref, dtype, 3, np.array([0, 1], dtype=np.int32), 1, 6, seed=11)
    psi = helpers.clone_chain(io_lib, psi_r)
    path = tmp_path / "m.hdf5"
    assert io_lib.dll.save_mps(str(path).encode(), psi.ptr) == 0
    f = hdf5_v0.H5File(str(path))
    b = f.buf
    import struct
    root_entry = 24 + 32
    btree, heap = struct.unpack_from("<QQ", b, root_entry + 24)
    seen = 0
    for name, addr in f._walk_btree(btree):
        msgs = {m[0]: (m[1], m[2]) for m in f._object_header(addr)}
        off, size = msgs[0x0003]
        assert b[off:off + len(expected) // 2].hex() == expected
        off, size = msgs[0x0005]
        assert b[off:off + 8].hex() == LIBHDF5_FILL_V2
        off, size = msgs[0x0008]
        assert b[off] == 3 and b[off + 1] == 1      # layout version 3, contiguous
        off, size = msgs[0x0001]
        assert b[off:off + 4].hex() == "01030100"    # dataspace version 1, rank 3, maximum dimensions present
        seen += 1
    assert seen == 3
    # superblock: version 0, 8-byte offsets and lengths, end-of-file address = file size
    assert b[8] == 0 and b[13] == 8 and b[14] == 8
    (eof,) = struct.unpack_from("<Q", b, 24 + 16)
    assert eof == len(b)
