"""TEST INFRASTRUCTURE -- minimal pure-Python reader for the HDF5 files of the reference's test-suite.

libhdf5 / h5py are absent from this image.  Every fixture under the reference's test/*/data and perf/ is the
simplest HDF5 dialect (written by h5py with default settings): superblock version 0 with 8-byte offsets, one
symbol-table root group (TREE / HEAP / SNOD), version-1 object headers, contiguous or compact dataset layout,
no filters; attributes (int64 / float64 arrays) sit on the root group; complex numbers are a compound {r, i}.
Only that dialect is understood here.  Used by tests/golden/make_golden.py (run in the build container, where
the reference tree is mounted) to turn the reference's own golden vectors into .npz files that travel.
"""
from __future__ import annotations

import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    def __init__(self, path: str):
        with open(path, "rb") as f:
            self.buf = f.read()
        b = self.buf
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError("not an HDF5 file")
        if b[8] != 0:
            raise ValueError(f"superblock version {b[8]} not supported")
        if b[13] != 8 or b[14] != 8:
            raise ValueError("only 8-byte offsets/lengths supported")
        # superblock v0: 8 sig, 8 version bytes, 2+2 group K, 4 flags, 4x8 addresses, then the root symbol-table entry
        root_entry = 24 + 4 * 8
        _, ohdr, cache_type, _ = struct.unpack_from("<QQII", b, root_entry)
        btree, heap = struct.unpack_from("<QQ", b, root_entry + 24)
        self.datasets: dict[str, np.ndarray] = {}
        self.attrs: dict[str, np.ndarray] = {}
        msgs = self._object_header(ohdr)
        for mtype, off, size in msgs:
            if mtype == 0x0011:   # symbol table message (when not cached in the entry)
                btree, heap = struct.unpack_from("<QQ", b, off)
            elif mtype == 0x000C:
                name, val = self._attribute(off)
                self.attrs[name] = val
        self._heap_data = self._local_heap(heap)
        for name, addr in self._walk_btree(btree):
            self.datasets[name] = self._dataset(addr)

    # ---- low level ----
    def _local_heap(self, addr: int) -> int:
        b = self.buf
        assert b[addr:addr + 4] == b"HEAP"
        (data_addr,) = struct.unpack_from("<Q", b, addr + 8 + 8 + 8)
        return data_addr

    def _heap_string(self, off: int) -> str:
        start = self._heap_data + off
        end = self.buf.index(b"\x00", start)
        return self.buf[start:end].decode()

    def _walk_btree(self, addr: int):
        b = self.buf
        if b[addr:addr + 4] == b"TREE":
            node_type, level, nent = struct.unpack_from("<BBH", b, addr + 4)
            pos = addr + 8 + 16      # skip left/right siblings
            # keys and children alternate: key0, child0, key1, child1, ... key_n
            for i in range(nent):
                pos += 8             # key
                (child,) = struct.unpack_from("<Q", b, pos)
                pos += 8
                yield from self._walk_btree(child)
        elif b[addr:addr + 4] == b"SNOD":
            (nsym,) = struct.unpack_from("<H", b, addr + 6)
            pos = addr + 8
            for i in range(nsym):
                name_off, ohdr = struct.unpack_from("<QQ", b, pos)
                yield self._heap_string(name_off), ohdr
                pos += 40
        else:
            raise ValueError("unexpected B-tree node signature")

    def _object_header(self, addr: int):
        """List of (type, data offset, size) of a version-1 object header incl. continuation blocks."""
        b = self.buf
        version, _, nmsg, _, hsize = struct.unpack_from("<BBHII", b, addr)
        if version != 1:
            raise ValueError("only version-1 object headers supported")
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", b, pos)
                data = pos + 8
                if mtype == 0x0010:
                    caddr, clen = struct.unpack_from("<QQ", b, data)
                    blocks.append((caddr, clen))
                out.append((mtype, data, msize))
                pos = data + msize
        return out

    @staticmethod
    def _pad8(n: int) -> int:
        return (n + 7) & ~7

    def _datatype(self, off: int):
        """(numpy dtype, encoded size) of a datatype message."""
        b = self.buf
        cv = b[off]
        cls, ver = cv & 0x0F, cv >> 4
        bits0 = b[off + 1]
        (size,) = struct.unpack_from("<I", b, off + 4)
        if cls == 0:      # fixed point
            signed = (bits0 >> 3) & 1
            return np.dtype(f"<{'i' if signed else 'u'}{size}"), 8 + 4
        if cls == 1:      # floating point
            return np.dtype(f"<f{size}"), 8 + 12
        if cls == 6:      # compound {r, i}
            nmemb = bits0 | (b[off + 2] << 8)
            pos = off + 8
            members = []
            for _ in range(nmemb):
                nend = b.index(b"\x00", pos)
                name = b[pos:nend].decode()
                if ver == 1 or ver == 2:
                    pos += self._pad8(nend - pos + 1)
                else:
                    pos = nend + 1
                if ver == 1:
                    (moff,) = struct.unpack_from("<I", b, pos)
                    pos += 4 + 1 + 3 + 4 + 4 + 16
                elif ver == 2:
                    (moff,) = struct.unpack_from("<I", b, pos)
                    pos += 4
                else:
                    nb = 1 if size < 256 else (2 if size < 65536 else 4)
                    moff = int.from_bytes(b[pos:pos + nb], "little")
                    pos += nb
                mdt, mlen = self._datatype(pos)
                pos += mlen
                members.append((name, mdt, moff))
            if len(members) == 2 and members[0][1] == members[1][1] and members[0][1].kind == "f":
                return np.dtype(f"<c{size}"), pos - off
            return np.dtype({"names": [m[0] for m in members], "formats": [m[1] for m in members], "offsets": [m[2] for m in members], "itemsize": size}), pos - off
        raise ValueError(f"datatype class {cls} not supported")

    def _dataspace(self, off: int):
        b = self.buf
        ver, rank, flags = b[off], b[off + 1], b[off + 2]
        pos = off + (8 if ver == 1 else 4)
        dims = struct.unpack_from(f"<{rank}Q", b, pos) if rank else ()
        return tuple(int(d) for d in dims)

    def _attribute(self, off: int):
        b = self.buf
        ver = b[off]
        name_size, dt_size, ds_size = struct.unpack_from("<HHH", b, off + 2)
        if ver != 1:
            raise ValueError("only version-1 attribute messages supported")
        pos = off + 8
        name = b[pos:pos + name_size].split(b"\x00")[0].decode()
        pos += self._pad8(name_size)
        dt, _ = self._datatype(pos)
        pos += self._pad8(dt_size)
        shape = self._dataspace(pos)
        pos += self._pad8(ds_size)
        n = int(np.prod(shape)) if shape else 1
        val = np.frombuffer(b, dtype=dt, count=n, offset=pos).reshape(shape).copy()
        return name, val

    def _dataset(self, addr: int) -> np.ndarray:
        b = self.buf
        dt = shape = None
        data_addr = data_size = None
        compact = None
        for mtype, off, size in self._object_header(addr):
            if mtype == 0x0001:
                shape = self._dataspace(off)
            elif mtype == 0x0003:
                dt, _ = self._datatype(off)
            elif mtype == 0x0008:
                ver, lclass = b[off], b[off + 1]
                if ver != 3:
                    raise ValueError("only version-3 layout messages supported")
                if lclass == 1:
                    data_addr, data_size = struct.unpack_from("<QQ", b, off + 2)
                elif lclass == 0:
                    (csize,) = struct.unpack_from("<H", b, off + 2)
                    compact = (off + 4, csize)
                else:
                    raise ValueError("chunked layout not supported")
        n = int(np.prod(shape)) if shape else 1
        if compact is not None:
            return np.frombuffer(b, dtype=dt, count=n, offset=compact[0]).reshape(shape).copy()
        if data_addr is None or data_addr == UNDEF:
            return np.zeros(shape, dtype=dt)
        return np.frombuffer(b, dtype=dt, count=n, offset=data_addr).reshape(shape).copy()


def load(path: str):
    """(datasets, attributes) of one fixture file."""
    f = H5File(path)
    return f.datasets, f.attrs
