"""A small reader for the subset of HDF5 that XDMFTensorOutput's data files use (test infrastructure).

There is no libhdf5 / h5py in this image.  This module parses the file format directly - superblock version 0, version 1
object headers (with continuation blocks), old-style groups (symbol table message -> version 1 B-tree of symbol table
nodes + local heap), simple dataspaces, fixed / floating point datatypes, and contiguous or chunked (version 3 layout,
version 1 chunk B-tree) datasets with the deflate filter - following the published "HDF5 File Format Specification
Version 2.0".  It reads the files libhdf5 wrote for the reference's gold results (tests/test_h5lite.py checks that against
the fixtures extracted earlier) and the files the host driver's own writer (host/shim/h5lite.C) produces, which is how that
writer is validated: one reader, both producers.
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    def __init__(self, path):
        self.b = open(path, "rb").read()
        b = self.b
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise ValueError("not an HDF5 file")
        self.sb_version = b[8]
        if self.sb_version != 0:
            raise ValueError(f"superblock version {self.sb_version} not supported by this reader")
        self.size_offsets, self.size_lengths = b[13], b[14]
        if (self.size_offsets, self.size_lengths) != (8, 8):
            raise ValueError("only 8-byte offsets / lengths")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", b, 16)
        self.base, self.free, self.eof, self.driver = struct.unpack_from("<QQQQ", b, 24)
        # root group symbol table entry
        self.root = self._symbol_entry(56)
        self.datasets = {}
        self._walk_group(self.root, "")

    # ---- low level
    def _symbol_entry(self, off):
        name_off, header, cache_type = struct.unpack_from("<QQI", self.b, off)
        scratch = self.b[off + 24:off + 40]
        return {"name_off": name_off, "header": header, "cache": cache_type, "scratch": scratch}

    def _messages(self, addr):
        """(type, flags, body) of every message of a version 1 object header, continuation blocks included."""
        b = self.b
        version, _, nmsg, _refcnt, hsize = struct.unpack_from("<BBHII", b, addr)
        if version != 1:
            raise ValueError(f"object header version {version} at {addr}")
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            start, size = blocks.pop(0)
            p, end = start, start + size
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, flags = struct.unpack_from("<HHB", b, p)
                body = b[p + 8:p + 8 + msize]
                out.append((mtype, flags, body))
                if mtype == 0x10:  # continuation
                    coff, clen = struct.unpack_from("<QQ", body, 0)
                    blocks.append((coff, clen))
                p += 8 + msize
        return out

    def _heap_name(self, heap_addr, off):
        b = self.b
        assert b[heap_addr:heap_addr + 4] == b"HEAP"
        data_addr = struct.unpack_from("<Q", b, heap_addr + 24)[0]
        end = b.index(b"\0", data_addr + off)
        return b[data_addr + off:end].decode()

    def _group_entries(self, btree, heap):
        b = self.b
        sig = b[btree:btree + 4]
        if sig == b"TREE":
            ntype, level, nent = struct.unpack_from("<BBH", b, btree + 4)
            assert ntype == 0
            p = btree + 24
            for i in range(nent):
                child = struct.unpack_from("<Q", b, p + 8)[0]  # key, child, key, child ...
                yield from self._group_entries(child, heap)
                p += 16
        elif sig == b"SNOD":
            nsym = struct.unpack_from("<H", b, btree + 6)[0]
            for i in range(nsym):
                e = self._symbol_entry(btree + 8 + 40 * i)
                yield self._heap_name(heap, e["name_off"]), e
        else:
            raise ValueError(f"unexpected signature {sig} at {btree}")

    def _walk_group(self, entry, prefix):
        btree = heap = None
        for mtype, _, body in self._messages(entry["header"]):
            if mtype == 0x11:
                btree, heap = struct.unpack_from("<QQ", body, 0)
        if btree is None:
            return
        for name, e in self._group_entries(btree, heap):
            msgs = self._messages(e["header"])
            if any(t == 0x11 for t, _, _ in msgs):
                self._walk_group(e, prefix + name + "/")
            else:
                self.datasets[prefix + name] = msgs

    # ---- datasets
    def keys(self):
        return sorted(self.datasets)

    def info(self, name):
        """Structural description of a dataset (shape, dtype, layout class, chunk shape, filters)."""
        d = {"filters": []}
        for mtype, _, body in self.datasets[name]:
            if mtype == 0x01:
                version, rank, flags = body[0], body[1], body[2]
                p = 8 if version == 1 else 4
                d["shape"] = struct.unpack_from("<" + "Q" * rank, body, p)
                d["dataspace_version"] = version
            elif mtype == 0x03:
                cls, size = body[0] & 0x0F, struct.unpack_from("<I", body, 4)[0]
                d["dtype_class"], d["dtype_size"], d["dtype_version"] = cls, size, body[0] >> 4
                if cls == 1:
                    d["dtype"] = {4: "<f4", 8: "<f8"}[size]
                elif cls == 0:
                    signed = bool(body[1] & 0x08)
                    d["dtype"] = ("<i" if signed else "<u") + str(size)
            elif mtype == 0x08:
                version = body[0]
                d["layout_version"] = version
                if version != 3:
                    raise ValueError(f"data layout version {version}")
                cls = body[1]
                d["layout_class"] = cls
                if cls == 1:
                    d["address"], d["nbytes"] = struct.unpack_from("<QQ", body, 2)
                elif cls == 2:
                    rank = body[2]
                    d["chunk_btree"] = struct.unpack_from("<Q", body, 3)[0]
                    d["chunk"] = struct.unpack_from("<" + "I" * rank, body, 11)  # last entry: element size
                elif cls == 0:
                    size = struct.unpack_from("<H", body, 2)[0]
                    d["compact"] = body[4:4 + size]
            elif mtype == 0x0B:
                version, nfilters = body[0], body[1]
                d["filter_version"] = version
                p = 8 if version == 1 else 2
                for _ in range(nfilters):
                    fid, namelen, flags, ncd = struct.unpack_from("<HHHH", body, p)
                    p += 8
                    if version == 1 or fid >= 256:
                        p += (namelen + 7) // 8 * 8 if version == 1 else namelen
                    cd = struct.unpack_from("<" + "I" * ncd, body, p)
                    p += 4 * ncd
                    if version == 1 and ncd % 2:
                        p += 4
                    d["filters"].append((fid, cd))
            elif mtype == 0x05:
                d["fill_version"] = body[0]
        return d

    def _chunks(self, addr, rank):
        b = self.b
        assert b[addr:addr + 4] == b"TREE", addr
        ntype, level, nent = struct.unpack_from("<BBH", b, addr + 4)
        assert ntype == 1
        p = addr + 24
        keysize = 8 + 8 * (rank + 1)
        for _ in range(nent):
            size, mask = struct.unpack_from("<II", b, p)
            offs = struct.unpack_from("<" + "Q" * (rank + 1), b, p + 8)
            child = struct.unpack_from("<Q", b, p + keysize)[0]
            if level == 0:
                yield offs[:rank], size, mask, child
            else:
                yield from self._chunks(child, rank)
            p += keysize + 8

    def read(self, name):
        d = self.info(name)
        shape, dt = d["shape"], np.dtype(d["dtype"])
        if d["layout_class"] == 1:
            return np.frombuffer(self.b, dt, int(np.prod(shape)), d["address"]).reshape(shape).copy()
        if d["layout_class"] == 0:
            return np.frombuffer(d["compact"], dt).reshape(shape).copy()
        rank = len(shape)
        cshape = d["chunk"][:rank]
        out = np.zeros(shape, dt)
        for offs, size, mask, addr in self._chunks(d["chunk_btree"], rank):
            raw = self.b[addr:addr + size]
            for i, (fid, _) in reversed(list(enumerate(d["filters"]))):
                if mask & (1 << i):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                else:
                    raise ValueError(f"filter {fid}")
            block = np.frombuffer(raw, dt).reshape(cshape)
            sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cshape, shape))
            out[sl] = block[tuple(slice(0, s.stop - s.start) for s in sl)]
        return out
