"""Minimal reader for the HDF5 files the reference ships (Data/**/*.h5) -- h5py is not a dependency of this package.

Scope (everything the reference's datasets use, Data/dataset.py:1-104 and the generators under Data/): superblock
version 0, the root group's symbol table (v1 B-tree + local heap + symbol-table nodes), version-1 object headers with
dataspace / datatype / data-layout messages, CONTIGUOUS little-endian fixed-point or IEEE floating-point data.  Anything
else (chunking, compression, nested groups, new-style groups) raises H5Error instead of guessing.

    f = H5File(path);  f.keys();  f["rhs"]  ->  numpy array (a copy; dtype / shape as stored)
"""
import struct

import numpy as np


class H5Error(RuntimeError):
    pass


UNDEF = 0xFFFFFFFFFFFFFFFF


class H5File:
    def __init__(self, path):
        self.path = path
        with open(path, "rb") as fh:
            self.buf = fh.read()
        b = self.buf
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise H5Error(f"{path}: not an HDF5 file")
        if b[8] != 0:
            raise H5Error(f"{path}: superblock version {b[8]} not supported (the reference's files are version 0)")
        if b[13] != 8 or b[14] != 8:
            raise H5Error("only 8-byte offsets / lengths are supported")
        # superblock v0: 8 sig, 8 versions/sizes, 2+2 group K, 4 flags, 4 x 8 addresses, then the root symbol-table entry
        self.base = struct.unpack_from("<Q", b, 24)[0]
        root = 24 + 32
        # symbol table entry: link name offset (8), object header address (8), cache type (4), reserved (4), scratch (16)
        cache_type = struct.unpack_from("<I", b, root + 16)[0]
        if cache_type != 1:
            raise H5Error("root group without cached B-tree / heap addresses")
        btree, heap = struct.unpack_from("<QQ", b, root + 24)
        self._datasets = {}
        # where the fields a fixture builder may patch live in the file: name -> dict(dims_off, layout_addr_off, ...)
        self.fields = {}
        heap_data = self._local_heap(heap)
        for name_off, ohdr in self._walk_btree(btree):
            name = self._cstr(heap_data + name_off)
            try:
                self._datasets[name] = self._dataset(name, ohdr)
            except H5Error:
                raise
        self.order = sorted(self._datasets, key=lambda k: self._datasets[k]["addr"])

    # ---- low level
    def _cstr(self, off):
        end = self.buf.index(b"\0", off)
        return self.buf[off:end].decode("ascii")

    def _local_heap(self, addr):
        b = self.buf
        if b[addr:addr + 4] != b"HEAP":
            raise H5Error("local heap signature missing")
        return struct.unpack_from("<Q", b, addr + 24)[0] + self.base  # address of the data segment

    def _walk_btree(self, addr):
        b = self.buf
        if b[addr:addr + 4] != b"TREE":
            raise H5Error("B-tree signature missing")
        node_type, level, used = b[addr + 4], b[addr + 5], struct.unpack_from("<H", b, addr + 6)[0]
        if node_type != 0:
            raise H5Error("not a group B-tree")
        pos = addr + 24  # after signature(4) type(1) level(1) entries(2) left(8) right(8): key0, child0, key1, ...
        for i in range(used):
            child = struct.unpack_from("<Q", b, pos + 8 + i * 16)[0] + self.base
            if level > 0:
                yield from self._walk_btree(child)
            else:
                yield from self._snod(child)

    def _snod(self, addr):
        b = self.buf
        if b[addr:addr + 4] != b"SNOD":
            raise H5Error("symbol table node signature missing")
        n = struct.unpack_from("<H", b, addr + 6)[0]
        for i in range(n):
            e = addr + 8 + i * 40
            name_off, ohdr = struct.unpack_from("<QQ", b, e)
            yield name_off, ohdr + self.base

    def _dataset(self, name, ohdr):
        b = self.buf
        version, nmsg = b[ohdr], struct.unpack_from("<H", b, ohdr + 2)[0]
        if version != 1:
            raise H5Error(f"{name}: object header version {version} not supported")
        hdr_size = struct.unpack_from("<I", b, ohdr + 8)[0]
        pos, end = ohdr + 16, ohdr + 16 + hdr_size
        info = {"name": name}
        blocks = [(pos, end)]
        seen = 0
        while blocks and seen < nmsg:
            pos, end = blocks.pop(0)
            while pos + 8 <= end and seen < nmsg:
                mtype, msize = struct.unpack_from("<HH", b, pos)
                body = pos + 8
                seen += 1
                if mtype == 0x0001:  # dataspace
                    ver, rank = b[body], b[body + 1]
                    doff = body + (8 if ver == 1 else 4)
                    info["dims"] = tuple(struct.unpack_from("<Q", b, doff + 8 * i)[0] for i in range(rank))
                    info["dims_off"] = doff
                elif mtype == 0x0003:  # datatype
                    cls, size = b[body] & 0x0F, struct.unpack_from("<I", b, body + 4)[0]
                    if b[body + 1] & 1:
                        raise H5Error(f"{name}: big-endian data not supported")
                    if cls == 1:
                        info["dtype"] = np.dtype(f"<f{size}")
                    elif cls == 0:
                        signed = (b[body + 1] >> 3) & 1
                        info["dtype"] = np.dtype(f"<{'i' if signed else 'u'}{size}")
                    else:
                        raise H5Error(f"{name}: datatype class {cls} not supported")
                elif mtype == 0x0008:  # data layout
                    ver = b[body]
                    if ver != 3 or b[body + 1] != 1:
                        raise H5Error(f"{name}: only contiguous (layout v3 class 1) datasets are supported")
                    info["addr"], info["size"] = struct.unpack_from("<QQ", b, body + 2)
                    info["addr_off"], info["size_off"] = body + 2, body + 10
                elif mtype == 0x0010:  # header continuation
                    caddr, clen = struct.unpack_from("<QQ", b, body)
                    blocks.append((caddr + self.base, caddr + self.base + clen))
                elif mtype == 0x000B:
                    raise H5Error(f"{name}: filtered (compressed) data not supported")
                pos = body + msize
        for k in ("dims", "dtype", "addr"):
            if k not in info:
                raise H5Error(f"{name}: not a simple contiguous dataset ({k} message missing)")
        if info["addr"] == UNDEF:
            raise H5Error(f"{name}: no data allocated")
        return info

    # ---- public
    def keys(self):
        return list(self._datasets)

    def __contains__(self, name):
        return name in self._datasets

    def shape(self, name):
        return self._datasets[name]["dims"]

    def __getitem__(self, name):
        d = self._datasets.get(name)
        if d is None:
            raise KeyError(name)
        count = int(np.prod(d["dims"])) if d["dims"] else 1
        if count * d["dtype"].itemsize != d["size"]:
            raise H5Error(f"{name}: stored size {d['size']} does not match shape {d['dims']}")
        return np.frombuffer(self.buf, dtype=d["dtype"], count=count, offset=d["addr"] + self.base).reshape(d["dims"]).copy()


def write_h5(path, arrays):
    """Write `arrays` (name -> numpy array of little-endian float / int data, at most 8 of them) as contiguous datasets in
    the root group, in the same on-disk format class as the reference's files (superblock 0, symbol-table root group,
    version-1 object headers) -- what Data/RHS/generate_rhs.py / python_fem.ipynb do through h5py."""
    names = sorted(arrays)
    if not 0 < len(names) <= 8:
        raise H5Error("1..8 datasets")
    U = struct.pack("<Q", UNDEF)
    # ---- local heap data segment: "" at offset 0, then the names, 8-byte aligned
    seg, name_off = bytearray(8), {}
    for n in names:
        name_off[n] = len(seg)
        raw = n.encode("ascii") + b"\0"
        seg += raw + b"\0" * (-len(raw) % 8)
    seg += b"\0" * (-len(seg) % 8) + struct.pack("<QQ", 1, 16)  # one free block closing the segment
    # ---- addresses
    SB, ROOT_OH, BTREE, HEAP = 0, 96, 136, 680
    heap_data = HEAP + 32
    snod = heap_data + len(seg)
    oh0 = snod + 8 + 8 * 40
    oh_size = 16 + (8 + 8 + 8 * 4) + (8 + 24) + (8 + 24)  # prefix, dataspace (rank <= 4), datatype, layout
    data0 = (oh0 + oh_size * len(names) + 2047) // 2048 * 2048
    out = bytearray(data0)
    info, pos = {}, data0
    for n in names:
        a = np.ascontiguousarray(arrays[n])
        if a.dtype.kind not in "fiu" or a.ndim < 1 or a.ndim > 4:
            raise H5Error(f"{n}: unsupported array")
        a = a.astype(a.dtype.newbyteorder("<"))
        info[n] = (a, pos)
        pos += a.nbytes + (-a.nbytes % 8)
    eof = pos
    # ---- superblock + root symbol table entry
    out[0:8] = b"\x89HDF\r\n\x1a\n"
    out[8:16] = bytes([0, 0, 0, 0, 0, 8, 8, 0])
    struct.pack_into("<HHI", out, 16, 4, 16, 0)
    struct.pack_into("<QQQQ", out, 24, 0, UNDEF, eof, UNDEF)
    struct.pack_into("<QQII", out, 56, 0, ROOT_OH, 1, 0)
    struct.pack_into("<QQ", out, 80, BTREE, HEAP)
    # ---- root object header: one symbol-table message
    struct.pack_into("<BBHII", out, ROOT_OH, 1, 0, 1, 1, 24)
    struct.pack_into("<HHBBBB", out, ROOT_OH + 16, 0x0011, 16, 0, 0, 0, 0)
    struct.pack_into("<QQ", out, ROOT_OH + 24, BTREE, HEAP)
    # ---- B-tree (one leaf entry -> the symbol table node) and local heap
    out[BTREE:BTREE + 4] = b"TREE"
    struct.pack_into("<BBH", out, BTREE + 4, 0, 0, 1)
    out[BTREE + 8:BTREE + 24] = U + U
    struct.pack_into("<QQQ", out, BTREE + 24, 0, snod, name_off[names[-1]])
    out[HEAP:HEAP + 4] = b"HEAP"
    struct.pack_into("<QQQ", out, HEAP + 8, len(seg), len(seg) - 16, heap_data)
    out[heap_data:heap_data + len(seg)] = seg
    # ---- symbol table node + dataset object headers
    out[snod:snod + 4] = b"SNOD"
    struct.pack_into("<BBH", out, snod + 4, 1, 0, len(names))
    for i, n in enumerate(names):
        a, addr = info[n]
        oh = oh0 + i * oh_size
        struct.pack_into("<QQII", out, snod + 8 + i * 40, name_off[n], oh, 0, 0)
        struct.pack_into("<BBHII", out, oh, 1, 0, 3, 1, oh_size - 16)
        m = oh + 16
        struct.pack_into("<HHBBBB", out, m, 0x0001, 8 + 8 * 4, 0, 0, 0, 0)  # dataspace v1
        struct.pack_into("<BBBB", out, m + 8, 1, a.ndim, 0, 0)
        for k, dim in enumerate(a.shape):
            struct.pack_into("<Q", out, m + 16 + 8 * k, dim)
        m += 8 + 8 + 8 * 4
        struct.pack_into("<HHBBBB", out, m, 0x0003, 24, 1, 0, 0, 0)  # datatype v1
        bits = 8 * a.dtype.itemsize
        if a.dtype.kind == "f":
            exp, man = (11, 52) if bits == 64 else (8, 23)
            struct.pack_into("<BBBBI", out, m + 8, 0x11, 0x20, bits - 1, 0, a.dtype.itemsize)
            struct.pack_into("<HHBBBBI", out, m + 16, 0, bits, man, exp, 0, man, (1 << (exp - 1)) - 1)
        else:
            struct.pack_into("<BBBBI", out, m + 8, 0x10, 0x08 if a.dtype.kind == "i" else 0, 0, 0, a.dtype.itemsize)
            struct.pack_into("<HH", out, m + 16, 0, bits)
        m += 8 + 24
        struct.pack_into("<HHBBBB", out, m, 0x0008, 24, 0, 0, 0, 0)  # contiguous layout v3
        struct.pack_into("<BBQQ", out, m + 8, 3, 1, addr, a.nbytes)
    for n in names:
        a, addr = info[n]
        out += b"\0" * (addr - len(out)) + a.tobytes()
    out += b"\0" * (eof - len(out))
    with open(path, "wb") as fh:
        fh.write(bytes(out))
    return path
