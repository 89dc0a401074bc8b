"""h5lite -- a small pure-Python reader/writer for the HDF5 subset PFFDTD uses.

Neither h5py nor libhdf5 exists in this image, yet the drop-in boundary of the simulation step is a
folder of HDF5 files: ``sim_consts.h5``, ``vox_out.h5``, ``comms_out.h5``,
``sim_mats.h5`` in, ``sim_outs.h5`` out (reference readers: c_cuda/fdtd_data.h:145-499,
python/fdtd/sim_fdtd.py:59-133; writers: fdtd_data.h:928-980, sim_fdtd.py:688-696).

Reader: superblock v0/v1 (optionally behind a user block), old-style groups (v1 B-tree + SNOD + local
heap), v1 object headers incl. continuation blocks, dataspace v1/v2, datatypes fixed-point / IEEE float /
enum-over-int (h5py's ``bool``), layout v3 compact / contiguous / chunked with deflate + shuffle filters.
Writer: superblock v0, one flat root group, contiguous or chunked+deflate datasets, numpy bool as the
same int8 enum h5py produces.  The byte layout of written files follows what libhdf5 emits for
``data/materials/*.h5`` (the only genuine fixtures available), so h5py/libhdf5 can read them back.

The API mimics the slice of h5py the reference touches: ``File(path, mode)``, ``f[name][...]``,
``f[name][()]``, ``f.create_dataset(name, data=..., compression=..., compression_opts=...)``,
``del f[name]``, ``f.close()``, ``name in f``, ``f.keys()``.
"""
from __future__ import annotations

import struct
import zlib
from pathlib import Path

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


# ----------------------------------------------------------------------------------------------
# reader
# ----------------------------------------------------------------------------------------------
class _Reader:
    def __init__(self, buf: bytes):
        self.b = buf
        self.base = 0
        self._parse_superblock()

    # -- primitives
    def u(self, off, n):
        return int.from_bytes(self.b[off:off + n], "little")

    def _parse_superblock(self):
        b = self.b
        off = 0
        while True:
            if b[off:off + 8] == SIG:
                break
            off = 512 if off == 0 else off * 2
            if off + 8 > len(b):
                raise H5Error("not an HDF5 file (no superblock signature)")
        self.sb_off = off
        ver = b[off + 8]
        if ver not in (0, 1):
            raise H5Error(f"superblock version {ver} not supported (only the libver='earliest' layout)")
        self.O = b[off + 13]
        self.L = b[off + 14]
        if self.O != 8 or self.L != 8:
            raise H5Error("only 8-byte offsets/lengths supported")
        self.leaf_k = self.u(off + 16, 2)
        self.int_k = self.u(off + 18, 2)
        p = off + 24 + (4 if ver == 1 else 0)
        self.base = self.u(p, 8)
        self.eof = self.u(p + 16, 8)
        # root symbol table entry
        p += 32
        self.root_header = self.u(p + 8, 8)
        cache_type = self.u(p + 16, 4)
        if cache_type == 1:
            self.root_btree = self.u(p + 24, 8)
            self.root_heap = self.u(p + 32, 8)
        else:
            msgs = self._object_header(self.root_header)
            st = [m for m in msgs if m[0] == 0x11]
            if not st:
                raise H5Error("root group has no symbol table (new-style groups unsupported)")
            self.root_btree = self.u(st[0][1], 8)
            self.root_heap = self.u(st[0][1] + 8, 8)

    def a(self, addr):
        """file address -> buffer offset"""
        return addr + self.base

    # -- groups
    def links(self):
        heap = self.a(self.root_heap)
        if self.b[heap:heap + 4] != b"HEAP":
            raise H5Error("bad local heap")
        heap_data = self.a(self.u(heap + 24, 8))
        out = {}
        self._walk_group_btree(self.a(self.root_btree), heap_data, out)
        return out

    def _walk_group_btree(self, off, heap_data, out):
        b = self.b
        if b[off:off + 4] != b"TREE":
            raise H5Error("bad group B-tree node")
        level = b[off + 5]
        used = self.u(off + 6, 2)
        p = off + 24 + 8  # skip key0
        for _ in range(used):
            child = self.a(self.u(p, 8))
            p += 16
            if level > 0:
                self._walk_group_btree(child, heap_data, out)
            else:
                if b[child:child + 4] != b"SNOD":
                    raise H5Error("bad symbol table node")
                nsym = self.u(child + 6, 2)
                q = child + 8
                for _ in range(nsym):
                    name_off = self.u(q, 8)
                    hdr = self.u(q + 8, 8)
                    s = heap_data + name_off
                    e = b.index(b"\0", s)
                    out[b[s:e].decode()] = hdr
                    q += 40

    # -- object headers
    def _object_header(self, addr):
        """returns list of (type, data_offset, size, flags)"""
        b = self.b
        off = self.a(addr)
        if b[off] != 1:
            raise H5Error(f"object header version {b[off]} not supported")
        nmsgs = self.u(off + 2, 2)
        size = self.u(off + 8, 4)
        blocks = [(off + 16, size)]
        msgs = []
        bi = 0
        while bi < len(blocks) and len(msgs) < nmsgs:
            p, sz = blocks[bi]
            end = p + sz
            while p + 8 <= end and len(msgs) < nmsgs:
                mtype = self.u(p, 2)
                msize = self.u(p + 2, 2)
                mflags = b[p + 4]
                d = p + 8
                if mtype == 0x10:  # continuation
                    blocks.append((self.a(self.u(d, 8)), self.u(d + 8, 8)))
                msgs.append((mtype, d, msize, mflags))
                p = d + msize
            bi += 1
        return msgs

    # -- datatype
    def _datatype(self, off):
        b = self.b
        cls = b[off] & 0x0F
        ver = b[off] >> 4
        bits = self.u(off + 1, 3)
        size = self.u(off + 4, 4)
        if bits & 1 and cls in (0, 1) and size > 1:
            order = ">"
        else:
            order = "<"
        if cls == 0:
            signed = (bits >> 3) & 1
            return np.dtype(f"{order}{'i' if signed else 'u'}{size}"), 8 + 4
        if cls == 1:
            if size not in (2, 4, 8):
                raise H5Error("unsupported float size")
            return np.dtype(f"{order}f{size}"), 8 + 12
        if cls == 8:
            nmemb = bits & 0xFFFF
            base, blen = self._datatype(off + 8)
            p = off + 8 + blen
            names = []
            for _ in range(nmemb):
                e = b.index(b"\0", p)
                names.append(b[p:e].decode())
                n = e - p + 1
                if ver < 3:
                    n = (n + 7) // 8 * 8
                p += n
            vals = np.frombuffer(b, dtype=base, count=nmemb, offset=p)
            p += nmemb * base.itemsize
            self._last_enum = dict(zip(names, vals.tolist()))
            return base, p - off
        raise H5Error(f"datatype class {cls} not supported")

    def read_dataset(self, addr):
        msgs = self._object_header(addr)
        b = self.b
        shape = None
        dtype = None
        layout = None
        filters = []
        is_bool = False
        for mtype, d, msize, _ in msgs:
            if mtype == 0x01:
                ver = b[d]
                rank = b[d + 1]
                p = d + (8 if ver == 1 else 4)
                shape = tuple(self.u(p + 8 * i, 8) for i in range(rank))
            elif mtype == 0x03:
                self._last_enum = None
                dtype, _ = self._datatype(d)
                if self._last_enum is not None and set(self._last_enum) == {"FALSE", "TRUE"}:
                    is_bool = True
            elif mtype == 0x08:
                ver = b[d]
                if ver in (1, 2):  # pre-1.6 encoding (e.g. MATLAB v7.3 files)
                    nd, cls = b[d + 1], b[d + 2]
                    p = d + 8
                    addr = None
                    if cls != 0:
                        addr = self.u(p, 8)
                        p += 8
                    dims = tuple(self.u(p + 4 * i, 4) for i in range(nd))
                    p += 4 * nd
                    if cls == 0:
                        layout = ("compact", p + 4, self.u(p, 4))
                    elif cls == 1:
                        layout = ("contiguous", addr, None)
                    else:
                        layout = ("chunked", addr, dims)
                    continue
                if ver != 3:
                    raise H5Error(f"data layout message version {ver} not supported")
                cls = b[d + 1]
                if cls == 0:
                    n = self.u(d + 2, 2)
                    layout = ("compact", d + 4, n)
                elif cls == 1:
                    layout = ("contiguous", self.u(d + 2, 8), self.u(d + 10, 8))
                elif cls == 2:
                    nd = b[d + 2]
                    bt = self.u(d + 3, 8)
                    dims = tuple(self.u(d + 11 + 4 * i, 4) for i in range(nd))
                    layout = ("chunked", bt, dims)
                else:
                    raise H5Error("unknown layout class")
            elif mtype == 0x0B:
                ver = b[d]
                nf = b[d + 1]
                p = d + (8 if ver == 1 else 2)
                for _ in range(nf):
                    fid = self.u(p, 2)
                    if ver == 1 or fid >= 256:
                        nlen = self.u(p + 2, 2)
                        p += 4
                    else:
                        nlen = 0
                        p += 2
                    ncd = self.u(p + 2, 2)
                    p += 4
                    p += (nlen + 7) // 8 * 8 if ver == 1 else nlen
                    cd = [self.u(p + 4 * i, 4) for i in range(ncd)]
                    p += 4 * ncd
                    if ver == 1 and ncd % 2:
                        p += 4
                    filters.append((fid, cd))
        if shape is None or dtype is None or layout is None:
            raise H5Error("object is not a simple dataset")
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if layout[0] == "compact":
            arr = np.frombuffer(b, dtype=dtype, count=count, offset=layout[1]).copy()
        elif layout[0] == "contiguous":
            if layout[1] == UNDEF or count == 0:
                arr = np.zeros(count, dtype=dtype)
            else:
                arr = np.frombuffer(b, dtype=dtype, count=count, offset=self.a(layout[1])).copy()
        else:
            arr = self._read_chunked(layout[1], layout[2], shape, dtype, filters)
        arr = arr.reshape(shape).astype(dtype.newbyteorder("="), copy=False)
        if is_bool:
            arr = arr.astype(np.bool_)
        return arr

    def _read_chunked(self, btree, cdims, shape, dtype, filters):
        rank = len(shape)
        chunk = tuple(cdims[:rank])
        out = np.zeros(shape, dtype=dtype)
        if btree == UNDEF:
            return out
        for offs, addr, nbytes, fmask in self._walk_chunk_btree(self.a(btree), rank):
            raw = self.b[self.a(addr):self.a(addr) + nbytes]
            for i, (fid, cd) in reversed(list(enumerate(filters))):
                if fmask & (1 << i):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else dtype.itemsize
                    n = len(raw) // es
                    raw = np.frombuffer(raw, dtype=np.uint8, count=n * es).reshape(es, n).T.tobytes() + raw[n * es:]
                elif fid == 3:  # fletcher32: drop the checksum
                    raw = raw[:-4]
                else:
                    raise H5Error(f"filter {fid} not supported")
            c = np.frombuffer(raw, dtype=dtype, count=int(np.prod(chunk))).reshape(chunk)
            sl_out = tuple(slice(o, min(o + c_, s)) for o, c_, s in zip(offs, chunk, shape))
            sl_in = tuple(slice(0, s.stop - s.start) for s in sl_out)
            out[sl_out] = c[sl_in]
        return out

    def _walk_chunk_btree(self, off, rank):
        b = self.b
        if b[off:off + 4] != b"TREE" or b[off + 4] != 1:
            raise H5Error("bad chunk B-tree node")
        level = b[off + 5]
        used = self.u(off + 6, 2)
        keysz = 8 + 8 * (rank + 1)
        p = off + 24
        for _ in range(used):
            nbytes = self.u(p, 4)
            fmask = self.u(p + 4, 4)
            offs = tuple(self.u(p + 8 + 8 * i, 8) for i in range(rank))
            child = self.u(p + keysz, 8)
            p += keysz + 8
            if level > 0:
                yield from self._walk_chunk_btree(self.a(child), rank)
            else:
                yield offs, child, nbytes, fmask


# ----------------------------------------------------------------------------------------------
# writer
# ----------------------------------------------------------------------------------------------
def _pad8(b: bytes) -> bytes:
    return b + b"\0" * (-len(b) % 8)


def _dtype_msg(dt: np.dtype, as_bool=False) -> bytes:
    """datatype message body (version 1 encodings, as libhdf5 writes for native types)"""
    if as_bool:
        base = _dtype_msg(np.dtype("i1"))
        body = bytes([0x18, 2, 0, 0]) + struct.pack("<I", 1) + base
        body += _pad8(b"FALSE\0") + _pad8(b"TRUE\0") + bytes([0, 1])
        return body
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00
        return bytes([0x10, bits0, 0, 0]) + struct.pack("<I", dt.itemsize) + struct.pack("<HH", 0, 8 * dt.itemsize)
    if dt.kind == "f" and dt.itemsize == 8:
        return bytes([0x11, 0x20, 0x3F, 0]) + struct.pack("<I", 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
    if dt.kind == "f" and dt.itemsize == 4:
        return bytes([0x11, 0x20, 0x1F, 0]) + struct.pack("<I", 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    raise H5Error(f"cannot write dtype {dt}")


def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    body = _pad8(body)
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


class _Alloc:
    """append-only file image with 8-byte aligned allocation"""

    def __init__(self):
        self.buf = bytearray()

    def alloc(self, n, align=8):
        pad = -len(self.buf) % align
        self.buf += b"\0" * pad
        off = len(self.buf)
        self.buf += b"\0" * n
        return off

    def put(self, off, data):
        self.buf[off:off + len(data)] = data


def _build_file(datasets: dict) -> bytes:
    """datasets: name -> (ndarray, chunks or None, gzip level or None)"""
    LEAF_K, INT_K = 4, 16
    names = sorted(datasets, key=lambda s: s.encode())
    if len(names) > 2 * LEAF_K * 2 * INT_K:
        raise H5Error("too many datasets for a single-level group B-tree")
    img = _Alloc()
    img.alloc(96)  # superblock (56 + 40-byte root entry)
    # root object header: one symbol-table message
    root_hdr = img.alloc(16 + 24)
    btree = img.alloc(24 + (2 * INT_K + 1) * 8 + 2 * INT_K * 8)
    # local heap
    heap_names = {}
    heap_data = bytearray(8)  # offset 0: empty string (root)
    for n in names:
        heap_names[n] = len(heap_data)
        heap_data += _pad8(n.encode() + b"\0")
    free_off = len(heap_data)
    heap_size = max(88, (free_off + 16 + 7) // 8 * 8)
    heap_data += struct.pack("<QQ", 1, heap_size - free_off)  # free block: next=1 (last), size
    heap_data += b"\0" * (heap_size - len(heap_data))
    heap = img.alloc(32)
    heap_data_addr = img.alloc(heap_size)
    img.put(heap, b"HEAP" + bytes(4) + struct.pack("<QQQ", heap_size, free_off, heap_data_addr))
    img.put(heap_data_addr, bytes(heap_data))

    # datasets: object header + raw data
    hdr_addr = {}
    for n in names:
        arr, chunks, gz = datasets[n]
        arr = np.asarray(arr)
        as_bool = arr.dtype == np.bool_
        store = arr.astype(np.int8) if as_bool else arr
        if store.dtype.byteorder == ">":
            store = store.astype(store.dtype.newbyteorder("<"))
        if store.ndim:
            store = np.ascontiguousarray(store)
        rank = store.ndim
        msgs = b""
        # dataspace v1
        if rank:
            dims = b"".join(struct.pack("<Q", s) for s in store.shape)
            msgs += _msg(0x01, bytes([1, rank, 1, 0, 0, 0, 0, 0]) + dims + dims)
        else:
            msgs += _msg(0x01, bytes([1, 0, 0, 0, 0, 0, 0, 0]))
        msgs += _msg(0x03, _dtype_msg(store.dtype, as_bool), flags=1)
        msgs += _msg(0x05, bytes([2, 2, 2, 1]) + struct.pack("<I", 0), flags=1)
        if chunks is not None and rank > 0 and store.size > 0:
            chunks = tuple(int(min(c, s)) if s else 1 for c, s in zip(chunks, store.shape))
            if gz is not None:
                # filter pipeline v1: deflate, client data = level
                fp = bytes([1, 1, 0, 0, 0, 0, 0, 0]) + struct.pack("<HHHH", 1, 8, 1, 1) + _pad8(b"deflate\0") \
                    + struct.pack("<I", int(gz)) + bytes(4)
                msgs += _msg(0x0B, fp, flags=1)
            # write chunks, then the chunk B-tree
            grid = [range(0, s, c) for s, c in zip(store.shape, chunks)]
            entries = []
            for offs in np.ndindex(*[len(g) for g in grid]):
                o = tuple(g[i] for g, i in zip(grid, offs))
                block = np.zeros(chunks, dtype=store.dtype)
                sl = tuple(slice(a, min(a + c, s)) for a, c, s in zip(o, chunks, store.shape))
                block[tuple(slice(0, s.stop - s.start) for s in sl)] = store[sl]
                raw = block.tobytes()
                if gz is not None:
                    raw = zlib.compress(raw, int(gz))
                addr = img.alloc(len(raw))
                img.put(addr, raw)
                entries.append((o, addr, len(raw)))
            keysz = 8 + 8 * (rank + 1)
            CH_K = 32  # indexed-storage internal node K (libhdf5 default)
            if len(entries) > 2 * CH_K:
                raise H5Error("too many chunks for a single-level chunk B-tree; use larger chunks")
            node = img.alloc(24 + (2 * CH_K + 1) * keysz + 2 * CH_K * 8)
            nb = bytearray(b"TREE" + bytes([1, 0]) + struct.pack("<H", len(entries)) + struct.pack("<QQ", UNDEF, UNDEF))
            for o, addr, nbytes in entries:
                nb += struct.pack("<II", nbytes, 0) + b"".join(struct.pack("<Q", v) for v in o) + struct.pack("<Q", 0)
                nb += struct.pack("<Q", addr)
            # final key: one past the last chunk
            nb += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", s) for s in store.shape) + struct.pack("<Q", 0)
            img.put(node, bytes(nb))
            lay = bytes([3, 2, rank + 1]) + struct.pack("<Q", node) + b"".join(struct.pack("<I", c) for c in chunks) \
                + struct.pack("<I", store.dtype.itemsize)
            msgs += _msg(0x08, lay)
        else:
            nbytes = store.size * store.dtype.itemsize
            if nbytes:
                data_addr = img.alloc(nbytes)
                img.put(data_addr, store.tobytes())
            else:
                data_addr = UNDEF
            msgs += _msg(0x08, bytes([3, 1]) + struct.pack("<QQ", data_addr, nbytes))
        # count messages by walking
        nmsg, p = 0, 0
        while p < len(msgs):
            nmsg += 1
            p += 8 + struct.unpack_from("<H", msgs, p + 2)[0]
        # pad the header with a NIL message to a comfortable size like libhdf5 (>= 256 bytes body)
        body = 256 if len(msgs) + 8 <= 256 else len(msgs)
        if body > len(msgs):
            msgs += struct.pack("<HHB3x", 0, body - len(msgs) - 8, 0) + bytes(body - len(msgs) - 8)
            nmsg += 1
        hdr = img.alloc(16 + len(msgs))
        img.put(hdr, struct.pack("<BBHII4x", 1, 0, nmsg, 1, len(msgs)) + msgs)
        hdr_addr[n] = hdr

    # symbol table nodes (sorted names, <= 2*LEAF_K per node)
    groups = [names[i:i + 2 * LEAF_K] for i in range(0, len(names), 2 * LEAF_K)] or [[]]
    snods = []
    for g in groups:
        s = img.alloc(8 + 2 * LEAF_K * 40)
        sb = bytearray(b"SNOD" + bytes([1, 0]) + struct.pack("<H", len(g)))
        for n in g:
            sb += struct.pack("<QQII16x", heap_names[n], hdr_addr[n], 0, 0)
        img.put(s, bytes(sb))
        snods.append(s)
    tb = bytearray(b"TREE" + bytes([0, 0]) + struct.pack("<H", len(groups)) + struct.pack("<QQ", UNDEF, UNDEF))
    tb += struct.pack("<Q", 0)
    for g, s in zip(groups, snods):
        tb += struct.pack("<Q", s) + struct.pack("<Q", heap_names[g[-1]] if g else 0)
    img.put(btree, bytes(tb))
    img.put(root_hdr, struct.pack("<BBHII4x", 1, 0, 1, 1, 24) + _msg(0x11, struct.pack("<QQ", btree, heap)))
    eof = len(img.buf)
    sb = SIG + bytes([0, 0, 0, 0, 0, 8, 8, 0]) + struct.pack("<HHI", LEAF_K, INT_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root_hdr, 1, 0) + struct.pack("<QQ", btree, heap)
    img.put(0, sb)
    return bytes(img.buf)


# ----------------------------------------------------------------------------------------------
# h5py-like facade
# ----------------------------------------------------------------------------------------------
class Dataset:
    def __init__(self, f: "File", name: str):
        self._f = f
        self._name = name

    def _value(self):
        return self._f._get(self._name)

    @property
    def shape(self):
        return self._value().shape

    @property
    def dtype(self):
        return self._value().dtype

    def __getitem__(self, key):
        v = self._value()
        if key is Ellipsis or key == ():
            return v[()] if v.ndim == 0 else v
        return v[key]

    def __setitem__(self, key, val):
        v = self._value().copy()
        if v.ndim == 0:
            v = np.asarray(val, dtype=v.dtype)
        else:
            v[key] = val
        self._f._set(self._name, v)

    def __array__(self, dtype=None, copy=None):
        v = self._value()
        return v.astype(dtype) if dtype is not None else v


class File:
    """h5py.File look-alike over a flat root group. Modes: 'r', 'w', 'r+', 'a'."""

    def __init__(self, path, mode="r"):
        self.path = Path(path)
        self.mode = mode
        self._cache = {}
        self._opts = {}
        self._reader = None
        self._links = {}
        self._dirty = False
        self._closed = False
        if mode in ("r", "r+") or (mode == "a" and self.path.exists()):
            self._reader = _Reader(self.path.read_bytes())
            self._links = self._reader.links()
        elif mode in ("w", "a", "x", "w-"):
            self._dirty = True
        else:
            raise ValueError(f"bad mode {mode}")

    # -- internals
    def _get(self, name):
        name = name.lstrip("/")
        if name in self._cache:
            return self._cache[name]
        if name not in self._links:
            raise KeyError(f"Unable to open object (object '{name}' doesn't exist)")
        v = self._reader.read_dataset(self._links[name])
        self._cache[name] = v
        return v

    def _set(self, name, arr, chunks=None, gz=None):
        if self.mode == "r":
            raise H5Error("file is read-only")
        name = name.lstrip("/")
        self._cache[name] = arr
        if chunks is not None or gz is not None:
            self._opts[name] = (chunks, gz)
        self._dirty = True

    # -- h5py surface
    def keys(self):
        return sorted(set(self._links) | set(self._cache))

    def __contains__(self, name):
        name = name.lstrip("/")
        return name in self._links or name in self._cache

    def __iter__(self):
        return iter(self.keys())

    def __getitem__(self, name):
        if name not in self:
            raise KeyError(f"Unable to open object (object '{name}' doesn't exist)")
        return Dataset(self, name.lstrip("/"))

    def __delitem__(self, name):
        if self.mode == "r":
            raise H5Error("file is read-only")
        name = name.lstrip("/")
        if name not in self:
            raise KeyError(name)
        self._links.pop(name, None)
        self._cache.pop(name, None)
        self._opts.pop(name, None)
        self._dirty = True

    def create_dataset(self, name, shape=None, dtype=None, data=None, compression=None, compression_opts=None,
                       chunks=None, **kw):
        name = name.lstrip("/")
        if name in self:
            raise ValueError(f"Unable to create dataset (name already exists): {name}")
        if data is None:
            data = np.zeros(shape if shape is not None else (), dtype=dtype or np.float64)
        arr = np.asarray(data)
        if dtype is not None:
            arr = arr.astype(dtype)
        if arr.dtype == object or arr.dtype.kind not in "iufb":
            raise H5Error(f"unsupported dtype {arr.dtype}")
        gz = None
        if compression in ("gzip", True) or isinstance(compression, int) and not isinstance(compression, bool):
            gz = 4 if compression_opts is None else int(compression_opts)
            if isinstance(compression, int) and not isinstance(compression, bool) and compression is not True:
                gz = int(compression)
        if arr.ndim == 0:
            gz, chunks = None, None  # scalars are never chunked (h5py raises; reference never asks)
        elif gz is not None and chunks is None:
            chunks = _guess_chunks(arr.shape, arr.dtype.itemsize)
        self._set(name, arr, chunks, gz)
        return Dataset(self, name)

    def flush(self):
        if self._dirty and self.mode != "r":
            names = self.keys()
            ds = {}
            for n in names:
                c, g = self._opts.get(n, (None, None))
                ds[n] = (self._get(n), c, g)
            self.path.write_bytes(_build_file(ds))
            self._dirty = False

    def close(self):
        if not self._closed:
            self.flush()
            self._closed = True

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def _guess_chunks(shape, itemsize):
    """few, large chunks (a single-level chunk B-tree holds 64): split the leading axis only"""
    nbytes = int(np.prod(shape)) * itemsize
    target = max(1 << 20, -(-nbytes // 48))
    row = max(1, int(np.prod(shape[1:])) * itemsize)
    lead = max(1, min(shape[0], target // row if row <= target else 1))
    while -(-shape[0] // lead) > 64:
        lead += 1
    if len(shape) > 1 and row > target and -(-shape[0] // lead) * 1 > 64:
        raise H5Error("dataset too large for h5lite chunked writer")
    return (lead,) + tuple(shape[1:])


def read_all(path) -> dict:
    """every dataset of a flat file as {name: ndarray}"""
    f = File(path, "r")
    return {k: f._get(k) for k in f.keys()}


def write_all(path, datasets: dict, compression=None):
    f = File(path, "w")
    for k, v in datasets.items():
        f.create_dataset(k, data=v, compression=("gzip" if compression is not None else None),
                         compression_opts=compression)
    f.close()
