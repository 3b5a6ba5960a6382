"""A minimal read-only HDF5 parser: enough to pull the arrays out of a Keras `.weights.h5` checkpoint
(audiomodel.py:278-283 writes `val_loss.weights.h5` through ModelCheckpoint(save_weights_only=True)) without h5py,
which this image does not have.

Scope -- what h5py / libhdf5 write with their default settings (libver "earliest"), which is what Keras uses:
  * superblock versions 0 and 1 (and 2 / 3 for files written with libver "latest"), any user block (the signature is
    searched at 0, 512, 1024, ...), 8-byte or 4-byte offsets and lengths;
  * version-1 object headers with continuation blocks; version-2 headers ("OHDR") without creation-order tracking;
  * groups as symbol tables (v1 B-tree of SNOD nodes + local heap) or as compact link messages;
  * datasets with a simple dataspace, fixed-point or IEEE floating-point types of 1 / 2 / 4 / 8 bytes in either byte
    order, and contiguous, compact or unfiltered chunked (v1 B-tree) storage.
Anything else (compression filters, variable-length strings, dense link storage, virtual datasets ...) raises
`H5Unsupported` naming the feature, rather than returning wrong numbers.  Attributes are skipped.

    f = h5lite.File("val_loss.weights.h5");  f.keys();  f["layers/conv2d/vars/0"]  ->  numpy array
    dict(f.walk())  ->  {"layers/conv2d/vars/0": array, ...}
"""
from __future__ import annotations

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = {4: 0xFFFFFFFF, 8: 0xFFFFFFFFFFFFFFFF}


class H5Error(ValueError):
    pass


class H5Unsupported(H5Error):
    pass


class _Object:
    def __init__(self):
        self.symtab = None      # (btree address, heap address): old-style group
        self.links = None       # {name: address}: new-style compact group
        self.shape = None
        self.dtype = None
        self.layout = None      # ("contiguous", addr, size) | ("compact", bytes) | ("chunked", btree addr, chunk dims)
        self.filters = False

    @property
    def is_group(self):
        return self.symtab is not None or self.links is not None

    @property
    def is_dataset(self):
        return self.layout is not None and self.dtype is not None and self.shape is not None


class File:
    def __init__(self, path_or_bytes):
        if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
            self.buf = bytes(path_or_bytes)
        else:
            with open(path_or_bytes, "rb") as fh:
                self.buf = fh.read()
        self._read_superblock()
        self._cache = {}

    # ---- primitive readers -------------------------------------------------------------------------------------------
    def _u(self, pos, size):
        if pos < 0 or pos + size > len(self.buf):
            raise H5Error(f"read of {size} bytes at {pos} runs past the end of the file ({len(self.buf)} bytes)")
        return int.from_bytes(self.buf[pos:pos + size], "little")

    def _addr(self, pos):
        """an `offset`-sized field -> absolute file position (None when undefined)"""
        v = self._u(pos, self.O)
        return None if v == UNDEF[self.O] else v + self.base

    def _len(self, pos):
        return self._u(pos, self.L)

    # ---- superblock ----------------------------------------------------------------------------------------------------
    def _read_superblock(self):
        pos = 0
        while True:
            if self.buf[pos:pos + 8] == SIGNATURE:
                break
            pos = 512 if pos == 0 else pos * 2
            if pos + 8 > len(self.buf):
                raise H5Error("not an HDF5 file (signature not found)")
        ver = self.buf[pos + 8]
        self.sb_pos, self.sb_version = pos, ver
        if ver in (0, 1):
            self.O, self.L = self.buf[pos + 13], self.buf[pos + 14]
            p = pos + 24 + (4 if ver == 1 else 0)
            self.base = self._u(p, self.O)            # base address: every other address is relative to it
            p += 4 * self.O                            # base, free-space, end-of-file, driver-info
            # root group symbol table entry: link name offset, object header address, cache type, reserved, scratch
            self.root_addr = self._addr(p + self.O)
        elif ver in (2, 3):
            self.O, self.L = self.buf[pos + 9], self.buf[pos + 10]
            p = pos + 12
            self.base = self._u(p, self.O)
            self.root_addr = self._addr(p + 3 * self.O)
        else:
            raise H5Unsupported(f"superblock version {ver}")
        if self.O not in (4, 8) or self.L not in (4, 8):
            raise H5Unsupported(f"offset / length sizes {self.O} / {self.L}")

    # ---- object headers ------------------------------------------------------------------------------------------------
    def _object(self, addr):
        if addr in self._cache:
            return self._cache[addr]
        obj = _Object()
        if self.buf[addr:addr + 4] == b"OHDR":
            self._header_v2(addr, obj)
        else:
            self._header_v1(addr, obj)
        self._cache[addr] = obj
        return obj

    def _header_v1(self, addr, obj):
        if self.buf[addr] != 1:
            raise H5Unsupported(f"object header version {self.buf[addr]} at {addr}")
        n_msgs = self._u(addr + 2, 2)
        size = self._u(addr + 8, 4)
        blocks = [(addr + 16, size)]                   # the first block starts 8-byte aligned after the 12-byte prefix
        seen = 0
        while blocks and seen < n_msgs:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and seen < n_msgs:
                mtype, msize, flags = self._u(p, 2), self._u(p + 2, 2), self.buf[p + 4]
                body = p + 8
                seen += 1
                self._message(mtype, body, msize, flags, obj, blocks)
                p = body + msize
        return obj

    def _header_v2(self, addr, obj):
        if self.buf[addr + 4] != 2:
            raise H5Unsupported(f"object header version {self.buf[addr + 4]}")
        flags = self.buf[addr + 5]
        p = addr + 6
        if flags & 0x20:
            p += 16                                     # access, modification, change, birth times
        if flags & 0x10:
            p += 4                                      # max compact / min dense attributes
        csize = 1 << (flags & 3)
        chunk0 = self._u(p, csize)
        p += csize
        track = bool(flags & 0x04)
        blocks = [(p, chunk0)]
        while blocks:
            p, left = blocks.pop(0)
            end = p + left
            while p + 4 + (2 if track else 0) <= end:
                mtype, msize, mflags = self.buf[p], self._u(p + 1, 2), self.buf[p + 3]
                body = p + 4 + (2 if track else 0)
                if mtype == 0 and msize == 0:
                    break
                cont = []
                self._message(mtype, body, msize, mflags, obj, cont)
                for (cp, cl) in cont:                   # continuation chunks start with "OCHK" and end with a checksum
                    blocks.append((cp + 4, cl - 8))
                p = body + msize

    def _message(self, mtype, body, size, flags, obj, blocks):
        if mtype == 0x0010:                             # continuation
            blocks.append((self._addr(body), self._len(body + self.O)))
        elif mtype == 0x0011:                           # symbol table: old-style group
            obj.symtab = (self._addr(body), self._addr(body + self.O))
        elif mtype == 0x0006:                           # link (new-style compact group)
            name, target = self._link(body)
            obj.links = obj.links or {}
            if target is not None:
                obj.links[name] = target
        elif mtype == 0x0002:                           # link info: dense storage if a fractal heap is present
            lflags = self.buf[body + 1]
            p = body + 2 + (8 if lflags & 1 else 0)
            if self._addr(p) is not None:
                raise H5Unsupported("group with dense (fractal heap) link storage")
            obj.links = obj.links or {}
        elif mtype == 0x0001:
            if flags & 2:
                raise H5Unsupported("shared dataspace message")
            obj.shape = self._dataspace(body)
        elif mtype == 0x0003:
            if flags & 2:
                raise H5Unsupported("shared (committed) datatype")
            obj.dtype = self._datatype(body)
        elif mtype == 0x0008:
            obj.layout = self._layout(body)
        elif mtype == 0x000B:
            obj.filters = True

    def _link(self, body):
        ver, lflags = self.buf[body], self.buf[body + 1]
        if ver != 1:
            raise H5Unsupported(f"link message version {ver}")
        p = body + 2
        ltype = 0
        if lflags & 0x08:
            ltype = self.buf[p]
            p += 1
        if lflags & 0x04:
            p += 8
        if lflags & 0x10:
            p += 1
        nsize = 1 << (lflags & 3)
        nlen = self._u(p, nsize)
        p += nsize
        name = self.buf[p:p + nlen].decode("utf-8")
        p += nlen
        return name, (self._addr(p) if ltype == 0 else None)   # soft / external links are not followed

    def _dataspace(self, body):
        ver, rank, flags = self.buf[body], self.buf[body + 1], self.buf[body + 2]
        if ver == 1:
            p = body + 8
        elif ver == 2:
            if self.buf[body + 3] == 2:
                return None                             # null dataspace
            p = body + 4
        else:
            raise H5Unsupported(f"dataspace version {ver}")
        return tuple(self._len(p + i * self.L) for i in range(rank))

    def _datatype(self, body):
        cls, ver = self.buf[body] & 0x0F, self.buf[body] >> 4
        bits0 = self.buf[body + 1]
        size = self._u(body + 4, 4)
        order = ">" if bits0 & 1 else "<"
        if cls == 1:                                    # IEEE float
            if size not in (2, 4, 8):
                raise H5Unsupported(f"{size}-byte floating point")
            # precision / exponent layout of a non-IEEE float would sit in the properties; bfloat16 shows up here as size 2
            exp_size = self.buf[body + 8 + 5]
            if size == 2 and exp_size != 5:
                raise H5Unsupported("2-byte float that is not IEEE half (bfloat16?)")
            return np.dtype(f"{order}f{size}")
        if cls == 0:                                    # fixed point
            if size not in (1, 2, 4, 8):
                raise H5Unsupported(f"{size}-byte integer")
            return np.dtype(f"{order}{'i' if bits0 & 0x08 else 'u'}{size}")
        names = {2: "time", 3: "string", 4: "bitfield", 5: "opaque", 6: "compound", 7: "reference", 8: "enum", 9: "variable-length",
                 10: "array"}
        return ("unsupported", names.get(cls, str(cls)), ver)

    def _layout(self, body):
        ver = self.buf[body]
        if ver == 3:
            cls = self.buf[body + 1]
            if cls == 0:
                n = self._u(body + 2, 2)
                return ("compact", self.buf[body + 4:body + 4 + n])
            if cls == 1:
                return ("contiguous", self._addr(body + 2), self._len(body + 2 + self.O))
            if cls == 2:
                nd = self.buf[body + 2]
                bt = self._addr(body + 3)
                dims = tuple(self._u(body + 3 + self.O + 4 * i, 4) for i in range(nd))
                return ("chunked", bt, dims)
            raise H5Unsupported(f"layout class {cls}")
        if ver in (1, 2):
            nd, cls = self.buf[body + 1], self.buf[body + 2]
            p = body + 8
            addr = None
            if cls != 0:
                addr = self._addr(p)
                p += self.O
            dims = tuple(self._u(p + 4 * i, 4) for i in range(nd))
            p += 4 * nd
            if cls == 1:
                return ("contiguous", addr, None)
            if cls == 2:
                return ("chunked", addr, dims + (self._u(p, 4),))
            n = self._u(p, 4)
            return ("compact", self.buf[p + 4:p + 4 + n])
        raise H5Unsupported(f"data layout message version {ver}")

    # ---- groups --------------------------------------------------------------------------------------------------------
    def _children(self, obj):
        if obj.links is not None and obj.symtab is None:
            return dict(obj.links)
        bt, heap = obj.symtab
        if self.buf[heap:heap + 4] != b"HEAP":
            raise H5Error("local heap signature missing")
        data = self._addr(heap + 8 + 2 * self.L)
        out = {}
        self._group_node(bt, data, out)
        return out

    def _group_node(self, addr, heap_data, out):
        sig = self.buf[addr:addr + 4]
        if sig == b"TREE":
            if self.buf[addr + 4] != 0:
                raise H5Error("expected a group B-tree node")
            used = self._u(addr + 6, 2)                 # (the node level does not matter: children are TREE or SNOD nodes)
            p = addr + 8 + 2 * self.O                   # past the sibling pointers
            for i in range(used):
                child = self._addr(p + self.L + i * (self.L + self.O))
                self._group_node(child, heap_data, out)
        elif sig == b"SNOD":
            n = self._u(addr + 6, 2)
            esize = 2 * self.O + 24
            for i in range(n):
                e = addr + 8 + i * esize
                name_off = self._u(e, self.O)
                end = self.buf.index(b"\x00", heap_data + name_off)
                out[self.buf[heap_data + name_off:end].decode("utf-8")] = self._addr(e + self.O)
        else:
            raise H5Error(f"unknown group node signature {sig!r} at {addr}")

    # ---- datasets ------------------------------------------------------------------------------------------------------
    def _read_dataset(self, obj, name):
        if isinstance(obj.dtype, tuple):
            raise H5Unsupported(f"{name}: datatype class {obj.dtype[1]}")
        if obj.filters:
            raise H5Unsupported(f"{name}: filtered (compressed) dataset")
        shape, dt = obj.shape or (), obj.dtype
        count = int(np.prod(shape)) if shape else 1
        nbytes = count * dt.itemsize
        kind = obj.layout[0]
        if kind == "compact":
            raw = obj.layout[1][:nbytes]
        elif kind == "contiguous":
            addr = obj.layout[1]
            if addr is None:
                return np.zeros(shape, dt.newbyteorder("="))     # never written: fill value
            if addr + nbytes > len(self.buf):
                raise H5Error(f"{name}: data runs past the end of the file")
            raw = self.buf[addr:addr + nbytes]
        else:
            return self._read_chunked(obj, name)
        return np.frombuffer(raw, dtype=dt, count=count).reshape(shape).astype(dt.newbyteorder("="))

    def _read_chunked(self, obj, name):
        _, bt, cdims = obj.layout
        shape, dt = obj.shape, obj.dtype
        chunk = cdims[:-1]
        if len(chunk) != len(shape) or cdims[-1] != dt.itemsize:
            raise H5Error(f"{name}: chunk rank / element size mismatch")
        out = np.zeros(shape, dt.newbyteorder("="))
        if bt is None:
            return out
        nd = len(shape)

        def walk(addr):
            if self.buf[addr:addr + 4] != b"TREE" or self.buf[addr + 4] != 1:
                raise H5Error(f"{name}: expected a chunk B-tree node")
            level, used = self.buf[addr + 5], self._u(addr + 6, 2)
            p = addr + 8 + 2 * self.O
            ksize = 8 + 8 * (nd + 1)
            for i in range(used):
                k = p + i * (ksize + self.O)
                csize, mask = self._u(k, 4), self._u(k + 4, 4)
                offs = tuple(self._u(k + 8 + 8 * d, 8) for d in range(nd))
                child = self._addr(k + ksize)
                if level > 0:
                    walk(child)
                    continue
                if mask:
                    raise H5Unsupported(f"{name}: filtered chunk")
                block = np.frombuffer(self.buf[child:child + csize], dtype=dt).reshape(chunk)
                sl = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, chunk, shape))
                out[sl] = block[tuple(slice(0, s.stop - s.start) for s in sl)]
        walk(bt)
        return out

    # ---- public ----------------------------------------------------------------------------------------------------------
    def _resolve(self, path):
        obj = self._object(self.root_addr)
        for part in [p for p in path.split("/") if p]:
            if not obj.is_group:
                raise KeyError(path)
            kids = self._children(obj)
            if part not in kids:
                raise KeyError(path)
            obj = self._object(kids[part])
        return obj

    def keys(self, path="/"):
        obj = self._resolve(path)
        if not obj.is_group:
            raise KeyError(f"{path} is not a group")
        return sorted(self._children(obj))

    def is_group(self, path):
        return self._resolve(path).is_group

    def __contains__(self, path):
        try:
            self._resolve(path)
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        obj = self._resolve(path)
        if obj.is_group and not obj.is_dataset:
            raise KeyError(f"{path} is a group; use keys()")
        return self._read_dataset(obj, path)

    def walk(self, path=""):
        """yield (path, array) for every dataset below `path`, depth first in name order"""
        obj = self._resolve(path)
        if obj.is_dataset:
            yield path.strip("/"), self._read_dataset(obj, path)
            return
        if obj.is_group:
            for name in sorted(self._children(obj)):
                yield from self.walk(f"{path}/{name}")
