"""Test infrastructure: writes a small HDF5 file in the "earliest" on-disk format libhdf5 / h5py use by default
(version-0 superblock, version-1 object headers, groups as symbol tables: local heap + v1 B-tree + SNOD nodes, contiguous
little-endian float datasets) so that audio_training_b200.h5lite and the Keras `.weights.h5` loader can be exercised
without h5py.  The reader is validated separately on a file libhdf5 itself wrote (scipy's MATLAB v7.3 test fixture)."""
import struct

import numpy as np

O = L = 8
LEAF_K, NODE_K = 32, 16
UNDEF = 0xFFFFFFFFFFFFFFFF


class _Buf:
    def __init__(self):
        self.b = bytearray()

    def alloc(self, n, align=8):
        while len(self.b) % align:
            self.b.append(0)
        pos = len(self.b)
        self.b.extend(b"\x00" * n)
        return pos

    def put(self, pos, data):
        self.b[pos:pos + len(data)] = data


def _msg(mtype, body):
    body = body + b"\x00" * (-len(body) % 8)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body


def _header(msgs):
    data = b"".join(msgs)
    return struct.pack("<BxHII4x", 1, len(msgs), 1, len(data)) + data


def _dataset(buf, arr):
    arr = np.ascontiguousarray(arr)
    kind = {"f": 1, "i": 0, "u": 0}[arr.dtype.kind]
    raw = arr.astype(arr.dtype.newbyteorder("<")).tobytes()
    data_pos = buf.alloc(max(len(raw), 1))
    buf.put(data_pos, raw)
    space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
    if kind == 1:
        bits = {2: (15, 10, 5, 0, 10, 15), 4: (31, 23, 8, 0, 23, 127), 8: (63, 52, 11, 0, 52, 1023)}[arr.itemsize]
        sign, eloc, esize, mloc, msize, bias = bits
        dtype = struct.pack("<BBBBI", 0x11, 0x20, sign, 0, arr.itemsize) + struct.pack("<HHBBBBI", 0, 8 * arr.itemsize, eloc, esize,
                                                                                     mloc, msize, bias)
    else:
        dtype = struct.pack("<BBBBI", 0x10, 0x08 if arr.dtype.kind == "i" else 0, 0, 0, arr.itemsize) + struct.pack("<HH", 0,
                                                                                                               8 * arr.itemsize)
    layout = struct.pack("<BB", 3, 1) + struct.pack("<QQ", data_pos, len(raw))
    hdr = _header([_msg(1, space), _msg(3, dtype), _msg(8, layout)])
    pos = buf.alloc(len(hdr))
    buf.put(pos, hdr)
    return pos


def _group(buf, tree):
    """tree: {name: ndarray | dict} -> object header address"""
    kids = {}
    for name in sorted(tree):
        v = tree[name]
        kids[name] = _group(buf, v) if isinstance(v, dict) else _dataset(buf, v)
    # local heap: offset 0 is the empty string; names 8-byte aligned
    heap = bytearray(b"\x00" * 8)
    off = {}
    for name in kids:
        off[name] = len(heap)
        enc = name.encode("utf-8") + b"\x00"
        heap.extend(enc + b"\x00" * (-len(enc) % 8))
    heap_data = buf.alloc(len(heap))
    buf.put(heap_data, bytes(heap))
    heap_pos = buf.alloc(8 + 2 * L + O)
    buf.put(heap_pos, b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), UNDEF, heap_data))
    names = list(kids)
    per = 2 * LEAF_K
    chunks = [names[i:i + per] for i in range(0, len(names), per)] or [[]]
    if len(chunks) > 2 * NODE_K:
        raise ValueError("too many links for a single-level B-tree")
    snods, keys = [], [0]
    for ch in chunks:
        pos = buf.alloc(8 + per * (2 * O + 24))
        body = b"SNOD" + struct.pack("<BxH", 1, len(ch))
        for n in ch:
            body += struct.pack("<QQII16x", off[n], kids[n], 0, 0)
        buf.put(pos, body)
        snods.append(pos)
        keys.append(off[ch[-1]] if ch else 0)
    bt = buf.alloc(8 + 2 * O + (2 * NODE_K + 1) * L + 2 * NODE_K * O)
    body = b"TREE" + struct.pack("<BBH", 0, 0, len(snods)) + struct.pack("<QQ", UNDEF, UNDEF)
    for i, s in enumerate(snods):
        body += struct.pack("<QQ", keys[i], s)
    body += struct.pack("<Q", keys[-1])
    buf.put(bt, body)
    hdr = _header([_msg(0x11, struct.pack("<QQ", bt, heap_pos))])
    pos = buf.alloc(len(hdr))
    buf.put(pos, hdr)
    return pos


def write(path, tree, userblock=0):
    """tree: nested dict of numpy arrays -> an HDF5 file (optionally behind a user block of 512 * 2^n bytes)."""
    buf = _Buf()
    sb = buf.alloc(24 + 4 * O + 2 * O + 24)        # superblock v0 + root symbol table entry
    root = _group(buf, tree)
    body = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, O, L, 0) + struct.pack("<HHI", LEAF_K, NODE_K, 0)
    body += struct.pack("<QQQQ", userblock, UNDEF, len(buf.b) + userblock, UNDEF)
    body += struct.pack("<QQII16x", 0, root, 0, 0)
    buf.put(sb, body)
    with open(path, "wb") as fh:
        fh.write(b"\x00" * userblock + bytes(buf.b))
