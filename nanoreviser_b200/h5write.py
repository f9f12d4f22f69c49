"""Minimal HDF5 WRITER: Keras-layout weight files of the training path (train.save_weights, the counterpart of
Model.save_weights in NanoReviser_train.py:173-174) and the test fixtures.

There is no h5py / libhdf5 in this image, and the fixtures the reference ships are all of one kind (Albacore 2.0.2 single-read
fast5, deflate).  This writer emits the same HDF5 subset the product's readers parse (nanoreviser_b200/h5mini.py and
csrc/nrv_ingest.cpp) so that variant inputs can be synthesised from a real fixture: other filters (VBZ,
id 32020, chunk bytes produced by the reference's own plugin), legacy event tables with float start / length columns
(nanorev_fast5_handeler.py:65-75), multi-read containers.

Structures written: superblock v0, object header v1, symbol-table groups (one B-tree v1 node over SNODs of <= 8 entries, local
heap), dataspace v1, datatypes int / float / fixed string / compound v1, layout v3 contiguous and rank-1 chunked (B-tree v1 node
type 1), filter pipeline v1, attribute v1 (scalars and fixed strings).
"""
from __future__ import annotations

import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIG = b"\x89HDF\r\n\x1a\n"


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * (-len(b) % 8)


def dtype_msg(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.names:
        out = struct.pack("<BBBBI", (1 << 4) | 6, len(dt.names) & 0xFF, len(dt.names) >> 8, 0, dt.itemsize)
        for nm in dt.names:
            sub, off = dt.fields[nm][0], dt.fields[nm][1]
            out += _pad8(nm.encode("ascii") + b"\x00")
            out += struct.pack("<IB3xII16x", off, 0, 0, 0)
            out += dtype_msg(sub)
        return out
    if dt.kind in "iu":
        return struct.pack("<BBBBIHH", (1 << 4) | 0, 0x08 if dt.kind == "i" else 0, 0, 0, dt.itemsize, 0, dt.itemsize * 8)
    if dt.kind == "f":
        if dt.itemsize == 4:
            return struct.pack("<BBBBIHHBBBBI", (1 << 4) | 1, 0x20, 31, 0, 4, 0, 32, 23, 8, 0, 23, 127)
        return struct.pack("<BBBBIHHBBBBI", (1 << 4) | 1, 0x20, 63, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
    if dt.kind == "S":
        return struct.pack("<BBBBI", (1 << 4) | 3, 0, 0, 0, dt.itemsize)
    raise ValueError("dtype %r" % dt)


def dataspace_msg(shape) -> bytes:
    shape = tuple(shape)
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", d) for d in shape)


def attr_msg(name: str, value) -> bytes:
    if isinstance(value, (bytes, str)):
        raw = value.encode() if isinstance(value, str) else value
        arr = np.array(raw, dtype="S%d" % max(len(raw), 1))
    else:
        arr = np.asarray(value)
    nm = name.encode("utf8") + b"\x00"
    dt, ds = dtype_msg(arr.dtype), dataspace_msg(arr.shape)
    return (struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds)) + _pad8(nm) + _pad8(dt) + _pad8(ds) + arr.tobytes())


class Writer:
    def __init__(self):
        self.buf = bytearray(96)          # superblock, filled in by finish()

    def alloc(self, data: bytes, align: int = 8) -> int:
        self.buf += b"\x00" * (-len(self.buf) % align)
        addr = len(self.buf)
        self.buf += data
        return addr

    def object_header(self, msgs) -> int:
        body = b""
        for mtype, data in msgs:
            data = _pad8(data)
            body += struct.pack("<HHB3x", mtype, len(data), 0) + data
        return self.alloc(struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body)

    def group(self, entries: dict, attrs: dict | None = None) -> int:
        """entries: name -> object header address"""
        names = sorted(entries)
        heap = bytearray(b"\x00" * 8)                      # offset 0: the empty name
        offs = {}
        for nm in names:
            offs[nm] = len(heap)
            heap += _pad8(nm.encode("utf8") + b"\x00")
        heap_data = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), UNDEF, heap_data))
        snods = []
        for i in range(0, max(len(names), 1), 8):
            part = names[i:i + 8]
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
            for nm in part:
                body += struct.pack("<QQII16x", offs[nm], entries[nm], 0, 0)
            body += b"\x00" * (40 * (8 - len(part)))
            snods.append((self.alloc(body), offs[part[-1]] if part else 0))
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF) + struct.pack("<Q", 0)
        for addr, last_key in snods:
            tree += struct.pack("<QQ", addr, last_key)
        tree_addr = self.alloc(tree)
        msgs = [(0x0011, struct.pack("<QQ", tree_addr, heap_addr))]
        for k, v in (attrs or {}).items():
            msgs.append((0x000C, attr_msg(k, v)))
        return self.object_header(msgs)

    def dataset(self, arr, chunk: int | None = None, filt=None, attrs: dict | None = None, encode=None) -> int:
        """arr: numpy array (rank <= 1 for chunked).  chunk: elements per chunk (None = contiguous).
        filt: None | ("deflate", level) | (filter_id, [cd_values], name) with encode(chunk_bytes) -> payload bytes"""
        arr = np.ascontiguousarray(arr)
        msgs = [(0x0001, dataspace_msg(arr.shape)), (0x0003, dtype_msg(arr.dtype))]
        raw = arr.tobytes()
        if chunk is None:
            addr = self.alloc(raw) if raw else UNDEF
            msgs.append((0x0008, struct.pack("<BBQQ", 3, 1, addr, len(raw))))
        else:
            assert arr.ndim == 1
            es = arr.dtype.itemsize
            if filt is not None:
                if filt[0] == "deflate":
                    fid, cd, name = 1, [filt[1]], b""
                    enc = lambda b: zlib.compress(b, filt[1])          # noqa: E731
                else:
                    fid, cd, name = filt[0], list(filt[1]), (filt[2].encode() + b"\x00" if len(filt) > 2 and filt[2] else b"")
                    enc = encode
                pm = struct.pack("<BB6x", 1, 1) + struct.pack("<HHHH", fid, len(_pad8(name)) if name else 0, 1, len(cd))
                pm += _pad8(name) if name else b""
                pm += b"".join(struct.pack("<I", c) for c in cd) + (b"\x00" * 4 if len(cd) % 2 else b"")
                msgs.append((0x000B, pm))
            else:
                enc = lambda b: b                                      # noqa: E731
            keys = []
            for off in range(0, len(arr), chunk):
                part = raw[off * es:(off + chunk) * es]
                part = part + b"\x00" * (chunk * es - len(part))       # HDF5 stores (and filters) full chunks
                payload = enc(part)
                keys.append((len(payload), off, self.alloc(payload)))
            node = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(keys), UNDEF, UNDEF)
            for size, off, addr in keys:
                node += struct.pack("<IIQQ", size, 0, off, 0) + struct.pack("<Q", addr)
            node += struct.pack("<IIQQ", 0, 0, (len(arr) + chunk - 1) // chunk * chunk, 0)
            baddr = self.alloc(node) if keys else UNDEF
            msgs.append((0x0008, struct.pack("<BBBQII", 3, 2, 2, baddr, chunk, es)))
        for k, v in (attrs or {}).items():
            msgs.append((0x000C, attr_msg(k, v)))
        return self.object_header(msgs)

    def finish(self, root_addr: int) -> bytes:
        sb = SIG + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack("<QQII16x", 0, root_addr, 0, 0)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def write_tree(tree: dict) -> bytes:
    """tree: nested dict; a value is a dict (group; the key '__attrs__' holds its attributes), or a tuple
    (array, kwargs) / an array (dataset)."""
    w = Writer()

    def build(node):
        if isinstance(node, dict):
            entries = {k: build(v) for k, v in node.items() if k != "__attrs__"}
            return w.group(entries, node.get("__attrs__"))
        if isinstance(node, tuple):
            return w.dataset(node[0], **node[1])
        return w.dataset(node)

    return w.finish(build(tree))
