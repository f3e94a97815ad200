"""TEST INFRASTRUCTURE: a tiny HDF5 *writer* (old-style groups, contiguous float32 datasets, superblock version 0) used only to
build Keras-layout ``.weights.h5`` fixtures for the h5lite reader -- there is no h5py / libhdf5 in this image.  It follows the same
published format specification as the reader, so a round trip through both checks the group / B-tree / heap traversal and the
Keras path mapping, NOT conformance with libhdf5 (that is checked against a real libhdf5-written file in test_h5lite_cpu.py)."""
import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class _Buf:
    def __init__(self):
        self.b = bytearray()

    def alloc(self, data: bytes) -> int:
        while len(self.b) % 8:
            self.b.append(0)
        addr = len(self.b)
        self.b += data
        return addr


def _msg(mtype: int, body: bytes) -> bytes:
    body = body + b"\0" * (-len(body) % 8)
    return struct.pack("<HHB3x", mtype, len(body), 0) + body


def _header(msgs) -> bytes:
    payload = b"".join(msgs)
    return struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(payload)) + payload


def _dataset(buf: _Buf, arr: np.ndarray) -> int:
    arr = np.ascontiguousarray(arr, dtype="<f4")
    data = buf.alloc(arr.tobytes())
    space = struct.pack("<BBB5x", 1, arr.ndim, 0) + b"".join(struct.pack("<Q", d) for d in arr.shape)
    dtype = struct.pack("<BBBBI", 0x11, 0x20, 0x1F, 0x00, 4) + struct.pack("<HHBBBBI", 0, 32, 23, 8, 0, 23, 127)
    layout = struct.pack("<BBQQ", 3, 1, data, arr.nbytes)
    return buf.alloc(_header([_msg(1, space), _msg(3, dtype), _msg(8, layout)]))


def _group(buf: _Buf, children: dict) -> tuple:
    """children: name -> ndarray | dict.  Returns (object header address, btree address, heap address)."""
    entries = []
    for name in sorted(children):
        v = children[name]
        if isinstance(v, dict):
            hdr, bt, hp = _group(buf, v)
            entries.append((name, hdr, 1, struct.pack("<QQ", bt, hp)))
        else:
            entries.append((name, _dataset(buf, v), 0, b"\0" * 16))
    heap_data = bytearray(b"\0" * 8)
    offs = []
    for name, *_ in entries:
        offs.append(len(heap_data))
        heap_data += name.encode() + b"\0"
        heap_data += b"\0" * (-len(heap_data) % 8)
    heap_seg = buf.alloc(bytes(heap_data))
    heap = buf.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap_data), UNDEF, heap_seg))
    snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(entries))
    for (name, hdr, ctype, scratch), off in zip(entries, offs):
        snod += struct.pack("<QQII", off, hdr, ctype, 0) + scratch
    snod_addr = buf.alloc(snod)
    last_key = offs[-1] if offs else 0
    tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF) + struct.pack("<QQQ", 0, snod_addr, last_key)
    bt = buf.alloc(tree)
    hdr = buf.alloc(_header([_msg(0x11, struct.pack("<QQ", bt, heap))]))
    return hdr, bt, heap


def write_h5(path, tree: dict, userblock: int = 0) -> None:
    """tree: nested dict of float arrays, e.g. {"conv_pre": {"vars": {"0": kernel, "1": bias}}}."""
    buf = _Buf()
    buf.b += b"\0" * 96                                  # superblock placeholder
    hdr, bt, heap = _group(buf, tree)
    sb = b"\x89HDF\r\n\x1a\n" + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, 64, 16, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, len(buf.b), UNDEF)
    sb += struct.pack("<QQII", 0, hdr, 1, 0) + struct.pack("<QQ", bt, heap)
    assert len(sb) == 96
    buf.b[:96] = sb
    with open(path, "wb") as f:
        f.write(b"\0" * userblock + bytes(buf.b))
