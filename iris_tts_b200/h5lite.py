"""Minimal pure-Python HDF5 *reader* -- just enough for Keras ``.weights.h5`` files (``HiFiGANVocoder.load_weights``,
reference src/iris/vocoder.py:167-170 hands the path to ``keras.Model.load_weights``; h5py is not installed in this image).

Scope (what h5py / libhdf5 write with default settings, "earliest" file format): superblock version 0 or 1 (optionally behind a
user block), version-1 object headers with continuation blocks, old-style groups (symbol-table message -> version-1 B-tree ->
symbol-table nodes -> local heap), datasets with a simple dataspace, fixed-point or IEEE floating-point datatype, and compact or
contiguous layout (version 3 layout message).  Chunked / filtered datasets, new-style (link-message / fractal-heap) groups and
superblock versions 2-3 raise ``H5Unsupported`` with the structure named.  Follows the published "HDF5 File Format Specification
Version 1.1 / 2.0" (The HDF Group); validated against a file written by the real HDF5 library (a MATLAB v7.3 ``.mat`` from scipy's
test data, tests/test_h5lite_cpu.py).  Read-only: nothing here can write HDF5.
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Tuple, Union

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Unsupported(ValueError):
    pass


class H5File:
    """``H5File(path).datasets()`` -> {"group/sub/name": ndarray}; ``.tree()`` -> nested dict of groups and arrays."""

    def __init__(self, path):
        with open(path, "rb") as f:
            self.buf = f.read()
        self.base = self._find_superblock()
        self._parse_superblock()

    # -- low level ------------------------------------------------------------
    def _u(self, off: int, size: int) -> int:
        return int.from_bytes(self.buf[off:off + size], "little")

    def _addr(self, off: int) -> int:
        v = self._u(off, self.so)
        return UNDEF if v == (1 << (8 * self.so)) - 1 else v + self.base

    def _find_superblock(self) -> int:
        off = 0
        while off + 8 <= len(self.buf):
            if self.buf[off:off + 8] == SIGNATURE:
                return off
            off = 512 if off == 0 else off * 2
        raise ValueError("not an HDF5 file (signature not found)")

    def _parse_superblock(self) -> None:
        p = self.base + 8
        ver = self.buf[p]
        if ver not in (0, 1):
            raise H5Unsupported(f"HDF5 superblock version {ver} (only the default 'earliest' format, versions 0/1, is read)")
        self.so, self.sl = self.buf[p + 5], self.buf[p + 6]
        p += 8 + 4 + 4                      # versions/sizes (8), group leaf/internal K (4), consistency flags (4)
        if ver == 1:
            p += 4
        self.so, self.sl = int(self.so), int(self.sl)
        # Addresses in the file are relative to the base address, which libhdf5 takes to be the superblock's own offset (a file
        # with a user block stores its superblock at 512, 1024, ... and everything after it relative to that point).
        p += 4 * self.so
        # root group symbol-table entry: link name offset, object header address, cache type, reserved, scratch
        self.root_header = self._addr(p + self.so)
        cache_type = self._u(p + 2 * self.so, 4)
        self.root_cache = None
        if cache_type == 1:
            sp = p + 2 * self.so + 8
            self.root_cache = (self._addr(sp), self._addr(sp + self.so))

    # -- object headers ---------------------------------------------------------
    def _messages(self, addr: int) -> List[Tuple[int, int, int]]:
        """(type, offset of the message body, size) of every message of a version-1 object header, continuations included."""
        if self.buf[addr:addr + 4] == b"OHDR":
            raise H5Unsupported("version-2 object header (file written with libver='latest'); only the default format is read")
        if self.buf[addr] != 1:
            raise H5Unsupported(f"object header version {self.buf[addr]}")
        nmsg = self._u(addr + 2, 2)
        size = self._u(addr + 8, 4)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, left = blocks.pop(0)
            end = p + left
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize = self._u(p, 2), self._u(p + 2, 2)
                body = p + 8
                if mtype == 0x10:           # continuation: offset, length
                    blocks.append((self._addr(body), self._u(body + self.so, self.sl)))
                out.append((mtype, body, msize))
                p = body + msize
        return out

    # -- groups -------------------------------------------------------------------
    def _heap_string(self, heap_addr: int, off: int) -> str:
        if self.buf[heap_addr:heap_addr + 4] != b"HEAP":
            raise ValueError("bad local heap signature")
        data = self._addr(heap_addr + 8 + 2 * self.sl)
        end = self.buf.index(b"\x00", data + off)
        return self.buf[data + off:end].decode("utf-8")

    def _btree_entries(self, node: int, heap: int) -> Iterator[Tuple[str, int]]:
        if self.buf[node:node + 4] == b"TREE":
            ntype, level, used = self.buf[node + 4], self.buf[node + 5], self._u(node + 6, 2)
            if ntype != 0:
                raise ValueError("expected a group B-tree node")
            p = node + 8 + 2 * self.so          # skip the sibling addresses
            for i in range(used):
                child = self._addr(p + self.sl + i * (self.sl + self.so))   # key0, child0, key1, child1, ...
                yield from self._btree_entries(child, heap)
        elif self.buf[node:node + 4] == b"SNOD":
            n = self._u(node + 6, 2)
            p = node + 8
            esize = 2 * self.so + 4 + 4 + 16
            for i in range(n):
                e = p + i * esize
                yield self._heap_string(heap, self._u(e, self.so)), self._addr(e + self.so)
        else:
            raise ValueError("bad group node signature")

    def _children(self, header: int):
        for mtype, body, _ in self._messages(header):
            if mtype == 0x11:                   # symbol table: B-tree address, local heap address
                return list(self._btree_entries(self._addr(body), self._addr(body + self.so)))
            if mtype in (0x02, 0x06):           # link info / link: new-style group
                raise H5Unsupported("new-style group (link messages); only symbol-table groups are read")
        return None

    # -- datasets -------------------------------------------------------------------
    def _dataset(self, header: int) -> np.ndarray:
        dims, dtype, data = None, None, None
        for mtype, body, msize in self._messages(header):
            if mtype == 0x01:                   # dataspace
                ver, rank, flags = self.buf[body], self.buf[body + 1], self.buf[body + 2]
                p = body + (8 if ver == 1 else 4)
                dims = tuple(self._u(p + i * self.sl, self.sl) for i in range(rank))
            elif mtype == 0x03:                 # datatype
                cls = self.buf[body] & 0x0F
                bits0 = self.buf[body + 1]
                size = self._u(body + 4, 4)
                order = ">" if bits0 & 1 else "<"
                if cls == 1:
                    dtype = np.dtype(f"{order}f{size}")
                elif cls == 0:
                    dtype = np.dtype(f"{order}{'i' if bits0 & 8 else 'u'}{size}")
                else:
                    raise H5Unsupported(f"datatype class {cls} (only fixed-point and floating-point datasets are read)")
            elif mtype == 0x0B:
                raise H5Unsupported("filtered (compressed) dataset")
            elif mtype == 0x08:                 # data layout
                ver = self.buf[body]
                if ver == 3:
                    lclass = self.buf[body + 1]
                    if lclass == 0:             # compact: size, then the raw data inside the header
                        data = (body + 4, self._u(body + 2, 2))
                    elif lclass == 1:           # contiguous: address, size
                        data = (self._addr(body + 2), self._u(body + 2 + self.so, self.sl))
                    else:
                        raise H5Unsupported("chunked dataset layout")
                elif ver in (1, 2):             # older writers: version, rank, class, 5 reserved, [address], rank x 4-byte dims, ...
                    rank, lclass = self.buf[body + 1], self.buf[body + 2]
                    if lclass == 1:
                        data = (self._addr(body + 8), None)
                    elif lclass == 0:
                        p2 = body + 8 + 4 * rank
                        data = (p2 + 4, self._u(p2, 4))
                    else:
                        raise H5Unsupported("chunked dataset layout")
                else:
                    raise H5Unsupported(f"data layout message version {ver}")
        if dims is None or dtype is None or data is None:
            raise ValueError("object is not a dataset")
        count = int(np.prod(dims)) if dims else 1
        addr, nbytes = data
        if addr == UNDEF or count == 0:
            return np.zeros(dims, dtype=dtype.newbyteorder("="))
        arr = np.frombuffer(self.buf, dtype=dtype, count=count, offset=addr).reshape(dims)
        return arr.astype(dtype.newbyteorder("="))

    # -- public ---------------------------------------------------------------------------
    def tree(self, header: int = None) -> Dict[str, Union[dict, np.ndarray]]:
        header = self.root_header if header is None else header
        out: Dict[str, Union[dict, np.ndarray]] = {}
        kids = self._children(header)
        if kids is None:
            raise ValueError("object is not a group")
        for name, addr in kids:
            if self._children(addr) is not None:
                out[name] = self.tree(addr)
            else:
                try:
                    out[name] = self._dataset(addr)
                except H5Unsupported as ex:
                    out[name] = ex            # reported to the caller only if that dataset is asked for
        return out

    def datasets(self) -> Dict[str, np.ndarray]:
        flat: Dict[str, np.ndarray] = {}

        def walk(prefix, node):
            for k, v in node.items():
                if isinstance(v, dict):
                    walk(prefix + k + "/", v)
                else:
                    flat[prefix + k] = v

        walk("", self.tree())
        return flat
