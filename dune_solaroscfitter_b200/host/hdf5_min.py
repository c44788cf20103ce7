"""Minimal reader for the h5py-written ``.jld2`` inputs of the reference.

The reference loads ``inputs/AGSS09_high_z.jld2``, ``inputs/fluxes.jld2`` and
``inputs/earth_uncertainty_cov.jld2`` with JLD2.jl (``solarModel.jl:3-39``,
``readConfig.jl:100-119``).  Those files are plain HDF5 (superblock v0, written
by h5py: ``utils/getProductionRegions.py:24-28``, ``utils/makeTestFlux.py:71-76``)
holding flat little-endian f8 datasets in the root group.  Neither Julia nor
h5py exist in this image, so this module parses exactly that subset: v0
superblock, v1 B-tree / local heap / symbol-table root group, v1 object
headers, contiguous (or compact) layout, IEEE f8/f4 and fixed-width integers.
Anything else raises ``NotImplementedError`` rather than guessing.
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


def _u(buf, off, n):
    return int.from_bytes(buf[off:off + n], "little")


class _File:
    def __init__(self, buf: bytes):
        if buf[:8] != _SIG:
            raise ValueError("not an HDF5 file")
        if buf[8] != 0:
            raise NotImplementedError("only superblock v0 is supported")
        self.buf = buf
        self.so = buf[13]  # size of offsets
        self.sl = buf[14]  # size of lengths
        if self.so != 8 or self.sl != 8:
            raise NotImplementedError("only 8-byte offsets/lengths")
        # v0 superblock: 24 bytes fixed + 4 addresses, then root symbol-table entry
        p = 24 + 4 * self.so
        self.root_entry = self._symtab_entry(p)

    def _symtab_entry(self, p):
        name_off = _u(self.buf, p, 8)
        ohdr = _u(self.buf, p + 8, 8)
        cache_type = _u(self.buf, p + 16, 4)
        scratch = self.buf[p + 24:p + 40]
        return dict(name_off=name_off, ohdr=ohdr, cache=cache_type, scratch=scratch)

    # -- groups ---------------------------------------------------------
    def _heap_data(self, addr):
        b = self.buf
        if b[addr:addr + 4] != b"HEAP":
            raise ValueError("bad local heap")
        return _u(b, addr + 24, 8)  # address of data segment

    def _walk_btree(self, addr, out):
        b = self.buf
        if b[addr:addr + 4] != b"TREE":
            raise ValueError("bad B-tree node")
        ntype, level, nent = b[addr + 4], b[addr + 5], _u(b, addr + 6, 2)
        if ntype != 0:
            raise NotImplementedError("only group B-trees")
        p = addr + 8 + 2 * self.so  # skip siblings
        # keys and children alternate: key0 child0 key1 child1 ... keyN
        p += self.sl  # key 0
        for _ in range(nent):
            child = _u(b, p, 8)
            p += self.so + self.sl
            if level > 0:
                self._walk_btree(child, out)
            else:
                self._read_snod(child, out)

    def _read_snod(self, addr, out):
        b = self.buf
        if b[addr:addr + 4] != b"SNOD":
            raise ValueError("bad symbol node")
        n = _u(b, addr + 6, 2)
        p = addr + 8
        for _ in range(n):
            out.append(self._symtab_entry(p))
            p += 40

    def root_members(self) -> Dict[str, int]:
        e = self.root_entry
        if e["cache"] == 1:
            btree = _u(e["scratch"], 0, 8)
            heap = _u(e["scratch"], 8, 8)
        else:
            btree = heap = None
            for mtype, body in self._messages(e["ohdr"]):
                if mtype == 0x11:
                    btree, heap = _u(body, 0, 8), _u(body, 8, 8)
            if btree is None:
                raise NotImplementedError("root group without symbol table")
        data = self._heap_data(heap)
        ents = []
        self._walk_btree(btree, ents)
        names = {}
        for en in ents:
            s = data + en["name_off"]
            end = self.buf.index(b"\x00", s)
            names[self.buf[s:end].decode()] = en["ohdr"]
        return names

    # -- object headers -------------------------------------------------
    def _messages(self, addr):
        b = self.buf
        if b[addr] != 1:
            raise NotImplementedError("only v1 object headers")
        nmsg = _u(b, addr + 2, 2)
        size = _u(b, addr + 8, 4)
        blocks = [(addr + 16, size)]
        out = []
        while blocks and len(out) < nmsg:
            p, sz = blocks.pop(0)
            end = p + sz
            while p + 8 <= end and len(out) < nmsg:
                mtype = _u(b, p, 2)
                msz = _u(b, p + 2, 2)
                body = b[p + 8:p + 8 + msz]
                if mtype == 0x10:  # continuation
                    blocks.append((_u(body, 0, 8), _u(body, 8, 8)))
                out.append((mtype, body))
                p += 8 + msz
        return out

    def read_dataset(self, addr) -> np.ndarray:
        shape = dtype = None
        data = None
        for mtype, body in self._messages(addr):
            if mtype == 0x1:  # dataspace
                ver, rank, flags = body[0], body[1], body[2]
                p = 8 if ver == 1 else 4
                shape = tuple(_u(body, p + 8 * i, 8) for i in range(rank))
            elif mtype == 0x3:  # datatype
                cls = body[0] & 0x0F
                bits0 = body[1]
                size = _u(body, 4, 4)
                if bits0 & 1:
                    raise NotImplementedError("big-endian data")
                if cls == 1:
                    dtype = {8: "<f8", 4: "<f4"}[size]
                elif cls == 0:
                    signed = (bits0 >> 3) & 1
                    dtype = ("<i" if signed else "<u") + str(size)
                else:
                    raise NotImplementedError(f"datatype class {cls}")
            elif mtype == 0x8:  # layout
                ver = body[0]
                if ver != 3:
                    raise NotImplementedError("only layout v3")
                lclass = body[1]
                if lclass == 1:  # contiguous
                    a, n = _u(body, 2, 8), _u(body, 10, 8)
                    data = b"" if a == _UNDEF else self.buf[a:a + n]
                elif lclass == 0:  # compact
                    n = _u(body, 2, 2)
                    data = body[4:4 + n]
                else:
                    raise NotImplementedError("chunked layout")
        if shape is None or dtype is None or data is None:
            raise ValueError("incomplete dataset header")
        arr = np.frombuffer(data, dtype=dtype)
        n = int(np.prod(shape)) if shape else 1
        arr = arr[:n].reshape(shape) if shape else arr[:1].reshape(())
        return np.array(arr)


def load(path: str) -> Dict[str, np.ndarray]:
    """Return ``{name: ndarray}`` for every dataset in the root group of ``path``."""
    with open(path, "rb") as fh:
        f = _File(fh.read())
    return {k: f.read_dataset(a) for k, a in f.root_members().items()}
