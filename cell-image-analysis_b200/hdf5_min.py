"""Minimal pure-Python HDF5 *reader*, enough for Keras ``model.weights.h5``.

Why it exists: the reference loads ``best_autoencoder.keras`` / ``encoder.keras``
with Keras (improved_detection.py:28-29); a ``.keras`` file is a zip whose weights
member is HDF5, and this image has neither h5py nor libhdf5.  The product must read
those files unchanged, so it carries its own reader for the subset h5py emits:

* superblock v0/v1 (h5py default) and v2/v3 (``libver='latest'``);
* object headers v1 (with continuation blocks) and v2 (``OHDR``/``OCHK``);
* old-style groups: symbol-table message -> B-tree v1 (``TREE``) -> ``SNOD`` +
  local ``HEAP``;  new-style groups with *compact* link messages;
* datasets: dataspace v1/v2, fixed-point / IEEE-float datatypes (LE or BE), layout
  v1-v3 compact / contiguous, and chunked (B-tree v1) *without* filters.

Anything else (dense link storage, filters, v4 layouts, variable-length types)
raises ``H5FormatError`` -- the loader refuses rather than guesses.
"""
from __future__ import annotations

import struct

import numpy as np

SIG = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(ValueError):
    pass


class H5File:
    def __init__(self, data: bytes):
        self.b = memoryview(data)
        try:
            self._parse_superblock()
        except IndexError as e:
            raise H5FormatError("truncated HDF5 superblock") from e

    # ---- primitives -------------------------------------------------------
    def _u(self, off, n):
        return int.from_bytes(self.b[off:off + n], "little")

    def _parse_superblock(self):
        base = 0
        while True:
            if base + 8 > len(self.b):
                raise H5FormatError("HDF5 signature not found")
            if bytes(self.b[base:base + 8]) == SIG:
                break
            base = 512 if base == 0 else base * 2
        ver = self.b[base + 8]
        self.sb_version = ver
        if ver in (0, 1):
            self.O = self.b[base + 13]
            self.L = self.b[base + 14]
            p = base + 24 + (4 if ver == 1 else 0)
            self.base_addr = self._u(p, self.O)
            p += 4 * self.O                       # base, free-space, eof, driver
            # root symbol-table entry
            self.root_header = self._u(p + self.O, self.O)
        elif ver in (2, 3):
            self.O = self.b[base + 9]
            self.L = self.b[base + 10]
            p = base + 12
            self.base_addr = self._u(p, self.O)
            self.root_header = self._u(p + 3 * self.O, self.O)
        else:
            raise H5FormatError(f"unsupported superblock version {ver}")
        if self.O != 8 or self.L != 8:
            raise H5FormatError("only 8-byte offsets/lengths supported")
        self.base_addr += 0 if self.base_addr != UNDEF else 0

    # ---- object headers ---------------------------------------------------
    def _messages(self, addr):
        """Yield (type, flags, data_offset, size) for every message of the object."""
        b = self.b
        addr += self.base_addr
        out = []
        if bytes(b[addr:addr + 4]) == b"OHDR":
            flags = b[addr + 5]
            p = addr + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            szn = 1 << (flags & 3)
            chunk0 = self._u(p, szn)
            p += szn
            chunks = [(p, chunk0)]
            track_order = bool(flags & 0x04)
            visited = 0
            while chunks:
                s, n = chunks.pop(0)
                e = s + n
                while s + 4 <= e:
                    mtype = b[s]
                    msize = self._u(s + 1, 2)
                    mflags = b[s + 3]
                    s += 4 + (2 if track_order else 0)
                    if s + msize > e:
                        break
                    if mtype == 0x10:
                        co, cl = self._u(s, 8) + self.base_addr, self._u(s + 8, 8)
                        if bytes(b[co:co + 4]) != b"OCHK":
                            raise H5FormatError("bad OCHK")
                        visited += 1
                        if visited > 4096:
                            raise H5FormatError("object header continuation chain does not end")
                        chunks.append((co + 4, cl - 8))
                    elif mtype != 0:
                        out.append((mtype, mflags, s, msize))
                    s += msize
            return out
        if b[addr] != 1:
            raise H5FormatError(f"unsupported object header version {b[addr]} at {addr}")
        nmsg = self._u(addr + 2, 2)
        size0 = self._u(addr + 8, 4)
        chunks = [(addr + 16, size0)]
        seen = 0
        while chunks and seen < nmsg:
            s, n = chunks.pop(0)
            e = s + n
            while s + 8 <= e and seen < nmsg:
                mtype = self._u(s, 2)
                msize = self._u(s + 2, 2)
                mflags = b[s + 4]
                s += 8
                seen += 1
                if mtype == 0x10:
                    chunks.append((self._u(s, 8) + self.base_addr, self._u(s + 8, 8)))
                elif mtype != 0:
                    out.append((mtype, mflags, s, msize))
                s += msize
        return out

    # ---- groups -----------------------------------------------------------
    def _heap_name(self, heap_addr, off):
        h = heap_addr + self.base_addr
        if bytes(self.b[h:h + 4]) != b"HEAP":
            raise H5FormatError("bad local heap")
        seg = self._u(h + 8 + 2 * self.L, self.O) + self.base_addr
        s = seg + off
        e = s
        while self.b[e] != 0:
            e += 1
        return bytes(self.b[s:e]).decode("utf-8")

    def _btree_group(self, node, heap, out):
        p = node + self.base_addr
        sig = bytes(self.b[p:p + 4])
        if sig == b"SNOD":
            n = self._u(p + 6, 2)
            q = p + 8
            for _ in range(n):
                name = self._heap_name(heap, self._u(q, 8))
                out[name] = self._u(q + 8, 8)
                q += 40
            return
        if sig != b"TREE" or self.b[p + 4] != 0:
            raise H5FormatError("bad group B-tree node")
        used = self._u(p + 6, 2)
        q = p + 8 + 2 * self.O
        for _ in range(used):
            q += self.L                         # key
            self._btree_group(self._u(q, 8), heap, out)
            q += self.O
        return

    def links(self, header_addr):
        """name -> object header address for a group object (empty for a dataset)."""
        out = {}
        for mtype, _f, off, size in self._messages(header_addr):
            if mtype == 0x11:
                bt, hp = self._u(off, 8), self._u(off + 8, 8)
                self._btree_group(bt, hp, out)
            elif mtype == 0x06:
                lf = self.b[off + 1]
                p = off + 2
                ltype = 0
                if lf & 0x08:
                    ltype = self.b[p]
                    p += 1
                if lf & 0x04:
                    p += 8
                if lf & 0x10:
                    p += 1
                ln = 1 << (lf & 3)
                nlen = self._u(p, ln)
                p += ln
                name = bytes(self.b[p:p + nlen]).decode("utf-8")
                p += nlen
                if ltype == 0:
                    out[name] = self._u(p, 8)
            elif mtype == 0x02:
                # link info: dense storage if a fractal heap address is set
                lf = self.b[off + 1]
                p = off + 2 + (8 if lf & 1 else 0)
                if self._u(p, 8) != UNDEF:
                    raise H5FormatError("dense (fractal-heap) link storage not supported")
        return out

    def is_dataset(self, header_addr):
        return any(m[0] == 0x08 for m in self._messages(header_addr))

    # ---- datasets ---------------------------------------------------------
    def _dtype(self, off):
        cv = self.b[off]
        cls, _ver = cv & 0x0F, cv >> 4
        bits0 = self.b[off + 1]
        size = self._u(off + 4, 4)
        bo = ">" if bits0 & 1 else "<"
        if cls == 1 and size in (2, 4, 8):
            return np.dtype(f"{bo}f{size}")
        if cls == 0 and size in (1, 2, 4, 8):
            return np.dtype(f"{bo}{'i' if bits0 & 0x08 else 'u'}{size}")
        raise H5FormatError(f"unsupported datatype class {cls} size {size}")

    def _shape(self, off):
        ver = self.b[off]
        rank = self.b[off + 1]
        flags = self.b[off + 2]
        p = off + (8 if ver == 1 else 4)
        if ver == 2 and self.b[off + 3] == 2:
            return None                              # null dataspace
        return tuple(self._u(p + 8 * i, 8) for i in range(rank))

    def _read_chunked(self, btree, rank, chunk_dims, dtype, shape):
        arr = np.zeros(shape, dtype)

        def walk(node):
            p = node + self.base_addr
            if bytes(self.b[p:p + 4]) != b"TREE" or self.b[p + 4] != 1:
                raise H5FormatError("bad chunk B-tree node")
            level = self.b[p + 5]
            used = self._u(p + 6, 2)
            q = p + 8 + 2 * self.O
            ksz = 8 + 8 * (rank + 1)
            for _ in range(used):
                csize = self._u(q, 4)
                fmask = self._u(q + 4, 4)
                offs = [self._u(q + 8 + 8 * i, 8) for i in range(rank)]
                child = self._u(q + ksz, 8)
                q += ksz + self.O
                if level > 0:
                    walk(child)
                    continue
                if fmask != 0 and False:
                    pass
                c = child + self.base_addr
                n = int(np.prod(chunk_dims)) * dtype.itemsize
                if csize != n:
                    raise H5FormatError("filtered/compressed chunks not supported")
                blk = np.frombuffer(self.b[c:c + n], dtype).reshape(chunk_dims)
                sl = tuple(slice(o, min(o + d, s)) for o, d, s in zip(offs, chunk_dims, shape))
                arr[sl] = blk[tuple(slice(0, s.stop - s.start) for s in sl)]
        if btree != UNDEF:
            walk(btree)
        return arr

    def read(self, header_addr) -> np.ndarray:
        dtype = shape = None
        layout = None
        for mtype, _f, off, size in self._messages(header_addr):
            if mtype == 0x01:
                shape = self._shape(off)
            elif mtype == 0x03:
                dtype = self._dtype(off)
            elif mtype == 0x08:
                layout = off
            elif mtype == 0x0B:
                raise H5FormatError("filter pipeline (compression) not supported")
        if dtype is None or layout is None:
            raise H5FormatError("object is not a dataset")
        if shape is None:
            return np.zeros((0,), dtype)
        n = int(np.prod(shape)) if shape else 1
        nbytes = n * dtype.itemsize
        ver = self.b[layout]
        if ver == 3:
            cls = self.b[layout + 1]
            if cls == 0:
                sz = self._u(layout + 2, 2)
                raw = self.b[layout + 4: layout + 4 + sz]
            elif cls == 1:
                addr = self._u(layout + 2, 8)
                if addr == UNDEF:
                    return np.zeros(shape, dtype)
                a = addr + self.base_addr
                raw = self.b[a:a + nbytes]
            elif cls == 2:
                rank1 = self.b[layout + 2]
                bt = self._u(layout + 3, 8)
                dims = [self._u(layout + 11 + 4 * i, 4) for i in range(rank1)]
                return self._read_chunked(bt, rank1 - 1, tuple(dims[:-1]), dtype, shape)
            else:
                raise H5FormatError(f"layout class {cls}")
        elif ver in (1, 2):
            rank = self.b[layout + 1]
            cls = self.b[layout + 2]
            p = layout + 8
            addr = None
            if cls != 0:
                addr = self._u(p, 8)
                p += 8
            dims = [self._u(p + 4 * i, 4) for i in range(rank)]
            p += 4 * rank
            if cls == 0:
                sz = self._u(p, 4)
                raw = self.b[p + 4:p + 4 + sz]
            elif cls == 1:
                a = addr + self.base_addr
                raw = self.b[a:a + nbytes]
            else:
                return self._read_chunked(addr, rank - 1, tuple(dims[:-1]), dtype, shape)
        else:
            raise H5FormatError(f"unsupported data layout version {ver}")
        if len(raw) < nbytes:
            raise H5FormatError("dataset extends past end of file")
        return np.frombuffer(bytes(raw[:nbytes]), dtype).reshape(shape).copy()

    # ---- convenience ------------------------------------------------------
    def datasets(self) -> dict:
        """Flat dict ``"a/b/c" -> ndarray`` of every dataset in the file."""
        out, seen = {}, set()

        def rec(addr, path):
            if addr in seen:
                return
            seen.add(addr)
            if self.is_dataset(addr):
                out[path] = self.read(addr)
                return
            for name, child in self.links(addr).items():
                rec(child, f"{path}/{name}" if path else name)
        try:
            rec(self.root_header, "")
        except H5FormatError:
            raise
        except (IndexError, ValueError, TypeError, KeyError, OverflowError, RecursionError, MemoryError,
                UnicodeDecodeError, struct.error) as e:
            # a corrupted address / size / name: callers (artifacts.py) report H5FormatError with the file name
            raise H5FormatError(f"malformed HDF5 structure: {type(e).__name__}: {e}") from e
        return out
