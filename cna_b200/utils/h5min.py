"""Minimal pure-Python HDF5 reader for old-style ``.h5ad`` files.

Only what the bundled demo fixture of the reference needs (``demo/data.h5ad``: superblock v0,
v1 object headers, symbol-table groups, contiguous or unfiltered-chunked numeric datasets); see
SURVEY.md section 8(c).  ``h5py``/``anndata`` are not available in this image, and the hot path
touches only ``obs`` columns and the ``connectivities`` CSR triplet, so this is enough to make the
demo configuration self-contained.

It is *not* a general HDF5 implementation: filters (gzip ...), v2 B-trees, fractal heaps,
variable-length strings and compound types raise ``NotImplementedError``.
"""
import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"


class H5File:
    def __init__(self, path):
        # memory-mapped: a row-block read of a multi-GB file touches only the pages it needs
        import mmap
        with open(path, "rb") as fh:
            self.buf = mmap.mmap(fh.fileno(), 0, access=mmap.ACCESS_READ)
        if self.buf[:8] != _SIG:
            raise ValueError("not an HDF5 file")
        ver = self.buf[8]
        if ver != 0:
            raise NotImplementedError(f"superblock version {ver}")
        if self.buf[13] != 8 or self.buf[14] != 8:
            raise NotImplementedError("only 8-byte offsets/lengths")
        # root group symbol-table entry sits after the 24 fixed bytes + 4 addresses
        self.root_header = struct.unpack_from("<Q", self.buf, 56 + 8)[0]

    # ---- object headers -------------------------------------------------------------------
    def _messages(self, addr):
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", self.buf, addr)
        if ver != 1:
            raise NotImplementedError(f"object header version {ver}")
        blocks = [(addr + 16, hsize)]
        out = []
        while blocks and len(out) < nmsg:
            pos, size = blocks.pop(0)
            end = pos + size
            while pos + 8 <= end and len(out) < nmsg:
                mtype, msize, _flags = struct.unpack_from("<HHB", self.buf, pos)
                body = pos + 8
                if mtype == 0x10:  # continuation block
                    off, ln = struct.unpack_from("<QQ", self.buf, body)
                    blocks.append((off, ln))
                out.append((mtype, body, msize))
                pos = body + msize
        return out

    # ---- groups ---------------------------------------------------------------------------
    def _heap_name(self, heap_addr, off):
        if self.buf[heap_addr:heap_addr + 4] != b"HEAP":
            raise ValueError("bad local heap")
        data_addr = struct.unpack_from("<Q", self.buf, heap_addr + 24)[0]
        start = data_addr + off
        stop = self.buf.find(b"\x00", start)
        return self.buf[start:stop].decode()

    def _group_walk(self, tree, heap, out):
        if self.buf[tree:tree + 4] != b"TREE":
            raise ValueError("bad group b-tree node")
        ntype, level, used = struct.unpack_from("<BBH", self.buf, tree + 4)
        if ntype != 0:
            raise ValueError("expected group node")
        pos = tree + 24
        for i in range(used):
            child = struct.unpack_from("<Q", self.buf, pos + 8 + 16 * i)[0]
            if level > 0:
                self._group_walk(child, heap, out)
                continue
            if self.buf[child:child + 4] != b"SNOD":
                raise ValueError("bad symbol node")
            nsym = struct.unpack_from("<H", self.buf, child + 6)[0]
            for s in range(nsym):
                name_off, hdr = struct.unpack_from("<QQ", self.buf, child + 8 + 40 * s)
                out[self._heap_name(heap, name_off)] = hdr

    def members(self, addr=None):
        """name -> object-header address for the group at ``addr`` (root if None)."""
        addr = self.root_header if addr is None else addr
        for mtype, body, _ in self._messages(addr):
            if mtype == 0x11:
                tree, heap = struct.unpack_from("<QQ", self.buf, body)
                out = {}
                self._group_walk(tree, heap, out)
                return out
        return None  # not a group

    def resolve(self, path):
        addr = self.root_header
        for part in [p for p in path.split("/") if p]:
            mem = self.members(addr)
            if mem is None or part not in mem:
                raise KeyError(path)
            addr = mem[part]
        return addr

    def tree(self, addr=None, prefix=""):
        """Recursively list datasets: path -> header address."""
        out = {}
        mem = self.members(addr)
        for name, hdr in (mem or {}).items():
            if self.members(hdr) is not None:
                out.update(self.tree(hdr, prefix + name + "/"))
            else:
                out[prefix + name] = hdr
        return out

    # ---- datasets -------------------------------------------------------------------------
    @staticmethod
    def _dtype(buf, body):
        cv, b0, _b1, _b2, size = struct.unpack_from("<BBBBI", buf, body)
        cls = cv & 0x0F
        order = ">" if (b0 & 1) else "<"
        if cls == 0:  # fixed point
            signed = bool(b0 & 0x08)
            return np.dtype(f"{order}{'i' if signed else 'u'}{size}")
        if cls == 1:  # float
            return np.dtype(f"{order}f{size}")
        if cls == 3:  # fixed-length string
            return np.dtype(f"S{size}")
        raise NotImplementedError(f"datatype class {cls}")

    def _chunks(self, node, rank, out):
        if self.buf[node:node + 4] != b"TREE":
            raise ValueError("bad chunk b-tree node")
        ntype, level, used = struct.unpack_from("<BBH", self.buf, node + 4)
        if ntype != 1:
            raise ValueError("expected chunk node")
        keysz = 8 + 8 * (rank + 1)
        pos = node + 24
        for i in range(used):
            kpos = pos + i * (keysz + 8)
            nbytes, fmask = struct.unpack_from("<II", self.buf, kpos)
            offs = struct.unpack_from(f"<{rank + 1}Q", self.buf, kpos + 8)
            child = struct.unpack_from("<Q", self.buf, kpos + keysz)[0]
            if level > 0:
                self._chunks(child, rank, out)
            else:
                if fmask != 0:
                    raise NotImplementedError("filtered chunks")
                out.append((offs[:rank], child, nbytes))

    def read(self, path):
        hdr = self.resolve(path) if isinstance(path, str) else path
        shape = dtype = layout = None
        for mtype, body, _ in self._messages(hdr):
            if mtype == 0x01:
                ver, rank, flags = struct.unpack_from("<BBB", self.buf, body)
                off = body + (8 if ver == 1 else 4)
                shape = struct.unpack_from(f"<{rank}Q", self.buf, off) if rank else ()
            elif mtype == 0x03:
                dtype = self._dtype(self.buf, body)
            elif mtype == 0x08:
                layout = body
            elif mtype == 0x0B:
                raise NotImplementedError("filter pipeline")
        if shape is None or dtype is None or layout is None:
            raise ValueError("not a dataset")
        ver, cls = struct.unpack_from("<BB", self.buf, layout)
        if ver != 3:
            raise NotImplementedError(f"layout version {ver}")
        count = int(np.prod(shape)) if shape else 1
        if cls == 1:  # contiguous
            addr, _size = struct.unpack_from("<QQ", self.buf, layout + 2)
            arr = np.frombuffer(self.buf, dtype=dtype, count=count, offset=addr)
            return arr.reshape(shape).copy()
        if cls == 0:  # compact
            size = struct.unpack_from("<H", self.buf, layout + 2)[0]
            arr = np.frombuffer(self.buf, dtype=dtype, count=count, offset=layout + 4)
            return arr.reshape(shape).copy()
        if cls == 2:  # chunked
            rank = self.buf[layout + 2] - 1
            btree = struct.unpack_from("<Q", self.buf, layout + 3)[0]
            cdims = struct.unpack_from(f"<{rank}I", self.buf, layout + 11)
            out = np.zeros(shape, dtype=dtype)
            chunks = []
            self._chunks(btree, rank, chunks)
            for offs, addr, nbytes in chunks:
                block = np.frombuffer(self.buf, dtype=dtype, count=int(np.prod(cdims)), offset=addr)
                block = block.reshape(cdims)
                sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, shape))
                out[sel] = block[tuple(slice(0, s.stop - s.start) for s in sel)]
            return out
        raise NotImplementedError(f"layout class {cls}")

    def dataset_info(self, path):
        """(shape, dtype, file offset or None) of a dataset; the offset is set for contiguous layouts."""
        hdr = self.resolve(path) if isinstance(path, str) else path
        shape = dtype = offset = None
        for mtype, body, _ in self._messages(hdr):
            if mtype == 0x01:
                ver, rank, flags = struct.unpack_from("<BBB", self.buf, body)
                off = body + (8 if ver == 1 else 4)
                shape = struct.unpack_from(f"<{rank}Q", self.buf, off) if rank else ()
            elif mtype == 0x03:
                dtype = self._dtype(self.buf, body)
            elif mtype == 0x08:
                ver, cls = struct.unpack_from("<BB", self.buf, body)
                if ver == 3 and cls == 1:
                    offset = struct.unpack_from("<Q", self.buf, body + 2)[0]
            elif mtype == 0x0B:
                raise NotImplementedError("filter pipeline (compressed dataset)")
        if shape is None or dtype is None:
            raise ValueError("not a dataset")
        return shape, dtype, offset

    def read_slice(self, path, start, stop, out=None):
        """Elements [start, stop) of a 1-D dataset, straight out of the mapped file for contiguous layouts
        (copied into ``out`` when given: e.g. a page-locked buffer)."""
        shape, dtype, offset = self.dataset_info(path)
        if len(shape) != 1:
            raise ValueError("read_slice: 1-D datasets only")
        if offset is None:  # chunked / compact: whole read, then slice
            arr = self.read(path)[start:stop]
        else:
            arr = np.frombuffer(self.buf, dtype=dtype, count=stop - start, offset=offset + start * dtype.itemsize)
        if out is None:
            return np.array(arr)
        out[...] = arr
        return out
