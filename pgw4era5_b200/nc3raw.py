"""
Raw access to NetCDF-3 files (classic, 64-bit offset and CDF-5) for the file pipeline of the
step_03 drop-in: the header is parsed here (it is a few hundred bytes), so that the big fields can
be moved between the file and pinned host memory with ``readinto`` / ``pwrite`` -- no decoding, no
intermediate numpy copies.  NetCDF-3 stores big-endian numbers; the byte order of the float32
fields is swapped on the GPU (``pgw_byteswap32``), next to the H2D / D2H copies.

Only what the pipeline needs: locate a variable (offset, shape, type) and read or overwrite the
values of one record.  Everything else (coordinates, attributes, the engine's level tables) goes
through ``ncio``.  File format: NetCDF Classic Format Specification (Unidata), all integers
big-endian:
    header  = magic numrecs dim_list gatt_list var_list
    var     = name nelems [dimid ...] vatt_list nc_type vsize begin
"""
import os
import struct

import numpy as np

NC_DIMENSION, NC_VARIABLE, NC_ATTRIBUTE = 0x0A, 0x0B, 0x0C
_TYPES = {1: ("i1", 1), 2: ("S1", 1), 3: (">i2", 2), 4: (">i4", 4), 5: (">f4", 4), 6: (">f8", 8),
          7: ("u1", 1), 8: (">u2", 2), 9: (">u4", 4), 10: (">i8", 8), 11: (">u8", 8)}
NC_FLOAT = 5


class NotNetCDF3(ValueError):
    pass


class _Reader:
    def __init__(self, buf, version):
        self.b, self.p, self.v = buf, 0, version

    def u32(self):
        (x,) = struct.unpack_from(">I", self.b, self.p)
        self.p += 4
        return x

    def u64(self):
        (x,) = struct.unpack_from(">Q", self.b, self.p)
        self.p += 8
        return x

    def size(self):                      # NON_NEG: 64 bit in CDF-5
        return self.u64() if self.v == 5 else self.u32()

    def name(self):
        n = self.size()
        s = bytes(self.b[self.p:self.p + n]).decode("utf-8")
        self.p += (n + 3) & ~3
        return s

    def skip_values(self, nc_type, nelems):
        self.p += ((_TYPES[nc_type][1] * nelems) + 3) & ~3

    def att_list(self):
        tag = self.u32()
        n = self.size()
        if tag == 0:
            return
        if tag != NC_ATTRIBUTE:
            raise NotNetCDF3("attribute list expected")
        for _ in range(n):
            self.name()
            t = self.u32()
            self.skip_values(t, self.size())


class RawVar:
    __slots__ = ("name", "dims", "shape", "nc_type", "dtype", "vsize", "begin", "is_record")

    def __init__(self, name, dims, shape, nc_type, vsize, begin, is_record):
        self.name, self.dims, self.shape, self.nc_type = name, dims, shape, nc_type
        self.dtype, self.vsize, self.begin, self.is_record = np.dtype(_TYPES[nc_type][0]), vsize, begin, is_record

    @property
    def record_shape(self):
        """Shape of the values of one record (record variables) or of the whole variable."""
        return self.shape[1:] if self.is_record else self.shape

    @property
    def nbytes(self):
        return int(np.prod(self.record_shape, dtype=np.int64)) * self.dtype.itemsize


class RawNC3:
    """Header of a NetCDF-3 file: ``dims`` (name -> length, record dimension -> numrecs), ``vars``."""

    def __init__(self, path, header_bytes=1 << 20):
        self.path = path
        with open(path, "rb") as f:
            head = f.read(header_bytes)
            if len(head) < 8 or head[:3] != b"CDF" or head[3] not in (1, 2, 5):
                raise NotNetCDF3("%s is not a NetCDF-3 file" % path)
            while True:
                try:
                    self._parse(head)
                    break
                except (struct.error, IndexError):
                    if len(head) < header_bytes:
                        raise NotNetCDF3("truncated header in %s" % path)
                    header_bytes *= 8
                    f.seek(0)
                    head = f.read(header_bytes)

    def _parse(self, head):
        version = head[3]
        r = _Reader(memoryview(head), version)
        r.p = 4
        numrecs = r.size()
        streaming = numrecs == (0xFFFFFFFFFFFFFFFF if version == 5 else 0xFFFFFFFF)
        tag, n = r.u32(), r.size()
        dims = []
        if tag == NC_DIMENSION:
            for _ in range(n):
                name = r.name()
                dims.append((name, r.size()))
        elif tag != 0:
            raise NotNetCDF3("dimension list expected")
        r.att_list()
        tag, n = r.u32(), r.size()
        self.vars, rec_vars = {}, []
        if tag == NC_VARIABLE:
            for _ in range(n):
                name = r.name()
                nd = r.size()
                dimids = [r.size() for _ in range(nd)]
                r.att_list()
                nc_type = r.u32()
                vsize = r.size()
                begin = r.u64() if version in (2, 5) else r.u32()
                is_rec = nd > 0 and dims[dimids[0]][1] == 0
                shape = tuple(dims[d][1] for d in dimids)
                v = RawVar(name, tuple(dims[d][0] for d in dimids), shape, nc_type, vsize, begin, is_rec)
                self.vars[name] = v
                if is_rec:
                    rec_vars.append(v)
        elif tag != 0:
            raise NotNetCDF3("variable list expected")
        self.version, self.header_size = version, r.p
        # record size: sum of the (padded) vsizes; a single record variable is not padded
        if len(rec_vars) == 1:
            self.recsize = int(np.prod(rec_vars[0].shape[1:], dtype=np.int64)) * rec_vars[0].dtype.itemsize
        else:
            self.recsize = sum(v.vsize for v in rec_vars)
        if streaming:
            size = os.path.getsize(self.path)
            first = min((v.begin for v in rec_vars), default=size)
            numrecs = (size - first) // self.recsize if rec_vars and self.recsize else 0
        self.numrecs = int(numrecs)
        self.dims = {name: (self.numrecs if length == 0 else length) for name, length in dims}
        for v in rec_vars:
            v.shape = (self.numrecs,) + v.shape[1:]

    def offset(self, name, record=0):
        v = self.vars[name]
        if v.is_record:
            if not 0 <= record < self.numrecs:
                raise IndexError("record %d of %d" % (record, self.numrecs))
            return v.begin + record * self.recsize
        return v.begin

    def read_into(self, fobj, name, buf, record=0):
        """Fill ``buf`` (writable buffer of exactly the variable's record size) with the raw
        big-endian bytes of ``name``; ``fobj`` is the file opened 'rb' (unbuffered is best)."""
        v = self.vars[name]
        mv = memoryview(buf).cast("B")
        if mv.nbytes != v.nbytes:
            raise ValueError("%s: buffer has %d bytes, variable %d" % (name, mv.nbytes, v.nbytes))
        off, done = self.offset(name, record), 0
        while done < mv.nbytes:                      # os.preadv may return short counts on large reads
            n = os.preadv(fobj.fileno(), [mv[done:]], off + done)
            if n <= 0:
                raise IOError("unexpected end of %s while reading %s" % (self.path, name))
            done += n

    def write_from(self, fd, name, buf, record=0):
        """Overwrite the values of ``name`` in the file open as descriptor ``fd`` with raw big-endian bytes."""
        v = self.vars[name]
        mv = memoryview(buf).cast("B")
        if mv.nbytes != v.nbytes:
            raise ValueError("%s: buffer has %d bytes, variable %d" % (name, mv.nbytes, v.nbytes))
        off, done = self.offset(name, record), 0
        while done < mv.nbytes:
            done += os.pwrite(fd, mv[done:], off + done)
