"""Minimal HDF5 reader / writer for Keras weight files, pure Python + numpy (h5py is not in the image).

The reference stores and loads generator weights as Keras HDF5 (``Generator.model().save("..._gen_N.h5")``,
simple_sr/models/sr_model.py:233-244; ``tf.keras.models.load_model(path)``, utils/models/model_builder.py:17-19) and reads
VGG19 weights from a Keras ``.h5`` (``original_vgg.load_weights(custom_weights_path)``, model_builder.py:222,249;
loss_functions/vgg_loss.py:95).  h5py writes those files in the library's "earliest" format; that subset is what this
module understands:

* superblock versions 0 / 1 (and 2 / 3 for reading), an optional user block (the signature is searched at 0, 512, ...),
  addresses relative to the superblock's base address;
* version-1 object headers with continuation blocks (version-2 ``OHDR`` headers are read too);
* old-style groups: symbol-table message -> v1 B-tree (``TREE``) -> symbol nodes (``SNOD``) -> names in a local heap
  (``HEAP``); compact new-style groups (link messages) for reading;
* datasets: compact, contiguous and chunked layout (v1 chunk B-tree; deflate and shuffle filters), fixed-point, IEEE
  float and fixed-length string element types, little or big endian;
* attributes (message versions 1-3) of those types plus variable-length strings (global heap, ``GCOL``), which is how
  h5py stores ``model_config``.

The reader is pinned to a file written by the HDF5 library itself (tests/test_h5lite.py reads a MATLAB v7.3 file that
ships with scipy's test data); the writer emits the same subset (superblock 0, v1 headers, symbol-table groups with a
proper B-tree, contiguous datasets, compact attributes) and is checked by reading its files back and by comparing the
structures it emits for the same content - datatype, dataspace and attribute messages, local heap, B-tree node, symbol
node - byte for byte with the ones the library wrote into that file.

HDF5 File Format Specification version 2.0 is the source for every structure below; section numbers in the comments
refer to it.
"""
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


def _pad8(n):
    return (n + 7) & ~7


# ------------------------------------------------------------------------------------------------------------- reader

class _Datatype:
    """A decoded datatype message (IV.A.2.d): numpy dtype for fixed-size classes, ``vlen_str`` for variable strings."""

    def __init__(self, buf, off):
        cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", buf, off)
        self.cls, self.version, self.size = cv & 15, cv >> 4, size
        self.vlen_str = False
        self.dtype = None
        self.consumed = 8
        if self.cls == 0:                                   # fixed point: bit 0 byte order, bit 3 signed
            self.dtype = np.dtype((">" if b0 & 1 else "<") + ("i" if b0 & 8 else "u") + str(size))
            self.consumed += 4
        elif self.cls == 1:                                 # floating point (IEEE layouts only)
            if size not in (2, 4, 8):
                raise H5Error(f"float of {size} bytes")
            self.dtype = np.dtype((">" if b0 & 1 else "<") + "f" + str(size))
            self.consumed += 12
        elif self.cls == 3:                                 # fixed-length string, padding type in bits 0-3
            self.dtype = np.dtype(f"S{size}")
            self.str_pad = b0 & 15
        elif self.cls == 9:                                 # variable length: bits 0-3 type (1 = string)
            base = _Datatype(buf, off + 8)
            self.consumed += base.consumed
            if (b0 & 15) != 1:
                raise H5Error("variable-length sequences are not supported (strings only)")
            self.vlen_str = True
        elif self.cls == 7:                                 # object reference: kept as raw addresses
            self.dtype = np.dtype(f"<u{size}")
        else:
            raise H5Error(f"datatype class {self.cls} is not supported")


def _dataspace(buf, off, L):
    """Dataspace message (IV.A.2.b) -> shape tuple (() for scalar, None for the null dataspace)."""
    version, rank, flags = struct.unpack_from("<BBB", buf, off)
    if version == 1:
        p = off + 8
    elif version == 2:
        if buf[off + 3] == 2:
            return None
        p = off + 4
    else:
        raise H5Error(f"dataspace version {version}")
    fmt = "<" + ("Q" if L == 8 else "I") * rank
    return tuple(struct.unpack_from(fmt, buf, p))


class _Object:
    """One object header, decoded into the messages this module uses."""

    def __init__(self, f, addr):
        self.f, self.addr = f, addr
        self.attr_raw = []          # (name, datatype, shape, bytes)
        self.shape = self.dt = self.layout = None
        self.filters = []
        self.symtab = None          # (btree address, local heap address)
        self.links = {}             # new-style hard links: name -> object header address
        self.dense_links = False
        self._parse()

    # -- header walking (IV.A.1.a / IV.A.1.b)
    def _parse(self):
        f = self.f
        buf = f.buf
        a = f.base + self.addr
        if buf[a:a + 4] == b"OHDR":
            self._parse_v2(a)
            return
        version, _, nmsg, _refs, hsize = struct.unpack_from("<BBHII", buf, a)
        if version != 1:
            raise H5Error(f"object header version {version} at {self.addr}")
        blocks = [(a + 16, hsize)]
        seen = 0
        while blocks and seen < nmsg:
            p, n = blocks.pop(0)
            end = p + n
            while p + 8 <= end and seen < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", buf, p)
                body = p + 8
                self._message(mtype, body, msize, mflags, blocks)
                p = body + msize
                seen += 1

    def _parse_v2(self, a):
        f = self.f
        buf = f.buf
        flags = buf[a + 5]
        p = a + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        nsz = 1 << (flags & 3)
        chunk0 = int.from_bytes(buf[p:p + nsz], "little")
        p += nsz
        track = 2 if flags & 4 else 0
        blocks = [(p, chunk0)]
        while blocks:
            p, n = blocks.pop(0)
            end = p + n
            while p + 4 + track <= end:
                mtype = buf[p]
                msize, mflags = struct.unpack_from("<HB", buf, p + 1)
                body = p + 4 + track
                if body + msize > end:
                    break
                cont = []
                self._message(mtype, body, msize, mflags, cont)
                for cp, cn in cont:                        # continuation chunks: "OCHK" + messages + checksum
                    blocks.append((cp + 4, cn - 8))
                p = body + msize

    def _message(self, mtype, body, msize, mflags, blocks):
        f = self.f
        buf = f.buf
        if mflags & 2 and mtype in (1, 3):
            raise H5Error("shared header messages are not supported")
        if mtype == 0x01:
            self.shape = _dataspace(buf, body, f.L)
        elif mtype == 0x03:
            self.dt = _Datatype(buf, body)
        elif mtype == 0x08:
            self.layout = (body, msize)
        elif mtype == 0x0B:
            self._filters(body)
        elif mtype == 0x0C:
            self._attribute(body, msize)
        elif mtype == 0x10:
            off, ln = f.read_offset(body), f.read_length(body + f.O)
            blocks.append((f.base + off, ln))
        elif mtype == 0x11:
            self.symtab = (f.read_offset(body), f.read_offset(body + f.O))
        elif mtype == 0x06:
            self._link(body)
        elif mtype == 0x02:                                 # link info: a fractal-heap address means dense storage
            flags = buf[body + 1]
            p = body + 2 + (8 if flags & 1 else 0)
            if f.read_offset(p) != f.undef:
                self.dense_links = True

    def _filters(self, body):
        buf = self.f.buf
        version, n = buf[body], buf[body + 1]
        p = body + (8 if version == 1 else 2)
        for _ in range(n):
            fid = struct.unpack_from("<H", buf, p)[0]
            p += 2
            nlen = 0
            if version == 1 or fid >= 256:
                nlen = struct.unpack_from("<H", buf, p)[0]
                p += 2
            _flags, ncd = struct.unpack_from("<HH", buf, p)
            p += 4
            p += _pad8(nlen) if version == 1 else nlen
            cd = struct.unpack_from("<" + "I" * ncd, buf, p)
            p += 4 * ncd
            if version == 1 and ncd & 1:
                p += 4
            self.filters.append((fid, cd))

    def _attribute(self, body, msize):
        buf = self.f.buf
        version = buf[body]
        nsz, tsz, ssz = struct.unpack_from("<HHH", buf, body + 2)
        p = body + 8
        if version == 3:
            p += 1
        elif version not in (1, 2):
            raise H5Error(f"attribute message version {version}")
        pad = _pad8 if version == 1 else (lambda n: n)
        name = bytes(buf[p:p + nsz]).split(b"\0")[0].decode("utf-8")
        p += pad(nsz)
        if version != 1 and buf[body + 1] & 3:
            raise H5Error("shared attribute datatype / dataspace")
        dt = _Datatype(buf, p)
        p += pad(tsz)
        shape = _dataspace(buf, p, self.f.L)
        p += pad(ssz)
        self.attr_raw.append((name, dt, shape, p, body + msize))

    def _link(self, body):
        f = self.f
        buf = f.buf
        flags = buf[body + 1]
        p = body + 2
        ltype = 0
        if flags & 8:
            ltype = buf[p]
            p += 1
        if flags & 4:
            p += 8
        if flags & 16:
            p += 1
        nsz = 1 << (flags & 3)
        nlen = int.from_bytes(buf[p:p + nsz], "little")
        p += nsz
        name = bytes(buf[p:p + nlen]).decode("utf-8")
        p += nlen
        if ltype == 0:
            self.links[name] = f.read_offset(p)


class _Node:
    """Common base of File / Group / Dataset: attributes."""

    def __init__(self, f, obj, name):
        self._f, self._obj, self.name = f, obj, name
        self._attrs = None

    @property
    def attrs(self):
        if self._attrs is None:
            self._attrs = {}
            for name, dt, shape, p, end in self._obj.attr_raw:
                self._attrs[name] = self._f.decode(dt, shape, p, end)
        return self._attrs


class Dataset(_Node):
    @property
    def shape(self):
        return self._obj.shape

    @property
    def dtype(self):
        return self._obj.dt.dtype

    def __getitem__(self, key):
        arr = self.read()
        return arr if key is Ellipsis or key == () else arr[key]

    def read(self):
        f, o = self._f, self._obj
        buf = f.buf
        if o.dt is None or o.layout is None:
            raise H5Error(f"{self.name}: not a dataset")
        if o.dt.vlen_str:
            raise H5Error(f"{self.name}: variable-length datasets are not supported")
        shape = o.shape if o.shape is not None else (0,)
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        nbytes = count * o.dt.size
        body, _ = o.layout
        version = buf[body]
        if version == 3:
            cls = buf[body + 1]
            if cls == 0:
                size = struct.unpack_from("<H", buf, body + 2)[0]
                raw = bytes(buf[body + 4:body + 4 + size])
            elif cls == 1:
                addr = f.read_offset(body + 2)
                raw = b"\0" * nbytes if addr == f.undef else bytes(buf[f.base + addr:f.base + addr + nbytes])
            elif cls == 2:
                ndim = buf[body + 2]
                bt = f.read_offset(body + 3)
                cdims = struct.unpack_from("<" + "I" * ndim, buf, body + 3 + f.O)
                return self._chunked(bt, cdims[:-1], shape)
            else:
                raise H5Error(f"layout class {cls}")
        elif version in (1, 2):
            ndim, cls = buf[body + 1], buf[body + 2]
            p = body + 8
            addr = None
            if cls != 0:
                addr = f.read_offset(p)
                p += f.O
            dims = struct.unpack_from("<" + "I" * ndim, buf, p)
            p += 4 * ndim
            if cls == 0:
                size = struct.unpack_from("<I", buf, p)[0]
                raw = bytes(buf[p + 4:p + 4 + size])
            elif cls == 1:
                raw = b"\0" * nbytes if addr == f.undef else bytes(buf[f.base + addr:f.base + addr + nbytes])
            else:
                return self._chunked(addr, dims[:-1] if version == 1 or len(dims) > len(shape) else dims, shape)
        else:
            raise H5Error(f"layout message version {version} (written by a newer library format)")
        if len(raw) < nbytes:
            raise H5Error(f"{self.name}: {len(raw)} bytes stored, {nbytes} expected")
        return np.frombuffer(raw[:nbytes], dtype=o.dt.dtype).reshape(shape).copy()

    def _chunked(self, btree, cdims, shape):
        f, o = self._f, self._obj
        out = np.zeros(shape, dtype=o.dt.dtype)
        if btree == f.undef:
            return out
        rank = len(shape)
        csize = int(np.prod(cdims)) * o.dt.size
        for nbytes, mask, offs, addr in f.chunk_leaves(btree, rank):
            raw = bytes(f.buf[f.base + addr:f.base + addr + nbytes])
            for i, (fid, cd) in reversed(list(enumerate(o.filters))):
                if mask & (1 << i):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    n = cd[0] if cd else o.dt.size
                    raw = np.frombuffer(raw, np.uint8).reshape(n, -1).T.tobytes()
                elif fid == 3:
                    raw = raw[:-4]                          # fletcher32 checksum trailer
                else:
                    raise H5Error(f"filter {fid} is not supported")
            chunk = np.frombuffer(raw[:csize], dtype=o.dt.dtype).reshape(cdims)
            sl = tuple(slice(offs[d], min(offs[d] + cdims[d], shape[d])) for d in range(rank))
            out[sl] = chunk[tuple(slice(0, s.stop - s.start) for s in sl)]
        return out


class Group(_Node):
    def __init__(self, f, obj, name):
        super().__init__(f, obj, name)
        self._members = None

    def _load(self):
        if self._members is None:
            o = self._obj
            if o.symtab is not None:
                self._members = dict(self._f.symbols(*o.symtab))
            elif o.dense_links:
                raise H5Error(f"{self.name}: dense (fractal-heap) group storage is not supported; "
                              "re-save the file with libver='earliest' (the h5py / Keras default)")
            else:
                self._members = dict(o.links)
        return self._members

    def keys(self):
        return list(self._load().keys())

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self._load())

    def items(self):
        return [(k, self[k]) for k in self.keys()]

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group):
                raise KeyError(path)
            members = node._load()
            if part not in members:
                raise KeyError(f"{path!r}: no member {part!r} in {node.name!r}")
            child = self._f.object(members[part])
            cname = node.name.rstrip("/") + "/" + part
            node = Dataset(self._f, child, cname) if child.dt is not None and child.layout is not None \
                else Group(self._f, child, cname)
        return node

    def visit_datasets(self, prefix=""):
        """[(path relative to this group, Dataset)] in name order, depth first."""
        out = []
        for k in sorted(self.keys()):
            n = self[k]
            if isinstance(n, Group):
                out.extend(n.visit_datasets(prefix + k + "/"))
            else:
                out.append((prefix + k, n))
        return out


class File(Group):
    """Read-only HDF5 file: ``File(path)["group/dataset"][...]``, ``.attrs``, ``.keys()`` like h5py."""

    def __init__(self, path):
        with open(path, "rb") as fh:
            self.buf = memoryview(fh.read())
        self.path = path
        pos = 0
        while True:                                         # II.A: the signature sits at 0, 512, 1024, ...
            if pos + 8 > len(self.buf):
                raise H5Error(f"{path}: not an HDF5 file (no signature)")
            if self.buf[pos:pos + 8] == SIGNATURE:
                break
            pos = 512 if pos == 0 else pos * 2
        buf = self.buf
        version = buf[pos + 8]
        if version in (0, 1):
            self.O, self.L = buf[pos + 13], buf[pos + 14]
            self.leaf_k, self.internal_k = struct.unpack_from("<HH", buf, pos + 16)
            p = pos + 24 + (4 if version == 1 else 0)
            self._set_sizes()
            self.base = self.read_offset_abs(p)
            # addresses are relative to the base address; a user block moves the superblock AND the base (MATLAB)
            root = p + 4 * self.O                           # root group symbol table entry (III.C)
            root_addr = self.read_offset_abs(root + self.O)
        elif version in (2, 3):
            self.O, self.L = buf[pos + 9], buf[pos + 10]
            self._set_sizes()
            self.base = self.read_offset_abs(pos + 12)
            root_addr = self.read_offset_abs(pos + 12 + 3 * self.O)
        else:
            raise H5Error(f"superblock version {version}")
        self._objects = {}
        self._f = self
        super().__init__(self, self.object(root_addr), "/")

    def _set_sizes(self):
        if self.O not in (4, 8) or self.L not in (4, 8):
            raise H5Error(f"offset / length sizes {self.O} / {self.L}")
        self.undef = (1 << (8 * self.O)) - 1

    def read_offset_abs(self, p):
        return int.from_bytes(self.buf[p:p + self.O], "little")

    read_offset = read_offset_abs

    def read_length(self, p):
        return int.from_bytes(self.buf[p:p + self.L], "little")

    def object(self, addr):
        if addr not in self._objects:
            self._objects[addr] = _Object(self, addr)
        return self._objects[addr]

    def close(self):
        self.buf = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- old-style groups (III.A.1, III.B, III.D)
    def _heap_data(self, heap_addr):
        a = self.base + heap_addr
        if self.buf[a:a + 4] != b"HEAP":
            raise H5Error(f"no local heap at {heap_addr}")
        return self.base + self.read_offset(a + 8 + 2 * self.L)

    def symbols(self, btree, heap):
        """(name, object header address) of every entry below a group B-tree, in tree (= name) order."""
        data = self._heap_data(heap)
        out = []
        self._walk_group(btree, data, out)
        return out

    def _walk_group(self, addr, heap_data, out):
        buf = self.buf
        a = self.base + addr
        sig = bytes(buf[a:a + 4])
        if sig == b"SNOD":
            n = struct.unpack_from("<H", buf, a + 6)[0]
            p = a + 8
            esz = 2 * self.O + 24
            for i in range(n):
                name_off = self.read_offset(p + i * esz)
                obj = self.read_offset(p + i * esz + self.O)
                q = heap_data + name_off
                end = q
                while buf[end] != 0:
                    end += 1
                out.append((bytes(buf[q:end]).decode("utf-8"), obj))
            return
        if sig != b"TREE" or buf[a + 4] != 0:
            raise H5Error(f"no group B-tree node at {addr}")
        used = struct.unpack_from("<H", buf, a + 6)[0]
        p = a + 8 + 2 * self.O + self.L                     # first child pointer (after key 0)
        for i in range(used):
            self._walk_group(self.read_offset(p + i * (self.O + self.L)), heap_data, out)

    def chunk_leaves(self, addr, rank):
        """(bytes stored, filter mask, chunk offsets, address) of every chunk below a v1 chunk B-tree."""
        buf = self.buf
        a = self.base + addr
        if bytes(buf[a:a + 4]) != b"TREE" or buf[a + 4] != 1:
            raise H5Error(f"no chunk B-tree node at {addr}")
        level = buf[a + 5]
        used = struct.unpack_from("<H", buf, a + 6)[0]
        ksz = 8 + 8 * (rank + 1)
        p = a + 8 + 2 * self.O
        for i in range(used):
            k = p + i * (ksz + self.O)
            nbytes, mask = struct.unpack_from("<II", buf, k)
            offs = struct.unpack_from("<" + "Q" * rank, buf, k + 8)
            child = self.read_offset(k + ksz)
            if level == 0:
                yield nbytes, mask, offs, child
            else:
                yield from self.chunk_leaves(child, rank)

    # -- attribute / element decoding
    def decode(self, dt, shape, p, end):
        buf = self.buf
        if shape is None:
            return None
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        if dt.vlen_str:
            vals = []
            for i in range(count):
                q = p + i * (4 + self.O + 4)
                ln = struct.unpack_from("<I", buf, q)[0]
                coll = self.read_offset(q + 4)
                idx = struct.unpack_from("<I", buf, q + 4 + self.O)[0]
                vals.append(self._global_heap_object(coll, idx)[:ln] if coll not in (0, self.undef) else b"")
            if not shape:
                return vals[0]
            return np.array(vals, dtype=object).reshape(shape)
        arr = np.frombuffer(bytes(buf[p:p + count * dt.size]), dtype=dt.dtype)
        if not shape:
            return arr[0]
        return arr.reshape(shape).copy()

    def _global_heap_object(self, coll, idx):
        buf = self.buf
        a = self.base + coll
        if bytes(buf[a:a + 4]) != b"GCOL":
            raise H5Error(f"no global heap collection at {coll}")
        size = self.read_length(a + 8)
        p = a + 8 + self.L
        end = a + size
        while p + 8 + self.L <= end:
            oidx = struct.unpack_from("<H", buf, p)[0]
            osz = self.read_length(p + 8)
            if oidx == 0:
                break
            if oidx == idx:
                return bytes(buf[p + 8 + self.L:p + 8 + self.L + osz])
            p += 8 + self.L + _pad8(osz)
        raise H5Error(f"global heap object {idx} not found in collection at {coll}")


# ------------------------------------------------------------------------------------------------------------- writer

def _dtype_message(dt):
    """Datatype message body (version 1) for a numpy dtype: fixed point, IEEE float, fixed-length string."""
    dt = np.dtype(dt)
    if dt.kind in "iu":
        bits = (8 if dt.kind == "i" else 0) | (1 if dt.byteorder == ">" else 0)
        return struct.pack("<BBBBIHH", 0x10, bits, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == "f":
        # class bits: byte order, mantissa normalisation 2 (implied msb) in bits 4-5, sign bit position in byte 1
        spec = {2: (15, 10, 5, 0, 10, 15), 4: (31, 23, 8, 0, 23, 127), 8: (63, 52, 11, 0, 52, 1023)}[dt.itemsize]
        sign, eloc, esz, mloc, msz, bias = spec
        b0 = 0x20 | (1 if dt.byteorder == ">" else 0)
        return struct.pack("<BBBBIHHBBBBI", 0x11, b0, sign, 0, dt.itemsize, 0, 8 * dt.itemsize, eloc, esz, mloc, msz,
                           bias)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 1, 0, 0, dt.itemsize)       # null-padded ASCII, as numpy 'S' arrays are
    raise H5Error(f"cannot store dtype {dt}")


def _dataspace_message(shape):
    shape = tuple(int(s) for s in shape)
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", s) for s in shape)


def _attr_message(name, value):
    if isinstance(value, str):
        value = value.encode("utf-8")
    if isinstance(value, bytes):
        value = np.array(value, dtype=f"S{max(1, len(value))}")
    value = np.asarray(value)
    if value.dtype.kind == "U":
        value = np.char.encode(value, "utf-8")
    nm = name.encode("utf-8") + b"\0"
    dtm, dsm = _dtype_message(value.dtype), _dataspace_message(value.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dtm), len(dsm))
    body += nm.ljust(_pad8(len(nm)), b"\0") + dtm.ljust(_pad8(len(dtm)), b"\0") + dsm.ljust(_pad8(len(dsm)), b"\0")
    body += np.ascontiguousarray(value).tobytes()
    if len(body) > 65000:
        raise H5Error(f"attribute {name!r} is {len(body)} bytes: the object-header limit is 64 KB "
                      "(Keras splits such lists into name0, name1, ...)")
    return 0x0C, body


class Writer:
    """Builds an HDF5 file of groups, contiguous datasets and compact attributes.

    ``w = Writer(); g = w.root.create_group("a"); g.attrs["x"] = ...; g.create_dataset("k", array); w.save(path)``
    """

    LEAF_K, INTERNAL_K = 4, 16        # library defaults: 8 symbols per SNOD, 32 children per B-tree node

    class _G:
        def __init__(self):
            self.attrs = {}
            self.children = {}        # name -> _G | numpy array wrapper

        def create_group(self, name):
            node = self
            for part in [p for p in name.split("/") if p]:
                nxt = node.children.get(part)
                if nxt is None:
                    nxt = node.children[part] = Writer._G()
                elif not isinstance(nxt, Writer._G):
                    raise H5Error(f"{part!r} is a dataset")
                node = nxt
            return node

        require_group = create_group

        def create_dataset(self, name, data, chunks=None, compression=None, shuffle=False):
            """Contiguous by default (what Keras writes).  ``chunks=(..)`` stores the array as a chunked dataset with a
            v1 chunk B-tree (at most 64 chunks), optionally with the shuffle and deflate (``compression="gzip"``)
            filters - the layout tools that compress weight files produce."""
            parts = [p for p in name.split("/") if p]
            node = self.create_group("/".join(parts[:-1])) if len(parts) > 1 else self
            if parts[-1] in node.children:
                raise H5Error(f"{name!r} exists")
            d = Writer._D(np.asarray(data).copy(order="C"))       # keeps 0-d arrays 0-d
            if chunks is not None:
                if len(chunks) != d.data.ndim or d.data.ndim == 0 or min(chunks) < 1:
                    raise H5Error("chunks must give one positive extent per dimension")
                d.chunks = tuple(int(c) for c in chunks)
                d.filters = ([2] if shuffle else []) + ([1] if compression in ("gzip", "deflate") else [])
            elif compression or shuffle:
                raise H5Error("filters need a chunked dataset")
            node.children[parts[-1]] = d
            return d

    class _D:
        chunks = None
        filters = ()

        def __init__(self, data):
            self.data = data
            self.attrs = {}

    def __init__(self):
        self.root = Writer._G()
        self._out = bytearray()

    # sequential allocator, every structure 8-byte aligned
    def _alloc(self, n):
        a = len(self._out)
        self._out.extend(b"\0" * _pad8(n))
        return a

    def _put(self, addr, data):
        self._out[addr:addr + len(data)] = data

    def _object_header(self, messages):
        body = b""
        for mtype, data in messages:
            data = data.ljust(_pad8(len(data)), b"\0")
            body += struct.pack("<HHB3x", mtype, len(data), 0) + data
        addr = self._alloc(16 + len(body))
        self._put(addr, struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(body)) + body)
        return addr

    def _write_dataset(self, d):
        arr = d.data
        dt = arr.dtype
        if dt.kind == "U":
            arr = np.char.encode(arr, "utf-8")
        if d.chunks is not None:
            return self._write_chunked(d, arr)
        data_addr = self._alloc(max(arr.nbytes, 1))
        self._put(data_addr, arr.tobytes())
        msgs = [(0x01, _dataspace_message(arr.shape)), (0x03, _dtype_message(arr.dtype)),
                # fill value as the library writes it for a default dataset (version 2: allocate late, write if set,
                # "defined" with size 0 = the default fill value), then layout v3 contiguous
                (0x05, struct.pack("<BBBBI", 2, 2, 2, 1, 0)),
                (0x08, struct.pack("<BBQQ", 3, 1, data_addr, arr.nbytes))]
        msgs += [_attr_message(k, v) for k, v in d.attrs.items()]
        return self._object_header(msgs)

    def _write_chunked(self, d, arr):
        """Chunked layout (IV.A.2.i class 2) with a one-node v1 chunk B-tree (III.A.1, node type 1) and the filter
        pipeline message (IV.A.2.l).  Edge chunks are stored at full chunk size, as the library does."""
        rank, cd, esz = arr.ndim, d.chunks, arr.dtype.itemsize
        grid = [range(0, max(n, 1), c) for n, c in zip(arr.shape, cd)]
        entries = []
        for offs in np.ndindex(*[len(g) for g in grid]) if arr.size else []:
            o = [grid[i][k] for i, k in enumerate(offs)]
            block = np.zeros(cd, arr.dtype)
            src = arr[tuple(slice(a, a + c) for a, c in zip(o, cd))]
            block[tuple(slice(0, n) for n in src.shape)] = src
            raw = block.tobytes()
            for fid in d.filters:
                if fid == 2:
                    raw = np.frombuffer(raw, np.uint8).reshape(-1, esz).T.tobytes()
                elif fid == 1:
                    raw = zlib.compress(raw, 4)
            a = self._alloc(len(raw))
            self._put(a, raw)
            entries.append((len(raw), o, a))
        if len(entries) > 64:
            raise H5Error("this writer keeps the chunk index in one B-tree node (at most 64 chunks)")
        ksz = 8 + 8 * (rank + 1)
        node = self._alloc(24 + 64 * (ksz + 8) + ksz)
        body = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(entries), UNDEF, UNDEF)
        for nbytes, o, a in entries:
            body += struct.pack("<II", nbytes, 0) + b"".join(struct.pack("<Q", x) for x in o + [0]) + struct.pack("<Q", a)
        body += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", x) for x in list(arr.shape) + [0])   # final key
        self._put(node, body)
        layout = struct.pack("<BBBQ", 3, 2, rank + 1, node) + b"".join(struct.pack("<I", c) for c in list(cd) + [esz])
        msgs = [(0x01, _dataspace_message(arr.shape)), (0x03, _dtype_message(arr.dtype)),
                (0x05, struct.pack("<BBBBI", 2, 3, 2, 1, 0)), (0x08, layout)]       # allocation time 3 = incremental
        if d.filters:
            pipe = struct.pack("<BB6x", 1, len(d.filters))
            for fid in d.filters:
                cdv = [esz] if fid == 2 else [4]
                pipe += struct.pack("<HHHH", fid, 0, 1 if fid == 2 else 0, len(cdv))
                pipe += b"".join(struct.pack("<I", v) for v in cdv) + (b"\0" * 4 if len(cdv) & 1 else b"")
            msgs.append((0x0B, pipe))
        msgs += [_attr_message(k, v) for k, v in d.attrs.items()]
        return self._object_header(msgs)

    def _write_group(self, g):
        """Children first, then local heap, symbol nodes, B-tree, object header.  Returns (header, btree, heap)."""
        entries = []
        for name in g.children:
            child = g.children[name]
            if isinstance(child, Writer._G):
                hdr, bt, hp = self._write_group(child)
                entries.append((name.encode("utf-8"), hdr, 1, bt, hp))
            else:
                entries.append((name.encode("utf-8"), self._write_dataset(child), 0, 0, 0))
        entries.sort(key=lambda e: e[0])                    # symbol nodes are ordered by strcmp of the names
        # local heap: offset 0 holds the empty string (key 0 of the leftmost node); names 8-byte aligned
        heap = bytearray(b"\0" * 8)
        offs = []
        for nm, *_ in entries:
            offs.append(len(heap))
            heap.extend(nm + b"\0")
            heap.extend(b"\0" * (_pad8(len(heap)) - len(heap)))
        free_off = len(heap)
        heap.extend(struct.pack("<QQ", 1, 16))             # one free block at the end: next = 1 (none), size 16
        heap_data = self._alloc(len(heap))
        self._put(heap_data, bytes(heap))
        heap_addr = self._alloc(32)
        self._put(heap_addr, b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, heap_data))
        # symbol nodes
        per = 2 * self.LEAF_K
        level = []                                          # (address, heap offset of the largest name below)
        for i in range(0, len(entries), per):               # an empty group gets a B-tree node with no children
            chunk = list(zip(entries[i:i + per], offs[i:i + per]))
            a = self._alloc(8 + per * 40)
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(chunk))
            for (nm, hdr, ctype, bt, hp), off in chunk:
                body += struct.pack("<QQII", off, hdr, ctype, 0) + (struct.pack("<QQ", bt, hp) if ctype == 1 else b"\0" * 16)
            self._put(a, body)
            level.append((a, chunk[-1][1]))
        # B-tree levels above
        fan = 2 * self.INTERNAL_K
        node_size = 24 + (2 * fan + 1) * 8
        depth = 0
        while True:
            nodes = []
            left_key = 0
            groups = [level[i:i + fan] for i in range(0, len(level), fan)] or [[]]
            addrs = [self._alloc(node_size) for _ in groups]
            for gi, grp in enumerate(groups):
                left = addrs[gi - 1] if gi > 0 else UNDEF
                right = addrs[gi + 1] if gi + 1 < len(groups) else UNDEF
                body = b"TREE" + struct.pack("<BBHQQ", 0, depth, len(grp), left, right) + struct.pack("<Q", left_key)
                for child, key in grp:
                    body += struct.pack("<QQ", child, key)
                self._put(addrs[gi], body)
                left_key = grp[-1][1] if grp else 0
                nodes.append((addrs[gi], left_key))
            level = nodes
            depth += 1
            if len(level) == 1:
                break
        btree_addr = level[0][0]
        msgs = [(0x11, struct.pack("<QQ", btree_addr, heap_addr))]
        msgs += [_attr_message(k, v) for k, v in g.attrs.items()]
        return self._object_header(msgs), btree_addr, heap_addr

    def tobytes(self):
        self._out = bytearray()
        sb = self._alloc(96)
        hdr, bt, hp = self._write_group(self.root)
        eof = len(self._out)
        head = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, self.LEAF_K, self.INTERNAL_K, 0)
        head += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        head += struct.pack("<QQII", 0, hdr, 1, 0) + struct.pack("<QQ", bt, hp)
        self._put(sb, head)
        return bytes(self._out)

    def save(self, path):
        data = self.tobytes()
        with open(path, "wb") as fh:
            fh.write(data)


# ------------------------------------------------------------------------------------------------ Keras weight layout

def _as_str(v):
    return v.decode("utf-8") if isinstance(v, (bytes, np.bytes_)) else str(v)


def _load_list_attr(node, name):
    """keras.engine.saving.load_attributes_from_hdf5_group: ``name`` or the chunks ``name0``, ``name1``, ..."""
    attrs = node.attrs
    if name in attrs:
        if attrs[name] is None:                              # null dataspace: an empty list
            return []
        return [_as_str(n) for n in np.atleast_1d(attrs[name])]
    out, i = [], 0
    while f"{name}{i}" in attrs:
        out.extend(_as_str(n) for n in np.atleast_1d(attrs[f"{name}{i}"]))
        i += 1
    if i == 0:
        raise H5Error(f"{node.name}: no attribute {name!r} (not a Keras weight file?)")
    return out


def load_keras_weights(path):
    """Read a Keras HDF5 file - ``model.save("x.h5")`` (weights under ``/model_weights``) or ``save_weights("x.h5")``.

    Returns ``(layers, meta)``: ``layers`` = [(layer name, [(weight name, fp32 array), ...])] in the file's
    ``layer_names`` / ``weight_names`` order, i.e. Keras creation order, layers without weights included;
    ``meta`` = the root attributes decoded to str (``model_config`` JSON, ``keras_version``, ``backend``) when present.
    """
    with File(path) as f:
        root = f["model_weights"] if "model_weights" in f.keys() else f
        meta = {}
        for k, v in f.attrs.items():
            if isinstance(v, (bytes, np.bytes_, str)):
                meta[k] = _as_str(v)
        layers = []
        for lname in _load_list_attr(root, "layer_names"):
            g = root[lname]
            ws = []
            for wname in _load_list_attr(g, "weight_names") if ("weight_names" in g.attrs or "weight_names0" in g.attrs) \
                    else []:
                ws.append((wname, np.asarray(g[wname].read())))
            layers.append((lname, ws))
    return layers, meta


def save_keras_weights(path, layers, model_config=None, under_model_weights=None):
    """Write ``layers`` = [(layer name, [(weight name, array), ...])] in the Keras HDF5 weight layout
    (keras.engine.saving.save_weights_to_hdf5_group): root attributes ``layer_names`` / ``backend`` / ``keras_version``,
    one group per layer with ``weight_names`` and one dataset per weight at ``<layer>/<weight name>``.  With
    ``model_config`` (a JSON string) the weights go under ``/model_weights`` as ``model.save`` does."""
    w = Writer()
    if under_model_weights is None:
        under_model_weights = model_config is not None
    root = w.root.create_group("model_weights") if under_model_weights else w.root
    if model_config is not None:
        mc = model_config.encode("utf-8")
        if len(mc) > 60000:
            raise H5Error("model_config exceeds the compact-attribute limit of this writer; pass model_config=None")
        w.root.attrs["model_config"] = mc
    for node in ({id(root): root, id(w.root): w.root}).values():
        node.attrs["backend"] = b"tensorflow"
        node.attrs["keras_version"] = b"2.3.0-tf"

    def put_list(node, name, items):
        items = [i.encode("utf-8") for i in items]
        arr = np.array(items, dtype=f"S{max([1] + [len(i) for i in items])}")
        if arr.nbytes <= 60000:
            node.attrs[name] = arr
            return
        per = max(1, 60000 // arr.dtype.itemsize)           # HDF5_OBJECT_HEADER_LIMIT chunking of Keras
        for ci, i in enumerate(range(0, len(items), per)):
            node.attrs[f"{name}{ci}"] = arr[i:i + per]

    put_list(root, "layer_names", [ln for ln, _ in layers])
    for lname, ws in layers:
        g = root.create_group(lname)
        put_list(g, "weight_names", [wn for wn, _ in ws])
        for wname, arr in ws:
            g.create_dataset(wname, np.asarray(arr))
    w.save(path)
