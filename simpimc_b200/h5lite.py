"""A small reader and writer for the subset of HDF5 the reference's files use -- no HDF5 library
exists in this image (no libhdf5, no h5py), and the reference keeps its pair-action tables
(src/actions/pair_action/ilkka_pair_action_class.h:266-418, david_pair_action_class.h:194-336, written by
scripts/pagen/IlkkaSquarer.py:118-163 / DavidParse.py:160-225 through h5py) and its block output
(include/scaffold/io/io_hdf5.h:106-212) in HDF5.

Implemented from the HDF5 file-format specification (version 1.8 family, what h5py's default
libver="earliest" and the reference's HDF5 1.8.13 write):

  read   superblock 0 / 1 at offset 0, 512, 1024, ... (user block); old-style groups (symbol-table message,
         v1 B-tree, SNOD nodes, local heap); version-1 object headers with continuation blocks; dataspace
         messages v1 / v2; datatypes: fixed-point, IEEE floating point, fixed-length strings, variable-length
         strings (global heap); data layout v1 / v2 / v3: compact, contiguous, chunked (v1 B-tree) with the
         deflate and shuffle filters; attributes (v1).  Superblock 2 / 3, fractal-heap ("new style") groups and
         version-2 object headers are recognised and rejected with a clear message.
  write  superblock 0, old-style groups, contiguous datasets of float64 / int32 / uint32 / int64 arrays and
         scalars, fixed-length strings -- the layout h5py itself produces for such data, so the files open in
         any HDF5 tool.

`read(path)` returns {"group/sub/dataset": numpy array | str}; `write(path, mapping)` takes the same.
Validated against the one file in this image that the HDF5 library itself wrote
(scipy/io/matlab/tests/data/testhdf5_7.4_GLNX86.mat, a MATLAB 7.3 file: HDF5 behind a 512-byte user block;
tests/test_h5lite_cpu.py) and by round trips.
"""
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(ValueError):
    pass


# ------------------------------------------------------------------------------------ reader
class _Reader:
    def __init__(self, buf):
        self.buf = buf
        self.base = None
        for off in [0] + [512 << i for i in range(0, 16)]:
            if buf[off:off + 8] == SIGNATURE:
                self.sb = off
                break
        else:
            raise H5Error("not an HDF5 file (no superblock signature)")
        p = self.sb + 8
        version = buf[p]
        if version >= 2:
            raise H5Error("HDF5 superblock version %d (libver='latest') is not supported; rewrite the file with the default format" % version)
        self.O, self.L = buf[p + 5], buf[p + 6]
        if self.O != 8 or self.L != 8:
            raise H5Error("only 8-byte offsets and lengths are supported")
        p += 8 + 2 + 2 + 4                       # versions + sizes, group leaf / internal K, consistency flags
        if version == 1:
            p += 4                               # indexed-storage K + reserved
        self.base, _free, self.eof, _drv = struct.unpack_from("<4Q", buf, p)
        p += 32
        # root group symbol-table entry
        _name_off, self.root_header, cache_type = struct.unpack_from("<QQI", buf, p)
        self.root_cache = struct.unpack_from("<QQ", buf, p + 24) if cache_type == 1 else None

    def at(self, addr):
        return self.base + addr

    # ---- object headers ----
    def messages(self, addr):
        """[(type, flags, bytes)] of a version-1 object header, continuation blocks followed."""
        buf = self.buf
        p = self.at(addr)
        if buf[p:p + 4] == b"OHDR":
            raise H5Error("version-2 object headers (libver='latest') are not supported")
        version, _, n_msg, _refs, size = struct.unpack_from("<BBHII", buf, p)
        if version != 1:
            raise H5Error("object header version %d" % version)
        out = []
        blocks = [(p + 16, size)]
        while blocks and len(out) < n_msg:
            q, left = blocks.pop(0)
            end = q + left
            while q + 8 <= end and len(out) < n_msg:
                mtype, msize, flags = struct.unpack_from("<HHB", buf, q)
                data = bytes(buf[q + 8:q + 8 + msize])
                q += 8 + msize
                if mtype == 0x0010:              # continuation
                    c_off, c_len = struct.unpack_from("<QQ", data, 0)
                    blocks.append((self.at(c_off), c_len))
                out.append((mtype, flags, data))
        return out

    # ---- groups ----
    def heap_string(self, heap_addr, offset):
        p = self.at(heap_addr)
        if self.buf[p:p + 4] != b"HEAP":
            raise H5Error("local heap signature missing")
        _size, _free, data_addr = struct.unpack_from("<QQQ", self.buf, p + 8)
        s = self.at(data_addr) + offset
        e = self.buf.index(b"\0", s)
        return bytes(self.buf[s:e]).decode("utf-8")

    def group_entries(self, btree_addr, heap_addr):
        """(name, object header address) of every link below a v1 group B-tree."""
        p = self.at(btree_addr)
        buf = self.buf
        if buf[p:p + 4] != b"TREE":
            raise H5Error("group B-tree signature missing")
        node_type, level, n_used = struct.unpack_from("<BBH", buf, p + 4)
        if node_type != 0:
            raise H5Error("not a group B-tree")
        q = p + 8 + 16                           # skip the sibling addresses
        out = []
        for i in range(n_used):
            child, = struct.unpack_from("<Q", buf, q + 8)        # key_i, child_i, key_i+1 ...
            q += 16
            if level > 0:
                out += self.group_entries(child, heap_addr)
            else:
                s = self.at(child)
                if buf[s:s + 4] != b"SNOD":
                    raise H5Error("symbol-table node signature missing")
                n_sym, = struct.unpack_from("<H", buf, s + 6)
                for k in range(n_sym):
                    name_off, header = struct.unpack_from("<QQ", buf, s + 8 + 40 * k)
                    out.append((self.heap_string(heap_addr, name_off), header))
        return out

    # ---- datatypes ----
    def datatype(self, data, p=0):
        """-> (kind, numpy dtype or None, element size, extra)"""
        cls_ver, b0, b1, _b2, size = struct.unpack_from("<BBBBI", data, p)
        cls = cls_ver & 0x0F
        if cls == 0:
            if b0 & 1:
                raise H5Error("big-endian integers are not supported")
            return "int", np.dtype("<%s%d" % ("i" if b0 & 8 else "u", size)), size, None
        if cls == 1:
            if b0 & 1:
                raise H5Error("big-endian floats are not supported")
            return "float", np.dtype("<f%d" % size), size, None
        if cls == 3:
            return "string", None, size, b0 & 0x0F      # padding type: 0 null-terminated, 1 null-padded, 2 space-padded
        if cls == 9:
            if (b0 & 0x0F) != 1:
                raise H5Error("variable-length sequences are not supported (only strings)")
            return "vlen_string", None, size, None
        raise H5Error("HDF5 datatype class %d is not supported" % cls)

    def dataspace(self, data):
        version, rank, flags = struct.unpack_from("<BBB", data, 0)
        p = 8 if version == 1 else 4
        if version == 2 and data[3] == 2:
            return None                           # null dataspace
        return tuple(struct.unpack_from("<%dQ" % rank, data, p)) if rank else ()

    # ---- raw data ----
    def chunks(self, btree_addr, n_dims):
        p = self.at(btree_addr)
        buf = self.buf
        if buf[p:p + 4] != b"TREE":
            raise H5Error("chunk B-tree signature missing")
        node_type, level, n_used = struct.unpack_from("<BBH", buf, p + 4)
        if node_type != 1:
            raise H5Error("not a chunk B-tree")
        key = 8 + 8 * n_dims                     # chunk size, filter mask, offsets (one per dimension incl. the element one)
        q = p + 8 + 16
        out = []
        for i in range(n_used):
            c_size, mask = struct.unpack_from("<II", buf, q)
            offs = struct.unpack_from("<%dQ" % n_dims, buf, q + 8)
            child, = struct.unpack_from("<Q", buf, q + key)
            q += key + 8
            if level > 0:
                out += self.chunks(child, n_dims)
            else:
                out.append((offs[:-1], c_size, mask, child))
        return out

    def read_dataset(self, msgs):
        space = dtype = layout = None
        filters = []
        for mtype, _flags, data in msgs:
            if mtype == 0x0001:
                space = self.dataspace(data)
            elif mtype == 0x0003:
                dtype = self.datatype(data)
            elif mtype == 0x0008:
                layout = data
            elif mtype == 0x000B:
                filters = self.filter_pipeline(data)
        if space is None or dtype is None or layout is None:
            return None
        kind, np_dtype, esize, extra = dtype
        n = int(np.prod(space)) if space else 1
        raw = self.raw_bytes(layout, space, esize, n, filters)
        if kind in ("int", "float"):
            arr = np.frombuffer(raw, dtype=np_dtype, count=n).copy()
            return arr.reshape(space) if space else arr.reshape(())[()]
        if kind == "string":
            items = [raw[i * esize:(i + 1) * esize].split(b"\0")[0].rstrip(b" " if extra == 2 else b"").decode("utf-8") for i in range(n)]
        else:
            items = []
            for i in range(n):
                length, addr, index = struct.unpack_from("<IQI", raw, 16 * i)
                items.append(self.global_heap_object(addr, index)[:length].decode("utf-8"))
        return items[0] if not space else np.array(items, dtype=object).reshape(space)

    def filter_pipeline(self, data):
        version, n_filters = data[0], data[1]
        p = 8 if version == 1 else 2
        out = []
        for _ in range(n_filters):
            fid, name_len, _flags, n_cd = struct.unpack_from("<HHHH", data, p)
            p += 8
            if version == 1 or fid >= 256:
                p += (name_len + 7) // 8 * 8 if version == 1 else name_len
            cd = struct.unpack_from("<%dI" % n_cd, data, p)
            p += 4 * n_cd
            if version == 1 and n_cd % 2:
                p += 4
            out.append((fid, cd))
        return out

    def raw_bytes(self, layout, space, esize, n, filters):
        buf = self.buf
        version = layout[0]
        if version == 3:
            cls = layout[1]
            if cls == 0:
                size, = struct.unpack_from("<H", layout, 2)
                return bytes(layout[4:4 + size])
            if cls == 1:
                addr, size = struct.unpack_from("<QQ", layout, 2)
                if addr == UNDEF:
                    return bytes(n * esize)       # never written: fill value 0
                return bytes(buf[self.at(addr):self.at(addr) + size])
            if cls == 2:
                n_dims = layout[2]
                btree, = struct.unpack_from("<Q", layout, 3)
                cdims = struct.unpack_from("<%dI" % n_dims, layout, 11)
                return self.assemble_chunks(btree, n_dims, cdims[:-1], space, esize, filters)
            raise H5Error("data layout class %d" % cls)
        if version in (1, 2):
            n_dims, cls = layout[1], layout[2]
            p = 8
            addr = None
            if cls != 0:
                addr, = struct.unpack_from("<Q", layout, p)
                p += 8
            dims = struct.unpack_from("<%dI" % n_dims, layout, p)
            p += 4 * n_dims
            if cls == 1:
                return bytes(buf[self.at(addr):self.at(addr) + n * esize])
            if cls == 2:
                return self.assemble_chunks(addr, n_dims, dims[:-1] if len(dims) == n_dims else dims, space, esize, filters)
            size, = struct.unpack_from("<I", layout, p)
            return bytes(layout[p + 4:p + 4 + size])
        raise H5Error("data layout message version %d" % version)

    def assemble_chunks(self, btree, n_dims, cdims, space, esize, filters):
        out = np.zeros(space, dtype="V%d" % esize)
        if btree == UNDEF:
            return out.tobytes()
        for offs, c_size, mask, addr in self.chunks(btree, n_dims):
            raw = bytes(self.buf[self.at(addr):self.at(addr) + c_size])
            for k, (fid, cd) in reversed(list(enumerate(filters))):
                if mask & (1 << k):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:                     # shuffle: bytes of every element were transposed
                    a = np.frombuffer(raw, dtype=np.uint8)
                    ne = len(a) // esize
                    raw = a[:ne * esize].reshape(esize, ne).T.tobytes()
                elif fid == 3:                     # fletcher32 checksum trails the chunk
                    raw = raw[:-4]
                else:
                    raise H5Error("HDF5 filter %d is not supported" % fid)
            chunk = np.frombuffer(raw, dtype="V%d" % esize, count=int(np.prod(cdims))).reshape(cdims)
            sl_dst = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cdims, space))
            sl_src = tuple(slice(0, d.stop - d.start) for d in sl_dst)
            out[sl_dst] = chunk[sl_src]
        return out.tobytes()

    def global_heap_object(self, addr, index):
        p = self.at(addr)
        buf = self.buf
        if buf[p:p + 4] != b"GCOL":
            raise H5Error("global heap signature missing")
        size, = struct.unpack_from("<Q", buf, p + 8)
        q, end = p + 16, p + size
        while q + 16 <= end:
            idx, _refs, _r, osize = struct.unpack_from("<HHIQ", buf, q)
            if idx == index:
                return bytes(buf[q + 16:q + 16 + osize])
            if idx == 0:
                break
            q += 16 + (osize + 7) // 8 * 8
        raise H5Error("global heap object %d not found" % index)

    def attributes(self, msgs):
        out = {}
        for mtype, _flags, data in msgs:
            if mtype != 0x000C or data[0] != 1:
                continue
            name_size, dt_size, ds_size = struct.unpack_from("<HHH", data, 2)
            pad = lambda v: (v + 7) // 8 * 8
            p = 8
            name = data[p:p + name_size].split(b"\0")[0].decode("utf-8")
            p += pad(name_size)
            kind, np_dtype, esize, extra = self.datatype(data, p)
            p += pad(dt_size)
            space = self.dataspace(data[p:p + ds_size])
            p += pad(ds_size)
            n = int(np.prod(space)) if space else 1
            raw = data[p:p + n * esize]
            if kind in ("int", "float"):
                arr = np.frombuffer(raw, dtype=np_dtype, count=n).copy()
                out[name] = arr.reshape(space) if space else arr.reshape(())[()]
            elif kind == "string":
                out[name] = raw.split(b"\0")[0].decode("utf-8")
        return out

    # ---- tree walk ----
    def walk(self, header_addr, prefix, out, attrs, depth=0):
        if depth > 64:
            raise H5Error("group nesting too deep (cycle?)")
        msgs = self.messages(header_addr)
        a = self.attributes(msgs)
        if a:
            attrs[prefix.rstrip("/")] = a
        for mtype, _flags, data in msgs:
            if mtype == 0x0011:                   # symbol-table message: this object is a group
                btree, heap = struct.unpack_from("<QQ", data, 0)
                for name, child in self.group_entries(btree, heap):
                    self.walk(child, prefix + name + "/", out, attrs, depth + 1)
                return
            if mtype in (0x0002, 0x0006):
                raise H5Error("fractal-heap ('new style') groups are not supported; write the file with libver='earliest'")
        value = self.read_dataset(msgs)
        if value is not None:
            out[prefix.rstrip("/")] = value


def read(path, with_attributes=False):
    """Every dataset of an HDF5 file as {"group/.../name": array | scalar | str}."""
    with open(path, "rb") as f:
        buf = f.read()
    r = _Reader(buf)
    out, attrs = {}, {}
    r.walk(r.root_header, "", out, attrs)
    return (out, attrs) if with_attributes else out


def is_hdf5(path):
    with open(path, "rb") as f:
        head = f.read(8)
        if head == SIGNATURE:
            return True
        for off in (512, 1024, 2048):
            f.seek(off)
            if f.read(8) == SIGNATURE:
                return True
    return False


# ------------------------------------------------------------------------------------ writer
class _Writer:
    """Superblock 0, old-style groups (one SNOD leaf per group, <= 2 * leaf K entries ... split into several
    leaves under one level-0 B-tree node), contiguous datasets."""

    LEAF_K, INTERNAL_K = 4, 16

    def __init__(self):
        self.buf = bytearray()

    def alloc(self, n, align=8):
        while len(self.buf) % align:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf.extend(bytes(n))
        return addr

    def put(self, addr, data):
        self.buf[addr:addr + len(data)] = data

    @staticmethod
    def msg(mtype, data, flags=0):
        data = bytes(data) + bytes((-len(data)) % 8)
        return struct.pack("<HHBBBB", mtype, len(data), flags, 0, 0, 0) + data

    def object_header(self, msgs):
        body = b"".join(msgs)
        hdr = struct.pack("<BBHII", 1, 0, len(msgs), 1, len(body)) + bytes(4)
        addr = self.alloc(len(hdr) + len(body))
        self.put(addr, hdr + body)
        return addr

    @staticmethod
    def datatype(value):
        """-> (datatype message bytes, element size, raw bytes, shape)"""
        if isinstance(value, (str, bytes)):
            b = value.encode("utf-8") if isinstance(value, str) else value
            size = len(b) + 1
            return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, size), size, b + b"\0", ()      # class 3 v1, null-terminated ASCII
        a = np.asarray(value)
        if a.dtype.kind == "f":
            a = a.astype("<f8")
            dt = struct.pack("<BBBBI", 0x11, 0x20, 0x3F, 0x00, 8) + struct.pack("<HHBBBBI", 0, 64, 52, 11, 0, 52, 1023)
        elif a.dtype.kind in "iub":
            signed = a.dtype.kind == "i"
            size = 8 if a.dtype.itemsize == 8 else 4
            a = a.astype("<%s%d" % ("i" if signed else "u", size))
            dt = struct.pack("<BBBBI", 0x10, 0x08 if signed else 0x00, 0, 0, size) + struct.pack("<HH", 0, 8 * size)
        else:
            raise H5Error("cannot write dtype %s" % a.dtype)
        return dt, a.dtype.itemsize, np.ascontiguousarray(a).tobytes(), a.shape

    def dataset(self, value):
        dt, _esize, raw, shape = self.datatype(value)
        rank = len(shape)
        space = struct.pack("<BBBB4x", 1, rank, 0, 0) + struct.pack("<%dQ" % rank, *shape) if rank else struct.pack("<BBBB4x", 1, 0, 0, 0)
        data_addr = self.alloc(len(raw)) if raw else UNDEF
        if raw:
            self.put(data_addr, raw)
        layout = struct.pack("<BBQQ", 3, 1, data_addr, len(raw))
        fill = struct.pack("<BBBB", 2, 2, 0, 0)            # fill value message v2: allocate late, write if set, undefined
        return self.object_header([self.msg(0x0001, space), self.msg(0x0003, dt, flags=1), self.msg(0x0005, fill),
                                   self.msg(0x0008, layout)])

    def group(self, tree):
        """tree: {name: subtree dict | value} -> object header address of the group."""
        names = sorted(tree)                       # B-tree keys are ordered by name
        children = {}
        for name in names:
            v = tree[name]
            children[name] = self.group(v) if isinstance(v, dict) else self.dataset(v)
        # local heap: the empty string at offset 0, then the names (each 8-byte aligned)
        heap = bytearray(8)
        offs = {}
        for name in names:
            offs[name] = len(heap)
            b = name.encode("utf-8") + b"\0"
            heap += b + bytes((-len(b)) % 8)
        free_off = len(heap)
        heap += struct.pack("<QQ", 1, 16)           # one free block: next = 1 (none), size 16
        heap_data = self.alloc(len(heap))
        self.put(heap_data, heap)
        heap_addr = self.alloc(32)
        self.put(heap_addr, b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), free_off, heap_data))
        # symbol-table nodes of <= 2 K entries under one B-tree node
        per = 2 * self.LEAF_K
        leaves = [names[i:i + per] for i in range(0, len(names), per)] or [[]]
        if len(leaves) > 2 * self.INTERNAL_K:
            raise H5Error("more than %d links in one group" % (per * 2 * self.INTERNAL_K))
        snods = []
        for leaf in leaves:
            addr = self.alloc(8 + 40 * per)
            body = b"SNOD" + struct.pack("<BBH", 1, 0, len(leaf))
            for name in leaf:
                body += struct.pack("<QQII16x", offs[name], children[name], 0, 0)
            self.put(addr, body)
            snods.append(addr)
        btree = self.alloc(24 + (2 * self.INTERNAL_K + 1) * 8 + 2 * self.INTERNAL_K * 8)
        node = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF) + struct.pack("<Q", 0)
        for leaf, addr in zip(leaves, snods):
            node += struct.pack("<QQ", addr, offs[leaf[-1]] if leaf else 0)
        self.put(btree, node)
        header = self.object_header([self.msg(0x0011, struct.pack("<QQ", btree, heap_addr))])
        self._last_group = (btree, heap_addr)
        return header

    def finish(self, tree):
        sb = self.alloc(96)
        root = self.group(tree)
        btree, heap = self._last_group
        eof = len(self.buf)
        s = SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", self.LEAF_K, self.INTERNAL_K, 0)
        s += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        s += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", btree, heap)
        self.put(sb, s)
        return bytes(self.buf)


def write(path, mapping):
    """{"group/sub/name": array | scalar | str} -> an HDF5 file (superblock 0, old-style groups, contiguous datasets)."""
    tree = {}
    for key, value in mapping.items():
        parts = [p for p in key.split("/") if p]
        node = tree
        for p in parts[:-1]:
            node = node.setdefault(p, {})
            if not isinstance(node, dict):
                raise H5Error("%s is both a dataset and a group" % p)
        node[parts[-1]] = value
    data = _Writer().finish(tree)
    with open(path, "wb") as f:
        f.write(data)
