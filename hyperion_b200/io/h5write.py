"""Minimal pure-Python HDF5 writer for ``.rtout`` files (and ``.rtin`` files in the tests).

The reference back end writes its output with the HDF5 Fortran library
(``src/main/main.f90:125-150,241-246,341-344``; datasets ``src/grid/grid_generic.f90:50-63``,
``src/images/image_type.f90:608-788``).  libhdf5 / h5py are not available where this engine
runs, so the host side carries its own writer for the small subset those files need:

* superblock version 0, 8-byte offsets and lengths
* groups in the compact "new style" (object header v1 with Link Info, Group Info and one Link
  message per child) -- the same form the reference's own output has for its root group, which
  holds the external link ``/Input`` when ``copy_input = no``
* hard links, soft links and external links
* datasets with contiguous layout; numpy dtypes: integers, IEEE floats, fixed-length strings,
  structured dtypes (compound, version 3) including sub-array members (array class, version 3)
* attributes (version 1 messages) of the same dtypes, scalar or n-dimensional

Counterpart of :mod:`hyperion_b200.io.h5min` (reader); ``tests/test_h5_roundtrip.py`` checks the
two against each other and the reader against the reference's own fixtures.
"""
from __future__ import annotations

import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"


def _pad8(b: bytes) -> bytes:
    return b + b"\x00" * ((-len(b)) % 8)


class ExternalLink:
    def __init__(self, filename, path):
        self.filename, self.path = filename, path


class SoftLink:
    def __init__(self, path):
        self.path = path


class _Node:
    def __init__(self):
        self.attrs = {}


class Dataset(_Node):
    def __init__(self, data):
        super().__init__()
        self.data = np.ascontiguousarray(data)


class Group(_Node):
    def __init__(self):
        super().__init__()
        self.children = {}

    # -- h5py-like construction API ------------------------------------------------------
    def create_group(self, path):
        node = self
        for part in path.strip("/").split("/"):
            nxt = node.children.get(part)
            if nxt is None:
                nxt = node.children[part] = Group()
            node = nxt
        return node

    def require_group(self, path):
        return self.create_group(path)

    def create_dataset(self, path, data):
        parts = path.strip("/").split("/")
        node = self.create_group("/".join(parts[:-1])) if len(parts) > 1 else self
        ds = node.children[parts[-1]] = Dataset(data)
        return ds

    def __setitem__(self, name, value):
        if isinstance(value, (ExternalLink, SoftLink, _Node)):
            parts = name.strip("/").split("/")
            node = self.create_group("/".join(parts[:-1])) if len(parts) > 1 else self
            node.children[parts[-1]] = value
        else:
            self.create_dataset(name, value)

    def __getitem__(self, path):
        node = self
        for part in path.strip("/").split("/"):
            node = node.children[part]
        return node


# ---------------------------------------------------------------------------------------------
# message encoders
# ---------------------------------------------------------------------------------------------
def _encode_datatype(dt: np.dtype) -> bytes:
    dt = np.dtype(dt)
    if dt.subdtype is not None:                      # array class, version 3
        base, shape = dt.subdtype
        out = struct.pack("<BBBBI", 0x30 | 10, 0, 0, 0, dt.itemsize)
        out += struct.pack("<B", len(shape)) + b"".join(struct.pack("<I", d) for d in shape)
        return out + _encode_datatype(base)
    if dt.names is not None:                         # compound, version 3
        n = len(dt.names)
        out = struct.pack("<BBBBI", 0x30 | 6, n & 0xFF, (n >> 8) & 0xFF, 0, dt.itemsize)
        nb = 1
        while (1 << (8 * nb)) <= dt.itemsize and nb < 8:
            nb += 1
        for name in dt.names:
            sub, off = dt.fields[name][:2]
            out += name.encode("utf-8") + b"\x00" + int(off).to_bytes(nb, "little") + _encode_datatype(sub)
        return out
    if dt.kind in "iu":
        bits0 = 0x08 if dt.kind == "i" else 0x00
        if dt.byteorder == ">":
            bits0 |= 1
        return struct.pack("<BBBBIHH", 0x10 | 0, bits0, 0, 0, dt.itemsize, 0, dt.itemsize * 8)
    if dt.kind == "b":
        return struct.pack("<BBBBIHH", 0x10 | 0, 0x00, 0, 0, 1, 0, 8)
    if dt.kind == "f":
        if dt.itemsize == 8:
            sign, eloc, esize, msize, bias = 63, 52, 11, 52, 1023
        elif dt.itemsize == 4:
            sign, eloc, esize, msize, bias = 31, 23, 8, 23, 127
        else:
            raise TypeError("unsupported float size %d" % dt.itemsize)
        bits0 = 0x20 | (1 if dt.byteorder == ">" else 0)
        return struct.pack("<BBBBIHHBBBBI", 0x10 | 1, bits0, sign, 0, dt.itemsize, 0, dt.itemsize * 8,
                           eloc, esize, 0, msize, bias)
    if dt.kind == "S":
        # null-padded ASCII (what h5py writes for numpy 'S' types)
        return struct.pack("<BBBBI", 0x10 | 3, 0x01, 0, 0, max(dt.itemsize, 1))
    raise TypeError("unsupported dtype %r" % (dt,))


def _encode_dataspace(shape) -> bytes:
    out = struct.pack("<BBBBI", 1, len(shape), 0, 0, 0)
    return out + b"".join(struct.pack("<Q", d) for d in shape)


def _message(mtype: int, data: bytes, flags: int = 0) -> bytes:
    data = _pad8(data)
    return struct.pack("<HHBBBB", mtype, len(data), flags, 0, 0, 0) + data


def _normalise_attr(value):
    if isinstance(value, str):
        value = value.encode("utf-8")
    if isinstance(value, bytes):
        return np.array(value, dtype="S%d" % max(len(value), 1))
    a = np.asarray(value)
    if a.dtype.kind == "U":
        a = np.char.encode(a, "utf-8")
    if a.dtype.kind == "O":
        raise TypeError("object attributes are not supported")
    if a.dtype == np.bool_:
        a = a.astype(np.int8)
    return np.ascontiguousarray(a)


def _attribute(name: str, value) -> bytes:
    a = _normalise_attr(value)
    nm = name.encode("utf-8") + b"\x00"
    dt = _encode_datatype(a.dtype)
    ds = _encode_dataspace(a.shape)
    body = struct.pack("<BBHHH", 1, 0, len(nm), len(dt), len(ds))
    body += _pad8(nm) + _pad8(dt) + _pad8(ds) + a.tobytes()
    if len(body) > 65000:
        raise ValueError("attribute %r is too large for a version-1 object header" % name)
    return _message(0x000C, body)


def _link(name: str, target) -> bytes:
    nm = name.encode("utf-8")
    if len(nm) > 255:
        raise ValueError("link name too long")
    if isinstance(target, int):            # hard link
        return _message(0x0006, struct.pack("<BBB", 1, 0x00, len(nm)) + nm + struct.pack("<Q", target))
    if isinstance(target, SoftLink):
        p = target.path.encode("utf-8")
        return _message(0x0006, struct.pack("<BBBB", 1, 0x08, 1, len(nm)) + nm + struct.pack("<H", len(p)) + p)
    if isinstance(target, ExternalLink):
        info = b"\x00" + target.filename.encode("utf-8") + b"\x00" + target.path.encode("utf-8") + b"\x00"
        return _message(0x0006, struct.pack("<BBBB", 1, 0x08, 64, len(nm)) + nm + struct.pack("<H", len(info)) + info)
    raise TypeError(type(target))


def _object_header(messages) -> bytes:
    body = b"".join(messages)
    return struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + b"\x00" * 4 + body


class File(Group):
    """Build the tree with the Group API, then ``write(filename)``.  Usable as a context
    manager: the file is written on exit."""

    def __init__(self, filename=None):
        super().__init__()
        self.filename = filename

    def __enter__(self):
        return self

    def __exit__(self, exc_type, *a):
        if exc_type is None and self.filename:
            self.write(self.filename)

    def write(self, filename=None):
        filename = filename or self.filename
        chunks = [b"\x00" * 96]           # superblock placeholder
        pos = [96]

        def emit(b: bytes) -> int:
            addr = pos[0]
            b = _pad8(b)
            chunks.append(b)
            pos[0] += len(b)
            return addr

        def write_node(node) -> int:
            msgs = []
            if isinstance(node, Dataset):
                a = node.data
                raw = a.tobytes()
                addr = emit(raw) if len(raw) else UNDEF
                msgs.append(_message(0x0001, _encode_dataspace(a.shape)))
                msgs.append(_message(0x0003, _encode_datatype(a.dtype), flags=1))
                msgs.append(_message(0x0005, struct.pack("<BBBB", 2, 1, 0, 0)))          # fill value: undefined
                msgs.append(_message(0x0008, struct.pack("<BBQQ", 3, 1, addr, len(raw))))  # contiguous layout
            else:
                links = []
                for name, child in node.children.items():
                    if isinstance(child, (ExternalLink, SoftLink)):
                        links.append(_link(name, child))
                    else:
                        links.append(_link(name, write_node(child)))
                msgs.append(_message(0x0002, struct.pack("<BBQQ", 0, 0, UNDEF, UNDEF)))    # link info
                msgs.append(_message(0x000A, struct.pack("<BBHH", 0, 1, 65535, 0)))        # group info
                msgs.extend(links)
            for k, v in node.attrs.items():
                msgs.append(_attribute(k, v))
            return emit(_object_header(msgs))

        root = write_node(self)
        eof = pos[0]
        sb = SIGNATURE + struct.pack("<BBBBBBBB", 0, 0, 0, 0, 0, 8, 8, 0) + struct.pack("<HHI", 4, 16, 0)
        sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
        sb += struct.pack("<QQII", 0, root, 0, 0) + b"\x00" * 16
        assert len(sb) == 96
        chunks[0] = sb
        with open(filename, "wb") as f:
            for c in chunks:
                f.write(c)
        return filename
