"""A small pure-Python HDF5 reader / writer for Keras weight files (no h5py, no libhdf5).

The reference saves and loads its models with Keras' `save_weights` / `load_weights`
(reference net.py:209-212,558-560,572), i.e. as HDF5 files written by h5py with its default (oldest)
file-format settings: version-0 superblock, groups as symbol tables (B-tree v1 + SNOD + local heap),
version-1 object headers, contiguous float32 datasets, string-array attributes `layer_names` /
`weight_names`.  Neither h5py nor libhdf5 is in this image, so this module restates exactly that subset
of the published HDF5 File Format Specification (v1.1/2.0):

reader  superblock v0/v1 (and v2/v3), user blocks, object headers v1 (+ continuation blocks) and v2
        ("OHDR"), old-style groups and compact new-style groups (link messages), dataspace v1/v2,
        datatypes fixed / float / fixed-length string / variable-length string (global heap) / enum /
        reference, data layout v1-v3 compact / contiguous / chunked (B-tree v1) with the deflate,
        shuffle and fletcher32 filters, attribute messages v1-v3.
        Validated against the one genuine libhdf5-written file in the image (scipy's MATLAB 7.3 test
        fixture, tests/test_hdf5_lite.py).
writer  the subset h5py's defaults produce for `model.save_weights`: nested groups, contiguous
        little-endian numeric datasets, attributes holding fixed-length byte strings or numeric arrays.
        Round-trips through the reader; it could not be opened with libhdf5 here, which DESIGN.md says.
"""
import struct
import zlib
from typing import Dict, Iterator, List, Optional, Tuple, Union

import numpy

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEFINED = 0xFFFFFFFFFFFFFFFF


class Hdf5FormatError(IOError):
    pass


# ---------------------------------------------------------------------------------------------------------
# reader
# ---------------------------------------------------------------------------------------------------------
class _Buffer:
    def __init__(self, data: bytes, base: int):
        self.data, self.base = data, base

    def at(self, address: int, size: int) -> bytes:
        start = self.base + address
        if address == UNDEFINED or start < 0 or start + size > len(self.data):
            raise Hdf5FormatError("address {:#x} (+{}) outside the file".format(address, size))
        return self.data[start:start + size]

    def u(self, address: int, size: int) -> int:
        return int.from_bytes(self.at(address, size), "little")


def _pad8(n: int) -> int:
    return (n + 7) & ~7


def _parse_datatype(raw: bytes) -> Tuple[object, int]:
    """-> (numpy dtype | ("vlen_str",) | ("vlen", base dtype), size in bytes of one element)"""
    cls, version = raw[0] & 0x0F, raw[0] >> 4
    bits = raw[1] | (raw[2] << 8) | (raw[3] << 16)
    size = struct.unpack_from("<I", raw, 4)[0]
    order = ">" if bits & 1 else "<"
    if cls == 0:
        return numpy.dtype("{}{}{}".format(order, "i" if bits & 8 else "u", size)), size
    if cls == 1:
        return numpy.dtype("{}f{}".format(order, size)), size
    if cls == 3:
        return numpy.dtype("S{}".format(size)), size
    if cls == 9:
        if bits & 0x0F == 1:
            return ("vlen_str",), size
        base, _ = _parse_datatype(raw[8:])
        return ("vlen", base), size
    if cls == 7:  # object / region reference: an address
        return numpy.dtype("<u{}".format(size)) if size in (1, 2, 4, 8) else numpy.dtype("V{}".format(size)), size
    if cls == 8:  # enum (h5py booleans): the base integer type
        base, _ = _parse_datatype(raw[8:])
        return base, size
    if cls in (2, 4, 5, 6, 10):  # time, bitfield, opaque, compound, array: raw bytes
        return numpy.dtype("V{}".format(size)), size
    raise Hdf5FormatError("datatype class {} (version {}) is not supported".format(cls, version))


def _parse_dataspace(raw: bytes) -> Optional[Tuple[int, ...]]:
    version, rank, flags = raw[0], raw[1], raw[2]
    if version == 1:
        start = 8
    elif version == 2:
        if raw[3] == 2:
            return None  # null dataspace
        start = 4
    else:
        raise Hdf5FormatError("dataspace version {}".format(version))
    return tuple(struct.unpack_from("<Q", raw, start + 8 * i)[0] for i in range(rank))


class _Object:
    """An object header, parsed into its messages."""

    def __init__(self, buf: _Buffer, address: int):
        self.buf, self.address = buf, address
        self.messages: List[Tuple[int, bytes, int]] = []  # (type, data, flags)
        head = buf.at(address, 4)
        if head == b"OHDR":
            self._parse_v2(address)
        elif head[0] == 1:
            self._parse_v1(address)
        else:
            raise Hdf5FormatError("no object header at {:#x}".format(address))

    def _parse_v1(self, address: int) -> None:
        count = self.buf.u(address + 2, 2)
        size = self.buf.u(address + 8, 4)
        blocks = [(address + 16, size)]
        while blocks and len(self.messages) < count:
            start, length = blocks.pop(0)
            position = start
            while position + 8 <= start + length and len(self.messages) < count:
                kind, msize = self.buf.u(position, 2), self.buf.u(position + 2, 2)
                flags = self.buf.u(position + 4, 1)
                data = self.buf.at(position + 8, msize)
                position += 8 + msize
                if kind == 0x10:
                    blocks.append(struct.unpack_from("<QQ", data))
                self.messages.append((kind, data, flags))

    def _parse_v2(self, address: int) -> None:
        flags = self.buf.u(address + 5, 1)
        position = address + 6
        if flags & 0x20:
            position += 16  # times
        if flags & 0x10:
            position += 4  # attribute storage phase change values
        width = 1 << (flags & 3)
        chunk0 = self.buf.u(position, width)
        position += width
        tracked = bool(flags & 0x04)
        blocks = [(position, chunk0)]
        while blocks:
            start, length = blocks.pop(0)
            position = start
            while position + 4 + (2 if tracked else 0) <= start + length:
                kind, msize, mflags = self.buf.u(position, 1), self.buf.u(position + 1, 2), self.buf.u(position + 3, 1)
                position += 4 + (2 if tracked else 0)
                data = self.buf.at(position, msize)
                position += msize
                if kind == 0x10:
                    offset, clen = struct.unpack_from("<QQ", data)
                    blocks.append((offset + 4, clen - 8))  # skip "OCHK", drop the checksum
                self.messages.append((kind, data, mflags))

    def find(self, kind: int) -> List[bytes]:
        return [data for k, data, _ in self.messages if k == kind]

    # ---- attributes
    def attributes(self, reader: "_Reader") -> Dict[str, object]:
        out: Dict[str, object] = {}
        for raw in self.find(0x0C):
            version = raw[0]
            name_size, type_size, space_size = struct.unpack_from("<HHH", raw, 2)
            position = 8
            if version == 3:
                position += 1
            pad = _pad8 if version == 1 else (lambda n: n)
            name = raw[position:position + name_size].split(b"\x00")[0].decode("utf8")
            position += pad(name_size)
            dtype, element = _parse_datatype(raw[position:position + type_size])
            position += pad(type_size)
            shape = _parse_dataspace(raw[position:position + space_size])
            position += pad(space_size)
            out[name] = reader.decode(raw[position:], dtype, element, shape)
        if self.find(0x15) and any(struct.unpack_from("<Q", d, 2 + (2 if d[1] & 1 else 0))[0] != UNDEFINED
                                   for d in self.find(0x15)):
            raise Hdf5FormatError("densely stored attributes (fractal heap) are not supported")
        return out


class Dataset:
    def __init__(self, reader: "_Reader", obj: _Object, name: str):
        self._reader, self._obj, self.name = reader, obj, name
        self._dtype, self._element = _parse_datatype(obj.find(0x03)[0])
        self.shape = _parse_dataspace(obj.find(0x01)[0])
        self.attrs = obj.attributes(reader)

    @property
    def dtype(self):
        return self._dtype if isinstance(self._dtype, numpy.dtype) else numpy.dtype(object)

    def read(self) -> numpy.ndarray:
        raw = self._obj.find(0x08)[0]
        count = int(numpy.prod(self.shape)) if self.shape else (0 if self.shape is None else 1)
        total = count * self._element
        buf = self._reader.buf
        version = raw[0]
        if version == 3:
            cls = raw[1]
            if cls == 0:
                size = struct.unpack_from("<H", raw, 2)[0]
                data = raw[4:4 + size]
            elif cls == 1:
                address, size = struct.unpack_from("<QQ", raw, 2)
                data = b"\x00" * total if address == UNDEFINED else buf.at(address, total)
            elif cls == 2:
                rank = raw[2]
                address = struct.unpack_from("<Q", raw, 3)[0]
                chunk = struct.unpack_from("<{}I".format(rank), raw, 11)
                data = self._read_chunked(address, chunk[:-1], total)
            else:
                raise Hdf5FormatError("data layout class {}".format(cls))
        elif version in (1, 2):
            rank, cls = raw[1], raw[2]
            position = 8
            address = UNDEFINED
            if cls != 0:
                address = struct.unpack_from("<Q", raw, position)[0]
                position += 8
            dims = struct.unpack_from("<{}I".format(rank), raw, position)
            position += 4 * rank
            if cls == 0:
                size = struct.unpack_from("<I", raw, position)[0]
                data = raw[position + 4:position + 4 + size]
            elif cls == 1:
                data = b"\x00" * total if address == UNDEFINED else buf.at(address, total)
            else:
                data = self._read_chunked(address, dims[:-1] if len(dims) > len(self.shape) else dims, total)
        else:
            raise Hdf5FormatError("data layout version {}".format(version))
        return self._reader.decode(data, self._dtype, self._element, self.shape)

    def __array__(self, dtype=None, copy=None):
        array = numpy.asarray(self.read())
        return array if dtype is None else array.astype(dtype)

    def _filters(self) -> List[Tuple[int, Tuple[int, ...]]]:
        found = self._obj.find(0x0B)
        if not found:
            return []
        raw = found[0]
        version, count = raw[0], raw[1]
        position = 8 if version == 1 else 2
        filters = []
        for _ in range(count):
            ident = struct.unpack_from("<H", raw, position)[0]
            position += 2
            name_length = 0
            if version == 1 or ident >= 256:
                name_length = struct.unpack_from("<H", raw, position)[0]
                position += 2
            _, values = struct.unpack_from("<HH", raw, position)
            position += 4
            position += _pad8(name_length) if version == 1 else name_length
            client = struct.unpack_from("<{}I".format(values), raw, position)
            position += 4 * values
            if version == 1 and values % 2:
                position += 4
            filters.append((ident, client))
        return filters

    def _read_chunked(self, address: int, chunk: Tuple[int, ...], total: int) -> bytes:
        if address == UNDEFINED:
            return b"\x00" * total
        buf, rank, element = self._reader.buf, len(self.shape), self._element
        out = numpy.zeros(self.shape, dtype="V{}".format(element))
        filters = self._filters()

        def walk(node: int) -> Iterator[Tuple[Tuple[int, ...], int, int, int]]:
            if buf.at(node, 4) != b"TREE" or buf.u(node + 4, 1) != 1:
                raise Hdf5FormatError("no chunk B-tree node at {:#x}".format(node))
            level, used = buf.u(node + 5, 1), buf.u(node + 6, 2)
            key = 8 + 8 * (rank + 1)
            position = node + 24
            for _ in range(used):
                size, mask = buf.u(position, 4), buf.u(position + 4, 4)
                offsets = struct.unpack_from("<{}Q".format(rank), buf.at(position + 8, 8 * rank))
                child = buf.u(position + key, 8)
                position += key + 8
                if level:
                    yield from walk(child)
                else:
                    yield offsets, size, mask, child

        for offsets, size, mask, where in walk(address):
            data = buf.at(where, size)
            for index, (ident, client) in reversed(list(enumerate(filters))):
                if mask & (1 << index):
                    continue
                if ident == 1:
                    data = zlib.decompress(data)
                elif ident == 2:
                    width = client[0] if client else element
                    n = len(data) // width
                    data = numpy.frombuffer(data[:n * width], dtype=numpy.uint8).reshape(width, n).T.tobytes() + data[n * width:]
                elif ident == 3:
                    data = data[:-4]
                else:
                    raise Hdf5FormatError("filter {} is not supported".format(ident))
            block = numpy.frombuffer(data, dtype="V{}".format(element), count=int(numpy.prod(chunk))).reshape(chunk)
            window = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offsets, chunk, self.shape))
            out[window] = block[tuple(slice(0, w.stop - w.start) for w in window)]
        return out.tobytes()


class Group:
    def __init__(self, reader: "_Reader", obj: _Object, name: str):
        self._reader, self._obj, self.name = reader, obj, name
        self.attrs = obj.attributes(reader)
        self._links: Optional[Dict[str, int]] = None

    def _load(self) -> Dict[str, int]:
        if self._links is not None:
            return self._links
        buf, links = self._reader.buf, {}
        for raw in self._obj.find(0x11):  # symbol table: B-tree v1 of symbol nodes + local heap of names
            tree, heap = struct.unpack_from("<QQ", raw)
            if buf.at(heap, 4) != b"HEAP":
                raise Hdf5FormatError("no local heap at {:#x}".format(heap))
            segment = buf.u(heap + 24, 8)

            def name_at(offset: int) -> str:
                chunk = b""
                position = segment + offset
                while b"\x00" not in chunk:
                    chunk += buf.at(position, 1)
                    position += 1
                return chunk[:-1].decode("utf8")

            def walk(node: int) -> None:
                signature = buf.at(node, 4)
                if signature == b"TREE":
                    used = buf.u(node + 6, 2)
                    for i in range(used):
                        walk(buf.u(node + 24 + 8 + 16 * i, 8))
                elif signature == b"SNOD":
                    for i in range(buf.u(node + 6, 2)):
                        entry = node + 8 + 40 * i
                        links[name_at(buf.u(entry, 8))] = buf.u(entry + 8, 8)
                else:
                    raise Hdf5FormatError("no group node at {:#x}".format(node))
            walk(tree)
        for raw in self._obj.find(0x06):  # link messages of a compact new-style group
            flags = raw[1]
            position = 2
            kind = 0
            if flags & 0x08:
                kind = raw[position]
                position += 1
            if flags & 0x04:
                position += 8
            if flags & 0x10:
                position += 1
            width = 1 << (flags & 3)
            length = int.from_bytes(raw[position:position + width], "little")
            position += width
            name = raw[position:position + length].decode("utf8")
            position += length
            if kind == 0:
                links[name] = struct.unpack_from("<Q", raw, position)[0]
        for raw in self._obj.find(0x02):  # link info: dense storage?
            flags = raw[1]
            position = 2 + (8 if flags & 1 else 0)
            if struct.unpack_from("<Q", raw, position)[0] != UNDEFINED:
                raise Hdf5FormatError("densely stored links (fractal heap) are not supported")
        self._links = links
        return links

    def keys(self) -> List[str]:
        return sorted(self._load())

    def __contains__(self, path: str) -> bool:
        try:
            self[path]
            return True
        except KeyError:
            return False

    def __iter__(self):
        return iter(self.keys())

    def __getitem__(self, path: str) -> Union["Group", Dataset]:
        node: Union[Group, Dataset] = self
        for part in [p for p in path.split("/") if p]:
            if not isinstance(node, Group) or part not in node._load():
                raise KeyError(path)
            node = self._reader.open(node._load()[part], (node.name.rstrip("/") + "/" + part))
        return node


class _Reader:
    def __init__(self, data: bytes):
        offset = 0
        while data[offset:offset + 8] != SIGNATURE:  # a user block of 512, 1024, ... bytes may come first
            offset = 512 if offset == 0 else offset * 2
            if offset + 8 > len(data):
                raise Hdf5FormatError("not an HDF5 file")
        version = data[offset + 8]
        if version in (0, 1):
            if data[offset + 13] != 8 or data[offset + 14] != 8:
                raise Hdf5FormatError("only 8-byte offsets and lengths are supported")
            position = offset + 24 + (4 if version == 1 else 0)
            base = struct.unpack_from("<Q", data, position)[0]
            self.buf = _Buffer(data, base)
            entry = position + 32
            self.root_address = struct.unpack_from("<Q", data, entry + 8)[0]
        elif version in (2, 3):
            if data[offset + 9] != 8 or data[offset + 10] != 8:
                raise Hdf5FormatError("only 8-byte offsets and lengths are supported")
            base = struct.unpack_from("<Q", data, offset + 12)[0]
            self.buf = _Buffer(data, base)
            self.root_address = struct.unpack_from("<Q", data, offset + 36)[0]
        else:
            raise Hdf5FormatError("superblock version {}".format(version))
        self._heaps: Dict[int, Dict[int, bytes]] = {}

    def open(self, address: int, name: str) -> Union[Group, Dataset]:
        obj = _Object(self.buf, address)
        return Dataset(self, obj, name) if obj.find(0x08) else Group(self, obj, name)

    def global_heap_object(self, address: int, index: int) -> bytes:
        if address not in self._heaps:
            if self.buf.at(address, 4) != b"GCOL":
                raise Hdf5FormatError("no global heap collection at {:#x}".format(address))
            size = self.buf.u(address + 8, 8)
            objects, position = {}, address + 16
            while position + 16 <= address + size:
                ident, length = self.buf.u(position, 2), self.buf.u(position + 8, 8)
                if ident == 0:
                    break
                objects[ident] = self.buf.at(position + 16, length)
                position += 16 + _pad8(length)
            self._heaps[address] = objects
        return self._heaps[address][index]

    def decode(self, data: bytes, dtype, element: int, shape: Optional[Tuple[int, ...]]):
        if shape is None:
            return None
        count = int(numpy.prod(shape)) if shape else 1
        if isinstance(dtype, numpy.dtype):
            array = numpy.frombuffer(data, dtype=dtype, count=count).reshape(shape)
            if dtype.kind == "S":
                array = numpy.array([s.split(b"\x00")[0] for s in array.reshape(-1)], dtype=dtype).reshape(shape)
            return array[()] if shape == () else array.copy()
        items = []
        for i in range(count):
            length, address, index = struct.unpack_from("<IQI", data, 16 * i)
            raw = self.global_heap_object(address, index) if address not in (0, UNDEFINED) or index else b""
            if dtype[0] == "vlen_str":
                items.append(raw[:length].decode("utf8", "replace") if length or raw else "")
            else:
                items.append(numpy.frombuffer(raw, dtype=dtype[1], count=length).copy())
        if shape == ():
            return items[0]
        out = numpy.empty(count, dtype=object)
        out[:] = items
        return out.reshape(shape)


class File(Group):
    """`hdf5_lite.File(path)`: read-only, the whole file is held in memory (weight files are a few hundred MB
    at most).  Indexing follows h5py: `f["group/dataset"]`, `.attrs`, `.keys()`, `numpy.asarray(dataset)`."""

    def __init__(self, path):
        with open(str(path), "rb") as handle:
            reader = _Reader(handle.read())
        super().__init__(reader, _Object(reader.buf, reader.root_address), "/")

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


# ---------------------------------------------------------------------------------------------------------
# writer (the subset h5py's defaults produce for Keras weight files)
# ---------------------------------------------------------------------------------------------------------
_LEAF_K = 32       # symbol-table leaf nodes hold up to 2 * 32 entries (a superblock field; libhdf5's default is 4)
_INTERNAL_K = 16   # B-tree v1 nodes hold up to 2 * 16 children


def _datatype_message(dtype: numpy.dtype) -> bytes:
    if dtype.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x00, 0, 0, max(dtype.itemsize, 1))  # class 3, null-terminated ASCII
    little = dtype.newbyteorder("<")
    size = little.itemsize
    if dtype.kind in "iu":
        bits0 = 0x08 if dtype.kind == "i" else 0x00
        return struct.pack("<BBBBIHH", 0x10, bits0, 0, 0, size, 0, 8 * size)
    if dtype.kind == "f" and size in (2, 4, 8):
        exponent, mantissa, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[size]
        # bit fields: little-endian, implied-1 mantissa normalisation (bits 4-5 = 2), sign at the top bit
        return struct.pack("<BBBBIHHBBBBI", 0x11, 0x20, 8 * size - 1, 0, size, 0, 8 * size, mantissa, exponent, 0,
                           mantissa, bias)
    raise ValueError("cannot store dtype {}".format(dtype))


def _dataspace_message(shape: Tuple[int, ...]) -> bytes:
    # version 1, max dims present (h5py always writes them)
    return struct.pack("<BBBBI", 1, len(shape), 1, 0, 0) + b"".join(struct.pack("<Q", n) for n in shape) * 2


def _message(kind: int, data: bytes, flags: int = 0) -> bytes:
    data = data + b"\x00" * (_pad8(len(data)) - len(data))
    return struct.pack("<HHBBBB", kind, len(data), flags, 0, 0, 0) + data


def _attribute_message(name: str, value) -> bytes:
    array = numpy.asarray(value)
    if array.dtype.kind == "U":
        array = numpy.char.encode(array, "utf8")
    if array.dtype.kind == "O":
        array = numpy.asarray([v if isinstance(v, bytes) else str(v).encode("utf8") for v in array.reshape(-1)]
                              ).reshape(array.shape)
    if array.dtype.kind not in "Siuf":
        raise ValueError("cannot store attribute {!r} of dtype {}".format(name, array.dtype))
    if array.dtype.kind != "S":
        array = array.astype(array.dtype.newbyteorder("<"))
    encoded = name.encode("utf8") + b"\x00"
    datatype, dataspace = _datatype_message(array.dtype), _dataspace_message(array.shape)
    body = struct.pack("<BBHHH", 1, 0, len(encoded), len(datatype), len(dataspace))
    for part in (encoded, datatype, dataspace):
        body += part + b"\x00" * (_pad8(len(part)) - len(part))
    return _message(0x0C, body + array.tobytes())


class _Writer:
    def __init__(self):
        self.out = bytearray(96)  # the superblock is filled in at the end

    def _append(self, blob: bytes) -> int:
        while len(self.out) % 8:
            self.out.append(0)
        address = len(self.out)
        self.out += blob
        return address

    def _object_header(self, messages: List[bytes]) -> int:
        body = b"".join(messages)
        return self._append(struct.pack("<BBHII", 1, 0, len(messages), 1, len(body)) + b"\x00" * 4 + body)

    def dataset(self, array: numpy.ndarray, attrs: Dict[str, object]) -> int:
        array = numpy.asarray(array, order="C")  # (ascontiguousarray would turn a scalar into shape (1,))
        if array.dtype.kind != "S":
            array = array.astype(array.dtype.newbyteorder("<"))
        data = self._append(array.tobytes()) if array.size else UNDEFINED
        layout = struct.pack("<BBQQ", 3, 1, data, array.nbytes)
        fill = struct.pack("<BBBB", 2, 2, 0, 0)  # version 2, late allocation, write at allocation, undefined value
        messages = [_message(0x01, _dataspace_message(array.shape)), _message(0x03, _datatype_message(array.dtype), 1),
                    _message(0x05, fill, 1), _message(0x08, layout)]
        messages += [_attribute_message(k, v) for k, v in attrs.items()]
        return self._object_header(messages)

    def group(self, links: Dict[str, int], attrs: Dict[str, object]) -> int:
        names = sorted(links, key=lambda n: n.encode("utf8"))
        if len(names) > 2 * _LEAF_K * 2 * _INTERNAL_K:
            raise ValueError("too many links in one group")
        # local heap: offset 0 holds the empty string every B-tree's first key points at
        heap_data = bytearray(b"\x00" * 8)
        offsets = {}
        for name in names:
            offsets[name] = len(heap_data)
            encoded = name.encode("utf8") + b"\x00"
            heap_data += encoded + b"\x00" * (_pad8(len(encoded)) - len(encoded))
        free = len(heap_data)
        heap_data += struct.pack("<QQ", 1, 16)  # one free block: next = 1 (last), its size
        segment = self._append(bytes(heap_data))
        heap = self._append(b"HEAP" + struct.pack("<BBBBQQQ", 0, 0, 0, 0, len(heap_data), free, segment))
        # symbol nodes (full size), then one B-tree node above them
        children, keys = [], [0]
        for start in range(0, len(names), 2 * _LEAF_K):
            part = names[start:start + 2 * _LEAF_K]
            node = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
            for name in part:
                node += struct.pack("<QQII", offsets[name], links[name], 0, 0) + b"\x00" * 16
            node += b"\x00" * (8 + 2 * _LEAF_K * 40 - len(node))
            children.append(self._append(node))
            keys.append(offsets[part[-1]])
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(children), UNDEFINED, UNDEFINED)
        for i, child in enumerate(children):
            tree += struct.pack("<QQ", keys[i], child)
        tree += struct.pack("<Q", keys[-1])
        tree += b"\x00" * (24 + (2 * _INTERNAL_K + 1) * 8 + 2 * _INTERNAL_K * 8 - len(tree))
        tree_address = self._append(tree)
        messages = [_message(0x11, struct.pack("<QQ", tree_address, heap))]
        messages += [_attribute_message(k, v) for k, v in attrs.items()]
        self._tables = getattr(self, "_tables", {})
        address = self._object_header(messages)
        self._tables[address] = (tree_address, heap)
        return address

    def finish(self, root: int) -> bytes:
        tree, heap = self._tables[root]
        while len(self.out) % 8:
            self.out.append(0)
        superblock = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _LEAF_K, _INTERNAL_K, 0)
        superblock += struct.pack("<QQQQ", 0, UNDEFINED, len(self.out), UNDEFINED)
        superblock += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", tree, heap)
        self.out[:96] = superblock
        return bytes(self.out)


def write(path, tree: Dict[str, object], attrs: Optional[Dict[str, Dict[str, object]]] = None) -> None:
    """tree: nested dicts of numpy arrays (a dict = a group); names may contain "/" as in h5py (intermediate
    groups are created).  attrs: {"/" or "group/path" or "group/dataset": {attribute name: value}}."""
    attrs = {k.strip("/"): v for k, v in (attrs or {}).items()}
    writer = _Writer()

    def expand(node: Dict[str, object]) -> Dict[str, object]:
        out: Dict[str, object] = {}
        for name, value in node.items():
            parts = [p for p in name.split("/") if p]
            target = out
            for part in parts[:-1]:
                target = target.setdefault(part, {})
                if not isinstance(target, dict):
                    raise ValueError("{} is both a dataset and a group".format(part))
            target[parts[-1]] = expand(value) if isinstance(value, dict) else value
        return out

    def emit(node: Dict[str, object], prefix: str) -> int:
        links = {}
        for name, value in node.items():
            where = (prefix + "/" + name).strip("/")
            links[name] = emit(value, where) if isinstance(value, dict) else \
                writer.dataset(numpy.asarray(value), attrs.get(where, {}))
        return writer.group(links, attrs.get(prefix.strip("/"), {}))

    root = emit(expand(tree), "")
    with open(str(path), "wb") as handle:
        handle.write(writer.finish(root))


def describe(path) -> List[str]:
    """One line per object of the file: `python -m speechless_b200.hdf5_lite weights-epoch12.h5`."""
    lines: List[str] = []

    def walk(group: Group, indent: int) -> None:
        for name, value in sorted(group.attrs.items()):
            lines.append("{}@{} = {!r}".format("  " * indent, name, value.tolist() if hasattr(value, "tolist") else value))
        for key in group.keys():
            node = group[key]
            if isinstance(node, Group):
                lines.append("{}{}/".format("  " * indent, key))
                walk(node, indent + 1)
            else:
                lines.append("{}{}  {} {}".format("  " * indent, key, node.dtype, node.shape))
    with File(path) as f:
        walk(f, 0)
    return lines


if __name__ == "__main__":
    import sys
    print("\n".join(describe(sys.argv[1])))
