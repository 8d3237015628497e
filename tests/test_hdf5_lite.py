"""speechless_b200.hdf5_lite: the pure-Python HDF5 subset that stands in for h5py (Keras weight files,
reference net.py:209-212,572)."""
import struct
from pathlib import Path

import numpy as np
import pytest

from speechless_b200 import hdf5_lite


def genuine_file():
    """The one file in this image that libhdf5 itself wrote: scipy's MATLAB 7.3 fixture (a 512-byte user block,
    then a version-0 superblock, symbol-table groups, version-1 object headers — the format h5py's defaults
    produce for Keras too)."""
    import scipy.io
    path = Path(scipy.io.__file__).parent / "matlab" / "tests" / "data" / "testhdf5_7.4_GLNX86.mat"
    if not path.exists():
        pytest.skip("scipy's HDF5 fixture is not installed")
    return path


def test_reads_a_genuine_libhdf5_file():
    with hdf5_lite.File(genuine_file()) as f:
        assert f.keys() == ["testdouble"]
        dataset = f["testdouble"]
        assert dataset.shape == (9, 1) and dataset.dtype == np.float64
        # scipy's fixtures hold testdouble = 0, pi/4, ..., 2 pi
        np.testing.assert_allclose(np.asarray(dataset)[:, 0], np.arange(9) * np.pi / 4, rtol=1e-15)
        assert dataset.attrs["MATLAB_class"] == b"double"
        assert "testdouble" in f and "nothing" not in f
        with pytest.raises(KeyError):
            f["testdouble/deeper"]


def test_writer_reproduces_the_genuine_files_structures():
    """Byte-level agreement of what the writer emits with what libhdf5 emitted for the same things: the float64
    datatype message, the B-tree / symbol-node / local-heap framing of a one-link group."""
    raw = genuine_file().read_bytes()[512:]
    reader = hdf5_lite._Reader(genuine_file().read_bytes())
    genuine = hdf5_lite._Object(reader.buf, reader.open(reader.root_address, "/")._load()["testdouble"])
    datatype = genuine.find(0x03)[0]
    assert hdf5_lite._datatype_message(np.dtype("<f8")) == datatype[:len(hdf5_lite._datatype_message(np.dtype("<f8")))]
    # genuine symbol node: signature, version 1, one entry = (name offset 8, object header address)
    assert raw[0x4e0:0x4e8] == b"SNOD\x01\x00\x01\x00" and struct.unpack_from("<Q", raw, 0x4e8)[0] == 8
    tree = raw[0x180:0x180 + 48]
    assert tree[:8] == b"TREE\x00\x00\x01\x00" and struct.unpack_from("<QQQ", tree, 24) == (0, 0x4e0, 8)
    # ... and the same three structures as the writer lays them out for a one-link group
    import io
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        path = Path(tmp) / "one.h5"
        hdf5_lite.write(path, {"testdouble": np.arange(9.0).reshape(9, 1)})
        mine = path.read_bytes()
    root_tree, root_heap = struct.unpack_from("<QQ", mine, 80)
    assert mine[root_tree:root_tree + 8] == b"TREE\x00\x00\x01\x00"
    key0, child, key1 = struct.unpack_from("<QQQ", mine, root_tree + 24)
    assert (key0, key1) == (0, 8) and mine[child:child + 8] == b"SNOD\x01\x00\x01\x00"
    assert struct.unpack_from("<Q", mine, child + 8)[0] == 8
    assert mine[root_heap:root_heap + 8] == b"HEAP\x00\x00\x00\x00"
    segment = struct.unpack_from("<Q", mine, root_heap + 24)[0]
    assert mine[segment:segment + 24] == raw[0x80:0x80 + 24]  # 8 zero bytes, then "testdouble" padded to 16


def test_round_trip_of_groups_datasets_and_attributes(tmp_path):
    rng = np.random.default_rng(0)
    tree = {
        "layer_a": {"layer_a/kernel:0": rng.standard_normal((3, 5, 7)).astype(np.float32),
                    "layer_a/bias:0": rng.standard_normal(7).astype(np.float32)},
        "numbers": {"i32": np.arange(-3, 9, dtype=np.int32).reshape(3, 4), "u8": np.arange(5, dtype=np.uint8),
                    "f64": rng.standard_normal((2, 2)), "f16": rng.standard_normal(6).astype(np.float16),
                    "big_endian": np.arange(4, dtype=">i4"), "empty": np.zeros((0, 3), dtype=np.float32),
                    "scalar": np.float32(2.5), "names": np.array([b"ab", b"c"])},
    }
    attrs = {"/": {"layer_names": [b"layer_a", b"numbers"], "backend": b"tensorflow", "version": np.int64(3)},
             "layer_a": {"weight_names": [b"layer_a/kernel:0", b"layer_a/bias:0"]},
             "numbers/i32": {"scale": np.float64(0.5), "text": "uniçode"}}
    path = tmp_path / "file.h5"
    hdf5_lite.write(path, tree, attrs)
    with hdf5_lite.File(path) as f:
        assert f.keys() == ["layer_a", "numbers"]
        assert f.attrs["layer_names"].tolist() == [b"layer_a", b"numbers"] and f.attrs["backend"] == b"tensorflow"
        assert f.attrs["version"] == 3
        group = f["layer_a"]
        names = [n.decode() for n in group.attrs["weight_names"]]
        assert names == ["layer_a/kernel:0", "layer_a/bias:0"]
        assert group.keys() == ["layer_a"]  # h5py semantics: a "/" in a dataset name makes an intermediate group
        np.testing.assert_array_equal(np.asarray(group[names[0]]), tree["layer_a"]["layer_a/kernel:0"])
        np.testing.assert_array_equal(np.asarray(f["layer_a/layer_a/bias:0"]), tree["layer_a"]["layer_a/bias:0"])
        for name, want in tree["numbers"].items():
            got = np.asarray(f["numbers"][name])
            assert got.shape == np.shape(want) and got.dtype.kind == np.asarray(want).dtype.kind, name
            np.testing.assert_array_equal(got, np.asarray(want), err_msg=name)
        assert f["numbers/i32"].attrs["scale"] == 0.5
        assert f["numbers/i32"].attrs["text"] == "uniçode".encode("utf8")


def test_groups_with_more_links_than_one_symbol_node_holds(tmp_path):
    tree = {"layer_{:03d}".format(i): {"w": np.full((2,), i, dtype=np.float32)} for i in range(150)}
    path = tmp_path / "many.h5"
    hdf5_lite.write(path, tree)
    with hdf5_lite.File(path) as f:
        assert f.keys() == sorted(tree)
        for i in (0, 63, 64, 127, 128, 149):
            assert np.asarray(f["layer_{:03d}/w".format(i)]).tolist() == [i, i]


def test_rejects_what_is_not_hdf5(tmp_path):
    path = tmp_path / "not.h5"
    path.write_bytes(b"PK\x03\x04" + b"\x00" * 2000)
    with pytest.raises(hdf5_lite.Hdf5FormatError):
        hdf5_lite.File(path)


def test_reads_object_headers_with_continuation_blocks(tmp_path):
    """libhdf5 moves messages into continuation blocks when attributes are added after an object was created —
    which is how Keras writes `layer_names` / `weight_names`.  The writer never needs them, so the file is
    assembled here: every object header keeps its first message and a continuation message, the rest of the
    messages live in a block elsewhere in the file."""
    class Splitting(hdf5_lite._Writer):
        def _object_header(self, messages):
            if len(messages) < 3:
                return super()._object_header(messages)
            rest = b"".join(messages[1:])
            block = self._append(rest)
            continuation = hdf5_lite._message(0x10, struct.pack("<QQ", block, len(rest)))
            first = messages[0] + continuation
            return self._append(struct.pack("<BBHII", 1, 0, len(messages) + 1, 1, len(first)) + b"\x00" * 4 + first)

    writer = Splitting()
    kernel = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    dataset = writer.dataset(kernel, {"note": b"kept in the continuation block"})
    group = writer.group({"kernel:0": dataset}, {"weight_names": [b"conv/kernel:0"], "extra": np.int32(7)})
    root = writer.group({"conv": group}, {"layer_names": [b"conv"], "backend": b"tensorflow"})
    path = tmp_path / "continued.h5"
    path.write_bytes(writer.finish(root))
    with hdf5_lite.File(path) as f:
        assert f.attrs["layer_names"].tolist() == [b"conv"] and f.attrs["backend"] == b"tensorflow"
        assert f["conv"].attrs["weight_names"].tolist() == [b"conv/kernel:0"] and f["conv"].attrs["extra"] == 7
        np.testing.assert_array_equal(np.asarray(f["conv/kernel:0"]), kernel)
        assert f["conv/kernel:0"].attrs["note"] == b"kept in the continuation block"


def test_describe_lists_groups_datasets_and_attributes(tmp_path):
    path = tmp_path / "w.h5"
    hdf5_lite.write(path, {"conv": {"conv/kernel:0": np.zeros((7, 3, 5), np.float32)}},
                    {"/": {"layer_names": [b"conv"]}, "conv": {"weight_names": [b"conv/kernel:0"]}})
    lines = hdf5_lite.describe(path)
    assert lines[0] == "@layer_names = [b'conv']" and "conv/" in lines
    assert any(line.strip() == "kernel:0  float32 (7, 3, 5)" for line in lines)
