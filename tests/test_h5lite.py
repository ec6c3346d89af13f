"""HDF5 / Keras weight files without h5py (simplesr_b200.h5lite, simplesr_b200.keras_h5) - host logic, no GPU.

The reader is pinned to a file the HDF5 library itself wrote: scipy ships ``testhdf5_7.4_GLNX86.mat`` (MATLAB v7.3 =
HDF5 behind a 512-byte user block; superblock 0, symbol-table groups, v1 object headers - the same "earliest" format
h5py uses for Keras files).  Its content is known: MATLAB's ``testdouble = 0:pi/4:2*pi`` (scipy/io/matlab/tests).
The writer is checked by reading its files back."""
import json
import math
import os

import numpy as np
import pytest

from simplesr_b200 import h5lite as H
from simplesr_b200 import keras_h5 as K


def _scipy_hdf5_file():
    try:
        import scipy.io.matlab
    except ImportError:
        return None
    p = os.path.join(os.path.dirname(scipy.io.matlab.__file__), "tests", "data", "testhdf5_7.4_GLNX86.mat")
    return p if os.path.isfile(p) else None


@pytest.mark.skipif(_scipy_hdf5_file() is None, reason="scipy's MATLAB v7.3 test file is not installed")
def test_reader_on_a_file_written_by_libhdf5():
    with H.File(_scipy_hdf5_file()) as f:
        assert (f.base, f.O, f.L, f.leaf_k, f.internal_k) == (512, 8, 8, 4, 16)
        assert f.keys() == ["testdouble"]
        d = f["testdouble"]
        assert d.shape == (9, 1) and d.dtype == np.dtype("<f8")
        np.testing.assert_array_equal(d[...].ravel(), np.arange(9) * (math.pi / 4))
        assert d.attrs["MATLAB_class"] == b"double"
        with pytest.raises(KeyError):
            f["missing"]


def test_not_hdf5(tmp_path):
    p = tmp_path / "x.h5"
    p.write_bytes(b"PK\x03\x04" + b"\0" * 100)
    with pytest.raises(H.H5Error):
        H.File(str(p))


def test_round_trip_types_groups_attributes(tmp_path):
    rng = np.random.default_rng(0)
    w = H.Writer()
    arrays = {"f32": rng.standard_normal((3, 3, 5, 7)).astype(np.float32), "f64": rng.standard_normal(11),
              "i32": rng.integers(-9, 9, size=(4, 2)).astype(np.int32), "u8": np.arange(256, dtype=np.uint8),
              "i64": np.array([2 ** 40, -3], np.int64), "f16": np.array([0.5, -2.0], np.float16),
              "be": np.arange(5, dtype=">f4"), "scalar": np.float32(2.5), "empty": np.zeros((0,), np.float32),
              "names": np.array([b"ab", b"c", b"defg"])}
    g = w.root.create_group("a/b")
    for k, v in arrays.items():
        g.create_dataset(k, v)
    w.root.create_dataset("top/left:0", arrays["f32"])           # intermediate groups are created, ':' is legal
    w.root.attrs["title"] = b"weights"
    w.root.attrs["version"] = "2.3.0-tf"
    g.attrs["list"] = np.array([b"x:0", b"yy:0"])
    g.attrs["pi"] = np.float64(math.pi)
    g.attrs["ints"] = np.arange(3, dtype=np.int32)
    g.attrs["none"] = np.zeros((0,), "S1")
    path = str(tmp_path / "t.h5")
    w.save(path)
    with H.File(path) as f:
        assert sorted(f.keys()) == ["a", "top"]
        assert f.attrs["title"] == b"weights" and f.attrs["version"] == b"2.3.0-tf"
        b = f["a/b"]
        assert sorted(b.keys()) == sorted(arrays)
        for k, v in arrays.items():
            got = b[k][...]
            assert got.shape == np.shape(v) and got.dtype == np.asarray(v).dtype, k
            np.testing.assert_array_equal(got, v, err_msg=k)
        np.testing.assert_array_equal(f["top/left:0"][...], arrays["f32"])
        np.testing.assert_array_equal(f["/a/b/f64"][2:5], arrays["f64"][2:5])
        assert list(b.attrs["list"]) == [b"x:0", b"yy:0"]
        assert b.attrs["pi"] == math.pi
        np.testing.assert_array_equal(b.attrs["ints"], [0, 1, 2])
        assert b.attrs["none"].shape == (0,)
        assert [p for p, _ in f.visit_datasets()][:2] == ["a/b/be", "a/b/empty"]


@pytest.mark.parametrize("count", [1, 8, 9, 257, 2100])
def test_large_groups_use_a_real_btree(tmp_path, count):
    """8 symbols per node, 32 children per B-tree node: 2100 members need 263 symbol nodes on two tree levels."""
    w = H.Writer()
    names = [f"conv2d_{i}" for i in range(count)]
    for i, n in enumerate(names):
        w.root.create_dataset(n, np.full((2,), i, np.int32))
    path = str(tmp_path / "big.h5")
    w.save(path)
    with H.File(path) as f:
        assert f.keys() == sorted(names)                          # tree order is strcmp order
        for i in (0, count // 2, count - 1):
            np.testing.assert_array_equal(f[names[i]][...], [i, i])
    raw = open(path, "rb").read()
    assert raw.count(b"SNOD") >= -(-count // 8)


def test_file_structure_matches_the_library_layout(tmp_path):
    """Field-by-field checks of what the writer emits against the format specification (the parts a foreign reader -
    libhdf5 under Keras - depends on and the round trip through our own reader could not catch)."""
    w = H.Writer()
    w.root.create_dataset("x", np.arange(4, dtype=np.float32))
    raw = w.tobytes()
    assert raw[:8] == H.SIGNATURE and raw[8] == 0 and raw[13] == 8 and raw[14] == 8
    eof = int.from_bytes(raw[40:48], "little")
    assert eof == len(raw) and len(raw) % 8 == 0
    root_hdr = int.from_bytes(raw[64:72], "little")
    assert int.from_bytes(raw[72:76], "little") == 1               # cache type 1: B-tree + heap cached in the entry
    bt, hp = int.from_bytes(raw[80:88], "little"), int.from_bytes(raw[88:96], "little")
    assert raw[bt:bt + 4] == b"TREE" and raw[hp:hp + 4] == b"HEAP" and raw[root_hdr] == 1
    # the B-tree node and the symbol node are allocated at their full size (the library reads whole nodes)
    assert len(raw) - bt >= 24 + (2 * 32 + 1) * 8
    heap_data = int.from_bytes(raw[hp + 24:hp + 32], "little")
    assert raw[heap_data:heap_data + 8] == b"\0" * 8 and raw[heap_data + 8:heap_data + 10] == b"x\0"


def _messages(raw, base, addr):
    """[(type, payload bytes)] of a version-1 object header (first block only)."""
    import struct
    a = base + addr
    _, _, n, _, size = struct.unpack_from("<BBHII", raw, a)
    p, out = a + 16, []
    while p + 8 <= a + 16 + size and len(out) < n:
        t, sz, _ = struct.unpack_from("<HHB", raw, p)
        out.append((t, bytes(raw[p + 8:p + 8 + sz])))
        p += 8 + sz
    return out


@pytest.mark.skipif(_scipy_hdf5_file() is None, reason="scipy's MATLAB v7.3 test file is not installed")
def test_writer_emits_the_structures_the_library_wrote():
    """Same content as the library-written file (one 9 x 1 float64 dataset with a 6-character string attribute in the
    root group): the datatype, dataspace and attribute messages, the name heap, the B-tree node and the symbol node of
    our writer equal the library's byte for byte (up to addresses, and the string padding flavour of the attribute:
    MATLAB null-terminates, numpy 'S' arrays are null-padded)."""
    import struct
    lib = open(_scipy_hdf5_file(), "rb").read()
    w = H.Writer()
    d = w.root.create_dataset("testdouble", (np.arange(9) * (math.pi / 4)).reshape(9, 1))
    d.attrs["MATLAB_class"] = b"double"
    ours = w.tobytes()

    def locate(raw, base):
        sb = base
        hdr = int.from_bytes(raw[sb + 64:sb + 72], "little")
        bt, hp = (int.from_bytes(raw[sb + o:sb + o + 8], "little") for o in (80, 88))
        heap_data = int.from_bytes(raw[base + hp + 24:base + hp + 32], "little")
        snod = int.from_bytes(raw[base + bt + 32:base + bt + 40], "little")
        entry = raw[base + snod + 8:base + snod + 48]
        return hdr, bt, heap_data, snod, entry

    lh, lbt, lheap, lsn, lentry = locate(lib, 512)
    oh, obt, oheap, osn, oentry = locate(ours, 0)
    # name heap: 8 zero bytes (the empty name of B-tree key 0), then the 8-byte aligned member name
    assert lib[512 + lheap:512 + lheap + 24] == ours[oheap:oheap + 24] == b"\0" * 8 + b"testdouble" + b"\0" * 6
    # B-tree node: signature, type 0, level 0, one entry, no siblings, key 0 = 0, key 1 = heap offset of the name
    assert lib[512 + lbt:512 + lbt + 32] == ours[obt:obt + 32]
    assert lib[512 + lbt + 40:512 + lbt + 48] == ours[obt + 40:obt + 48] == struct.pack("<Q", 8)
    # symbol node header and entry (name offset 8, cache type 0, empty scratch pad); object addresses differ
    assert lib[512 + lsn:512 + lsn + 8] == ours[osn:osn + 8] == b"SNOD\x01\x00\x01\x00"
    assert lentry[:8] == oentry[:8] and lentry[16:] == oentry[16:]
    # root group header: the symbol-table message
    assert _messages(lib, 512, lh)[0][0] == _messages(ours, 0, oh)[0][0] == 0x11
    lm = dict(_messages(lib, 512, int.from_bytes(lentry[8:16], "little")))
    om = dict(_messages(ours, 0, int.from_bytes(oentry[8:16], "little")))
    assert lm[0x03] == om[0x03]                                   # datatype: IEEE little-endian float64
    assert lm[0x01] == om[0x01]                                   # dataspace: rank 2, 9 x 1
    la, oa = bytearray(lm[0x0C]), bytearray(om[0x0C])
    assert la[25] == 0 and oa[25] == 1                            # string padding: null-terminated vs null-padded
    la[25] = oa[25] = 0
    assert bytes(la) == bytes(oa)                                 # attribute: name, string type, scalar space, value
    assert om[0x08][:2] == b"\x03\x01"                            # contiguous layout, message version 3 (h5py's)
    with H.File(_scipy_hdf5_file()) as f1:
        got = f1["testdouble"][...]
    import tempfile
    with tempfile.TemporaryDirectory() as dd:
        w.save(dd + "/o.h5")
        with H.File(dd + "/o.h5") as f2:
            np.testing.assert_array_equal(f2["testdouble"][...], got)


def _layers(named):
    return K.variables_to_layers(named)


def test_keras_layout_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    layers = [("input_1", []),
              ("conv2d", [("conv2d/kernel:0", rng.standard_normal((3, 3, 3, 8)).astype(np.float32)),
                          ("conv2d/bias:0", rng.standard_normal(8).astype(np.float32))]),
              ("leaky_re_lu", []),
              ("batch_normalization", [(f"batch_normalization/{k}:0", rng.standard_normal(8).astype(np.float32))
                                       for k in ("gamma", "beta", "moving_mean", "moving_variance")]),
              ("p_re_lu", [("p_re_lu/alpha:0", rng.standard_normal((1, 1, 8)).astype(np.float32))])]
    for as_model in (False, True):
        path = str(tmp_path / f"k{int(as_model)}.h5")
        H.save_keras_weights(path, layers, model_config=json.dumps({"class_name": "Model"}) if as_model else None)
        got, meta = H.load_keras_weights(path)
        assert [ln for ln, _ in got] == [ln for ln, _ in layers]
        for (_, a), (_, b) in zip(layers, got):
            assert [n for n, _ in a] == [n for n, _ in b]
            for (_, x), (_, y) in zip(a, b):
                np.testing.assert_array_equal(x, y)
        assert meta["backend"] == "tensorflow"
        assert ("model_config" in meta) == as_model
        with H.File(path) as f:
            assert ("model_weights" in f.keys()) == as_model
            root = f["model_weights"] if as_model else f
            assert root["conv2d"]["conv2d"]["kernel:0"].shape == (3, 3, 3, 8)      # Keras nests <layer>/<layer>/<weight>
    trainable, moving = K.layers_to_weight_list(got)
    assert [t.shape for t in trainable] == [(3, 3, 3, 8), (8,), (8,), (8,), (8,)] and len(moving) == 2


def test_long_layer_lists_are_chunked_like_keras(tmp_path):
    """More than 64 KB of names: ``layer_names0``, ``layer_names1``, ... (HDF5_OBJECT_HEADER_LIMIT handling)."""
    layers = [(f"a_rather_long_keras_layer_name_number_{i:05d}", []) for i in range(2500)]
    layers[7] = (layers[7][0], [("w:0", np.ones((2, 2), np.float32))])
    path = str(tmp_path / "long.h5")
    H.save_keras_weights(path, layers)
    with H.File(path) as f:
        assert "layer_names" not in f.attrs and "layer_names0" in f.attrs and "layer_names1" in f.attrs
    got, _ = H.load_keras_weights(path)
    assert [ln for ln, _ in got] == [ln for ln, _ in layers]
    np.testing.assert_array_equal(got[7][1][0][1], np.ones((2, 2)))


def _rrdb_variables(sf, nf, blocks, dense, convs, rng):
    gc = nf // 2
    out = []

    def add(name, cin, cout):
        out.append((f"{name}/kernel:0", rng.standard_normal((3, 3, cin, cout)).astype(np.float32)))
        out.append((f"{name}/bias:0", rng.standard_normal(cout).astype(np.float32)))

    add("fea", 3, nf)
    for b in range(blocks):
        for d in range(dense):
            for k in range(convs):
                add(f"rrdb{b}_db{d}_conv{k}", nf + k * gc, gc)
            add(f"rrdb{b}_db{d}_out", nf + convs * gc, nf)
    add("trunk", nf, nf)
    for u in range(int(math.log2(sf))):
        add(f"up{u}", nf, 4 * nf)
    add("hr", nf, nf)
    add("last", nf, 3)
    return out


def _srresnet_variables(sf, nf, blocks, bn, rng):
    train, moving = [], []

    def add(name, ks, cin, cout, prelu, with_bn=False, up=False):
        train.append((f"{name}/kernel:0", rng.standard_normal((ks, ks, cin, cout)).astype(np.float32)))
        train.append((f"{name}/bias:0", rng.standard_normal(cout).astype(np.float32)))
        if with_bn and bn:
            train.append((f"{name}_bn/gamma:0", rng.standard_normal(cout).astype(np.float32)))
            train.append((f"{name}_bn/beta:0", rng.standard_normal(cout).astype(np.float32)))
            moving.append((f"{name}_bn/moving_mean:0", rng.standard_normal(cout).astype(np.float32)))
            moving.append((f"{name}_bn/moving_variance:0", rng.uniform(0.5, 2, cout).astype(np.float32)))
        if prelu:
            train.append((f"{name}_prelu/alpha:0", rng.standard_normal(cout // 4 if up else cout).astype(np.float32)))

    add("first", 9, 3, nf, True)
    for b in range(blocks):
        add(f"res{b}_conv0", 3, nf, nf, True, with_bn=True)
        add(f"res{b}_conv1", 3, nf, nf, False, with_bn=True)
    add("trunk", 3, nf, nf, False, with_bn=True)
    for u in range(int(math.log2(sf))):
        add(f"up{u}", 3, nf, 4 * nf, True, up=True)
    add("last", 9, nf, 3, False)
    return train, moving


@pytest.mark.parametrize("sf,nf,blocks,dense,convs", [(4, 64, 2, 3, 4), (2, 32, 1, 1, 4), (8, 32, 5, 1, 2), (4, 32, 1, 3, 3)])
def test_rrdb_architecture_is_recovered_from_the_weights(tmp_path, sf, nf, blocks, dense, convs):
    named = _rrdb_variables(sf, nf, blocks, dense, convs, np.random.default_rng(2))
    path = str(tmp_path / "rrdb_gen_3.h5")

    class V:
        def __init__(self, n, a):
            self.name, self._a = n, a

        def numpy(self):
            return self._a

    K.write_model_file(path, [V(n, a) for n, a in named])
    arch, kw, trainable, moving = K.read_generator_file(path)
    assert arch == "rrdb" and moving == []
    assert kw["upsample_factor"] == sf and kw["num_filters"] == nf and kw["num_convs"] == convs
    assert kw["num_rrdb_blocks"] * kw["num_dense_blocks"] == blocks * dense       # only the product is observable
    if (blocks * dense) % 3 == 0:
        assert kw["num_dense_blocks"] == 3
    assert len(trainable) == len(named)
    for (_, a), b in zip(named, trainable):
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("sf,nf,blocks,bn", [(4, 64, 16, True), (2, 32, 1, False), (8, 16, 3, True)])
def test_srresnet_architecture_is_recovered_from_the_weights(tmp_path, sf, nf, blocks, bn):
    train, moving = _srresnet_variables(sf, nf, blocks, bn, np.random.default_rng(3))
    layers = K.variables_to_layers(train + moving)
    names = [ln for ln, _ in layers]
    assert names[:3] == ["first", "first_prelu", "res0_conv0"]
    if bn:
        i = names.index("res0_conv0_bn")
        assert [K._kind(w) for w, _ in layers[i][1]] == ["gamma", "beta", "moving_mean", "moving_variance"]
    assert layers[1][1][0][1].shape == (1, 1, nf)                 # PReLU(shared_axes=[1, 2]) slopes as Keras keeps them
    cfg = json.dumps({"class_name": "Model", "config": {"layers": [
        {"class_name": "Conv2D", "config": {}}, {"class_name": "BatchNormalization", "config": {"momentum": 0.65}}]}})
    path = str(tmp_path / "resnet_gen_1.h5")
    H.save_keras_weights(path, layers, model_config=cfg if bn else None, under_model_weights=True)
    arch, kw, t2, m2 = K.read_generator_file(path)
    assert arch == "srresnet"
    assert kw == dict(upsample_factor=sf, num_filters=nf, num_res_blocks=blocks, batch_normalization=bn,
                      momentum=0.65 if bn else 0.8)
    assert len(t2) == len(train) and len(m2) == len(moving)
    for (_, a), b in zip(train + moving, t2 + m2):
        np.testing.assert_array_equal(a, b)


def test_infer_generator_rejects_other_files():
    with pytest.raises(ValueError):
        K.infer_generator([("dense", [("dense/kernel:0", np.zeros((4, 4), np.float32))])])
    rng = np.random.default_rng(0)
    named = _rrdb_variables(4, 32, 1, 1, 4, rng)
    named[2] = (named[2][0], np.zeros((3, 3, 32, 24), np.float32))           # growth channels != num_filters // 2
    with pytest.raises(ValueError):
        K.infer_generator(K.variables_to_layers(named))


def test_vgg19_file_in_the_stock_keras_naming(tmp_path):
    """The stock file names its weights ``block1_conv1_W_1:0`` / ``block1_conv1_b_1:0`` and lists the pooling layers
    and the input layer without weights."""
    from simplesr_b200.vgg import VGG19_LAYERS
    rng = np.random.default_rng(4)
    layers, want = [("input_1", [])], []
    for layer in VGG19_LAYERS:
        if len(layer) == 3:
            name, cin, cout = layer
            cin, cout = min(cin, 8), min(cout, 8)                 # shapes are not this function's business
            k = rng.standard_normal((3, 3, cin, cout)).astype(np.float32)
            b = rng.standard_normal(cout).astype(np.float32)
            layers.append((name, [(f"{name}_W_1:0", k), (f"{name}_b_1:0", b)]))
            want += [k, b]
        else:
            layers.append((layer[0], []))
    path = str(tmp_path / "vgg19_weights_tf_dim_ordering_tf_kernels_notop.h5")
    H.save_keras_weights(path, layers)
    conv_names = [l[0] for l in VGG19_LAYERS if len(l) == 3]
    got = K.read_vgg19_file(path, conv_names)
    assert len(got) == 32
    for a, b in zip(want, got):
        np.testing.assert_array_equal(a, b)
    # a save_weights of a functional copy with auto-named layers: matched by position
    renamed = [(f"conv2d_{i}" if ws else ln, ws) for i, (ln, ws) in enumerate(layers)]
    path2 = str(tmp_path / "custom_vgg.h5")
    H.save_keras_weights(path2, renamed)
    for a, b in zip(want, K.read_vgg19_file(path2, conv_names)):
        np.testing.assert_array_equal(a, b)
    assert K.is_hdf5(path2) and not K.is_hdf5(str(tmp_path / "weights.npz"))


def test_round_trip_random_trees(tmp_path):
    """Seeded random group trees (names with ':' and digits, 0 to 40 members per group, shapes of rank 0 to 4 including
    empty ones, four element types): everything written comes back, in name order, through the B-tree walk."""
    rng = np.random.default_rng(7)
    dtypes = [np.float32, np.float64, np.int32, np.uint8]
    for trial in range(12):
        w = H.Writer()
        want = {}

        def fill(group, prefix, depth):
            for i in range(int(rng.integers(0, 41 if depth == 0 else 12))):
                name = f"n{int(rng.integers(0, 10 ** 6))}_{i}" + (":0" if rng.random() < 0.3 else "")
                if depth < 2 and rng.random() < 0.25:
                    fill(group.create_group(name), prefix + name + "/", depth + 1)
                else:
                    shape = tuple(int(x) for x in rng.integers(0, 5, size=int(rng.integers(0, 5))))
                    arr = (rng.standard_normal(shape) * 100).astype(dtypes[int(rng.integers(0, 4))])
                    group.create_dataset(name, arr)
                    want[prefix + name] = arr
            group.attrs["count"] = np.int32(len(group.children))

        fill(w.root, "", 0)
        path = str(tmp_path / f"r{trial}.h5")
        w.save(path)
        with H.File(path) as f:
            got = dict(f.visit_datasets())
            assert sorted(got) == sorted(want)
            for k, arr in want.items():
                g = got[k][...]
                assert g.shape == arr.shape and g.dtype == arr.dtype, k
                np.testing.assert_array_equal(g, arr, err_msg=k)
            assert f.attrs["count"] == len(f.keys())


@pytest.mark.parametrize("filters", [dict(), dict(compression="gzip"), dict(compression="gzip", shuffle=True),
                                     dict(shuffle=True)])
def test_chunked_and_filtered_datasets(tmp_path, filters):
    """Chunked layout with partial edge chunks, deflate and shuffle: the chunk B-tree walk, the filter pipeline in
    reverse order and the placement of every chunk."""
    rng = np.random.default_rng(3)
    a = rng.standard_normal((3, 3, 10, 7)).astype(np.float32)
    b = rng.integers(0, 1000, size=(33,)).astype(np.int64)
    c = np.zeros((0, 4), np.float32)
    w = H.Writer()
    w.root.create_dataset("g/a", a, chunks=(2, 3, 4, 7), **filters)
    w.root.create_dataset("g/b", b, chunks=(8,), **filters)
    w.root.create_dataset("g/c", c, chunks=(1, 4), **filters)
    w.root.create_dataset("g/plain", a)
    path = str(tmp_path / "chunked.h5")
    w.save(path)
    with H.File(path) as f:
        np.testing.assert_array_equal(f["g/a"][...], a)
        np.testing.assert_array_equal(f["g/b"][...], b)
        assert f["g/c"][...].shape == (0, 4)
        np.testing.assert_array_equal(f["g/plain"][...], a)
        assert [fid for fid, _ in f["g/a"]._obj.filters] == ([2] if filters.get("shuffle") else []) + \
            ([1] if filters.get("compression") else [])
    if filters.get("compression"):
        assert os.path.getsize(path) > 0
    with pytest.raises(H.H5Error):
        H.Writer().root.create_dataset("x", a, compression="gzip")
    with pytest.raises(H.H5Error):
        H.Writer().root.create_dataset("x", a, chunks=(1, 1))


def test_convert_weights_tool(tmp_path, capsys):
    """tools/convert_weights.py: Keras HDF5 -> the self-describing .npz of GeneratorModel.save -> Keras HDF5, and --info."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "convert_weights", os.path.join(os.path.dirname(os.path.dirname(__file__)), "tools", "convert_weights.py"))
    tool = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tool)
    train, moving = _srresnet_variables(4, 32, 2, True, np.random.default_rng(5))
    h5a, npz, h5b = (str(tmp_path / n) for n in ("resnet_gen_9.h5", "gen.npz", "back.h5"))
    H.save_keras_weights(h5a, K.variables_to_layers(train + moving), under_model_weights=True)
    assert tool.main([h5a, npz]) == 0
    with np.load(npz) as z:
        assert str(z["__architecture__"]) == "srresnet" and int(z["__upsample_factor__"]) == 4
        cfg = json.loads(str(z["__config__"]))
        assert cfg["num_res_blocks"] == 2 and cfg["batch_norm"] is True and cfg["num_filters"] == 32
        keys = sorted(k for k in z.files if "|" in k)
        assert len(keys) == len(train) + len(moving)
        for k, (name, arr) in zip(keys, train + moving):
            assert k.split("|", 1)[1] == name
            np.testing.assert_array_equal(z[k], arr)
    assert tool.main([npz, h5b]) == 0
    a, _ = H.load_keras_weights(h5a)
    b, _ = H.load_keras_weights(h5b)
    assert [ln for ln, _ in a] == [ln for ln, _ in b]
    for (_, wa), (_, wb) in zip(a, b):
        for (na, xa), (nb_, xb) in zip(wa, wb):
            assert na == nb_
            np.testing.assert_array_equal(xa, xb)
    assert tool.main(["--info", h5a]) == 0
    out = capsys.readouterr().out
    assert "architecture: srresnet" in out and "res1_conv1_bn" in out and "parameters" in out
    assert tool.main([h5a]) == 2
