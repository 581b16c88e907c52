"""TensorFlow tensor-bundle reader / writer without TensorFlow (monopsr_b200/core/tf_checkpoint.py) and the variable
mapping of the reference's two restore paths.  No TF-written checkpoint exists in the reference checkout, so the format
is checked by round trips through the writer, known CRC-32C / varint / snappy vectors and hand-assembled bytes."""
import os
import struct

import numpy as np
import pytest

from monopsr_b200.core import model_spec as ms
from monopsr_b200.core import tf_checkpoint as C


def test_primitives_known_answers():
    assert C.crc32c(b"123456789") == 0xE3069283                     # the standard CRC-32C check value
    assert C.crc32c(b"") == 0 and C.crc32c(b"\x00" * 32) == 0x8A9136AA          # RFC 3720 B.4 test vector
    assert C.unmask_crc(C.mask_crc(0xDEADBEEF)) == 0xDEADBEEF
    assert C._put_varint(300) == b"\xac\x02" and C._varint(b"\xac\x02", 0) == (300, 2)
    # snappy: literal "abcd" + copy(offset 4, length 8) -> "abcdabcdabcd"
    assert C.snappy_decompress(bytes([12, (4 - 1) << 2]) + b"abcd" + bytes([((8 - 4) << 2) | 1, 4])) == b"abcdabcdabcd"
    with pytest.raises(ValueError):
        C.snappy_decompress(bytes([4, 1 | (0 << 2), 9]))            # copy before any output


def _tensors(seed=0, n_small=150):
    rng = np.random.RandomState(seed)
    T = {"a/weights": rng.randn(3, 3, 4, 8).astype(np.float32), "a/BatchNorm/gamma": rng.randn(8).astype(np.float32),
         "global_step": np.array(123, np.int64), "dbl": rng.randn(5, 7), "flags": rng.rand(9) < 0.5,
         "half": rng.randn(6).astype(np.float16), "empty": np.zeros((0, 3), np.float32)}
    for i in range(n_small):                                         # several index blocks
        T["layer_%03d/biases" % i] = rng.randn(i % 7 + 1).astype(np.float32)
    return T


def test_roundtrip_and_checksums(tmp_path):
    T = _tensors()
    prefix = str(tmp_path / "model.ckpt-5")
    C.write_bundle(prefix, T, block_entries=16)
    R = C.read_bundle(prefix, verify_data=True)
    assert set(R) == set(T)
    for k in T:
        assert R[k].dtype == T[k].dtype and R[k].shape == T[k].shape and np.array_equal(R[k], T[k]), k
    sub = C.read_bundle(prefix, names={"a/weights", "global_step"})
    assert set(sub) == {"a/weights", "global_step"} and int(sub["global_step"]) == 123
    # footer: magic in the last 8 bytes; a flipped byte in an index block or in the data file is detected
    raw = bytearray(open(prefix + ".index", "rb").read())
    assert struct.unpack("<Q", raw[-8:])[0] == C.MAGIC
    raw[10] ^= 0xFF
    open(prefix + ".index", "wb").write(bytes(raw))
    with pytest.raises(ValueError):
        C.read_bundle(prefix)
    C.write_bundle(prefix, T)
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[5] ^= 0x01
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    with pytest.raises(ValueError):
        C.read_bundle(prefix, verify_data=True)
    with pytest.raises(ValueError):
        open(prefix + ".index", "wb").write(b"not a table" * 10)
        C.read_bundle(prefix)


def test_reads_prefix_compressed_and_snappy_blocks(tmp_path):
    """a block as LevelDB's BlockBuilder writes it (shared key prefixes between restart points), once uncompressed
    and once as a snappy literal, assembled by hand"""
    a = np.arange(6, dtype=np.float32).reshape(2, 3)
    b = np.arange(4, dtype=np.float32)
    raw = a.tobytes() + b.tobytes()
    e1 = C._entry_proto(a, 0, C.crc32c(a.tobytes()))
    e2 = C._entry_proto(b, a.nbytes, C.crc32c(b.tobytes()))
    header = b"\x08\x01\x10\x00"
    k1, k2 = b"scope/conv/biases", b"scope/conv/weights"
    shared = len(os.path.commonprefix([k1, k2]))
    body = (b"\x00\x00" + C._put_varint(len(header)) + header +
            b"\x00" + C._put_varint(len(k1)) + C._put_varint(len(e2)) + k1 + e2 +
            C._put_varint(shared) + C._put_varint(len(k2) - shared) + C._put_varint(len(e1)) + k2[shared:] + e1 +
            struct.pack("<II", 0, 1))
    for compressed in (False, True):
        if compressed:                     # one snappy literal holding the whole block
            n = len(body) - 1
            stored = C._put_varint(len(body)) + bytes([61 << 2]) + struct.pack("<H", n) + body
            ctype = b"\x01"
        else:
            stored, ctype = body, b"\x00"
        blk = stored + ctype + struct.pack("<I", C.mask_crc(C.crc32c(ctype, C.crc32c(stored))))
        meta_body, meta_tr = C._block([])
        idx_body, idx_tr = C._block([(k2, C._put_varint(0) + C._put_varint(len(stored)))], restart_interval=1)
        out = blk + meta_body + meta_tr
        mh = C._put_varint(len(blk)) + C._put_varint(len(meta_body))
        ih = C._put_varint(len(out)) + C._put_varint(len(idx_body))
        out += idx_body + idx_tr
        out += mh + ih + b"\x00" * (40 - len(mh) - len(ih)) + struct.pack("<Q", C.MAGIC)
        prefix = str(tmp_path / ("hand%d" % compressed))
        open(prefix + ".index", "wb").write(out)
        open(prefix + ".data-00000-of-00001", "wb").write(raw)
        R = C.read_bundle(prefix, verify_data=True)
        assert np.array_equal(R["scope/conv/weights"], a) and np.array_equal(R["scope/conv/biases"], b)


def test_monopsr_checkpoint_mapping(tmp_path):
    table = ms.param_table()[:40] + [t for t in ms.param_table() if t[0].startswith("squash")]
    rng = np.random.RandomState(1)
    ck = {n: rng.randn(*s).astype(np.float32) for n, s, _ in table}
    ck.update({n + C.EMA_SUFFIX: v * 0.5 for n, v in list(ck.items())[:10]})
    ck["global_step"] = np.array(7, np.int64)
    wrong = table[3][0]
    ck[wrong] = np.zeros((1, 2, 3), np.float32)                      # same name, different shape: skipped
    del ck[table[5][0]]                                              # absent: reported missing
    prefix = str(tmp_path / "model.ckpt-7")
    C.write_bundle(prefix, ck)
    params, rep = C.load_checkpoint(prefix, table, kind="monopsr")
    assert wrong in rep["shape_mismatch"] and table[5][0] in rep["missing"]
    assert set(rep["loaded"]) == set(params) and len(params) == len(table) - 2
    assert np.array_equal(params[table[0][0]], ck[table[0][0]])
    ema, _ = C.load_checkpoint(prefix, table, kind="monopsr", use_ema=True)
    assert np.array_equal(ema[table[0][0]], ck[table[0][0] + C.EMA_SUFFIX])          # shadow wins where present
    assert np.array_equal(ema[table[20][0]], ck[table[20][0]])
    with pytest.raises(ValueError):
        C.load_checkpoint(prefix, table, kind="bogus")


def test_detection_checkpoint_feeds_both_encoders(tmp_path):
    """'FirstStageFeatureExtractor/x' -> '..._crop/x' and '..._full/x' (core/checkpoint_utils.py:64-117)"""
    table = ms.param_table()
    crop = [t for t in table if t[0].startswith(ms.ENCODERS[0] + "/")][:12]
    full = [t for t in table if t[0].startswith(ms.ENCODERS[1] + "/")][:12]
    rng = np.random.RandomState(2)
    ck = {"FirstStageFeatureExtractor/" + n[len(ms.ENCODERS[0]) + 1:]: rng.randn(*s).astype(np.float32) for n, s, _ in crop}
    ck["SecondStageBoxPredictor/weights"] = np.zeros((4, 4), np.float32)
    prefix = str(tmp_path / "model.ckpt")
    C.write_bundle(prefix, ck)
    params, rep = C.load_checkpoint(prefix, crop + full, kind="detection")
    assert len(params) == 24 and not rep["missing"] and not rep["shape_mismatch"]
    for (nc, _, _), (nf, _, _) in zip(crop, full):
        assert np.array_equal(params[nc], params[nf])
        assert np.array_equal(params[nc], ck["FirstStageFeatureExtractor/" + nc[len(ms.ENCODERS[0]) + 1:]])


def test_primitives_against_the_tensorflow_ecosystem_code_in_tensorboard():
    """tensorboard ships TensorFlow's record checksum (CRC-32C + the rotate-and-add mask) and the generated protobuf
    classes of tensor_shape.proto / types.proto / versions.proto: an independent implementation, written by the
    TensorFlow authors, of three pieces the bundle format is made of."""
    pw = pytest.importorskip("tensorboard.compat.tensorflow_stub.pywrap_tensorflow")
    shape_pb2 = pytest.importorskip("tensorboard.compat.proto.tensor_shape_pb2")
    types_pb2 = pytest.importorskip("tensorboard.compat.proto.types_pb2")
    rs = np.random.RandomState(0)
    for n in (0, 1, 7, 8, 9, 63, 64, 65, 1000, 4097):
        data = rs.randint(0, 256, n).astype(np.uint8).tobytes()
        assert C.crc32c(data) == pw.crc32c(data)
        assert C.mask_crc(C.crc32c(data)) == pw.masked_crc32c(data)
    # the enum values the reader / writer use for dtypes
    for np_dtype, name in ((np.float32, "DT_FLOAT"), (np.float64, "DT_DOUBLE"), (np.int32, "DT_INT32"),
                           (np.int64, "DT_INT64"), (np.uint8, "DT_UINT8"), (np.float16, "DT_HALF"), (np.bool_, "DT_BOOL")):
        if np.dtype(np_dtype) in C.DTYPE_ENUM:
            assert C.DTYPE_ENUM[np.dtype(np_dtype)] == getattr(types_pb2, name), name
            assert np.dtype(C.DTYPES[getattr(types_pb2, name)]) == np.dtype(np_dtype)
    assert C.DTYPE_ENUM[np.dtype(np.float32)] == types_pb2.DT_FLOAT and C.DTYPE_ENUM[np.dtype(np.int64)] == types_pb2.DT_INT64
    # BundleEntryProto.shape (field 2) as written here parses with the generated TensorShapeProto class, and a shape
    # serialised BY that class parses with the reader
    for shape in ((), (7,), (3, 3, 64, 256), (1, 0, 5), (18432, 1024)):
        arr = np.zeros(shape, np.float32) if int(np.prod(shape, dtype=np.int64)) < 10 ** 6 else np.lib.stride_tricks.as_strided(
            np.zeros(1, np.float32), shape=shape, strides=(0,) * len(shape))
        msg = C._entry_proto(arr, 12, 0xABCDEF)
        fields = {f: v for f, _, v in C._proto_fields(msg)}
        assert fields[1] == types_pb2.DT_FLOAT and fields[4] == 12 and fields[5] == arr.size * 4
        parsed = shape_pb2.TensorShapeProto.FromString(bytes(fields.get(2, b"")))
        assert tuple(d.size for d in parsed.dim) == tuple(shape)
        theirs = shape_pb2.TensorShapeProto(dim=[shape_pb2.TensorShapeProto.Dim(size=s) for s in shape]).SerializeToString()
        assert C._parse_shape(theirs) == tuple(shape)
    with pytest.raises(ValueError):
        C._parse_shape(shape_pb2.TensorShapeProto(unknown_rank=True).SerializeToString())


def test_detection_checkpoint_mapping_equals_the_reference_restore_logic():
    """which model variable is fed from which checkpoint variable: map_detection_checkpoint against the record of the
    reference's own restore code (checkpoint_utils.restore_obj_detection_api_weights + get_variable_restore_map +
    variables_helper.get_variables_available_in_checkpoint, executed: tests/golden/make_restore_golden.py) on the same
    synthetic detection checkpoint -- both towers from 'FirstStageFeatureExtractor/', block4 and the RPN / predictor
    variables untouched, a variable with a foreign shape skipped"""
    import importlib.util
    import json
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_restore_golden", os.path.join(here, "golden", "make_restore_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    gold = json.load(open(os.path.join(here, "golden", "restore_golden.json")))["restored"]
    mv = gen.model_variables()
    ck_shapes = gen.detection_checkpoint(mv)
    ck = {n: np.zeros(s, np.float32) for n, s in ck_shapes.items()}
    from monopsr_b200.core import model_spec as ms
    loaded, report = C.map_detection_checkpoint(ck, ms.param_table())
    want = {}
    for saver in gold:                                      # checkpoint name -> model variable, per tower
        assert saver["other"] == {}                         # nothing but '<first stage>/x' -> '<tower>/x' pairs
        for suffix in saver["suffixes"]:
            want[saver["model_prefix"] + suffix] = saver["ckpt_prefix"] + suffix
    assert sorted(s["model_prefix"] for s in gold) == sorted(e + "/" for e in ms.ENCODERS)
    table = {n for n, _, _ in ms.param_table()}
    # the reference also "restores" block4 of nothing (not in the first-stage scope) -- and this engine has no block4
    want_here = {m: c for m, c in want.items() if m in table}
    assert set(loaded) == set(want_here)
    assert all(c == "FirstStageFeatureExtractor/" + m.split("/", 1)[1] for m, c in want_here.items())
    assert not [m for m in want if "/block4/" in m]                     # block4 lives in the SECOND stage of the checkpoint
    assert len(gold[0]["suffixes"]) == len(gold[1]["suffixes"]) == 469
    for enc in ms.ENCODERS:
        assert enc + "/resnet_v1_101/conv1/weights" not in loaded       # foreign shape: skipped by both
        assert enc + "/resnet_v1_101/conv1/BatchNorm/gamma" in loaded
    assert not any(n.startswith(("squash/", "map_decoder/", "output/")) for n in loaded)


def test_training_resume_restores_adam_and_ema_slots(tmp_path):
    """a training checkpoint carries '<var>/Adam', '<var>/Adam_1', '<var>/ExponentialMovingAverage' (tf.train.Saver saves
    the optimizer's slot variables; core/trainer.py:85,148-153 restores them): map_slots finds them, also below the
    train-op's name scope, and load_checkpoint(with_slots=True) reports them with the int32 global step"""
    rng = np.random.RandomState(3)
    table = [("x/weights", (3, 3, 4, 8), "weights"), ("x/BatchNorm/gamma", (8,), "gamma"), ("y/weights", (16, 5), "weights")]
    T = {}
    for n, s, _ in table:
        T[n] = rng.randn(*s).astype(np.float32)
    for n, s, _ in table[:2]:
        for sfx in (C.ADAM_M_SUFFIX, C.ADAM_V_SUFFIX, C.EMA_SUFFIX):
            T[n + sfx] = rng.randn(*s).astype(np.float32)
    for sfx in (C.ADAM_M_SUFFIX, C.ADAM_V_SUFFIX, C.EMA_SUFFIX):          # the third variable: slots under 'train_op/'
        T["train_op/y/weights" + sfx] = rng.randn(16, 5).astype(np.float32)
    T["global_step"] = np.array(4000, np.int32)
    prefix = str(tmp_path / "m-00004000")
    C.write_bundle(prefix, T)
    assert sorted(os.listdir(tmp_path)) == ["m-00004000.data-00000-of-00001", "m-00004000.index"]     # no temporary files left
    params, report = C.load_checkpoint(prefix, table, with_slots=True)
    assert report["global_step"] == 4000 and not report["missing"]
    assert sorted(report["slots"]) == sorted(n for n, _, _ in table)
    for n, _, _ in table:
        m, v, e = report["slots"][n]
        src = n if n != "y/weights" else "train_op/" + n
        assert np.array_equal(m, T[src + C.ADAM_M_SUFFIX]) and np.array_equal(v, T[src + C.ADAM_V_SUFFIX])
        assert np.array_equal(e, T[src + C.EMA_SUFFIX]) and np.array_equal(params[n], T[n])
    # the shadows win only on request, and are found below the name scope too
    ema, _ = C.load_checkpoint(prefix, table, use_ema=True)
    assert np.array_equal(ema["y/weights"], T["train_op/y/weights" + C.EMA_SUFFIX])
    raw, rep = C.load_checkpoint(prefix, table)
    assert np.array_equal(raw["y/weights"], T["y/weights"]) and "slots" not in rep
    # a variable with an incomplete slot set is left to re-initialisation
    del T["x/weights" + C.ADAM_V_SUFFIX]
    C.write_bundle(prefix, T)
    assert "x/weights" not in C.load_checkpoint(prefix, table, with_slots=True)[1]["slots"]


def test_a_reader_never_sees_a_half_written_checkpoint(tmp_path, monkeypatch):
    """data and index are written under temporary names and renamed into place, the index last"""
    order = []
    real = os.replace
    monkeypatch.setattr(os, "replace", lambda a, b: (order.append(os.path.basename(b)), real(a, b))[1])
    C.write_bundle(str(tmp_path / "c"), {"v": np.arange(4, dtype=np.float32)})
    assert order == ["c.data-00000-of-00001", "c.index"]
    assert np.array_equal(C.read_bundle(str(tmp_path / "c"))["v"], np.arange(4, dtype=np.float32))
