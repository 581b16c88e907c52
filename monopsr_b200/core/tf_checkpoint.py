"""TensorFlow checkpoint (tensor bundle, "V2" format) reader / writer without TensorFlow, and the variable mapping
of MonoPSR's two restore paths.

TensorFlow is not installable here and no TF-written checkpoint exists in the reference checkout, so the FILE FORMAT
below is restated from its published definition and is UNPINNED against a real TF file (tests round-trip it through
the writer and check hand-assembled bytes):
  <prefix>.index   a LevelDB-format sorted string table (tensorflow/core/lib/io/table*.cc, format.cc):
                   data blocks | metaindex block | index block | 48-byte footer (two BlockHandles as varint64 pairs,
                   zero padding, magic 0xdb4775248b80fb57 little endian).  A block is a run of prefix-compressed entries
                   (varint32 shared, non_shared, value_len; key suffix; value), a restart array (uint32 each + count)
                   and a 5-byte trailer (compression type, masked crc32c).  BundleWriter writes uncompressed blocks;
                   snappy-compressed blocks are decoded too.
                   key ""  -> BundleHeaderProto {1: num_shards, 2: endianness, 3: version}
                   key var -> BundleEntryProto  {1: dtype, 2: TensorShapeProto, 3: shard_id, 4: offset, 5: size,
                                                 6: crc32c (fixed32, masked), 7: slices}       (tensor_bundle.proto)
  <prefix>.data-SSSSS-of-NNNNN   raw little-endian tensor bytes at [offset, offset + size)
The VARIABLE MAPPING follows the reference: core/checkpoint_utils.py:64-117 (the object-detection-API ResNet-101
checkpoint feeds BOTH encoders: 'FirstStageFeatureExtractor/...' -> '..._crop/...' and '..._full/...'),
object_detection/utils/variables_helper.py:99-144 (only variables present with the same shape are restored) and
tf.train.Saver over MovingAverageOptimizer variables (core/trainer.py:122-167: the EMA shadow of variable v is stored
as 'v/ExponentialMovingAverage').  Layouts need no conversion: the engine's public parameter dict is already in TF
layout (HWIO convolutions, [in, out] FC, HWC flatten order)."""
import os
import struct

import numpy as np

MAGIC = 0xDB4775248B80FB57
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
          17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
DTYPE_ENUM = {np.dtype(v): k for k, v in DTYPES.items()}


# ------------------------------------------------------------------------------------------------- primitives
def _varint(buf, pos):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _put_varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _proto_fields(buf):
    """[(field number, wire type, value)] of one protobuf message; nested messages stay bytes"""
    pos, out = 0, []
    while pos < len(buf):
        key, pos = _varint(buf, pos)
        f, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        out.append((f, wt, v))
    return out


_CRC_TABLE = None
_NATIVE = None


def _native_crc():
    """mpb_crc32c of the built library (slicing-by-8, ~1 GB/s); None when the library has not been built"""
    global _NATIVE
    if _NATIVE is None:
        try:
            import ctypes
            from .. import lib as _lib
            fn = _lib.load().mpb_crc32c
            fn.argtypes, fn.restype = [ctypes.c_void_p, ctypes.c_ulonglong, ctypes.c_uint], ctypes.c_uint
            _NATIVE = fn
        except Exception:       # noqa: BLE001 -- any failure to load simply selects the Python loop below
            _NATIVE = False
    return _NATIVE or None


def crc32c(data, crc=0):
    """CRC-32C (Castagnoli): the library's host routine for tensor data, a byte-wise Python table otherwise (index
    blocks are a few KB)"""
    global _CRC_TABLE
    data = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
    if len(data) >= 4096:
        fn = _native_crc()
        if fn is not None:
            import ctypes
            buf = (ctypes.c_char * len(data)).from_buffer_copy(data)
            return int(fn(ctypes.cast(buf, ctypes.c_void_p), len(data), crc))
    if _CRC_TABLE is None:
        tab = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tab.append(c)
        _CRC_TABLE = tab
    c = crc ^ 0xFFFFFFFF
    tab = _CRC_TABLE
    for b in bytes(data):
        c = tab[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def mask_crc(c):
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def unmask_crc(m):
    rot = (m - 0xA282EAD8) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


def snappy_decompress(buf):
    """raw snappy block format (LevelDB block compression type 1)"""
    n, pos = _varint(buf, 0)
    out = bytearray()
    while pos < len(buf):
        tag = buf[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:                                   # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(buf[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += buf[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | buf[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 2], "little")
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(buf[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt snappy stream")
        for _ in range(ln):                             # overlapping copies are byte-serial by definition
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy length mismatch")
    return bytes(out)


# ------------------------------------------------------------------------------------------------- table reader
def _read_block(data, offset, size, verify=True):
    body, trailer = data[offset:offset + size], data[offset + size:offset + size + 5]
    if len(trailer) != 5:
        raise ValueError("truncated table block")
    if verify and unmask_crc(struct.unpack("<I", trailer[1:])[0]) != crc32c(trailer[:1], crc32c(body)):
        raise ValueError("table block checksum mismatch")
    if trailer[0] == 1:
        body = snappy_decompress(body)
    elif trailer[0] != 0:
        raise ValueError("unknown block compression %d" % trailer[0])
    return body


def _block_entries(block):
    nrestarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * nrestarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(block[pos:pos + vlen])))
        pos += vlen
    return out


def read_table(path, verify=True):
    """[(key bytes, value bytes)] of a LevelDB-format table file, in key order"""
    data = open(path, "rb").read()
    if len(data) < 48 or struct.unpack("<Q", data[-8:])[0] != MAGIC:
        raise ValueError("%s is not a TensorFlow checkpoint index (bad magic)" % path)
    footer = data[-48:]
    _, p = _varint(footer, 0)
    _, p = _varint(footer, p)              # metaindex handle (unused)
    ioff, p = _varint(footer, p)
    isz, p = _varint(footer, p)
    out = []
    for _, handle in _block_entries(_read_block(data, ioff, isz, verify)):
        off, q = _varint(handle, 0)
        sz, q = _varint(handle, q)
        out += _block_entries(_read_block(data, off, sz, verify))
    return out


# ------------------------------------------------------------------------------------------------- bundle reader
def _parse_shape(buf):
    dims = []
    for f, _, v in _proto_fields(buf):
        if f == 2:
            size = 0
            for g, _, w in _proto_fields(v):
                if g == 1:
                    size = w - (1 << 64) if w >= (1 << 63) else w
            dims.append(size)
        elif f == 3 and v:
            raise ValueError("tensor of unknown rank in checkpoint")
    return tuple(dims)


def read_bundle_index(prefix, verify=True):
    """{variable name: dict(dtype, shape, shard_id, offset, size, crc32c)}, header dict"""
    entries, header = {}, {"num_shards": 1, "endianness": 0}
    for key, value in read_table(prefix + ".index", verify):
        fields = _proto_fields(value)
        if key == b"":
            for f, _, v in fields:
                if f == 1:
                    header["num_shards"] = v
                elif f == 2:
                    header["endianness"] = v
            if header["endianness"] != 0:
                raise ValueError("big-endian checkpoints are not supported")
            continue
        e = {"dtype": 0, "shape": (), "shard_id": 0, "offset": 0, "size": 0, "crc32c": None, "sliced": False}
        for f, _, v in fields:
            if f == 1:
                e["dtype"] = v
            elif f == 2:
                e["shape"] = _parse_shape(v)
            elif f == 3:
                e["shard_id"] = v
            elif f == 4:
                e["offset"] = v
            elif f == 5:
                e["size"] = v
            elif f == 6:
                e["crc32c"] = v
            elif f == 7:
                e["sliced"] = True
        entries[key.decode("utf-8")] = e
    return entries, header


def read_bundle(prefix, names=None, verify_data=False):
    """{name: ndarray} of a checkpoint `prefix` (.index + .data-*); `names` restricts what is read"""
    entries, header = read_bundle_index(prefix)
    out, files = {}, {}
    for name, e in entries.items():
        if names is not None and name not in names:
            continue
        if e["sliced"]:
            raise ValueError("partitioned variable %r is not supported" % name)
        if e["dtype"] not in DTYPES:
            raise ValueError("variable %r has unsupported dtype enum %d" % (name, e["dtype"]))
        sid = e["shard_id"]
        if sid not in files:
            files[sid] = np.memmap("%s.data-%05d-of-%05d" % (prefix, sid, header["num_shards"]), dtype=np.uint8, mode="r")
        raw = files[sid][e["offset"]:e["offset"] + e["size"]]
        dt = np.dtype(DTYPES[e["dtype"]])
        if e["size"] != int(np.prod(e["shape"], dtype=np.int64)) * dt.itemsize:
            raise ValueError("variable %r: size does not match dtype x shape" % name)
        if verify_data and e["crc32c"] is not None and unmask_crc(e["crc32c"]) != crc32c(raw):
            raise ValueError("variable %r: data checksum mismatch" % name)
        out[name] = np.frombuffer(bytes(raw), dtype=dt.newbyteorder("<")).reshape(e["shape"]).astype(dt)
    return out


# ------------------------------------------------------------------------------------------------- writer
def _block(pairs, restart_interval=16):
    """uncompressed block with trailer; keys stored without prefix sharing (valid, and what every restart point is)"""
    body, restarts = bytearray(), []
    for i, (k, v) in enumerate(pairs):
        if i % restart_interval == 0:
            restarts.append(len(body))
        # between restart points the format ALLOWS sharing; shared = 0 is always a legal encoding
        body += _put_varint(0) + _put_varint(len(k)) + _put_varint(len(v)) + k + v
    if not restarts:
        restarts = [0]
    for r in restarts:
        body += struct.pack("<I", r)
    body += struct.pack("<I", len(restarts))
    trailer = b"\x00" + struct.pack("<I", mask_crc(crc32c(b"\x00", crc32c(body))))
    return bytes(body), trailer


def _entry_proto(arr, offset, crc):
    shape = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(int(s)) for s in arr.shape))
    msg = b"\x08" + _put_varint(DTYPE_ENUM[arr.dtype]) + b"\x12" + _put_varint(len(shape)) + shape
    msg += b"\x20" + _put_varint(offset) + b"\x28" + _put_varint(arr.nbytes) + b"\x35" + struct.pack("<I", mask_crc(crc))
    return msg


def write_bundle(prefix, tensors, block_entries=64, with_data_crc=True):
    """write {name: ndarray} as <prefix>.index + <prefix>.data-00000-of-00001 (one shard, uncompressed blocks)"""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    pairs, offset = [], 0
    # like TensorFlow's BundleWriter: both files are written under temporary names and renamed into place, the index
    # LAST -- a reader that globs '*.index' (the evaluator's polling loop, other data-parallel ranks) never sees a
    # half-written checkpoint
    tmp = "%s.tmp%d" % (prefix, os.getpid())
    with open(tmp + ".data", "wb") as f:
        for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
            a = np.asarray(tensors[name]).copy(order="C")        # (ascontiguousarray would turn a scalar into shape (1,))
            a = a.astype(a.dtype.newbyteorder("<")) if a.dtype.byteorder == ">" else a
            raw = a.tobytes()
            f.write(raw)
            pairs.append((name.encode("utf-8"), _entry_proto(a, offset, crc32c(raw) if with_data_crc else 0)))
            offset += len(raw)
    header = b"\x08\x01" + b"\x10\x00" + b"\x1a\x02\x08\x01"        # num_shards 1, little endian, version {producer 1}
    pairs = [(b"", header)] + pairs
    out, index = bytearray(), []
    for i in range(0, len(pairs), block_entries):
        chunk = pairs[i:i + block_entries]
        body, trailer = _block(chunk)
        index.append((chunk[-1][0], _put_varint(len(out)) + _put_varint(len(body))))
        out += body + trailer
    meta_body, meta_trailer = _block([])
    meta_handle = _put_varint(len(out)) + _put_varint(len(meta_body))
    out += meta_body + meta_trailer
    idx_body, idx_trailer = _block(index, restart_interval=1)
    idx_handle = _put_varint(len(out)) + _put_varint(len(idx_body))
    out += idx_body + idx_trailer
    footer = meta_handle + idx_handle
    out += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC)
    with open(tmp + ".index", "wb") as f:
        f.write(bytes(out))
    os.replace(tmp + ".data", prefix + ".data-00000-of-00001")
    os.replace(tmp + ".index", prefix + ".index")


# ------------------------------------------------------------------------------------------------- variable mapping
EMA_SUFFIX = "/ExponentialMovingAverage"
ADAM_M_SUFFIX, ADAM_V_SUFFIX = "/Adam", "/Adam_1"          # tf.train.AdamOptimizer slot variables
SLOT_SCOPES = ("", "train_op/")                              # slot variables may live under the train-op's name scope


def _find(ckpt_vars, name):
    for sc in SLOT_SCOPES:
        if sc + name in ckpt_vars:
            return sc + name
    return None


def map_slots(ckpt_vars, param_table):
    """Optimizer state of a MonoPSR training checkpoint: {name: (adam_m, adam_v, ema)} for the variables whose three slot
    tensors ('<var>/Adam', '<var>/Adam_1', '<var>/ExponentialMovingAverage', optionally below 'train_op/') are all
    present with the variable's shape.  tf.train.Saver restores them on resume (core/trainer.py:148-153)."""
    out = {}
    for name, shape, _ in param_table:
        keys = [_find(ckpt_vars, name + sfx) for sfx in (ADAM_M_SUFFIX, ADAM_V_SUFFIX, EMA_SUFFIX)]
        if all(k is not None and tuple(ckpt_vars[k].shape) == tuple(shape) for k in keys):
            out[name] = tuple(np.asarray(ckpt_vars[k], np.float32) for k in keys)
    return out


def map_monopsr_checkpoint(ckpt_vars, param_table, use_ema=False):
    """A MonoPSR training checkpoint -> engine parameter dict.  Returns (params, report): only variables that exist
    in the checkpoint WITH THE SAME SHAPE are taken (variables_helper.py:99-144); with use_ema the shadow
    'name/ExponentialMovingAverage' wins where present (evaluation restores the averaged weights)."""
    params, report = {}, {"loaded": [], "missing": [], "shape_mismatch": [], "unused": []}
    used = set()
    for name, shape, _ in param_table:
        ema_key = _find(ckpt_vars, name + EMA_SUFFIX) if use_ema else None
        src = ema_key if ema_key is not None else name
        if src not in ckpt_vars:
            report["missing"].append(name)
            continue
        if tuple(ckpt_vars[src].shape) != tuple(shape):
            report["shape_mismatch"].append(name)
            continue
        params[name] = np.asarray(ckpt_vars[src], np.float32)
        report["loaded"].append(name)
        used.add(src)
    report["unused"] = sorted(set(ckpt_vars) - used)
    return params, report


def map_detection_checkpoint(ckpt_vars, param_table, encoders=("FirstStageFeatureExtractor_crop",
                                                               "FirstStageFeatureExtractor_full")):
    """The pre-trained object-detection-API checkpoint (faster_rcnn_resnet101_kitti) initialises BOTH encoders:
    a variable 'FirstStageFeatureExtractor_crop/x' (and '..._full/x') is restored from 'FirstStageFeatureExtractor/x'
    when that exists with the same shape (core/checkpoint_utils.py:64-117)."""
    renamed = {}
    for name, shape, _ in param_table:
        for enc in encoders:
            if name.startswith(enc + "/"):
                src = "FirstStageFeatureExtractor/" + name[len(enc) + 1:]
                if src in ckpt_vars:
                    renamed[name] = ckpt_vars[src]
    return map_monopsr_checkpoint(renamed, [t for t in param_table if t[0] in renamed])


def load_checkpoint(prefix, param_table, kind="monopsr", use_ema=False, with_slots=False):
    """read `prefix` and map it; kind 'monopsr' (a MonoPSR training checkpoint) or 'detection' (the pre-trained
    object-detection-API ResNet-101).  The result feeds Engine.load_params (missing variables keep their values).
    with_slots (kind 'monopsr'): report['slots'] = map_slots(...) and report['global_step'] for a training resume."""
    want = None
    if kind == "monopsr":
        names = {n for n, _, _ in param_table}
        want = set(names)
        for sc in SLOT_SCOPES:
            want |= {sc + n + EMA_SUFFIX for n in names}
            if with_slots:
                want |= {sc + n + sfx for n in names for sfx in (ADAM_M_SUFFIX, ADAM_V_SUFFIX)}
        want.add("global_step")
    ckpt = read_bundle(prefix, names=want) if want is not None else read_bundle(prefix)
    if kind == "monopsr":
        params, report = map_monopsr_checkpoint(ckpt, param_table, use_ema)
        if with_slots:
            report["slots"] = map_slots(ckpt, param_table)
            report["global_step"] = int(ckpt["global_step"]) if "global_step" in ckpt else None
        return params, report
    if kind == "detection":
        return map_detection_checkpoint(ckpt, param_table)
    raise ValueError("Invalid checkpoint kind", kind)
