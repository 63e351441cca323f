"""Input pipeline, SURVEY 8(f)-2: the reference's two TFRecord schemas read WITHOUT TensorFlow.

Restates dyn_mult_view/multi_view_model/utils/read_tf_records.py:15-112 (``build_tfrecord_input``: features
image0/image1/depth0/depth1 as raw uint8 bytes + displacement float[2]) and read_tf_records_multobj.py:20-140
(``Build_tfrecord_input``: the twelve image tensors of a two-object scene + displacement); the writers are
collect_data_node.py:126-133 and render_multiobj.py:554-569.

File format (public TensorFlow formats, restated): a TFRecord file is a sequence of
    uint64 length | uint32 masked_crc32c(length) | byte data[length] | uint32 masked_crc32c(data)
(little endian; masked = ((crc >> 15 | crc << 17) + 0xa282ead8) mod 2^32), and each record is a serialised
``tf.train.Example``: Example{1: Features{1: repeated map entry{1: key, 2: Feature}}}, Feature = oneof
{1: BytesList{1: repeated bytes}, 2: FloatList{1: packed float}, 3: Int64List{1: packed varint}}.

Semantics kept from the reference: files = sorted glob of conf['data_dir'] split at floor(train_val_split * n)
(:27-36), shuffling unless 'test_mode' (:38-45), images uint8 / 255 as float32 NHWC (:88-111; the crop and bicubic
resize there are identities at the native size and are not re-implemented: another size raises), batches of
conf['batch_size'].  Host-side NumPy: input decoding is not part of the measured GPU path; the uint8 -> float
conversion happens after the (4x smaller) uint8 batch has been copied to the device when ``device`` is given.
"""
import glob
import os
import struct

import numpy as np

# ----------------------------------------------------------------------------- crc32c (Castagnoli), table driven
_POLY = 0x82F63B78
_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ (_POLY if _c & 1 else 0)
    _TABLE.append(_c)
_TABLE = np.array(_TABLE, dtype=np.uint32)


def crc32c(data):
    crc = 0xFFFFFFFF
    tab = _TABLE
    for b in bytes(data):
        crc = int(tab[(crc ^ b) & 0xFF]) ^ (crc >> 8)
    return crc ^ 0xFFFFFFFF


def masked_crc(data):
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ----------------------------------------------------------------------------- protobuf wire format (subset)
def _varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _fields(buf):
    """Yield (field_number, wire_type, value) of one message; length-delimited values are memoryviews."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fn, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = bytes(buf[pos:pos + 4]); pos += 4
        elif wt == 1:
            v = bytes(buf[pos:pos + 8]); pos += 8
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield fn, wt, v


def parse_example(record):
    """Serialised tf.train.Example -> {name: bytes | np.float32 array | np.int64 array}."""
    out = {}
    buf = memoryview(record)
    for fn, _, features in _fields(buf):
        if fn != 1:
            continue
        for fn2, _, entry in _fields(features):
            if fn2 != 1:
                continue
            key, feat = None, None
            for fn3, _, v in _fields(entry):
                if fn3 == 1:
                    key = bytes(v).decode()
                elif fn3 == 2:
                    feat = v
            for kind, _, lst in _fields(feat):
                if kind == 1:        # BytesList
                    vals = [bytes(v) for f, _, v in _fields(lst) if f == 1]
                    out[key] = vals[0] if len(vals) == 1 else vals
                elif kind == 2:      # FloatList: packed (wire type 2) or repeated fixed32
                    vals = []
                    for f, wt, v in _fields(lst):
                        if f == 1:
                            vals.append(np.frombuffer(bytes(v), dtype="<f4"))
                    out[key] = np.concatenate(vals) if vals else np.zeros(0, np.float32)
                elif kind == 3:      # Int64List
                    vals = []
                    for f, wt, v in _fields(lst):
                        if f != 1:
                            continue
                        if wt == 0:
                            vals.append(v)
                        else:
                            p, b = 0, v
                            while p < len(b):
                                x, p = _varint(b, p)
                                vals.append(x)
                    out[key] = np.array(vals, dtype=np.int64)
    return out


def _enc_varint(x):
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        out.append(b | (0x80 if x else 0))
        if not x:
            return bytes(out)


def _ld(fn, payload):
    return _enc_varint((fn << 3) | 2) + _enc_varint(len(payload)) + payload


def encode_example(features):
    """{name: bytes | float sequence} -> serialised tf.train.Example (what the reference's writers emit)."""
    entries = b""
    for k in sorted(features):
        v = features[k]
        if isinstance(v, (bytes, bytearray)):
            feat = _ld(1, _ld(1, bytes(v)))
        else:
            feat = _ld(2, _ld(1, np.asarray(v, dtype="<f4").tobytes()))
        entries += _ld(1, _ld(1, k.encode()) + _ld(2, feat))
    return _ld(1, entries)


def write_tfrecord(path, examples):
    with open(path, "wb") as f:
        for ex in examples:
            data = encode_example(ex)
            hdr = struct.pack("<Q", len(data))
            f.write(hdr + struct.pack("<I", masked_crc(hdr)) + data + struct.pack("<I", masked_crc(data)))


def read_tfrecord(path, check_crc=True):
    with open(path, "rb") as f:
        while True:
            hdr = f.read(8)
            if not hdr:
                return
            if len(hdr) < 8:
                raise IOError("truncated TFRecord header in %s" % path)
            (n,) = struct.unpack("<Q", hdr)
            (hcrc,) = struct.unpack("<I", f.read(4))
            data = f.read(n)
            (dcrc,) = struct.unpack("<I", f.read(4))
            if len(data) < n:
                raise IOError("truncated TFRecord in %s" % path)
            if check_crc and (hcrc != masked_crc(hdr) or dcrc != masked_crc(data)):
                raise IOError("corrupt TFRecord (crc mismatch) in %s" % path)
            yield data


# ----------------------------------------------------------------------------- the two schemas
SINGLE = {"image0": 3, "image1": 3, "depth0": 1, "depth1": 1}                       # read_tf_records.py:50-63
MULTI = {"image0": 3, "image0_mask0": 1, "image0_mask1": 1, "image1": 3, "image1_only0": 3, "image1_only1": 3,
         "image1_mask0": 1, "image1_mask1": 1, "depth0": 1, "depth1": 1, "depth1_only0": 1, "depth1_only1": 1}   # _multobj.py:51-80


def _filenames(conf, training):
    names = sorted(glob.glob(os.path.join(conf["data_dir"], "*")))
    if not names:
        raise RuntimeError("No data_files files found.")         # read_tf_records.py:28-29
    if "test_mode" in conf:
        return names, False
    index = int(np.floor(conf["train_val_split"] * len(names)))
    return (names[:index] if training else names[index:]), True


class TFRecordInput(object):
    """Iterator over batches: dict name -> uint8 NHWC array (images) / float32 [B,2] (displacement).
    ``float_batch`` converts to the reference's float32 / 255 (:111), on the device when one is given."""

    def __init__(self, conf, training=True, schema=None, seed=0, check_crc=False):
        self.conf = conf
        self.schema = dict(schema or SINGLE)
        self.batch_size = int(conf["batch_size"])
        self.files, self.shuffle = _filenames(conf, training)
        if not self.files:
            raise RuntimeError("train_val_split leaves no files for this split")
        self.rng = np.random.default_rng(seed)
        self.size = int(conf.get("image_size", 128))
        self.check_crc = check_crc
        self._gen = self._examples()

    def _examples(self):
        while True:                                              # string_input_producer: endless epochs
            order = list(self.files)
            if self.shuffle:
                self.rng.shuffle(order)
            for path in order:
                for rec in read_tfrecord(path, self.check_crc):
                    yield parse_example(rec)

    def _image(self, ex, name, chan):
        """tf.decode_raw + reshape to the STORED size [ORIGINAL_HEIGHT, ORIGINAL_WIDTH, C] (read_tf_records.py:92-102; the
        reference hard-codes 128 x 128, here conf['original_height'/'original_width'] or, for square records, the byte
        count decides).  The crop + bicubic resize to the model's size and the / 255 (:103-111) run on the device
        (dmv_u8_crop_resize_bicubic, one kernel) so that uint8 pixels are what crosses PCIe."""
        raw = np.frombuffer(ex[name], dtype=np.uint8)
        h0, w0 = self.conf.get("original_height"), self.conf.get("original_width")
        if h0 is None or w0 is None:
            side = int(round(np.sqrt(raw.size // chan)))
            h0 = w0 = side
        if raw.size != h0 * w0 * chan:
            raise ValueError("%s holds %d bytes, which is not %dx%dx%d; set conf['original_height'/'original_width']" % (name, raw.size, h0, w0, chan))
        return raw.reshape(h0, w0, chan)

    def next_batch(self):
        cols = {k: [] for k in self.schema}
        disp = []
        for _ in range(self.batch_size):
            ex = next(self._gen)
            for k, c in self.schema.items():
                cols[k].append(self._image(ex, k, c))
            disp.append(np.asarray(ex["displacement"], dtype=np.float32).reshape(2))
        out = {k: np.stack(v, 0) for k, v in cols.items()}
        out["displacement"] = np.stack(disp, 0)
        return out

    def float_batch(self, device):
        """The reader's output as the reference's graph sees it: float32 images in [0,1] at the model's size
        (read_tf_records.py:103-111: central crop, bicubic resize, / 255 -- one CUDA kernel), displacement float32."""
        import torch
        from . import functional as F
        out = {}
        for k, v in self.next_batch().items():
            t = torch.from_numpy(v).to(device, non_blocking=True)
            out[k] = F.prepare_images(t, self.size) if v.dtype == np.uint8 else t
        return out

    __next__ = next_batch

    def __iter__(self):
        return self


def build_tfrecord_input(conf, training=True, **kw):
    """read_tf_records.py:15 -- single-object schema (image0, image1, depth0, depth1, displacement)."""
    return TFRecordInput(conf, training, SINGLE, **kw)


def Build_tfrecord_input(conf, training=True, **kw):
    """read_tf_records_multobj.py:20 -- two-object schema (twelve image tensors + displacement)."""
    return TFRecordInput(conf, training, MULTI, **kw)
