"""CPU tests of the TFRecord input pipeline: CRC-32C known answers, the tf.train.Example wire format,
write -> read round trips for both reference schemas, the train/val split and shuffling rules."""
import os
import struct

import numpy as np
import pytest

from dynamic_multiview_3d_b200 import read_tf_records as R


def test_crc32c_known_answers():
    # RFC 3720 appendix B.4 test vectors
    assert R.crc32c(b"") == 0
    assert R.crc32c(b"123456789") == 0xE3069283
    assert R.crc32c(bytes(32)) == 0x8A9136AA
    assert R.crc32c(bytes([0xFF] * 32)) == 0x62A8AB43
    assert R.crc32c(bytes(range(32))) == 0x46DD794E


def test_example_wire_format_golden_bytes():
    """A tf.train.Example with one bytes and one float[2] feature, hand-assembled from the protobuf wire rules."""
    img = bytes([1, 2, 3])
    flt = struct.pack("<2f", 0.5, -2.0)
    feat_b = b"\x0a" + bytes([len(img) + 2]) + b"\x0a" + bytes([len(img)]) + img                 # Feature{bytes_list{value}}
    feat_f = b"\x12" + bytes([len(flt) + 2]) + b"\x0a" + bytes([len(flt)]) + flt                 # Feature{float_list{packed}}
    e1 = b"\x0a\x01a" + b"\x12" + bytes([len(feat_b)]) + feat_b
    e2 = b"\x0a\x01d" + b"\x12" + bytes([len(feat_f)]) + feat_f
    feats = b"\x0a" + bytes([len(e1)]) + e1 + b"\x0a" + bytes([len(e2)]) + e2
    golden = b"\x0a" + bytes([len(feats)]) + feats
    assert R.encode_example({"a": img, "d": [0.5, -2.0]}) == golden
    ex = R.parse_example(golden)
    assert ex["a"] == img and np.array_equal(ex["d"], np.float32([0.5, -2.0]))
    # unpacked repeated floats (fixed32 wire type) parse to the same values
    unpacked = b"\x12" + bytes([10]) + b"\x0d" + flt[:4] + b"\x0d" + flt[4:]
    e3 = b"\x0a\x01d" + b"\x12" + bytes([len(unpacked)]) + unpacked
    f3 = b"\x0a" + bytes([len(e3)]) + e3
    assert np.array_equal(R.parse_example(b"\x0a" + bytes([len(f3)]) + f3)["d"], np.float32([0.5, -2.0]))


def _write(dirname, schema, n_files, per_file, side, rng):
    os.makedirs(dirname, exist_ok=True)
    truth = []
    for f in range(n_files):
        exs = []
        for _ in range(per_file):
            ex = {k: rng.integers(0, 256, size=side * side * c, dtype=np.uint8).tobytes() for k, c in schema.items()}
            ex["displacement"] = rng.standard_normal(2).astype(np.float32)
            exs.append(ex)
        R.write_tfrecord(os.path.join(dirname, "traj_%03d.tfrecords" % f), exs)
        truth.append(exs)
    return truth


@pytest.mark.parametrize("schema,builder", [(R.SINGLE, R.build_tfrecord_input), (R.MULTI, R.Build_tfrecord_input)])
def test_roundtrip_split_and_order(tmp_path, schema, builder):
    rng = np.random.default_rng(0)
    side = 16
    truth = _write(str(tmp_path / "d"), schema, n_files=4, per_file=3, side=side, rng=rng)
    conf = {"data_dir": str(tmp_path / "d"), "train_val_split": 0.75, "batch_size": 3, "image_size": side, "test_mode": ""}
    inp = builder(conf, check_crc=True)                     # test_mode: every file, in order
    for f in range(4):
        b = inp.next_batch()
        for k, c in schema.items():
            assert b[k].shape == (3, side, side, c) and b[k].dtype == np.uint8
            assert all(b[k][i].tobytes() == truth[f][i][k] for i in range(3))
        assert np.array_equal(b["displacement"], np.stack([truth[f][i]["displacement"] for i in range(3)]))
    fb = builder(conf).float_batch()
    assert fb["image0"].dtype == np.float32 and 0.0 <= fb["image0"].min() and fb["image0"].max() <= 1.0
    del conf["test_mode"]
    tr, va = builder(conf, training=True), builder(conf, training=False)
    assert len(tr.files) == 3 and len(va.files) == 1 and tr.shuffle       # floor(0.75 * 4) = 3 (read_tf_records.py:31-35)
    assert set(tr.files).isdisjoint(va.files)
    seen = {tr.next_batch()["image0"][0].tobytes() for _ in range(12)}
    assert len(seen) == 3                                                  # only the three training files are visited


def test_errors(tmp_path):
    with pytest.raises(RuntimeError):
        R.build_tfrecord_input({"data_dir": str(tmp_path / "none"), "train_val_split": 0.9, "batch_size": 1})
    rng = np.random.default_rng(1)
    _write(str(tmp_path / "d"), R.SINGLE, 1, 1, 8, rng)
    p = str(tmp_path / "d" / "traj_000.tfrecords")
    raw = bytearray(open(p, "rb").read())
    raw[20] ^= 0xFF
    open(p, "wb").write(bytes(raw))
    conf = {"data_dir": str(tmp_path / "d"), "train_val_split": 1.0, "batch_size": 1, "image_size": 8, "test_mode": ""}
    with pytest.raises(IOError):
        R.build_tfrecord_input(conf, check_crc=True).next_batch()
    with pytest.raises(ValueError):
        R.build_tfrecord_input(dict(conf, image_size=16)).next_batch()
