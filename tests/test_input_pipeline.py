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
    # records of another size than the model's are read at their STORED size (crop + bicubic resize happen on the device)
    stored = builder(dict(conf, image_size=2 * side)).next_batch()
    assert stored["image0"].shape == (3, side, side, schema["image0"])
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
    with pytest.raises(ValueError):          # the byte count contradicts the stated stored size
        R.build_tfrecord_input(dict(conf, original_height=8, original_width=12)).next_batch()


def test_image_preparation_oracle_identity_and_direct_kernel_evaluation():
    """oracle.tf_ops.process_image (read_tf_records.py:100-111; TF-1.3 resize_bicubic [TF-upstream, recalled]):
    at the reference's own setting (stored size == model size) it is exactly uint8 / 255; a resize equals the direct
    evaluation of the Keys kernel (A = -0.75) at the table-quantised legacy coordinates; constants stay constant."""
    from oracle import tf_ops as T
    rng = np.random.default_rng(3)
    x = rng.integers(0, 256, (2, 16, 16, 3), dtype=np.uint8)
    assert np.array_equal(T.process_image(x, 16), x.astype(np.float32) / np.float32(255.0))
    wide = rng.integers(0, 256, (1, 12, 20, 1), dtype=np.uint8)           # central crop to 12 x 12, columns 4..15
    assert np.array_equal(T.process_image(wide, 12), wide[:, :, 4:16].astype(np.float32) / np.float32(255.0))
    assert np.abs(T.process_image(np.full((1, 12, 12, 1), 200, np.uint8), 21) - 200 / 255).max() < 1e-6

    def keys(d, A=-0.75):
        d = abs(d)
        return (A + 2) * d ** 3 - (A + 3) * d ** 2 + 1 if d <= 1 else (A * d ** 3 - 5 * A * d ** 2 + 8 * A * d - 4 * A if d < 2 else 0.0)
    H, S = 11, 17
    img = rng.integers(0, 256, (1, H, H, 1), dtype=np.uint8)
    y = T.resize_bicubic_tf13(img, S, S)[0, :, :, 0]
    sc = np.float32(H) / np.float32(S)
    for i in (0, 5, 16):
        for j in (0, 7, 16):
            py, px = float(np.float32(sc * np.float32(i))), float(np.float32(sc * np.float32(j)))
            iy, ix = int(py), int(px)
            fy, fx = round((py - iy) * 1024) / 1024, round((px - ix) * 1024) / 1024
            acc = sum(float(img[0, min(H - 1, max(0, iy + a)), min(H - 1, max(0, ix + b)), 0]) * keys(a - fy) * keys(b - fx)
                      for a in range(-1, 3) for b in range(-1, 3))
            assert abs(y[i, j] - acc) < 1e-3, (i, j)


@pytest.mark.gpu
@pytest.mark.parametrize("C,H0,W0,S", [(3, 128, 128, 224), (1, 128, 128, 224), (3, 96, 130, 64), (4, 40, 40, 40), (3, 300, 260, 224)])
def test_image_preparation_kernel_bit_equal_to_oracle(C, H0, W0, S):
    """dmv_u8_crop_resize_bicubic == oracle.tf_ops.process_image, bit for bit (up- and down-scaling, non-square records)."""
    import torch
    from oracle import tf_ops as T
    from dynamic_multiview_3d_b200 import functional as F
    rng = np.random.default_rng(C + H0 + S)
    x = rng.integers(0, 256, (3, H0, W0, C), dtype=np.uint8)
    out = F.prepare_images(torch.from_numpy(x).cuda(), S).cpu().numpy()
    assert np.array_equal(out, T.process_image(x, S))
    # leading axes (source frames x batch) are one batch for the kernel
    x5 = rng.integers(0, 256, (2, 2, H0, W0, C), dtype=np.uint8)
    o5 = F.prepare_images(torch.from_numpy(x5).cuda(), S).cpu().numpy()
    assert o5.shape == (2, 2, S, S, C) and np.array_equal(o5[1], T.process_image(x5[1], S))


@pytest.mark.gpu
def test_graphed_step_resizes_stored_records_on_the_device():
    """128 x 128 uint8 records into a 64 x 64 model: the captured step's input copy does the reader's crop + bicubic
    resize + / 255 (read_tf_records.py:103-111) in one kernel and lands on the oracle's pixels."""
    import torch
    import dynamic_multiview_3d_b200 as pkg
    from oracle import tf_ops as T
    from dynamic_multiview_3d_b200.train import GraphedTrainStep
    rng = np.random.default_rng(0)
    B, S = 2, 64
    model = pkg.AppearanceFlowModel({"batch_size": B, "learning_rate": 1e-4, "image_size": S, "viewpoint_dim": 2})
    raw = {k: rng.integers(0, 256, (B, 128, 128, 3), dtype=np.uint8) for k in ("image0", "image1")}
    batch = {k: torch.from_numpy(v).pin_memory() for k, v in raw.items()}
    batch["disp"] = torch.from_numpy(rng.standard_normal((B, 2)).astype(np.float32)).pin_memory()
    step = GraphedTrainStep(model, warmup=1)
    loss = float(step(batch))
    assert np.isfinite(loss)
    assert np.array_equal(step.static["image0"].cpu().numpy(), T.process_image(raw["image0"], S))
