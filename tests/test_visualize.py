"""CPU tests of the evaluation outputs: the mosaic layout and value scaling of save_images
(tf_utils.py:101-147) and a PNG that decodes back to the same pixels."""
import struct
import zlib

import numpy as np

from dynamic_multiview_3d_b200 import visualize as V


def _read_png(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, {}
    while pos < len(raw):
        n, tag = struct.unpack(">I", raw[pos:pos + 4])[0], raw[pos + 4:pos + 8]
        data = raw[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(tag + data) & 0xFFFFFFFF
        chunks.setdefault(tag, b"")
        chunks[tag] += data
        pos += 12 + n
    w, h, depth, ctype = struct.unpack(">IIBB", chunks[b"IHDR"][:10])
    px = zlib.decompress(chunks[b"IDAT"])
    bpp = (3 if ctype == 2 else 1) * depth // 8
    rows = [px[y * (w * bpp + 1) + 1:(y + 1) * (w * bpp + 1)] for y in range(h)]
    dt = ">u2" if depth == 16 else np.uint8
    a = np.frombuffer(b"".join(rows), dtype=dt).reshape((h, w, 3) if ctype == 2 else (h, w))
    return a


def test_mosaic_layout_matches_reference_indexing():
    imgs = np.stack([np.full((2, 3, 3), k, np.float32) for k in range(7)])
    m = V.mosaic(imgs, [2, 3])                       # 2 rows x 3 columns, first 6 images
    assert m.shape == (4, 9, 3)
    for idx in range(6):
        i, j = idx % 3, idx // 3                     # tf_utils.py:110-111
        assert (m[j * 2:(j + 1) * 2, i * 3:(i + 1) * 3] == idx).all()


def test_save_images_scaling_and_png_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    imgs = rng.random((4, 8, 8, 3)).astype(np.float32)
    p = str(tmp_path / "c.png")
    V.save_images(imgs, [2, 2], p)
    a = _read_png(p)
    exp = np.clip((V.mosaic(imgs, [2, 2]) / 1.5 + 0.5) * 255, 0, 255).astype(np.uint8)     # rescale_image, tf_utils.py:139
    assert np.array_equal(a, exp)
    d = rng.random((4, 8, 8, 1)).astype(np.float32)
    p = str(tmp_path / "d.png")
    V.save_images(d, [2, 2], p, color=False)
    a = _read_png(p)
    exp = np.clip((V.mosaic(d, [2, 2]) / 1.5 + 0.5) * 65535, 0, 65535).astype(np.uint16)   # rescale_dm, :144
    assert np.array_equal(a, exp)


def test_flow_color_wheel():
    f = np.zeros((2, 2, 2), np.float32)
    f[0, 0] = (1, 0); f[0, 1] = (0, 1); f[1, 0] = (-1, 0)
    c = V.flow_to_color(f)
    assert tuple(c[0, 0]) == (255, 0, 0)             # +x: hue 0 = red
    assert tuple(c[1, 1]) == (255, 255, 255)         # zero flow: unsaturated
    assert c[0, 1, 1] == 255 and c[1, 0, 2] == 255   # +y: hue 1/4 (green channel full), -x: hue 1/2 (cyan)


def test_quiver_and_correspondence_figures():
    """The reference's two matplotlib figures (appearance_flow_model.py:151-179), rasterised: arrows point along
    (U, V) with y up, probes are joined from the generated image to the sampled source location."""
    ii, jj = np.meshgrid(np.arange(32.0), np.arange(32.0), indexing="ij")
    q = V.quiver_image(np.stack([np.ones_like(ii), np.zeros_like(ii)], -1), cell=10, stride=4)      # U = 1, V = 0: arrows to the right
    ink = np.argwhere((q < 255).any(-1))
    assert q.dtype == np.uint8 and q.shape == (8 * 10 + 10, 8 * 10 + 10, 3) and len(ink) > 8 * 8 * 5
    rows = np.unique(ink[:, 0])
    assert len(rows) < 8 * 9                             # horizontal arrows: ink only on the shaft rows and the barbs next to them
    a, b = np.zeros((16, 16, 3)), np.ones((16, 16, 3))
    c = V.corr_plot_image(a, b, [(4, 5)], [(10, 2)])
    assert c.shape == (16, 16 * 2 + 8, 3)
    assert tuple(c[10, 2]) == (230, 30, 30) and tuple(c[4, 16 + 8 + 5]) == (230, 30, 30)          # both end points carry the probe's colour
    assert tuple(c[0, 0]) == (0, 0, 0) and tuple(c[0, 40 - 1]) == (255, 255, 255)


def test_panel_layout():
    p = V.panel([np.ones((4, 6, 3)), None, np.zeros((4, 6, 1))], 1, 3)
    assert p.shape == (4 + 8, 3 * (6 + 4) + 4, 3)
    assert tuple(p[5, 5]) == (255, 255, 255) and tuple(p[5, 4 + 2 * 10 + 1]) == (0, 0, 0)
