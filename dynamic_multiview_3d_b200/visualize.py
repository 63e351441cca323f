"""Evaluation outputs, SURVEY 8(f)-4: the image grids of ``save_images`` (dyn_mult_view/mv3d/utils/tf_utils.py:101-147)
and the ``visualize`` method of the models (appearance_flow_model.py:132-179) without matplotlib / scipy.misc.

``save_images`` lays a batch out exactly like the reference (row-major ``size = [rows, cols]`` mosaic, first
rows*cols images, ``rescale_image`` = (x / 1.5 + 0.5) * 255 for colour and ``rescale_dm`` = (x / 1.5 + 0.5) * 65535 as
16-bit grey for depth maps) and writes PNG through zlib.  The reference's quiver and correspondence PLOTS are
replaced by data products that carry the same information: the flow field as an HSV colour-wheel image and the
sampled-location table of the six random correspondence probes (appearance_flow_model.py:164-170).
Host-side NumPy on tensors copied back from the device; nothing here is on the measured path.
"""
import math
import os
import re
import struct
import zlib

import numpy as np


def rescale_image(image):
    """tf_utils.py:139-141"""
    return (image / 1.5 + 0.5) * 255


def rescale_dm(image):
    """tf_utils.py:144-146"""
    return (image / 1.5 + 0.5) * 65535


def mosaic(images, size):
    """tf_utils.py:101-119: [N,h,w(,3)] -> [h*rows, w*cols(,3)], image idx at row idx // cols, column idx % cols."""
    images = np.asarray(images)
    h, w = images.shape[1], images.shape[2]
    color = images.ndim == 4 and images.shape[3] == 3
    if images.ndim == 4 and not color:
        images = images[..., 0]
    img = np.zeros((h * size[0], w * size[1], 3) if color else (h * size[0], w * size[1]), np.float64)
    for idx, image in enumerate(images[:size[0] * size[1]]):
        i, j = idx % size[1], int(math.floor(idx / size[1]))
        img[j * h:j * h + h, i * w:i * w + w] = image
    return img


def write_png(path, arr):
    """uint8 [H,W,3] / [H,W] or uint16 [H,W] -> PNG (filter 0 scanlines, one zlib stream)."""
    arr = np.ascontiguousarray(arr)
    if arr.dtype == np.uint16:
        depth, ctype, raw = 16, 0, arr.astype(">u2").tobytes()
        row = arr.shape[1] * 2
    else:
        arr = arr.astype(np.uint8)
        depth, ctype = 8, (2 if arr.ndim == 3 else 0)
        raw, row = arr.tobytes(), arr.shape[1] * (3 if arr.ndim == 3 else 1)
    lines = b"".join(b"\x00" + raw[y * row:(y + 1) * row] for y in range(arr.shape[0]))

    def chunk(tag, data):
        c = struct.pack(">I", len(data)) + tag + data
        return c + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", arr.shape[1], arr.shape[0], depth, ctype, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(lines, 6)) + chunk(b"IEND", b""))


def save_images(images, size, image_path, color=True):
    """tf_utils.py:101-123 (scipy.misc.toimage(..., cmin=0, cmax=255|65535) clips to the range)."""
    img = mosaic(images, size)
    if color:
        write_png(image_path, np.clip(rescale_image(img), 0, 255).astype(np.uint8))
    else:
        write_png(image_path, np.clip(rescale_dm(img), 0, 65535).astype(np.uint16))


def flow_to_color(flow):
    """[h,w,2] displacement field -> uint8 RGB: hue = direction, saturation = magnitude / max magnitude."""
    fx, fy = flow[..., 0].astype(np.float64), flow[..., 1].astype(np.float64)
    mag = np.hypot(fx, fy)
    hsv_h = (np.arctan2(fy, fx) / (2 * np.pi)) % 1.0
    s = mag / max(float(mag.max()), 1e-12)
    i = np.floor(hsv_h * 6).astype(int) % 6
    f = hsv_h * 6 - np.floor(hsv_h * 6)
    p, q, t = 1 - s, 1 - s * f, 1 - s * (1 - f)
    one = np.ones_like(s)
    r = np.choose(i, [one, q, p, p, t, one])
    g = np.choose(i, [t, one, one, q, p, p])
    b = np.choose(i, [p, p, t, one, one, q])
    return (np.stack([r, g, b], -1) * 255).astype(np.uint8)


def visualize(model, image0, image1, disp, iter_num=None, seed=0):
    """appearance_flow_model.py:132-179 for a model with (image0, disp) -> gen.  Writes output_/tr_gt_/tr_input_<iter>.png
    (8x8 grids), flow_<iter>.png (sample 0; stands in for quiver_<iter>.pdf) and returns
    {'loss', 'max_resample_coord', 'correspondences': [(output_pt, sampled_location)] * 6} (the corr_plot data)."""
    import torch
    if iter_num is None:
        m = re.match(".*?([0-9]+)$", str(model.conf.get("visualize", "0")))
        iter_num = m.group(1) if m else "0"
    path = model.conf.get("output_dir", ".")
    os.makedirs(path, exist_ok=True)
    with torch.no_grad():
        out = model.forward(image0, disp)
        loss = float(model.build_loss(image1).detach())
    gen = out["gen"].detach().float().cpu().numpy() if "gen" in out else model.gen.detach().float().cpu().numpy()
    warp_pts = model.warp_pts.cpu().numpy()
    save_images(gen, [8, 8], os.path.join(path, "output_%s.png" % iter_num))
    save_images(image1.detach().cpu().numpy(), [8, 8], os.path.join(path, "tr_gt_%s.png" % iter_num))
    save_images(image0.detach().cpu().numpy(), [8, 8], os.path.join(path, "tr_input_%s.png" % iter_num))
    write_png(os.path.join(path, "flow_%s.png" % iter_num), flow_to_color(model.flow_field[0].detach().float().cpu().numpy()))
    H = gen.shape[1]
    rng = np.random.RandomState(seed)
    pts = rng.randint(int(0.3125 * H), int(0.6875 * H), size=(6, 2))           # randint(40, 88) at 128 (:166)
    corr = [(tuple(int(v) for v in p), tuple(int(v) for v in warp_pts[0, p[0], p[1], :].astype("uint32"))) for p in pts]
    return {"loss": loss, "max_resample_coord": float(np.max(warp_pts)), "correspondences": corr}
