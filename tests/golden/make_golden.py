"""Generates tests/golden/*.npz.  Run in the authoring container (needs /root/reference for the
rectangle fixture; nothing else reads the reference).  The GPU box only reads the .npz files.

  sampler_kat.npz     hand-derived known answers of SURVEY.md 8(c) + oracle outputs on seeded
                      inputs (forward, indices, masks, both gradients) for C in {1,3,4}
  rectangle.npz       the reference's only fixture for this path
                      (dyn_mult_view/multi_view_model/tests/rectangle.png, restated from
                      test_resampler.py:20-44): rectangle bounds, SHA-256 of the oracle's uint8
                      output, and the rotated rectangle's principal-axis angle
  layers.npz          oracle conv / deconv / linear outputs on seeded bf16-representable inputs
  graphs.npz          oracle outputs of the whole model graphs at 32x32 (single-view app-flow, colour+depth,
                      multi-object, 3-view fusion) on seeded inputs and seed-0 parameters: flow / generated images
                      (float16-rounded to keep the file small) and the losses -- a regression pin of oracle/graph.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import tf_ops as T  # noqa: E402


def bf16_round(x):
    """round-to-nearest-even to bfloat16, returned as float32"""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) & 0xFFFF0000).view(np.float32)


def rectangle_case():
    from PIL import Image
    png = "/root/reference/dyn_mult_view/multi_view_model/tests/rectangle.png"
    im = np.array(Image.open(png))
    dark = im[..., 1] == 0
    ys, xs = np.nonzero(dark)
    y0, y1, x0, x1 = ys.min(), ys.max(), xs.min(), xs.max()
    assert dark.sum() == (y1 - y0 + 1) * (x1 - x0 + 1)
    fg, bg = im[y0, x0], im[0, 0]
    out, angle = rectangle_run(im.shape[0], im.shape[1], (y0, y1, x0, x1), fg, bg)
    np.savez(os.path.join(HERE, "rectangle.npz"), shape=np.array(im.shape), bounds=np.array([y0, y1, x0, x1]),
             fg=fg, bg=bg, sha256=np.frombuffer(hashlib.sha256(out.tobytes()).digest(), np.uint8), angle_deg=angle,
             rect_count=np.array(np.all(out == fg, axis=-1).sum()))
    print("rectangle: angle", angle)


def rectangle_image(H, W, bounds, fg, bg):
    y0, y1, x0, x1 = bounds
    im = np.empty((H, W, 3), np.uint8)
    im[:] = bg
    im[y0:y1 + 1, x0:x1 + 1] = fg
    return im


def rectangle_warp(H, W, angle=10):
    """test_resampler.py:20-40 restated."""
    rads = np.radians(angle)
    c, s = np.cos(rads), np.sin(rads)
    rot = np.array([[c, -s], [s, c]], np.float32)
    y = np.arange(H, dtype=np.float32)
    x = np.arange(W, dtype=np.float32)
    Y, X = np.meshgrid(y, x, indexing="ij")
    pts = np.stack([X.reshape(-1), Y.reshape(-1)], axis=1)
    wp = (pts @ rot).astype(np.float32)
    wx = np.clip(wp[:, 0].reshape(H, W), 0, W)
    wy = np.clip(wp[:, 1].reshape(H, W), 0, H)
    return np.stack([wx, wy], axis=2)[None].astype(np.float32)


def principal_angle(mask):
    ys, xs = np.nonzero(mask)
    xs = xs - xs.mean()
    ys = ys - ys.mean()
    cov = np.cov(np.stack([xs, ys]))
    w, v = np.linalg.eigh(cov)
    major = v[:, np.argmax(w)]
    return float(np.degrees(np.arctan2(major[1], major[0])))


def rectangle_run(H, W, bounds, fg, bg):
    im = rectangle_image(H, W, bounds, fg, bg)
    out = T.resampler(im[None].astype(np.float32), rectangle_warp(H, W))
    out8 = out.astype(np.uint8)[0]          # tf.cast(float -> uint8) truncates
    # pixels that kept the rectangle's exact colour (zero-filled out-of-range samples are (0,0,0))
    rect = np.all(out8 == np.asarray(fg, np.uint8), axis=-1)
    return out8, principal_angle(rect) % 180.0


def sampler_cases():
    d = {}
    img = np.array([[1, 2, 3], [4, 5, 6]], np.float32).reshape(1, 2, 3, 1)
    pts = np.array([[0, 0], [2, 1], [0.5, 0.5], [1.25, 0.75], [-0.5, 0], [-1, 0], [2.5, 1], [3, 1], [2, 1.5], [2, 2],
                    [-0.999, -0.999]], np.float32)
    d["kat_img"], d["kat_pts"] = img, pts
    d["kat_expect"] = np.array([1, 6, 3, 4.5, 0.5, 0, 3, 0, 3, 0, np.nan], np.float32)  # last: ~1e-6, checked loosely
    rng = np.random.default_rng(7)
    for C in (1, 3, 4):
        B, H, W = 2, 9, 13
        data = rng.random((B, H, W, C), dtype=np.float32)
        warp = rng.uniform(-2, [W + 1, H + 1], size=(B, 7, 11, 2)).astype(np.float32)
        bset = np.array([-1, -1 + 1e-6, -0.5, 0, W - 1, W - 0.5, W - 1e-4, W], np.float32)
        warp[0, 0, :8, 0] = bset
        warp[0, 1, :8, 1] = np.array([-1, -1 + 1e-6, -0.5, 0, H - 1, H - 0.5, H - 1e-4, H], np.float32)
        warp[1, 0, 0] = np.nan
        go = rng.standard_normal((B, 7, 11, C)).astype(np.float32)
        out = T.resampler(data, warp)
        fx, fy, cx, cy, mask = T.resampler_indices(data.shape, warp)
        gd, gw = T.resampler_grad(data, warp, go)
        d.update({"c%d_data" % C: data, "c%d_warp" % C: warp, "c%d_go" % C: go, "c%d_out" % C: out,
                  "c%d_idx" % C: np.stack([fx, fy, cx, cy], -1), "c%d_mask" % C: mask, "c%d_gd" % C: gd, "c%d_gw" % C: gw})
    # quirk KAT: zero flow through coords() -> transposed image
    ramp = np.arange(16, dtype=np.float32).reshape(1, 4, 4, 1)
    d["quirk_img"] = ramp
    d["quirk_out"] = T.resampler(ramp, T.warp_pts_layer(np.zeros((1, 4, 4, 2), np.float32)))
    np.savez(os.path.join(HERE, "sampler_kat.npz"), **d)


def layer_cases():
    rng = np.random.default_rng(11)
    d = {}
    cfgs = [("k5s2", 5, 2, 12, 3, 8), ("k5s1", 5, 1, 10, 8, 8), ("k3s2", 3, 2, 14, 16, 8), ("k3s1", 3, 1, 7, 8, 16)]
    for name, k, s, H, cin, cout in cfgs:
        x = bf16_round(rng.standard_normal((2, H, H, cin)).astype(np.float32))
        w = bf16_round((rng.standard_normal((k, k, cin, cout)) * T.conv_stddev(k, k, cin)).astype(np.float32))
        b = rng.standard_normal(cout).astype(np.float32)
        y = T.conv2d_same(x, w, b, s, s)
        gy = bf16_round(rng.standard_normal(y.shape).astype(np.float32))
        gx, gw, gb = T.conv2d_same_grads(x, w, gy, s, s)
        d.update({"conv_%s_x" % name: x, "conv_%s_w" % name: w, "conv_%s_b" % name: b, "conv_%s_y" % name: y,
                  "conv_%s_gy" % name: gy, "conv_%s_gx" % name: gx, "conv_%s_gw" % name: gw, "conv_%s_gb" % name: gb})
        # deconv: small side [2,h,h,cin] -> [2,H,H,cout]
        h = -(-H // s)
        xd = bf16_round(rng.standard_normal((2, h, h, cin)).astype(np.float32))
        wd = bf16_round((rng.standard_normal((k, k, cout, cin)) * T.deconv_stddev(k, k, cin, s, s)).astype(np.float32))
        yd = T.conv2d_transpose_same(xd, wd, (2, H, H, cout), s, s)
        gyd = bf16_round(rng.standard_normal(yd.shape).astype(np.float32))
        gxd, gwd = T.conv2d_transpose_same_grads(xd, wd, gyd, s, s)
        d.update({"deconv_%s_x" % name: xd, "deconv_%s_w" % name: wd, "deconv_%s_y" % name: yd, "deconv_%s_gy" % name: gyd,
                  "deconv_%s_gx" % name: gxd, "deconv_%s_gw" % name: gwd})
    x = bf16_round(rng.standard_normal((5, 40)).astype(np.float32))
    m = bf16_round((rng.standard_normal((40, 24)) * T.linear_stddev(40)).astype(np.float32))
    b = rng.standard_normal(24).astype(np.float32)
    y = T.linear(x, m, b)
    gy = bf16_round(rng.standard_normal(y.shape).astype(np.float32))
    gx, gm, gb = T.linear_grads(x, m, gy)
    d.update(lin_x=x, lin_m=m, lin_b=b, lin_y=y, lin_gy=gy, lin_gx=gx, lin_gm=gm, lin_gb=gb)
    np.savez(os.path.join(HERE, "layers.npz"), **d)


def graph_inputs(H=32, B=1, V=2, Vw=3):
    """Seeded inputs shared by graph_cases() and tests/test_oracle.py::test_graph_golden_vectors."""
    rng = np.random.default_rng(77)
    names = ["image0", "image0_mask0", "image0_mask1", "image1", "image1_only0", "image1_only1", "image1_mask0", "image1_mask1",
             "depth0", "depth1", "depth1_only0", "depth1_only1"]
    mo = {n: rng.random((B, H, H, 3 if (n.startswith("image") and "mask" not in n) else 1), dtype=np.float32) for n in names}
    mo["displacement"] = rng.standard_normal((B, V)).astype(np.float32)
    views = {"image0": rng.random((Vw, B, H, H, 3), dtype=np.float32), "depth0": rng.random((Vw, B, H, H, 1), dtype=np.float32),
             "image0_mask0": rng.random((Vw, B, H, H, 1), dtype=np.float32), "image0_mask1": rng.random((Vw, B, H, H, 1), dtype=np.float32),
             "displacement": rng.standard_normal((Vw, B, V)).astype(np.float32), "image1": mo["image1"]}
    return mo, views


GRAPH_CONFS = {
    "colordepth": {"use_color": "", "use_depth": "", "depth_lr_factor": 0.1},
    "multiobject": {"use_color": "", "use_depth": 0.1, "combination_image": "", "gen_sep_images": "", "predict_target_masks": 0.1,
                    "masked_image_loss": ""},
    "multiobject_fc": {"use_color": "", "combination_image": "", "fully_conv": ""},
}


def graph_outputs():
    from oracle import graph as G
    H, B, V = 32, 1, 2
    mo, views = graph_inputs(H, B, V)
    ops = G.NumpyOps()
    out = {}
    P = G.init_params(G.appflow_param_shapes(H, V, "base"), 0)
    o = G.appearance_flow_forward(ops, P, mo["image0"], mo["displacement"], "base")
    out["appflow_flow"], out["appflow_gen"] = o["flow_field"], o["gen"]
    out["appflow_loss"] = np.float64(G.appearance_flow_loss(ops, o, mo["image1"]))
    c = GRAPH_CONFS["colordepth"]
    P = G.init_params(G.colordepth_param_shapes(H, V, c), 0)
    o = G.colordepth_forward(ops, P, c, mo["image0"], mo["depth0"], mo["displacement"])
    out["colordepth_gen_image1"], out["colordepth_gen_dimage1"] = o["gen_image1"], o["gen_dimage1"]
    out["colordepth_loss"] = np.float64(G.colordepth_loss(ops, o, c, mo["image1"], mo["depth1"]))
    for key in ("multiobject", "multiobject_fc"):
        c = GRAPH_CONFS[key]
        P = G.init_params(G.multiobject_param_shapes(H, V, c), 0)
        o = G.multiobject_forward(ops, P, c, mo)
        for k, v in o.items():
            out["%s_%s" % (key, k)] = v
        out[key + "_loss"] = np.float64(G.multiobject_loss(ops, o, c, mo))
    c = {"use_depth": 0.1}
    P = G.init_params(G.multiview_param_shapes(H, V, c), 0)
    o = G.multiview_forward(ops, P, c, views)
    out["multiview_fused"], out["multiview_logits"] = o["fused"], o["logits"]
    out["multiview_loss"] = np.float64(G.multiview_loss(ops, o, views["image1"]))
    return out


def graph_cases():
    out = graph_outputs()
    np.savez_compressed(os.path.join(HERE, "graphs.npz"),
                        **{k: (v.astype(np.float16) if getattr(v, "ndim", 0) > 0 else v) for k, v in out.items()})
    print("graphs:", len(out), "arrays")


if __name__ == "__main__":
    sampler_cases()
    layer_cases()
    graph_cases()
    rectangle_case()
