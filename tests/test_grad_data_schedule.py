"""CPU model of the schedule of sampler_grad_data_kernel (dynamic_multiview_3d_b200/csrc/sampler.cu) against the oracle.

The kernel computes the resampler's gradient wrt the source by an owner-computes scatter: a pre-pass writes the tap box of
every 32x32 output tile and of each of its 32 warp chunks (8x4 patches; runs of 32 samples for 1-D lists); a CTA that owns a
32x16 SOURCE tile visits only the chunks whose box intersects it.  The culling rule is the part of that kernel that could
silently lose contributions, and it is pure index logic -- so it is restated here in NumPy, step by step as the kernel does
it (boxes from floor coordinates clipped to the image, two-level intersection test, hits dealt round-robin to eight
warp-private accumulators, lanes sharing a floor cell taking turns, warps summed in order), and held against
oracle.tf_ops.resampler_grad on flows that stress it: jitter, large displacements, samples outside the image, NaN,
minification (many outputs per source cell) and ragged sizes.  The GPU tests (tests/test_sampler_gpu.py) check the CUDA
code itself against the same oracle."""
import numpy as np
import pytest

from oracle import tf_ops

TILE = 32                      # output tile side (images)
SRC_W, SRC_H = 32, 16          # source tile owned by one CTA
WARPS = 8
BIG = 0x7FFFFFFF


def chunk_pixels(q, tw_shift):
    """(di, dj) of the 32 lanes of chunk q inside its tile: chunk_pixel() of sampler.cu."""
    lane = np.arange(32)
    if tw_shift == 5:
        return (q >> 2) * 4 + (lane >> 3), (q & 3) * 8 + (lane & 7)
    p = q * 32 + lane
    return p >> tw_shift, p & ((1 << tw_shift) - 1)


def model_grad_data(warp, go, H, W):
    """warp [Ho, Wo, 2] absolute (x, y) float32, go [Ho, Wo, C] float32 -> grad_data [H, W, C] float32 (one image)."""
    Ho, Wo, C = go.shape
    tw_shift = 10 if Ho == 1 else 5
    th, tw = 1024 >> tw_shift, 1 << tw_shift
    tiles_y, tiles_x = -(-Ho // th), -(-Wo // tw)
    x, y = warp[..., 0], warp[..., 1]
    with np.errstate(invalid="ignore"):
        valid = (x > -1) & (y > -1) & (x < W) & (y < H)
    fx = np.where(valid, np.floor(np.where(valid, x, 0)), 0).astype(np.int64)
    fy = np.where(valid, np.floor(np.where(valid, y, 0)), 0).astype(np.int64)

    # ---- pre-pass: sampler_box_kernel
    n_tiles = tiles_y * tiles_x
    cbox = np.empty((n_tiles, 32, 4), np.int64)
    tbox = np.empty((n_tiles, 4), np.int64)
    pix = {}
    for t in range(n_tiles):
        ty, tx = divmod(t, tiles_x)
        for q in range(32):
            di, dj = chunk_pixels(q, tw_shift)
            i, j = ty * th + di, tx * tw + dj
            inb = (i < Ho) & (j < Wo)
            ii, jj = np.where(inb, i, 0), np.where(inb, j, 0)
            v = inb & valid[ii, jj]
            pix[t, q] = (ii, jj, v)
            if v.any():
                cbox[t, q] = (np.maximum(fx[ii, jj][v], 0).min(), np.minimum(fx[ii, jj][v] + 1, W - 1).max(),
                              np.maximum(fy[ii, jj][v], 0).min(), np.minimum(fy[ii, jj][v] + 1, H - 1).max())
            else:
                cbox[t, q] = (BIG, -BIG, BIG, -BIG)
        tbox[t] = (cbox[t, :, 0].min(), cbox[t, :, 1].max(), cbox[t, :, 2].min(), cbox[t, :, 3].max())

    out = np.zeros((H, W, C), np.float32)
    visits = 0
    for sy0 in range(0, H, SRC_H):
        for sx0 in range(0, W, SRC_W):
            sx1, sy1 = min(sx0 + SRC_W, W) - 1, min(sy0 + SRC_H, H) - 1

            def misses(b):
                return b[0] > sx1 or b[1] < sx0 or b[2] > sy1 or b[3] < sy0

            acc = np.zeros((WARPS, SRC_H, SRC_W, C), np.float32)
            listed = [t for t in range(n_tiles) if not misses(tbox[t])]                  # (A), in tile order
            for cb in range(0, len(listed), WARPS):                                      # (B): eight tiles per batch
                entries = [(t, q) for t in listed[cb:cb + WARPS] for q in range(32) if not misses(cbox[t, q])]
                for e, (t, q) in enumerate(entries):                                     # (C): entry e -> warp e % 8
                    w = e % WARPS
                    visits += 1
                    ii, jj, v = pix[t, q]
                    lfx, lfy = fx[ii, jj], fy[ii, jj]
                    sxx = np.where(v, x[ii, jj], 0).astype(np.float32)
                    syy = np.where(v, y[ii, jj], 0).astype(np.float32)
                    dx = ((lfx + 1).astype(np.float32) - sxx).astype(np.float32)
                    dy = ((lfy + 1).astype(np.float32) - syy).astype(np.float32)
                    one = np.float32(1)
                    taps = [(lfx, lfy, dx * dy), (lfx + 1, lfy + 1, (one - dx) * (one - dy)),
                            (lfx, lfy + 1, dx * (one - dy)), (lfx + 1, lfy, (one - dx) * dy)]
                    inside = [v & (u >= sx0) & (u <= sx1) & (r >= sy0) & (r <= sy1) for u, r, _ in taps]
                    mine = inside[0] | inside[1] | inside[2] | inside[3]
                    if not mine.any():
                        continue
                    # lanes that share a floor cell take turns in lane order; one turn = all four taps
                    rank = np.zeros(32, np.int64)
                    seen = {}
                    for lane in range(32):
                        if mine[lane]:
                            key = (int(lfx[lane]), int(lfy[lane]))
                            rank[lane] = seen.get(key, 0)
                            seen[key] = rank[lane] + 1
                    for r in range(max(seen.values())):
                        for (u, rr, wgt), ins in zip(taps, inside):
                            sel = np.nonzero(ins & mine & (rank == r))[0]
                            cells = list(zip(rr[sel] - sy0, u[sel] - sx0))
                            assert len(set(cells)) == len(cells), "two lanes of one turn hit the same cell"
                            for lane, (cr, cc) in zip(sel, cells):
                                acc[w, cr, cc] += go[ii[lane], jj[lane]] * wgt[lane]
            tile = np.zeros((SRC_H, SRC_W, C), np.float32)
            for w in range(WARPS):                                                       # warps summed in order
                tile = tile + acc[w]
            out[sy0:sy1 + 1, sx0:sx1 + 1] = tile[:sy1 - sy0 + 1, :sx1 - sx0 + 1]
    return out, visits


def _flows(kind, rng, Ho, Wo, H, W):
    ii, jj = np.meshgrid(np.arange(Ho, dtype=np.float32), np.arange(Wo, dtype=np.float32), indexing="ij")
    if kind == "jitter":                       # the reference's (Y,X) grid: x = row + f0, y = column + f1
        return np.stack([ii + rng.uniform(-3, 3, (Ho, Wo)), jj + rng.uniform(-3, 3, (Ho, Wo))], -1).astype(np.float32)
    if kind == "jitter_xy":
        return np.stack([jj + rng.uniform(-3, 3, (Ho, Wo)), ii + rng.uniform(-3, 3, (Ho, Wo))], -1).astype(np.float32)
    if kind == "far":
        return np.stack([rng.uniform(-20, W + 20, (Ho, Wo)), rng.uniform(-20, H + 20, (Ho, Wo))], -1).astype(np.float32)
    if kind == "minify":                       # many outputs per source cell
        return np.stack([jj / 5.0, ii / 5.0], -1).astype(np.float32)
    if kind == "boundary":
        w = np.empty((Ho, Wo, 2), np.float32)
        w[..., 0] = rng.choice(np.array([-1.0, -1 + 1e-6, -0.5, 0.0, 0.5, W - 1.0, W - 0.5, W, 1e6, np.nan], np.float32), (Ho, Wo))
        w[..., 1] = rng.choice(np.array([-1.0, -0.25, 0.0, 0.75, H - 1.0, H - 0.25, H, -1e6], np.float32), (Ho, Wo))
        return w
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["jitter", "jitter_xy", "far", "minify", "boundary"])
def test_schedule_model_matches_oracle(kind):
    rng = np.random.default_rng(len(kind) * 7 + 1)
    H, W, Ho, Wo, C = 44, 52, 40, 72, 3                    # ragged tiles on both sides, several source tiles
    warp = _flows(kind, rng, Ho, Wo, H, W)
    go = rng.standard_normal((Ho, Wo, C)).astype(np.float32)
    data = rng.random((1, H, W, C), dtype=np.float32)
    gd, visits = model_grad_data(warp, go, H, W)
    ref, _ = tf_ops.resampler_grad(data, warp[None], go[None])
    scale = max(1.0, float(np.abs(ref).max()))
    assert float(np.abs(gd - ref[0]).max()) <= 1e-5 * scale
    if kind in ("jitter", "jitter_xy"):
        # the point of the chunk boxes: few visits per output pixel (every 1024-pixel tile processed whole would be ~6x here)
        assert visits * 32 <= 4.0 * Ho * Wo


def test_schedule_model_on_sample_list():
    """[N, 2] sample lists use 1 x 1024 tiles and runs of 32 samples as chunks."""
    rng = np.random.default_rng(3)
    H, W, C, N = 20, 36, 3, 2500
    warp = np.stack([rng.uniform(-2, W + 1, N), rng.uniform(-2, H + 1, N)], -1).astype(np.float32)[None]      # [1, N, 2]
    go = rng.standard_normal((1, N, C)).astype(np.float32)
    data = rng.random((1, H, W, C), dtype=np.float32)
    gd, _ = model_grad_data(warp, go, H, W)
    ref, _ = tf_ops.resampler_grad(data, warp, go)
    assert float(np.abs(gd - ref[0]).max()) <= 1e-5 * max(1.0, float(np.abs(ref).max()))


def test_workspace_holds_one_box_per_tile_and_chunk():
    from dynamic_multiview_3d_b200 import _lib
    lib = _lib.load()
    assert lib.dmv_sampler_bwd_workspace_size(2, 44, 52, 3, 40, 72) == 2 * (2 * 3) * 33 * 16
    assert lib.dmv_sampler_bwd_workspace_size(1, 20, 36, 3, 1, 2500) == 3 * 33 * 16


def test_schedule_model_on_random_ragged_shapes():
    """Random source / output sizes (not multiples of the tiles) and displacement scales, C in {1, 3, 4}."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=12, deadline=None, derandomize=True)
    @given(H=st.integers(3, 50), W=st.integers(3, 70), Ho=st.integers(2, 45), Wo=st.integers(2, 70), C=st.sampled_from([1, 3, 4]),
           amp=st.sampled_from([0.5, 3.0, 12.0]), xy=st.booleans(), seed=st.integers(0, 1000))
    def check(H, W, Ho, Wo, C, amp, xy, seed):
        rng = np.random.default_rng(seed)
        ii, jj = np.meshgrid(np.arange(Ho, dtype=np.float32), np.arange(Wo, dtype=np.float32), indexing="ij")
        a, b = (jj, ii) if xy else (ii, jj)
        warp = np.stack([a * (W / max(Ho, Wo)) + rng.uniform(-amp, amp, (Ho, Wo)), b * (H / max(Ho, Wo)) + rng.uniform(-amp, amp, (Ho, Wo))],
                        -1).astype(np.float32)
        go = rng.standard_normal((Ho, Wo, C)).astype(np.float32)
        data = rng.random((1, H, W, C), dtype=np.float32)
        gd, _ = model_grad_data(warp, go, H, W)
        ref, _ = tf_ops.resampler_grad(data, warp[None], go[None])
        assert float(np.abs(gd - ref[0]).max()) <= 1e-5 * max(1.0, float(np.abs(ref).max()))

    check()
