// tcgen05 weight-gradient kernel (conv / deconv / linear wgrad) for sm_100a.
//
//     dW[(tap, ci), co] = sum_{pixels} Big_tap[pixel, ci] * Small[pixel, co]
// is a GEMM whose contraction runs over PIXELS.  Both operands are read straight from the
// NHWC activation tensors as MN-major tcgen05 operands -- no transposed copies:
//   A (M = tap x ci): per (tap, 32/64-channel chunk) one TMA box of the tap-shifted input
//       window [pixels][channels]; boxes are stacked along M until 128 rows are reached
//       (e.g. 4 taps x 32 channels), the LBO of the matrix descriptor being the box size.
//   B (N = co): the matching box of the gradient tensor [pixels][co].
// Stride-2 layers use the same parity-view tensor map as the forward kernel; TMA zero fill
// supplies TF-SAME padding and the ragged image edge.
// One CTA owns a group of M tiles (as many fp32 accumulators as fit the 512 TMEM columns), an
// N tile (<= 256 output channels) and a slice of the pixels; it streams its pixel chunks
// through a TMA/mbarrier pipeline and finally stores its accumulators.  Partial results of
// the pixel slices go to the workspace and are summed in slice order by a second kernel
// (deterministic split-K); a single slice writes dW directly.
// Warp roles: 0 = TMA producer, 1 = MMA issuer (+ TMEM allocator), 2..5 = epilogue.
#include "tc_common.cuh"
#include "conv_impl.h"

namespace {
using namespace dmv;
using namespace dmv::tc;

constexpr int kThreads = 192;
constexpr int kMaxBoxes = 200;   // (tap, channel chunk) boxes along M

struct ABox {
    short dh, dw, ph, c_off;   // shift in the parity view, parity plane, channel offset inside the view
    int row0;                  // first dW row of this box: tap_id * Cbig + chunk * CB
};
struct WgradParams {
    int n_boxes, boxes_per_tile, m_tiles, tiles_per_group, groups;   // M decomposition
    int CB, row_bytes_a;            // channels per A box, bytes per smem row of A
    int CBs, row_bytes_b, b_boxes;  // same for B; boxes per N tile
    int n_tile, n_tiles, Cs;        // output channels per N tile (UMMA N), number of N tiles, total
    int BW, BH, NB, rows;           // pixel chunk box; rows = BW*BH*NB (multiple of 16)
    int tiles_w, tiles_h, img_groups, chunks, slices, chunks_per_slice;
    int stages;
    long long out_elems;            // rows_total * Cs (size of one partial)
    float* out;                     // partials (slices > 1) or dW itself
    ABox box[kMaxBoxes];
};

__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(const __grid_constant__ CUtensorMap map_a,
                                                             const __grid_constant__ CUtensorMap map_b,
                                                             const __grid_constant__ WgradParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const int a_box_bytes = p.rows * p.row_bytes_a;
    const int a_bytes = p.boxes_per_tile * a_box_bytes;
    const int b_box_bytes = p.rows * p.row_bytes_b;
    const int b_bytes = p.b_boxes * b_box_bytes;
    const int stage_bytes = a_bytes + b_bytes;
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* done_bar = empty_bar + p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work item of this CTA
    int w = blockIdx.x;
    const int slice = w % p.slices; w /= p.slices;
    const int nt = w % p.n_tiles; w /= p.n_tiles;
    const int grp = w;
    const int mt0 = grp * p.tiles_per_group;
    const int mt1 = min(p.m_tiles, mt0 + p.tiles_per_group);
    const int n_mt = mt1 - mt0;
    const int ch0 = slice * p.chunks_per_slice;
    const int ch1 = min(p.chunks, ch0 + p.chunks_per_slice);

    uint32_t tmem_cols = 32;
    while (tmem_cols < (uint32_t)(p.tiles_per_group * p.n_tile)) tmem_cols <<= 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_b);
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(done_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int ch = ch0; ch < ch1; ++ch) {
                int t = ch;
                const int tw = t % p.tiles_w; t /= p.tiles_w;
                const int th = t % p.tiles_h; t /= p.tiles_h;
                const int g = t;
                for (int mt = mt0; mt < mt1; ++mt) {
                    const int b0 = mt * p.boxes_per_tile;
                    const int nb = min(p.boxes_per_tile, p.n_boxes - b0);
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + (size_t)stage * stage_bytes;
                    mbar_expect_tx(&full_bar[stage], (uint32_t)(nb * a_box_bytes + b_bytes));
                    for (int j = 0; j < nb; ++j) {
                        const ABox& bx = p.box[b0 + j];
                        tma_load_5d(sa + j * a_box_bytes, &map_a, &full_bar[stage], bx.c_off, tw * p.BW + bx.dw, bx.ph,
                                    th * p.BH + bx.dh, g * p.NB);
                    }
                    for (int j = 0; j < p.b_boxes; ++j)
                        tma_load_5d(sa + a_bytes + j * b_box_bytes, &map_b, &full_bar[stage], nt * p.n_tile + j * p.CBs, tw * p.BW, 0,
                                    th * p.BH, g * p.NB);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // D = f32, A = B = bf16, both MN-major (bits 15, 16), N = n_tile, M = 128
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(p.n_tile >> 3) << 17) |
                                   ((128u >> 4) << 24);
            int stage = 0;
            uint32_t phase = 0;
            for (int ch = ch0; ch < ch1; ++ch) {
                for (int i = 0; i < n_mt; ++i) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t adesc = make_mnmajor_desc(sa, p.row_bytes_a, a_box_bytes);
                    const uint64_t bdesc = make_mnmajor_desc(sa + a_bytes, p.row_bytes_b, b_box_bytes);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(i * p.n_tile);
                    for (int k = 0; k < p.rows / 16; ++k)   // 16 pixels per MMA: advance 16 rows in both operands
                        tc_mma_bf16(d_tmem, adesc + (uint64_t)((k * 16 * p.row_bytes_a) >> 4), bdesc + (uint64_t)((k * 16 * p.row_bytes_b) >> 4),
                                    idesc, (ch > ch0 || k > 0) ? 1u : 0u);
                    tc_commit(&empty_bar[stage]);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
            tc_commit(done_bar);
        }
    } else {
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;
        mbar_wait(done_bar, 0);
        tc_fence_after();
        float* outp = p.out + (long long)slice * p.out_elems;
        const bool have = ch1 > ch0;
        for (int i = 0; i < n_mt; ++i) {
            const int bi = (mt0 + i) * p.boxes_per_tile + m / p.CB;
            const bool ok = bi < p.n_boxes;
            const long long row = ok ? (long long)p.box[bi].row0 + (m % p.CB) : 0;
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(i * p.n_tile);
            for (int c0 = 0; c0 < p.n_tile; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(taddr + (uint32_t)c0, v);
                tmem_ld_wait();
                if (ok) {
                    float* o = outp + row * p.Cs + nt * p.n_tile + c0;
#pragma unroll
                    for (int k = 0; k < 16; k += 4)
                        *reinterpret_cast<float4*>(o + k) =
                            have ? make_float4(__uint_as_float(v[k]), __uint_as_float(v[k + 1]), __uint_as_float(v[k + 2]), __uint_as_float(v[k + 3]))
                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

// pixel chunk box: rows = BW*BH*NB a multiple of 16, <= max_rows, maximising the useful fraction
void choose_chunk(int Jh, int Jw, int N, int max_rows, int& BW, int& BH, int& NB) {
    double best = -1.0;
    BW = 16; BH = 1; NB = 1;
    const int wmax = Jw + 15;
    for (int bw = 1; bw <= wmax && bw <= max_rows; ++bw)
        for (int bh = 1; bh <= Jh + 15 && bw * bh <= max_rows; ++bh) {
            int nb = 1;
            if (bw >= Jw && bh >= Jh) nb = max_rows / (bw * bh);
            if (nb > N) nb = N;
            if (nb < 1) nb = 1;
            const int rows = bw * bh * nb;
            if (rows % 16) continue;
            const long long chunks = (long long)ceil_div(Jw, bw) * ceil_div(Jh, bh) * ceil_div(N, nb);
            const double eff = (double)Jh * Jw * N / ((double)chunks * rows);
            // on ties prefer more rows per chunk (fewer pipeline steps), then wider boxes
            if (eff > best + 1e-9 || (eff > best - 1e-9 && (rows > BW * BH * NB || (rows == BW * BH * NB && bw > BW)))) {
                best = eff; BW = bw; BH = bh; NB = nb;
            }
        }
}

struct WProblem {
    const void* big; int N, Hb, Wb, Cb, stride;      // tensor read through tap shifts (conv: x; deconv: dy)
    const void* small; int Hs, Ws, Cs;               // tensor read unshifted (conv: dy; deconv: x)
    int kh, kw, pt, pl;
    float* dw;                                       // [kh*kw][Cb][Cs] fp32
};

size_t plan_wgrad(const WProblem& q, WgradParams& p) {
    memset(&p, 0, sizeof(p));
    p.CB = (q.Cb % 64 == 0) ? 64 : 32;
    p.row_bytes_a = p.CB * 2;
    p.CBs = (q.Cs % 64 == 0) ? 64 : 32;
    p.row_bytes_b = p.CBs * 2;
    p.Cs = q.Cs;
    p.n_tile = q.Cs > 256 ? 256 : q.Cs;
    p.n_tiles = ceil_div(q.Cs, p.n_tile);
    p.b_boxes = p.n_tile / p.CBs;
    const int chunks_c = q.Cb / p.CB;
    p.n_boxes = q.kh * q.kw * chunks_c;
    p.boxes_per_tile = 128 / p.CB;
    p.m_tiles = ceil_div(p.n_boxes, p.boxes_per_tile);
    p.tiles_per_group = 512 / p.n_tile;
    if (p.tiles_per_group > p.m_tiles) p.tiles_per_group = p.m_tiles;
    p.groups = ceil_div(p.m_tiles, p.tiles_per_group);
    const int max_rows = (p.n_tile >= 128) ? 64 : 128;
    choose_chunk(q.Hs, q.Ws, q.N, max_rows, p.BW, p.BH, p.NB);
    p.rows = p.BW * p.BH * p.NB;
    p.tiles_w = ceil_div(q.Ws, p.BW);
    p.tiles_h = ceil_div(q.Hs, p.BH);
    p.img_groups = ceil_div(q.N, p.NB);
    p.chunks = p.tiles_w * p.tiles_h * p.img_groups;
    const int base_ctas = p.groups * p.n_tiles;
    int slices = ceil_div(num_sms(), base_ctas);
    if (slices > p.chunks) slices = p.chunks;
    if (slices < 1) slices = 1;
    p.chunks_per_slice = ceil_div(p.chunks, slices);
    p.slices = ceil_div(p.chunks, p.chunks_per_slice);
    p.out_elems = (long long)q.kh * q.kw * q.Cb * q.Cs;
    return p.slices > 1 ? (size_t)p.slices * p.out_elems * sizeof(float) : 0;
}

bool wgrad_eligible(int Cb, int Cs, int taps) {
    if (Cb % 32 || Cs % 32) return false;
    const int CB = (Cb % 64 == 0) ? 64 : 32;
    if (taps * (Cb / CB) > kMaxBoxes) return false;
    if (Cs > 256 && Cs % 256) return false;
    return true;
}

int run_wgrad(const WProblem& q, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!wgrad_eligible(q.Cb, q.Cs, q.kh * q.kw)) return fail(DMV_E_UNSUPPORTED_SHAPE, "tc wgrad: channel counts not covered");
    if (q.stride != 1 && q.stride != 2) return fail(DMV_E_UNSUPPORTED_SHAPE, "tc wgrad: stride");
    if (q.stride == 2 && ((q.Hb & 1) || (q.Wb & 1))) return fail(DMV_E_UNSUPPORTED_SHAPE, "tc wgrad: stride-2 source needs even H and W");
    if (((uintptr_t)q.big & 15) || ((uintptr_t)q.small & 15) || ((uintptr_t)q.dw & 15))
        return fail(DMV_E_UNSUPPORTED_SHAPE, "tc wgrad: buffers must be 16-byte aligned");
    WgradParams p;
    const size_t need = plan_wgrad(q, p);
    if (need > 0 && (!ws || ws_bytes < need)) return fail(DMV_E_WORKSPACE, "tc wgrad: workspace too small");
    p.out = p.slices > 1 ? reinterpret_cast<float*>(ws) : q.dw;
    // A boxes: (tap, channel chunk)
    const int chunks_c = q.Cb / p.CB;
    int n = 0;
    for (int r = 0; r < q.kh; ++r)
        for (int s = 0; s < q.kw; ++s)
            for (int c = 0; c < chunks_c; ++c) {
                ABox& b = p.box[n++];
                const int dy = r - q.pt, dx = s - q.pl;
                int ph = 0, pw = 0, dh = dy, dwv = dx;
                if (q.stride == 2) {
                    ph = ((dy % 2) + 2) % 2; dh = (dy - ph) / 2;
                    pw = ((dx % 2) + 2) % 2; dwv = (dx - pw) / 2;
                }
                b.dh = (short)dh; b.dw = (short)dwv; b.ph = (short)ph; b.c_off = (short)(pw * q.Cb + c * p.CB);
                b.row0 = (r * q.kw + s) * q.Cb + c * p.CB;
            }
    CUtensorMap map_a, map_b;
    {
        const int s = q.stride;
        cuuint64_t dims[5] = {(cuuint64_t)(s * q.Cb), (cuuint64_t)(q.Wb / s), (cuuint64_t)s, (cuuint64_t)(q.Hb / s), (cuuint64_t)q.N};
        const cuuint64_t pix = (cuuint64_t)q.Cb * 2;
        cuuint64_t strides[4] = {(cuuint64_t)s * pix, (cuuint64_t)q.Wb * pix, (cuuint64_t)s * q.Wb * pix, (cuuint64_t)q.Hb * q.Wb * pix};
        cuuint32_t box[5] = {(cuuint32_t)p.CB, (cuuint32_t)p.BW, 1u, (cuuint32_t)p.BH, (cuuint32_t)p.NB};
        int rc = encode_map(&map_a, q.big, 5, dims, strides, box, p.row_bytes_a);
        if (rc) return rc;
    }
    {
        cuuint64_t dims[5] = {(cuuint64_t)q.Cs, (cuuint64_t)q.Ws, 1, (cuuint64_t)q.Hs, (cuuint64_t)q.N};
        const cuuint64_t pix = (cuuint64_t)q.Cs * 2;
        cuuint64_t strides[4] = {pix, (cuuint64_t)q.Ws * pix, (cuuint64_t)q.Ws * pix, (cuuint64_t)q.Hs * q.Ws * pix};
        cuuint32_t box[5] = {(cuuint32_t)p.CBs, (cuuint32_t)p.BW, 1u, (cuuint32_t)p.BH, (cuuint32_t)p.NB};
        int rc = encode_map(&map_b, q.small, 5, dims, strides, box, p.row_bytes_b);
        if (rc) return rc;
    }
    const int stage_bytes = p.boxes_per_tile * p.rows * p.row_bytes_a + p.b_boxes * p.rows * p.row_bytes_b;
    int stages = (176 * 1024) / stage_bytes;
    if (stages > 6) stages = 6;
    if (stages < 2) return fail(DMV_E_UNSUPPORTED_SHAPE, "tc wgrad: stage does not fit shared memory");
    p.stages = stages;
    const size_t smem = (size_t)stages * stage_bytes + (2 * stages + 1) * sizeof(uint64_t) + 16 + 1024;
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
        set_error("tc wgrad: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return DMV_E_CUDA;
    }
    const int grid = p.groups * p.n_tiles * p.slices;
    wgrad_kernel<<<grid, kThreads, smem, st>>>(map_a, map_b, p);
    count_tc_launch();
    int rc = check_launch("wgrad_tc");
    if (rc) return rc;
    if (p.slices > 1) {
        rc = reduce_partials(reinterpret_cast<const float*>(ws), q.dw, p.out_elems, p.slices, st);
    }
    return rc;
}

}  // namespace

namespace dmv {

size_t tc_wgrad_workspace(int taps, int Cin, int Cout, long long pixels) {
    // mirrors plan_wgrad: one fp32 copy of dW per pixel slice (none when a single slice writes dW directly)
    if (!wgrad_eligible(Cin, Cout, taps)) return 0;
    const int CB = (Cin % 64 == 0) ? 64 : 32;
    const int n_tile = Cout > 256 ? 256 : Cout;
    const int n_tiles = ceil_div(Cout, n_tile);
    const int m_tiles = ceil_div(taps * (Cin / CB), 128 / CB);
    int tpg = 512 / n_tile;
    if (tpg > m_tiles) tpg = m_tiles;
    const int base = ceil_div(m_tiles, tpg) * n_tiles;
    long long slices = ceil_div(num_sms(), base);
    const long long max_chunks = ceil_div_ll(pixels, 16);
    if (slices > max_chunks) slices = max_chunks;
    if (slices <= 1) return 0;
    return (size_t)slices * (size_t)taps * Cin * Cout * sizeof(float) + 256;
}

int tc_conv_wgrad(const void* x, int xdt, const void* dy, float* dw, float* db, int B, int H, int W, int Cin, int Cout, int kh, int kw,
                  int stride, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (xdt != DMV_DT_BF16) return fail(DMV_E_UNSUPPORTED_SHAPE, "tc_conv_wgrad: bf16 input only");
    if (db) return fail(DMV_E_UNSUPPORTED_SHAPE, "tc_conv_wgrad: bias gradient is taken by the caller");
    const SamePad ph = same_pad(H, kh, stride), pw = same_pad(W, kw, stride);
    WProblem q;
    q.big = x; q.N = B; q.Hb = H; q.Wb = W; q.Cb = Cin; q.stride = stride;
    q.small = dy; q.Hs = ph.out; q.Ws = pw.out; q.Cs = Cout;
    q.kh = kh; q.kw = kw; q.pt = ph.before; q.pl = pw.before; q.dw = dw;
    return run_wgrad(q, ws, ws_bytes, st);
}

int tc_deconv_wgrad(const void* x, const void* dy, int dydt, float* dw, int B, int Hout, int Wout, int Cin, int Cout, int kh, int kw,
                    int stride, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (dydt != DMV_DT_BF16) return fail(DMV_E_UNSUPPORTED_SHAPE, "tc_deconv_wgrad: bf16 gradient only");
    const SamePad ph = same_pad(Hout, kh, stride), pw = same_pad(Wout, kw, stride);
    WProblem q;
    q.big = dy; q.N = B; q.Hb = Hout; q.Wb = Wout; q.Cb = Cout; q.stride = stride;
    q.small = x; q.Hs = ph.out; q.Ws = pw.out; q.Cs = Cin;
    q.kh = kh; q.kw = kw; q.pt = ph.before; q.pl = pw.before; q.dw = dw;
    return run_wgrad(q, ws, ws_bytes, st);
}

int tc_linear_wgrad(const void* x, const void* dy, float* dw, float* db, int M, int K, int N, void* ws, size_t ws_bytes,
                    cudaStream_t st) {
    if (db) return fail(DMV_E_UNSUPPORTED_SHAPE, "tc_linear_wgrad: bias gradient is taken by the caller");
    WProblem q;
    q.big = x; q.N = M; q.Hb = 1; q.Wb = 1; q.Cb = K; q.stride = 1;
    q.small = dy; q.Hs = 1; q.Ws = 1; q.Cs = N;
    q.kh = 1; q.kw = 1; q.pt = 0; q.pl = 0; q.dw = dw;
    return run_wgrad(q, ws, ws_bytes, st);
}


int tc_thin_wgrad(const void* thin, int thin_dtype, const void* wide, float* dw, int N, int Hb, int Wb, int Ct, int Cw, int kh, int kw,
                  int stride, void* ws, size_t ws_bytes, cudaStream_t st) {
    const SamePad ph = same_pad(Hb, kh, stride), pw = same_pad(Wb, kw, stride);
    const int rows = kh * kw * Ct, Kp = thin_patch_cols(kh * kw, Ct);
    if (!wgrad_eligible(Kp, Cw, 1)) return fail(DMV_E_UNSUPPORTED_SHAPE, "tc_thin_wgrad: channel counts not covered");
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t Pb = al((size_t)N * ph.out * pw.out * Kp * 2), Db = al((size_t)Kp * Cw * 4);
    const size_t parts = tc_wgrad_workspace(1, Kp, Cw, (long long)N * ph.out * pw.out);
    if (!ws || ws_bytes < Pb + Db + parts || ((uintptr_t)ws & 255)) return fail(DMV_E_WORKSPACE, "tc_thin_wgrad: workspace too small or unaligned");
    uint8_t* base = reinterpret_cast<uint8_t*>(ws);
    int rc = thin_im2col(thin, thin_dtype, base, N, Hb, Wb, Ct, kh, kw, stride, st);
    if (rc) return rc;
    float* dwt = reinterpret_cast<float*>(base + Pb);
    WProblem q;
    q.big = base; q.N = N; q.Hb = ph.out; q.Wb = pw.out; q.Cb = Kp; q.stride = 1;
    q.small = wide; q.Hs = ph.out; q.Ws = pw.out; q.Cs = Cw;
    q.kh = 1; q.kw = 1; q.pt = 0; q.pl = 0; q.dw = dwt;
    rc = run_wgrad(q, base + Pb + Db, ws_bytes - Pb - Db, st);
    if (rc) return rc;
    return copy_f32(dwt, dw, rows * Cw, st);      // drop the zero-padded rows
}

}  // namespace dmv

