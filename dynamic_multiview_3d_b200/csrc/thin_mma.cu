// The two thin layers of the graph as bandwidth kernels.
//
// e0 (conv 5x5 stride 2, 3 -> 32 channels on the 224^2 image) and the flow head (deconv 5x5 stride 2, 32 -> 2 channels)
// have a channel count on the image side far below tcgen05 / TMA granularity.  Through the space-to-depth route
// (thin_s2d_*, conv_tc.cu) they pay a prep pass and 4x zero-padded K on N = 32 tensor-core tiles.  Their arithmetic is
// tiny (4 GFLOP each) next to their HBM traffic (77 - 90 MB), so here each is ONE kernel that stages the thin tensor's
// window in shared memory (fp32 -> bf16 on the way, channels padded to 4) and contracts with warp-level mma.sync
// (m16n8k16, bf16 in, fp32 accumulate) straight out of that window -- no patch matrix, no prep kernel:
//
//   thin_conv_kernel    out[n,oh,ow,0:32] = act(bias + sum_{r,s,c} thin[n, 2oh+r-pt, 2ow+s-pl, c] * w[r,s,c,:])
//                       e0 forward (thin = image), flow-head input gradient (thin = dL/dflow; same formula)
//   thin_deconv_kernel  y[n,2q+p,:] = sum over the 3x3 source window of x[n,q+d,0:32] * w[r(p,d),s(p,d),c,:]   (sub-pixel form)
//                       flow-head forward
//   thin_wgrad_kernel   dW[r,s,c,0:32] = sum_{n,oh,ow} thin[n, 2oh+r-pt, 2ow+s-pl, c] * wide[n,oh,ow,0:32]
//                       e0 weight gradient (thin = image, wide = dPre), flow-head weight gradient (thin = dflow, wide = x)
//
// Shapes: 5x5 kernel, stride 2, 32 wide channels, <= 4 thin channels, wide side a multiple of 8 x 16 pixels; anything else
// returns DMV_E_UNSUPPORTED_SHAPE and the caller falls through to the space-to-depth path.
//
// K ordering of the window contraction: k = r * 20 + s * 4 + c (5 rows of 5 taps x 4 padded channels), 100 -> 112 = 7
// k16 steps.  For an output pixel the 20 values of a row r are CONTIGUOUS in the staged window, so an A-fragment register
// (two consecutive k) is one 32-bit shared load, conflict-free because neighbouring output pixels are 16 bytes apart.
#include <stdlib.h>

#include "common.cuh"
#include "conv_impl.h"

namespace {
using namespace dmv;
typedef __nv_bfloat16 bf16;

constexpr int TH = 8, TW = 16;      // tile of wide-side pixels: 8 rows x 16 columns = one m16 tile per warp
constexpr int WR = 2 * TH + 4;      // window rows (19 used + 1 zero row reached by the K padding)
constexpr int WC = 2 * TW + 4;      // window pixels per row (35 used)
constexpr int KS = 7;               // k16 steps: 5 * 20 = 100 -> 112
constexpr int XPITCH = 40;          // bf16 per staged pixel of a 32-channel tensor (80 B: conflict-free ldmatrix rows)

__device__ __forceinline__ void mma_bf16(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(unsigned& r0, unsigned& r1, unsigned& r2, unsigned& r3, const void* p) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(unsigned& r0, unsigned& r1, unsigned& r2, unsigned& r3, const void* p) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void ldsm_x2_t(unsigned& r0, unsigned& r1, const void* p) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a));
}
__device__ __forceinline__ unsigned pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<unsigned*>(&v);
}
__device__ __forceinline__ unsigned pack_raw(bf16 lo, bf16 hi) {
    return (unsigned)__bfloat16_as_ushort(lo) | ((unsigned)__bfloat16_as_ushort(hi) << 16);
}

struct ThinGeo {
    int N, Hb, Wb, Hs, Ws, pt, pl, tiles_x, tiles_y;
    long long tiles;
};

// Window of the thin tensor for the wide-side tile (n, ty, tx): rows 2*oh0 - pt .. + 18, pixels 2*ow0 - pl .. + 34, fp32 ->
// bf16, channels padded to 4 (the pad lanes, row 19 and pixel 35 are zeroed once per CTA and never written).  Staged through
// registers in two halves so that the loads of the NEXT tile are in flight while the current one is contracted:
// window_load issues them (NR per thread, independent), window_store converts and writes them after the barrier.
template <int CT, int NTHR>
struct WindowRegs {
    static constexpr int ROWE = 35 * CT, TOTAL = 19 * ROWE, NR = (TOTAL + NTHR - 1) / NTHR;
    float v[NR];
    int goff[NR];      // element offset from the window's origin in the thin tensor (row * Wb * CT + e); -1: no element
    int meta[NR];      // shared-memory element index | row << 12 | pixel << 17
    // the element -> (row, pixel, channel) map does not depend on the tile: computed once per thread (the kernels are bound
    // by instruction issue, not by memory: profiles/r02_thin_ncu.txt)
    __device__ __forceinline__ void init(const ThinGeo& g, int tid) {
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            const int i = tid + j * NTHR;
            const int row = i / ROWE, e = i - row * ROWE, px = e / CT, c = e - px * CT;
            goff[j] = i < TOTAL ? row * g.Wb * CT + e : -1;
            meta[j] = ((row * WC + px) * 4 + c) | (row << 12) | (px << 17);
        }
    }
    __device__ __forceinline__ void load(const float* __restrict__ thin, const ThinGeo& g, int n, int oh0, int ow0) {
        const int y0 = 2 * oh0 - g.pt, x0 = 2 * ow0 - g.pl;
        const float* base = thin + (((long long)n * g.Hb + y0) * g.Wb + x0) * CT;
        if (y0 >= 0 && x0 >= 0 && y0 + 19 <= g.Hb && x0 + 35 <= g.Wb) {          // interior tile: no bounds checks
#pragma unroll
            for (int j = 0; j < NR; ++j) v[j] = goff[j] >= 0 ? __ldg(base + goff[j]) : 0.f;
        } else {
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                const int gy = y0 + ((meta[j] >> 12) & 31), gx = x0 + (meta[j] >> 17);
                v[j] = (goff[j] >= 0 && gy >= 0 && gy < g.Hb && gx >= 0 && gx < g.Wb) ? __ldg(base + goff[j]) : 0.f;
            }
        }
    }
    __device__ __forceinline__ void store(bf16* win) const {
#pragma unroll
        for (int j = 0; j < NR; ++j)
            if (goff[j] >= 0) win[meta[j] & 4095] = __float2bfloat16_rn(v[j]);
    }
};

__device__ __forceinline__ void cp_async16_zfill(void* dst, const void* src, bool valid) {
    const unsigned sz = valid ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

__device__ __forceinline__ void tile_coords(const ThinGeo& g, long long tile, int& n, int& oh0, int& ow0) {
    const unsigned per = (unsigned)(g.tiles_x * g.tiles_y), tl = (unsigned)tile;        // tiles < 2^31 (checked on the host)
    n = (int)(tl / per);
    const unsigned rem = tl - (unsigned)n * per;
    const unsigned ty = rem / (unsigned)g.tiles_x;
    oh0 = (int)ty * TH;
    ow0 = (int)(rem - ty * (unsigned)g.tiles_x) * TW;
}

// ----------------------------------------------------------------------------------------------------------------
// thin conv forward: 256 threads = 8 warps, warp w owns output row oh0 + w (16 pixels = one m16 tile), N = 32 channels = 4 n8
// tiles whose columns are permuted (MMA column j of tile nt <-> channel 8 (j / 2) + 2 nt + (j % 2)) so that thread t of a
// quad ends up with the 8 consecutive channels 8t .. 8t+7 of its pixels: one 16-byte store per pixel, a quad writes the
// pixel's whole 64-byte row.  The weight fragments sit in shared memory in fragment order (one conflict-free 8-byte load
// per MMA); persistent CTAs, the next tile's window is loaded into registers while the current one is contracted.
template <int CT>
__global__ void __launch_bounds__(256, 2)
thin_conv_kernel(const float* __restrict__ thin, const bf16* __restrict__ w, const float* __restrict__ bias, bf16* __restrict__ out, ThinGeo g,
                 int act) {
    __shared__ __align__(16) bf16 win[WR * WC * 4];
    __shared__ __align__(16) uint2 bsm[KS * 4 * 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, t = lane & 3;
    for (int i = tid; i < WR * WC * 2; i += 256) reinterpret_cast<unsigned*>(win)[i] = 0u;
    for (int i = tid; i < KS * 4 * 32; i += 256) {           // fragment (ks, nt) of lane l: B[k = 16 ks + 2 tt (+1, +8, +9)][n = gg]
        const int l = i & 31, nt = (i >> 5) & 3, ks = i >> 7, gg = l >> 2, tt = l & 3;
        const int co = 8 * (gg >> 1) + 2 * nt + (gg & 1);
        unsigned f[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            bf16 v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = 16 * ks + 8 * h + 2 * tt + e;
                const int r = k / 20, rem = k - r * 20, sx = rem >> 2, c = rem & 3;
                v[e] = (r < 5 && c < CT) ? w[((r * 5 + sx) * CT + c) * 32 + co] : __float2bfloat16_rn(0.f);
            }
            f[h] = pack_raw(v[0], v[1]);
        }
        bsm[i] = make_uint2(f[0], f[1]);
    }
    float bv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) bv[e] = bias ? __ldg(bias + 8 * t + e) : 0.f;

    // element offsets of this thread's k pairs inside the window, relative to the pixel's origin (row 2*oh_l, pixel 2*ow_l)
    int koff[KS][2];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = 16 * ks + 8 * h + 2 * t;
            koff[ks][h] = (k / 20) * (WC * 4) + (k % 20);
        }
    WindowRegs<CT, 256> wr;
    wr.init(g, tid);
    long long tile = blockIdx.x;
    int n = 0, oh0 = 0, ow0 = 0;
    if (tile < g.tiles) {
        tile_coords(g, tile, n, oh0, ow0);
        wr.load(thin, g, n, oh0, ow0);
    }
    for (; tile < g.tiles; tile += gridDim.x) {
        __syncthreads();                                   // the previous tile's fragments have been read
        wr.store(win);
        __syncthreads();
        const int cn = n, coh0 = oh0, cow0 = ow0;
        if (tile + gridDim.x < g.tiles) {                  // the next tile's loads fly during this tile's contraction
            tile_coords(g, tile + gridDim.x, n, oh0, ow0);
            wr.load(thin, g, n, oh0, ow0);
        }
        float acc[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
        const bf16* p0 = win + ((2 * warp) * WC + 2 * gq) * 4;          // pixel gq of this warp's row
        const bf16* p1 = p0 + 16 * 4;                                      // pixel gq + 8
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
            const int o0 = koff[ks][0], o1 = koff[ks][1];
            unsigned a[4];
            a[0] = *reinterpret_cast<const unsigned*>(p0 + o0);
            a[1] = *reinterpret_cast<const unsigned*>(p1 + o0);
            a[2] = *reinterpret_cast<const unsigned*>(p0 + o1);
            a[3] = *reinterpret_cast<const unsigned*>(p1 + o1);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const uint2 bq = bsm[(ks * 4 + nt) * 32 + lane];
                mma_bf16(acc[nt], a, bq.x, bq.y);
            }
        }
        const int oh = coh0 + warp;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ow = cow0 + gq + 8 * h;
            float v[8];
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                v[2 * nt] = apply_act(acc[nt][2 * h] + bv[2 * nt], act);
                v[2 * nt + 1] = apply_act(acc[nt][2 * h + 1] + bv[2 * nt + 1], act);
            }
            uint4 pk = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
            *reinterpret_cast<uint4*>(out + (((long long)cn * g.Hs + oh) * g.Ws + ow) * 32 + 8 * t) = pk;
        }
    }
}

// ----------------------------------------------------------------------------------------------------------------
// thin deconv forward (sub-pixel form).  Source pixel q produces the 2x2 output block (2q + p); output parity p reads source
// rows q + d with kernel row r = p + pt - 2d (0 <= r < 5), likewise for columns.  Contraction: K = 9 source taps x 32
// channels = 18 k16 steps, N = 4 parities x CT channels (8 for the flow head: one n8 tile).  The A fragments come from the
// staged source window with ldmatrix (rows = 16 consecutive pixels, 80-byte pitch).  An accumulator column pair (2t, 2t+1)
// of the flow head is (parity t, channels 0..1): one 8-byte fp32 store per output pixel, 16 consecutive pixels per row.
template <int CT, typename OutT>
__global__ void __launch_bounds__(256, 2)
thin_deconv_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w, OutT* __restrict__ y, ThinGeo g, int act) {
    constexpr int NT = (4 * CT + 7) / 8;
    __shared__ __align__(16) bf16 xw[2 * (TH + 2) * (TW + 2) * XPITCH];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, t = lane & 3;

    unsigned bfrag[18][NT][2];
#pragma unroll
    for (int ks = 0; ks < 18; ++ks) {
        const int tap = ks >> 1, di = tap / 3, dj = tap % 3;             // source offsets d = di - 1, dj - 1
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int nn = 8 * nt + gq, par = nn / CT, c = nn - par * CT;
            const int py = par >> 1, px = par & 1;
            const int r = py + g.pt - 2 * (di - 1), s = px + g.pl - 2 * (dj - 1);
            const bool ok = par < 4 && r >= 0 && r < 5 && s >= 0 && s < 5;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int ci = (ks & 1) * 16 + 8 * h + 2 * t;
                bfrag[ks][nt][h] = ok ? *reinterpret_cast<const unsigned*>(w + ((r * 5 + s) * CT + c) * 32 + ci) : 0u;
            }
        }
    }
    const int mat = lane >> 3, rr = lane & 7;
    // source windows arrive by 16-byte cp.async (zero fill outside the image) into two buffers: tile i+1 is in flight while
    // tile i is contracted
    constexpr int NU = ((TH + 2) * (TW + 2) * 4 + 255) / 256;          // 16-byte units per thread
    int uoff[NU], umeta[NU];                                              // global offset (uint4) from the window origin; smem unit | wy << 12 | wx << 17
#pragma unroll
    for (int j = 0; j < NU; ++j) {
        const int i = tid + j * 256, pix = i >> 2, u = i & 3, wy = pix / (TW + 2), wx = pix - wy * (TW + 2);
        uoff[j] = i < (TH + 2) * (TW + 2) * 4 ? (wy * g.Ws + wx) * 4 + u : -1;
        umeta[j] = (pix * (XPITCH / 8) + u) | (wy << 12) | (wx << 17);
    }
    auto issue = [&](long long tl, int buf) {
        if (tl < g.tiles) {
            int n, q0y, q0x;
            tile_coords(g, tl, n, q0y, q0x);
            const uint4* base = reinterpret_cast<const uint4*>(x + (((long long)n * g.Hs + q0y - 1) * g.Ws + q0x - 1) * 32);
            uint4* dst = reinterpret_cast<uint4*>(xw + buf * (TH + 2) * (TW + 2) * XPITCH);
#pragma unroll
            for (int j = 0; j < NU; ++j) {
                if (uoff[j] < 0) continue;
                const int sy = q0y - 1 + ((umeta[j] >> 12) & 31), sx = q0x - 1 + (umeta[j] >> 17);
                const bool ok = sy >= 0 && sy < g.Hs && sx >= 0 && sx < g.Ws;
                cp_async16_zfill(dst + (umeta[j] & 4095), ok ? (const void*)(base + uoff[j]) : (const void*)x, ok);
            }
        }
        cp_async_commit();
    };
    issue(blockIdx.x, 0);
    int buf = 0;
    for (long long tile = blockIdx.x; tile < g.tiles; tile += gridDim.x, buf ^= 1) {
        int n, q0y, q0x;
        tile_coords(g, tile, n, q0y, q0x);
        issue(tile + gridDim.x, buf ^ 1);                   // (the barrier that ended the previous tile freed that buffer)
        cp_async_wait<1>();
        __syncthreads();
        const bf16* xb = xw + buf * (TH + 2) * (TW + 2) * XPITCH;
        float acc[NT][4];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[nt][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 18; ++ks) {
            const int tap = ks >> 1, di = tap / 3, dj = tap % 3;
            unsigned a[4];
            // matrices: 0 = pixels 0-7 / k 0-7, 1 = pixels 8-15 / k 0-7, 2 = pixels 0-7 / k 8-15, 3 = pixels 8-15 / k 8-15
            ldsm_x4(a[0], a[1], a[2], a[3], xb + ((warp + di) * (TW + 2) + dj + rr + (mat & 1) * 8) * XPITCH + (ks & 1) * 16 + (mat >> 1) * 8);
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) mma_bf16(acc[nt], a, bfrag[ks][nt][0], bfrag[ks][nt][1]);
        }
        const int qy = q0y + warp;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int qx = q0x + gq + 8 * h;
                if (CT == 2) {
                    const int py = t >> 1, px = t & 1;               // column pair (2t, 2t+1) = parity t, channels 0..1
                    OutT* dst = y + (((long long)n * g.Hb + 2 * qy + py) * g.Wb + 2 * qx + px) * 2;
                    const float v0 = apply_act(acc[nt][2 * h], act), v1 = apply_act(acc[nt][2 * h + 1], act);
                    if (sizeof(OutT) == 4) *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
                    else *reinterpret_cast<unsigned*>(dst) = pack_bf16(v0, v1);
                } else {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int nn = 8 * nt + 2 * t + e, par = nn / CT, c = nn - par * CT;
                        if (par < 4)
                            store_from_float(y + (((long long)n * g.Hb + 2 * qy + (par >> 1)) * g.Wb + 2 * qx + (par & 1)) * CT + c,
                                             apply_act(acc[nt][2 * h + e], act));
                    }
                }
            }
        __syncthreads();                                   // this buffer is refilled two tiles from now
    }
    cp_async_wait<0>();
}

// ----------------------------------------------------------------------------------------------------------------
// thin weight gradient.  Contraction over pixels: D[cw][(r, s, c)] += sum_pixels wide[pixel][cw] * window[pixel][(r, s, c)].
// Both operands are stored pixel-major, so both fragments come from ldmatrix.trans.  For kernel row r the 20 window values of
// an output pixel are contiguous and start 16 bytes after its left neighbour's: the "matrix" an ldmatrix reads is 8 overlapping
// 16-byte rows -- three n8 tiles per r (taps 0-1, 2-3, 4 + 4 discarded columns).  160 threads = 5 warps, warp r owns kernel row
// r: 2 m16 tiles (cw) x 3 n8 tiles, 24 accumulators, over the CTA's tiles (persistent); the per-CTA partials are summed in CTA
// order by the shared deterministic reduction, then gathered into dW[r][s][c][cw].
template <int CT>
__global__ void __launch_bounds__(160, 3)
thin_wgrad_mma_kernel(const float* __restrict__ thin, const bf16* __restrict__ wide, float* __restrict__ part, ThinGeo g) {
    __shared__ __align__(16) bf16 win[WR * WC * 4];
    __shared__ __align__(16) bf16 wd[TH * TW * XPITCH];
    const int tid = threadIdx.x, lane = tid & 31, r = tid >> 5;
    const int gq = lane >> 2, t = lane & 3, mat = lane >> 3, rr = lane & 7;
    for (int i = tid; i < WR * WC * 2; i += 160) reinterpret_cast<unsigned*>(win)[i] = 0u;
    float acc[2][3][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;

    WindowRegs<CT, 160> wr;
    wr.init(g, tid);
    uint4 wq[4];                                             // this thread's 16-byte units of the wide tile (512 per tile)
    int woff[4];                                             // their offsets (in uint4) from the tile's first pixel
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = tid + j * 160, pix = i >> 2, oy = pix / TW, ox = pix - oy * TW;
        woff[j] = i < TH * TW * 4 ? (oy * g.Ws + ox) * 4 + (i & 3) : -1;
    }
    auto load_wide = [&](int n, int oh0, int ow0) {
        const uint4* base = reinterpret_cast<const uint4*>(wide + (((long long)n * g.Hs + oh0) * g.Ws + ow0) * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (woff[j] >= 0) wq[j] = __ldg(base + woff[j]);
    };
    long long tile = blockIdx.x;
    int n = 0, oh0 = 0, ow0 = 0;
    if (tile < g.tiles) {
        tile_coords(g, tile, n, oh0, ow0);
        wr.load(thin, g, n, oh0, ow0);
        load_wide(n, oh0, ow0);
    }
    for (; tile < g.tiles; tile += gridDim.x) {
        __syncthreads();
        wr.store(win);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = tid + j * 160;
            if (woff[j] >= 0) *reinterpret_cast<uint4*>(wd + (i >> 2) * XPITCH + (i & 3) * 8) = wq[j];
        }
        __syncthreads();
        if (tile + gridDim.x < g.tiles) {                    // next tile's loads fly during this tile's contraction
            tile_coords(g, tile + gridDim.x, n, oh0, ow0);
            wr.load(thin, g, n, oh0, ow0);
            load_wide(n, oh0, ow0);
        }
#pragma unroll 2
        for (int row = 0; row < TH; ++row) {                // one k16 step per output row: its 16 pixels
            unsigned a[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)                   // A = wide^T: stored [pixel][cw]
                ldsm_x4_t(a[mt][0], a[mt][1], a[mt][2], a[mt][3], wd + (row * TW + rr + (mat >> 1) * 8) * XPITCH + mt * 16 + (mat & 1) * 8);
            const bf16* wrow = win + ((2 * row + r) * WC) * 4;
            unsigned b[3][2];
            // matrices 0/1: pixels 0-7 / 8-15 of n-tile 0; 2/3: the same of n-tile 1
            ldsm_x4_t(b[0][0], b[0][1], b[1][0], b[1][1], wrow + (2 * (rr + (mat & 1) * 8)) * 4 + (mat >> 1) * 8);
            ldsm_x2_t(b[2][0], b[2][1], wrow + (2 * (rr + (mat & 1) * 8)) * 4 + 16);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 3; ++nt) mma_bf16(acc[mt][nt], a[mt], b[nt][0], b[nt][1]);
        }
    }
    // partial of this CTA: part[cta][r][n (24)][cw (32)]
    float* dst = part + ((long long)blockIdx.x * 5 + r) * 24 * 32;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int cw = 16 * mt + gq + 8 * (e >> 1), nn = 8 * nt + 2 * t + (e & 1);
                dst[nn * 32 + cw] = acc[mt][nt][e];
            }
}

// dW[r][s][c][cw] <- sum[r][s * 4 + c][cw]
__global__ void thin_wgrad_gather_kernel(const float* __restrict__ sum, float* __restrict__ dw, int CT) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 25 * CT * 32) return;
    const int cw = i & 31, row = i >> 5, c = row % CT, tap = row / CT, r = tap / 5, s = tap % 5;
    dw[i] = sum[(r * 24 + s * 4 + c) * 32 + cw];
}

int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

// persistent grid: every CTA resident at once (SMs x occupancy), never more CTAs than tiles
template <typename K>
long long resident_grid(K kernel, int threads, long long tiles) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
    long long grid = (long long)num_sms() * per_sm;
    return grid > tiles ? tiles : grid;
}

bool make_geo(ThinGeo& g, int N, int Hb, int Wb, int Ct, int Cw, int kh, int kw, int stride) {
    if (getenv("DMV_NO_THIN_MMA")) return false;
    if (stride != 2 || kh != 5 || kw != 5 || Cw != 32 || Ct < 1 || Ct > 4 || (Hb & 1) || (Wb & 1)) return false;
    g.N = N; g.Hb = Hb; g.Wb = Wb; g.Hs = Hb / 2; g.Ws = Wb / 2;
    if (g.Hs % TH || g.Ws % TW) return false;
    g.pt = same_pad(Hb, kh, stride).before;
    g.pl = same_pad(Wb, kw, stride).before;
    if (g.pt != 1 || g.pl != 1) return false;              // even sizes, k = 5, stride 2: TF SAME pads 1 before, 2 after
    g.tiles_x = g.Ws / TW; g.tiles_y = g.Hs / TH;
    g.tiles = (long long)N * g.tiles_x * g.tiles_y;
    return g.tiles < (1ll << 31) && (long long)N * Hb * Wb * 4 < (1ll << 31);      // 32-bit tile / offset arithmetic in the kernels
}
}  // namespace

namespace dmv {

bool thin_mma_eligible(int N, int Hb, int Wb, int Ct, int Cw, int kh, int kw, int stride) {
    ThinGeo g;
    return make_geo(g, N, Hb, Wb, Ct, Cw, kh, kw, stride);
}

size_t thin_mma_wgrad_workspace(int N, int Hb, int Wb, int Ct, int Cw, int kh, int kw, int stride) {
    if (!thin_mma_eligible(N, Hb, Wb, Ct, Cw, kh, kw, stride)) return 0;
    return ((size_t)num_sms() * 8 + 1) * 5 * 24 * 32 * sizeof(float) + 256;      // partials of at most 8 resident CTAs per SM
}

int thin_mma_conv_fwd(const void* thin, int thin_dtype, const void* w_bf16, const float* bias, void* out, int out_dtype, int N, int Hb, int Wb,
                      int Ct, int Cw, int kh, int kw, int stride, int act, cudaStream_t st) {
    ThinGeo g;
    if (thin_dtype != DMV_DT_F32 || out_dtype != DMV_DT_BF16 || !make_geo(g, N, Hb, Wb, Ct, Cw, kh, kw, stride))
        return fail(DMV_E_UNSUPPORTED_SHAPE, "thin_mma_conv_fwd: shape not covered");
    if (((uintptr_t)out & 15) || ((uintptr_t)w_bf16 & 3)) return fail(DMV_E_ALIGN, "thin_mma_conv_fwd: alignment");
    long long grid = 0;
#define DMV_THIN_CONV(CTV)                                                                                                      \
    do {                                                                                                                        \
        grid = resident_grid(thin_conv_kernel<CTV>, 256, g.tiles);                                                              \
        thin_conv_kernel<CTV><<<(unsigned)grid, 256, 0, st>>>((const float*)thin, (const bf16*)w_bf16, bias, (bf16*)out, g, act); \
    } while (0)
    switch (Ct) {
        case 1: DMV_THIN_CONV(1); break;
        case 2: DMV_THIN_CONV(2); break;
        case 3: DMV_THIN_CONV(3); break;
        default: DMV_THIN_CONV(4); break;
    }
#undef DMV_THIN_CONV
    count_tc_launch();
    return check_launch("thin_conv");
}

int thin_mma_deconv_fwd(const void* x_bf16, const void* w_bf16, void* y, int y_dtype, int N, int Hb, int Wb, int Ct, int Cw, int kh, int kw,
                        int stride, int act, cudaStream_t st) {
    ThinGeo g;
    if (!make_geo(g, N, Hb, Wb, Ct, Cw, kh, kw, stride)) return fail(DMV_E_UNSUPPORTED_SHAPE, "thin_mma_deconv_fwd: shape not covered");
    if (((uintptr_t)x_bf16 & 15) || ((uintptr_t)w_bf16 & 3) || ((uintptr_t)y & 7)) return fail(DMV_E_ALIGN, "thin_mma_deconv_fwd: alignment");
    long long grid = 0;
#define DMV_THIN_DECONV(CTV)                                                                                                          \
    do {                                                                                                                              \
        if (y_dtype == DMV_DT_F32) {                                                                                                  \
            grid = resident_grid(thin_deconv_kernel<CTV, float>, 256, g.tiles);                                                       \
            thin_deconv_kernel<CTV, float><<<(unsigned)grid, 256, 0, st>>>((const bf16*)x_bf16, (const bf16*)w_bf16, (float*)y, g, act); \
        } else {                                                                                                                      \
            grid = resident_grid(thin_deconv_kernel<CTV, bf16>, 256, g.tiles);                                                        \
            thin_deconv_kernel<CTV, bf16><<<(unsigned)grid, 256, 0, st>>>((const bf16*)x_bf16, (const bf16*)w_bf16, (bf16*)y, g, act);   \
        }                                                                                                                             \
    } while (0)
    switch (Ct) {
        case 1: DMV_THIN_DECONV(1); break;
        case 2: DMV_THIN_DECONV(2); break;
        case 3: DMV_THIN_DECONV(3); break;
        default: DMV_THIN_DECONV(4); break;
    }
#undef DMV_THIN_DECONV
    count_tc_launch();
    return check_launch("thin_deconv");
}

int thin_mma_wgrad(const void* thin, int thin_dtype, const void* wide_bf16, float* dw, int N, int Hb, int Wb, int Ct, int Cw, int kh, int kw,
                   int stride, void* ws, size_t ws_bytes, cudaStream_t st) {
    ThinGeo g;
    if (thin_dtype != DMV_DT_F32 || !make_geo(g, N, Hb, Wb, Ct, Cw, kh, kw, stride))
        return fail(DMV_E_UNSUPPORTED_SHAPE, "thin_mma_wgrad: shape not covered");
    if (((uintptr_t)wide_bf16 & 15)) return fail(DMV_E_ALIGN, "thin_mma_wgrad: alignment");
    const size_t per = (size_t)5 * 24 * 32;
    long long grid = 0;
    switch (Ct) {
        case 1: grid = resident_grid(thin_wgrad_mma_kernel<1>, 160, g.tiles); break;
        case 2: grid = resident_grid(thin_wgrad_mma_kernel<2>, 160, g.tiles); break;
        case 3: grid = resident_grid(thin_wgrad_mma_kernel<3>, 160, g.tiles); break;
        default: grid = resident_grid(thin_wgrad_mma_kernel<4>, 160, g.tiles); break;
    }
    if (grid > (long long)num_sms() * 8) grid = (long long)num_sms() * 8;
    uint8_t* base = reinterpret_cast<uint8_t*>(((uintptr_t)ws + 255) & ~(uintptr_t)255);
    if (!ws || (size_t)(base - (uint8_t*)ws) + ((size_t)grid + 1) * per * sizeof(float) > ws_bytes)
        return fail(DMV_E_WORKSPACE, "thin_mma_wgrad: workspace too small");
    float* part = reinterpret_cast<float*>(base);
    float* sum = part + (size_t)grid * per;
    switch (Ct) {
        case 1: thin_wgrad_mma_kernel<1><<<(unsigned)grid, 160, 0, st>>>((const float*)thin, (const bf16*)wide_bf16, part, g); break;
        case 2: thin_wgrad_mma_kernel<2><<<(unsigned)grid, 160, 0, st>>>((const float*)thin, (const bf16*)wide_bf16, part, g); break;
        case 3: thin_wgrad_mma_kernel<3><<<(unsigned)grid, 160, 0, st>>>((const float*)thin, (const bf16*)wide_bf16, part, g); break;
        default: thin_wgrad_mma_kernel<4><<<(unsigned)grid, 160, 0, st>>>((const float*)thin, (const bf16*)wide_bf16, part, g); break;
    }
    count_tc_launch();
    int rc = check_launch("thin_wgrad_mma");
    if (rc) return rc;
    rc = reduce_partials(part, sum, (long long)per, (int)grid, st);
    if (rc) return rc;
    thin_wgrad_gather_kernel<<<ceil_div(25 * Ct * 32, 256), 256, 0, st>>>(sum, dw, Ct);
    return check_launch("thin_wgrad_gather");
}

}  // namespace dmv
