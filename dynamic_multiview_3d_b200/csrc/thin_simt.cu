// CUDA-core kernels for the two "thin" layers of the graph, whose channel count on the
// image side (3 for e0, 2 for the flow head) is far below tensor-core and TMA granularity
// ("treat as bandwidth kernels").
//
// thin wgrad:  dW[(tap, ct), cw] = sum_{pixels} Thin[n, oh*st + r - pt, ow*st + s - pl, ct] * Wide[n, oh, ow, cw]
//   (conv e0: Thin = fp32 image, Wide = dY;  deconv flow head: Thin = fp32 flow gradient, Wide = x)
// One WARP owns the complete [rows x 32] output in registers: lane = (row group, column
// group) holds R x 8 accumulators and streams pixels, loading its 8 wide channels with one
// 16-byte load and its R thin values through L1.  Warps/CTAs take interleaved pixels; CTA
// partials are summed in CTA order by the shared split-K reduce (deterministic).
#include <stdlib.h>

#include "common.cuh"
#include "conv_impl.h"

namespace {
using namespace dmv;
typedef __nv_bfloat16 bf16;

struct ThinGeom {
    int N, Hb, Wb, Ct, Hs, Ws, kh, kw, st, pt, pl, rows;   // rows = kh*kw*Ct
};

template <typename TT, int R>
__global__ void __launch_bounds__(256, 2) thin_wgrad_kernel(const TT* __restrict__ thin, const bf16* __restrict__ wide,
                                                          float* __restrict__ part, ThinGeom g, long long pixels_per_cta) {
    __shared__ float red[8][32][9];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rg = lane >> 2, cg = lane & 3;
    // per-lane row table: element offset of the tap/channel from the window origin, and (dy+64, dx+64, valid) packed
    int r_off[R], r_pk[R];
#pragma unroll
    for (int i = 0; i < R; ++i) {
        const int row = rg * R + i;
        const bool ok = row < g.rows;
        const int tap = ok ? row / g.Ct : 0;
        const int c = ok ? row - tap * g.Ct : 0;
        const int dy = tap / g.kw - g.pt, dx = tap % g.kw - g.pl;
        r_off[i] = (dy * g.Wb + dx) * g.Ct + c;
        r_pk[i] = (dy + 64) | ((dx + 64) << 8) | (ok ? (1 << 16) : 0);
    }
    float acc[R][8];
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[i][k] = 0.f;

    const long long total = (long long)g.N * g.Hs * g.Ws;
    const long long p0 = (long long)blockIdx.x * pixels_per_cta;
    const long long p1 = min(total, p0 + pixels_per_cta);
    for (long long p = p0 + warp; p < p1; p += 8) {
        const int ow = (int)(p % g.Ws);
        const int oh = (int)((p / g.Ws) % g.Hs);
        const int n = (int)(p / ((long long)g.Ws * g.Hs));
        const uint4 wv = __ldg(reinterpret_cast<const uint4*>(wide + p * 32 + cg * 8));
        float wf[8];
        {
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&wv);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float2 t = __bfloat1622float2(h[k]);
                wf[2 * k] = t.x;
                wf[2 * k + 1] = t.y;
            }
        }
        const int by = oh * g.st, bx = ow * g.st;
        const TT* base = thin + (((long long)n * g.Hb + by) * g.Wb + bx) * g.Ct;
        const bool interior = (by - g.pt >= 0) && (bx - g.pl >= 0) && (by + g.kh - 1 - g.pt < g.Hb) && (bx + g.kw - 1 - g.pl < g.Wb);
#pragma unroll
        for (int i = 0; i < R; ++i) {
            float tv = 0.f;
            if (r_pk[i] >> 16) {
                const int y = by + (r_pk[i] & 255) - 64, x = bx + ((r_pk[i] >> 8) & 255) - 64;
                if (interior || (y >= 0 && y < g.Hb && x >= 0 && x < g.Wb)) tv = load_as_float(base + r_off[i]);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[i][k] = fmaf(tv, wf[k], acc[i][k]);
        }
    }
    // fixed-order reduction over the 8 warps, row by row
    float* out = part + (long long)blockIdx.x * g.rows * 32;
#pragma unroll
    for (int i = 0; i < R; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) red[warp][lane][k] = acc[i][k];
        __syncthreads();
        if (warp == 0) {
            const int row = rg * R + i;
            if (row < g.rows) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float s = 0.f;
#pragma unroll
                    for (int w = 0; w < 8; ++w) s += red[w][lane][k];
                    out[(long long)row * 32 + cg * 8 + k] = s;
                }
            }
        }
        __syncthreads();
    }
}

__global__ void thin_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, int n, int parts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int z = 0; z < parts; ++z) s += part[(long long)z * n + i];
    out[i] = s;
}

// Patch matrix of a thin tensor: P[pixel][k], k = (tap, ct) in reference weight order, zero padded to Kp columns.
// One CTA builds the patches of ONE output row: the kh input rows it touches are staged in shared memory
// (coalesced, zero padded left/right/top/bottom exactly like TF-SAME), then every thread assembles 16-byte
// chunks of P through a per-column offset table.  HBM-bound: reads the thin tensor once, writes P once.
template <typename TT>
__global__ void __launch_bounds__(256) thin_im2col_kernel(const TT* __restrict__ thin, bf16* __restrict__ P, ThinGeom g, int Kp) {
    extern __shared__ float s_rows[];             // [kh][pitch] then int koff[Kp]
    const int pr = max((g.Ws - 1) * g.st + g.kw - g.Wb - g.pl, 0);
    const int pitch = (g.pl + g.Wb + pr) * g.Ct;
    int* koff = reinterpret_cast<int*>(s_rows + g.kh * pitch);
    const int oh = blockIdx.x % g.Hs, n = blockIdx.x / g.Hs;
    for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
        int off = -1;
        if (k < g.rows) {
            const int tap = k / g.Ct, c = k - tap * g.Ct;
            off = (tap / g.kw) * pitch + (tap % g.kw) * g.Ct + c;
        }
        koff[k] = off;
    }
    const int row_elems = g.Wb * g.Ct;
    for (int r = 0; r < g.kh; ++r) {
        const int y = oh * g.st + r - g.pt;
        float* dst = s_rows + r * pitch;
        const bool in = (y >= 0 && y < g.Hb);
        const TT* src = thin + ((long long)n * g.Hb + (in ? y : 0)) * row_elems;
        for (int e = threadIdx.x; e < pitch; e += blockDim.x) {
            const int j = e - g.pl * g.Ct;
            dst[e] = (in && j >= 0 && j < row_elems) ? load_as_float(src + j) : 0.f;
        }
    }
    __syncthreads();
    const int groups = Kp / 8;
    bf16* out = P + ((long long)n * g.Hs + oh) * g.Ws * Kp;
    for (int i = threadIdx.x; i < g.Ws * groups; i += blockDim.x) {
        const int ow = i / groups, gq = i - ow * groups;
        const int base = ow * g.st * g.Ct;
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int off = koff[gq * 8 + k];
            v[k] = off >= 0 ? s_rows[base + off] : 0.f;
        }
        uint4 q;
        __nv_bfloat162 h;
        h = __floats2bfloat162_rn(v[0], v[1]); q.x = *reinterpret_cast<uint32_t*>(&h);
        h = __floats2bfloat162_rn(v[2], v[3]); q.y = *reinterpret_cast<uint32_t*>(&h);
        h = __floats2bfloat162_rn(v[4], v[5]); q.z = *reinterpret_cast<uint32_t*>(&h);
        h = __floats2bfloat162_rn(v[6], v[7]); q.w = *reinterpret_cast<uint32_t*>(&h);
        *reinterpret_cast<uint4*>(out + (long long)i * 8) = q;
    }
}

// weights w[rows][Cw] (reference order) -> packed[Cw][Kp] K-major, zero padded
__global__ void thin_pack_kernel(const bf16* __restrict__ w, bf16* __restrict__ out, int rows, int Cw, int Kp) {
    const int n = Cw * Kp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int k = i % Kp, co = i / Kp;
        out[i] = k < rows ? w[(long long)k * Cw + co] : __float2bfloat16_rn(0.f);
    }
}

__global__ void copy_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

template <typename TT>
int run_thin(const TT* thin, const bf16* wide, float* dw, const ThinGeom& g, void* ws, size_t ws_bytes, cudaStream_t st) {
    const long long total = (long long)g.N * g.Hs * g.Ws;
    int ctas = 148 * 4;
    if (ctas > (int)ceil_div_ll(total, 64)) ctas = (int)ceil_div_ll(total, 64);
    if (ctas < 1) ctas = 1;
    const long long per = ceil_div_ll(total, ctas);
    ctas = (int)ceil_div_ll(total, per);
    const size_t need = (size_t)ctas * g.rows * 32 * sizeof(float);
    if (!ws || ws_bytes < need) return fail(DMV_E_WORKSPACE, "thin wgrad: workspace too small");
    float* part = reinterpret_cast<float*>(ws);
    const int R = ceil_div(g.rows, 8);
    if (R <= 7) thin_wgrad_kernel<TT, 7><<<ctas, 256, 0, st>>>(thin, wide, part, g, per);
    else if (R <= 10) thin_wgrad_kernel<TT, 10><<<ctas, 256, 0, st>>>(thin, wide, part, g, per);
    else return fail(DMV_E_UNSUPPORTED_SHAPE, "thin wgrad: too many rows");
    int rc = check_launch("thin_wgrad");
    if (rc) return rc;
    const int n = g.rows * 32;
    thin_reduce_kernel<<<ceil_div(n, 256), 256, 0, st>>>(part, dw, n, ctas);
    return check_launch("thin_wgrad reduce");
}

// ---- space-to-depth path of the thin stride-2 layers -------------------------------------------------------
// X2[n][y][x][(ph*2 + pw)*Ct + c] = thin[n][2y + ph][2x + pw][c], channels 4*Ct .. 31 zero.  A CTA stages one pair of
// thin rows (contiguous in memory) in shared memory with 16-byte loads; four threads build one X2 pixel, one 16-byte
// channel group each, so the stores of a warp are 512 contiguous bytes.
template <typename TT>
__global__ void __launch_bounds__(256) thin_s2d_prep_kernel(const TT* __restrict__ thin, uint4* __restrict__ X2, int N, int Hb, int Wb, int Ct) {
    extern __shared__ float s_rows[];             // [2][Wb * Ct]
    const int H2 = Hb >> 1, W2 = Wb >> 1, row_elems = Wb * Ct, q = threadIdx.x & 3;
    int off[8];                                   // source offset of channel 8q + k relative to pixel 2x of row 2y; -1: zero
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int ch = 8 * q + k, blk = ch / Ct, c = ch - blk * Ct;
        off[k] = (ch < 4 * Ct) ? (blk >> 1) * row_elems + (blk & 1) * Ct + c : -1;
    }
    for (long long rp = blockIdx.x; rp < (long long)N * H2; rp += gridDim.x) {
        const TT* src = thin + rp * 2 * row_elems;        // rows 2y and 2y + 1 of image n are adjacent
        if (sizeof(TT) == 4 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0) && ((2 * row_elems) & 3) == 0) {
            const float4* s4 = reinterpret_cast<const float4*>(src);
            for (int e = threadIdx.x; e < (2 * row_elems) >> 2; e += 256) reinterpret_cast<float4*>(s_rows)[e] = __ldg(s4 + e);
        } else {
            for (int e = threadIdx.x; e < 2 * row_elems; e += 256) s_rows[e] = load_as_float(src + e);
        }
        __syncthreads();
        for (int o = threadIdx.x; o < W2 * 4; o += 256) {
            const int base = (o >> 2) * 2 * Ct;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) v[k] = off[k] >= 0 ? s_rows[base + off[k]] : 0.f;
            uint4 u;
            __nv_bfloat162 h;
            h = __floats2bfloat162_rn(v[0], v[1]); u.x = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2bfloat162_rn(v[2], v[3]); u.y = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2bfloat162_rn(v[4], v[5]); u.z = *reinterpret_cast<uint32_t*>(&h);
            h = __floats2bfloat162_rn(v[6], v[7]); u.w = *reinterpret_cast<uint32_t*>(&h);
            X2[rp * W2 * 4 + o] = u;
        }
        __syncthreads();
    }
}

// reference tap (r, s) and thin channel of entry (shift j, ch') of the s2d weight matrix; false: structural zero
__device__ __forceinline__ bool s2d_entry(int j, int ch, int Ct, int kh, int kw, int pt, int pl, dmv::S2dGeom g, int& tap, int& c) {
    if (ch >= 4 * Ct) return false;
    const int dh = g.dh0 + j / g.kw2, dw = g.dw0 + j % g.kw2;
    const int blk = ch / Ct;
    c = ch - blk * Ct;
    const int r = 2 * dh + (blk >> 1) + pt, s = 2 * dw + (blk & 1) + pl;
    if (r < 0 || r >= kh || s < 0 || s >= kw) return false;
    tap = r * kw + s;
    return true;
}

__global__ void thin_s2d_pack_kernel(const bf16* __restrict__ w, bf16* __restrict__ out, int Ct, int Cw, int kh, int kw, int pt, int pl,
                                     dmv::S2dGeom g) {
    const int K = g.kh2 * g.kw2 * 32, total = Cw * K;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int cw = i / K, kk = i - cw * K;
        int tap, c;
        bf16 v = __float2bfloat16_rn(0.f);
        if (s2d_entry(kk >> 5, kk & 31, Ct, kh, kw, pt, pl, g, tap, c)) v = w[((long long)tap * Ct + c) * Cw + cw];
        out[i] = v;
    }
}

__global__ void thin_s2d_gather_kernel(const float* __restrict__ dw2, float* __restrict__ dw, int Ct, int Cw, int kh, int kw, int pt, int pl,
                                       dmv::S2dGeom g) {
    const int total = g.kh2 * g.kw2 * 32 * Cw;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int cw = i % Cw, kk = i / Cw;
        int tap, c;
        if (s2d_entry(kk >> 5, kk & 31, Ct, kh, kw, pt, pl, g, tap, c)) dw[((long long)tap * Ct + c) * Cw + cw] = dw2[i];
    }
}

}  // namespace

namespace dmv {


static int floor_div2(int a) { return a >= 0 ? a / 2 : -((-a + 1) / 2); }

S2dGeom thin_s2d_geom(int Hb, int Wb, int kh, int kw) {
    const SamePad ph = same_pad(Hb, kh, 2), pw = same_pad(Wb, kw, 2);
    S2dGeom g;
    g.dh0 = floor_div2(-ph.before);
    g.kh2 = floor_div2(kh - 1 - ph.before) - g.dh0 + 1;
    g.dw0 = floor_div2(-pw.before);
    g.kw2 = floor_div2(kw - 1 - pw.before) - g.dw0 + 1;
    return g;
}

bool thin_s2d_eligible(int Hb, int Wb, int Ct, int Cw, int kh, int kw, int stride) {
    if (getenv("DMV_NO_S2D")) return false;
    if (stride != 2 || (Hb & 1) || (Wb & 1) || 4 * Ct > 32 || Ct < 1) return false;
    if (!(Cw == 32 || Cw == 64)) return false;
    if ((Hb / 2) * (Wb / 2) < 256) return false;       // below this the halo kernels are not used
    const S2dGeom g = thin_s2d_geom(Hb, Wb, kh, kw);
    return g.kh2 * g.kw2 > 1 && g.kh2 <= 3 && g.kw2 <= 3;
}

int thin_s2d_prep(const void* thin, int thin_dtype, void* X2, int N, int Hb, int Wb, int Ct, cudaStream_t st) {
    long long blocks = (long long)N * (Hb / 2);
    if (blocks > 148 * 8) blocks = 148 * 8;
    const size_t smem = (size_t)2 * Wb * Ct * sizeof(float);
    if (smem > 48 * 1024) return fail(DMV_E_UNSUPPORTED_SHAPE, "thin_s2d_prep: row pair does not fit shared memory");
    if (thin_dtype == DMV_DT_F32)
        thin_s2d_prep_kernel<float><<<(int)blocks, 256, smem, st>>>((const float*)thin, (uint4*)X2, N, Hb, Wb, Ct);
    else
        thin_s2d_prep_kernel<bf16><<<(int)blocks, 256, smem, st>>>((const bf16*)thin, (uint4*)X2, N, Hb, Wb, Ct);
    return check_launch("thin_s2d_prep");
}

int thin_s2d_pack_weights(const void* w, void* packed, int Ct, int Cw, int kh, int kw, int pt, int pl, S2dGeom g, cudaStream_t st) {
    thin_s2d_pack_kernel<<<ceil_div(Cw * g.kh2 * g.kw2 * 32, 256), 256, 0, st>>>((const bf16*)w, (bf16*)packed, Ct, Cw, kh, kw, pt, pl, g);
    return check_launch("thin_s2d_pack");
}

int thin_s2d_gather_dw(const float* dw2, float* dw, int Ct, int Cw, int kh, int kw, int pt, int pl, S2dGeom g, cudaStream_t st) {
    thin_s2d_gather_kernel<<<ceil_div(Cw * g.kh2 * g.kw2 * 32, 256), 256, 0, st>>>(dw2, dw, Ct, Cw, kh, kw, pt, pl, g);
    return check_launch("thin_s2d_gather");
}

int thin_patch_cols(int taps, int Ct) { return ceil_div(taps * Ct, 32) * 32; }

int thin_im2col(const void* thin, int thin_dtype, void* P, int N, int Hb, int Wb, int Ct, int kh, int kw, int stride, cudaStream_t st) {
    const SamePad ph = same_pad(Hb, kh, stride), pw = same_pad(Wb, kw, stride);
    ThinGeom g;
    g.N = N; g.Hb = Hb; g.Wb = Wb; g.Ct = Ct; g.Hs = ph.out; g.Ws = pw.out; g.kh = kh; g.kw = kw; g.st = stride;
    g.pt = ph.before; g.pl = pw.before; g.rows = kh * kw * Ct;
    const int Kp = thin_patch_cols(kh * kw, Ct);
    const int prr = (g.Ws - 1) * stride + kw - Wb - g.pl;
    const size_t smem = (size_t)kh * (g.pl + Wb + (prr > 0 ? prr : 0)) * Ct * sizeof(float) + (size_t)Kp * sizeof(int);
    if (smem > 200 * 1024) return fail(DMV_E_UNSUPPORTED_SHAPE, "thin_im2col: input rows do not fit shared memory");
    const int blocks = N * g.Hs;
    if (thin_dtype == DMV_DT_F32) {
        cudaFuncSetAttribute(thin_im2col_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        thin_im2col_kernel<float><<<blocks, 256, smem, st>>>((const float*)thin, (bf16*)P, g, Kp);
    } else {
        cudaFuncSetAttribute(thin_im2col_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        thin_im2col_kernel<bf16><<<blocks, 256, smem, st>>>((const bf16*)thin, (bf16*)P, g, Kp);
    }
    return check_launch("thin_im2col");
}

int thin_pack_weights(const void* w_bf16, void* packed, int rows, int Cw, int Kp, cudaStream_t st) {
    thin_pack_kernel<<<ceil_div(Cw * Kp, 256), 256, 0, st>>>((const bf16*)w_bf16, (bf16*)packed, rows, Cw, Kp);
    return check_launch("thin_pack");
}

int copy_f32(const float* src, float* dst, int n, cudaStream_t st) {
    copy_rows_kernel<<<ceil_div(n, 256), 256, 0, st>>>(src, dst, n);
    return check_launch("copy_f32");
}

size_t thin_wgrad_workspace(int taps, int Ct) { return (size_t)148 * 4 * taps * Ct * 32 * sizeof(float) + 256; }

bool thin_wgrad_eligible(int taps, int Ct, int Cw) { return Cw == 32 && Ct <= 4 && taps * Ct <= 80; }

// thin_dtype: DMV_DT_F32 or DMV_DT_BF16
int thin_wgrad(const void* thin, int thin_dtype, const void* wide_bf16, float* dw, int N, int Hb, int Wb, int Ct, int kh, int kw, int stride,
               void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!thin_wgrad_eligible(kh * kw, Ct, 32)) return fail(DMV_E_UNSUPPORTED_SHAPE, "thin wgrad: shape not covered");
    if ((uintptr_t)wide_bf16 & 15) return fail(DMV_E_UNSUPPORTED_SHAPE, "thin wgrad: wide tensor must be 16-byte aligned");
    const SamePad ph = same_pad(Hb, kh, stride), pw = same_pad(Wb, kw, stride);
    ThinGeom g;
    g.N = N; g.Hb = Hb; g.Wb = Wb; g.Ct = Ct; g.Hs = ph.out; g.Ws = pw.out; g.kh = kh; g.kw = kw; g.st = stride;
    g.pt = ph.before; g.pl = pw.before; g.rows = kh * kw * Ct;
    if (thin_dtype == DMV_DT_F32) return run_thin<float>((const float*)thin, (const bf16*)wide_bf16, dw, g, ws, ws_bytes, st);
    return run_thin<bf16>((const bf16*)thin, (const bf16*)wide_bf16, dw, g, ws, ws_bytes, st);
}

}  // namespace dmv
