// Weight gradient of a linear layer consumed in place by TF-Adam (include/dmv3d.h: dmv_linear_wgrad_adam).
//
// The FC matrices are 98 % of the parameters of the appearance-flow graph and every byte of them is streamed from HBM
// once per pass.  Separately, the weight gradient writes 4 B/parameter that Adam reads back 4 B/parameter later; here
// the tile of dW = x^T dy (contraction over the batch, 64 samples) is formed in registers from the two small bf16
// operands (both L2-resident) and ApplyAdam runs on the accumulators: 26 B/parameter instead of 34 (theta, m, v read;
// theta, m, v, bf16 copy written).  The kernel is HBM-bound, so the contraction uses warp-level mma.sync (bf16 in, fp32
// accumulate) -- its ~17 GFLOP per step are noise next to the 3.5 GB stream, and the accumulator layout is chosen for
// the stream (see "Column ownership" below).
//
// Two kernels: a streaming form for the training shapes (persistent CTAs, parameters arrive by bulk async copies into a
// shared-memory ring) and a generic form (any M, tails).  Operands sit in shared memory as [sample][k] and [sample][n] and
// are read with ldmatrix.trans (both are stored with the contraction index outermost).
#include <stdlib.h>

#include "common.cuh"

namespace {
using namespace dmv;

constexpr int TK = 64, TN = 128, TC = 64;     // tile rows (K), tile columns (N), contraction chunk (samples)
constexpr int XP = TK + 8, DP = TN + 8;       // shared pitches in elements (+16 B: conflict-free ldmatrix rows)

__device__ __forceinline__ void ldsm_x4_t(unsigned& r0, unsigned& r1, unsigned& r2, unsigned& r3, const void* p) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---------------------------------------------------------------------------------------------------------------
// Column ownership.  An m16n8k16 accumulator gives thread (g, t) columns 2t, 2t+1 of an 8-column tile.  The dy tile is
// stored with its columns permuted inside every 16-column group (actual column 4q + 2j + e at position 8j + 2q + e), so
// that the two MMA tiles j = 0, 1 of a group hand thread t the four consecutive columns 4t .. 4t+3: every global access
// of the parameter stream is a float4, a quad covers 64 contiguous bytes and a warp instruction 16 full 32-byte sectors.
// A 16-byte unit of a dy row (8 columns, h = unit & 1 inside its group) scatters as four 4-byte words:
__device__ __forceinline__ void store_dy_unit(__nv_bfloat16* row, int u, uint4 v) {
    unsigned* d = reinterpret_cast<unsigned*>(row + (u >> 1) * 16) + 2 * (u & 1);     // word index 4 (w % 2) + 2 h + w / 2
    d[0] = v.x; d[4] = v.y; d[1] = v.z; d[5] = v.w;
}

__device__ __forceinline__ void adam4_store(float4 p, float4 m, float4 v, const float (&gr)[4], float* theta, float* mom, float* vel,
                                            __nv_bfloat16* half, float* dw_out, long long o, float lr_t, float omb1, float omb2, float eps,
                                            float gscale) {
    if (dw_out) *reinterpret_cast<float4*>(dw_out + o) = make_float4(gr[0], gr[1], gr[2], gr[3]);
    adam_one(p.x, __fmul_rn(gr[0], gscale), m.x, v.x, lr_t, omb1, omb2, eps);
    adam_one(p.y, __fmul_rn(gr[1], gscale), m.y, v.y, lr_t, omb1, omb2, eps);
    adam_one(p.z, __fmul_rn(gr[2], gscale), m.z, v.z, lr_t, omb1, omb2, eps);
    adam_one(p.w, __fmul_rn(gr[3], gscale), m.w, v.w, lr_t, omb1, omb2, eps);
    *reinterpret_cast<float4*>(theta + o) = p;
    *reinterpret_cast<float4*>(mom + o) = m;
    *reinterpret_cast<float4*>(vel + o) = v;
    if (half) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(p.x, p.y), hi = __floats2bfloat162_rn(p.z, p.w);
        uint2 pk;
        pk.x = *reinterpret_cast<unsigned*>(&lo);
        pk.y = *reinterpret_cast<unsigned*>(&hi);
        *reinterpret_cast<uint2*>(half + o) = pk;
    }
}

// Generic form: any M (sample chunks of 64), row / column tails.  One 64 x 128 tile per CTA; the parameter stream goes
// through registers.  PF = how many 16-column groups ahead a thread requests its theta / m / v (each group: 2 rows x 3
// arrays x float4 = 24 registers): PF = 4 puts the whole tile's 24 loads in flight before the contraction starts.
template <int PF>
__global__ void __launch_bounds__(256, PF >= 2 ? 1 : 2)
fc_wgrad_adam_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* __restrict__ theta,
                     float* __restrict__ mom, float* __restrict__ vel, __nv_bfloat16* __restrict__ half, float* __restrict__ dw_out,
                     int M, int K, int N, const float* __restrict__ state, float omb1, float omb2, float eps, float gscale) {
    __shared__ __align__(16) __nv_bfloat16 xs[TC * XP];
    __shared__ __align__(16) __nv_bfloat16 ds[TC * DP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int kr = warp & 3, nh = warp >> 2;
    const int g = lane >> 2, t = lane & 3;
    const int k0 = blockIdx.y * TK, n0 = blockIdx.x * TN;
    const int mat = lane >> 3, r = lane & 7;

    float4 P[4][2], Mo[4][2], Ve[4][2];
    auto ok = [&](int G, int h) { return n0 + nh * 64 + G * 16 + 4 * t < N && k0 + kr * 16 + g + 8 * h < K; };
    auto off = [&](int G, int h) { return (long long)(k0 + kr * 16 + g + 8 * h) * N + n0 + nh * 64 + G * 16 + 4 * t; };
    auto request = [&](int G) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
            if (ok(G, h)) {
                const long long o = off(G, h);
                P[G][h] = *reinterpret_cast<const float4*>(theta + o);
                Mo[G][h] = *reinterpret_cast<const float4*>(mom + o);
                Ve[G][h] = *reinterpret_cast<const float4*>(vel + o);
            }
    };
#pragma unroll
    for (int G = 0; G < 4; ++G)
        if (G < PF) request(G);

    float acc[4][2][4];
#pragma unroll
    for (int G = 0; G < 4; ++G)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[G][j][e] = 0.f;

    for (int c0 = 0; c0 < M; c0 += TC) {
        const int rows = min(TC, M - c0);
        if (c0) __syncthreads();
        for (int i = tid; i < TC * (TK / 8); i += 256) {            // x tile: [sample][64 k]
            const int c = i >> 3, q = i & 7;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (c < rows && k0 + q * 8 < K) v = __ldg(reinterpret_cast<const uint4*>(x + (long long)(c0 + c) * K + k0 + q * 8));
            *reinterpret_cast<uint4*>(xs + c * XP + q * 8) = v;
        }
        for (int i = tid; i < TC * (TN / 8); i += 256) {            // dy tile: [sample][128 n], permuted
            const int c = i >> 4, u = i & 15;
            uint4 v = make_uint4(0, 0, 0, 0);
            if (c < rows && n0 + u * 8 < N) v = __ldg(reinterpret_cast<const uint4*>(dy + (long long)(c0 + c) * N + n0 + u * 8));
            store_dy_unit(ds + c * DP, u, v);
        }
        __syncthreads();
        const int ksteps = (rows + 15) >> 4;
        for (int s = 0; s < ksteps; ++s) {
            unsigned a[4];
            ldsm_x4_t(a[0], a[1], a[2], a[3], xs + (s * 16 + r + (mat >> 1) * 8) * XP + kr * 16 + (mat & 1) * 8);   // A = x^T
#pragma unroll
            for (int G = 0; G < 4; ++G) {
                unsigned b0, b1, b2, b3;
                ldsm_x4_t(b0, b1, b2, b3, ds + (s * 16 + r + (mat & 1) * 8) * DP + nh * 64 + G * 16 + (mat >> 1) * 8);
                mma_bf16(acc[G][0], a, b0, b1);
                mma_bf16(acc[G][1], a, b2, b3);
            }
        }
    }

    const float lr_t = __ldg(state + 2);
#pragma unroll
    for (int G = 0; G < 4; ++G) {
        if (G + PF < 4 || PF == 0) request(PF == 0 ? G : G + PF);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            if (!ok(G, h)) continue;
            const float gr[4] = {acc[G][0][2 * h], acc[G][0][2 * h + 1], acc[G][1][2 * h], acc[G][1][2 * h + 1]};
            adam4_store(P[G][h], Mo[G][h], Ve[G][h], gr, theta, mom, vel, half, dw_out, off(G, h), lr_t, omb1, omb2, eps, gscale);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Streaming form (the training shapes: M <= 64, K % 32 == 0, N % 128 == 0).  Persistent CTAs walk 32 x 128 tiles down
// column strips (the permuted dy tile of a strip is loaded once and reused for every tile of the strip).  theta, m, v
// tiles and the x tile arrive by 16-byte cp.async copies (L2 -> shared, no registers held) into a 3-stage ring: two
// tiles (104 KB) are in flight per SM while the third is consumed.  Row pitches are padded (576 B / 80 B) so that the
// float4 reads and the ldmatrix rows are conflict-free.  (Row-wise cp.async.bulk copies were measured first: 160 small
// bulk requests per tile ran at 2 TB/s.)
//
// STAGES = 3 (194 KB) is the fastest form in isolation; STAGES = 2 (138 KB, one tile in flight while one is consumed) leaves
// 88 KB of the SM to a co-resident CTA of the input-gradient chain (DMV_FC_ADAM_STAGES; measurements in
// profiles/r02_fc_coresidency.txt).
constexpr int SR = 32, SC = 128, STREAM_THREADS = 512;
constexpr int PP = SC + 16;                   // floats per parameter row in shared memory (576 B)
constexpr int SXP = SR + 8;                   // bf16 per x row (80 B)
constexpr size_t STREAM_D_BYTES = (size_t)TC * DP * sizeof(__nv_bfloat16);
constexpr size_t stream_par_bytes(int stages) { return (size_t)stages * 3 * SR * PP * sizeof(float); }
constexpr size_t stream_x_bytes(int stages) { return (size_t)stages * TC * SXP * sizeof(__nv_bfloat16); }
constexpr size_t stream_smem(int stages) { return stream_par_bytes(stages) + stream_x_bytes(stages) + STREAM_D_BYTES; }

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N_) : "memory"); }

template <int STAGES>
__global__ void __launch_bounds__(STREAM_THREADS, 1)
fc_wgrad_adam_stream_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy, float* __restrict__ theta,
                            float* __restrict__ mom, float* __restrict__ vel, __nv_bfloat16* __restrict__ half, float* __restrict__ dw_out,
                            int M, int K, int N, const float* __restrict__ state, float omb1, float omb2, float eps, float gscale) {
    constexpr size_t STREAM_PAR_BYTES = stream_par_bytes(STAGES), STREAM_X_BYTES = stream_x_bytes(STAGES);
    extern __shared__ __align__(128) unsigned char smem[];
    float* par = reinterpret_cast<float*>(smem);                                                  // [STAGES][3][SR][PP]
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(smem + STREAM_PAR_BYTES);                  // [STAGES][64][SXP]
    __nv_bfloat16* ds = reinterpret_cast<__nv_bfloat16*>(smem + STREAM_PAR_BYTES + STREAM_X_BYTES);   // [64][DP]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wk = warp & 1, wn = warp >> 1;            // 2 x 8 warps: 16 rows x 16 columns each (one column group)
    const int g = lane >> 2, t = lane & 3, mat = lane >> 3, r = lane & 7;
    const int KT = K / SR;
    const long long T = (long long)(N / SC) * KT;
    const long long first = T * blockIdx.x / gridDim.x, last = T * (blockIdx.x + 1) / gridDim.x;
    const int ntiles = (int)(last - first);
    if (ntiles <= 0) return;

    // sample rows >= M of the x tiles stay zero for the whole kernel (the copies only write rows < M)
    for (int i = tid; i < (int)(STREAM_X_BYTES / 16); i += STREAM_THREADS) reinterpret_cast<uint4*>(xs)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();

    // a tile's copies: 3 arrays x 32 rows x 32 chunks of 16 B (6 per thread) + M rows x 4 chunks of the x tile
    const int prow = tid >> 5, pch = tid & 31;          // this thread copies chunk pch of rows prow + 16 j
    auto issue = [&](int i) {
        if (i < ntiles) {
            const long long tl = first + i;
            const int strip = (int)(tl / KT), kt = (int)(tl % KT), s = i % STAGES;
            const long long base = (long long)kt * SR * N + (long long)strip * SC + pch * 4;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int row = prow + 16 * j;
                const long long o = base + (long long)row * N;
                float* d = par + (s * 3 * SR + row) * PP + pch * 4;
                cp_async16(d, theta + o);
                cp_async16(d + SR * PP, mom + o);
                cp_async16(d + 2 * SR * PP, vel + o);
            }
            const int c = tid >> 2, q = tid & 3;
            if (tid < 4 * TC && c < M) cp_async16(xs + (s * TC + c) * SXP + q * 8, x + (long long)c * K + kt * SR + q * 8);
        }
        cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < STAGES; ++i) issue(i);

    const float lr_t = __ldg(state + 2);
    const int ksteps = (M + 15) >> 4;
    int cur_strip = -1;
    for (int i = 0; i < ntiles; ++i) {
        const long long tl = first + i;
        const int strip = (int)(tl / KT), kt = (int)(tl % KT), s = i % STAGES;
        if (strip != cur_strip) {                  // (every warp passed the __syncthreads that ended the previous tile)
            cur_strip = strip;
            for (int j = tid; j < TC * (SC / 8); j += STREAM_THREADS) {
                const int c = j >> 4, u = j & 15;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (c < M) v = __ldg(reinterpret_cast<const uint4*>(dy + (long long)c * N + strip * SC + u * 8));
                store_dy_unit(ds + c * DP, u, v);
            }
        }
        cp_async_wait<STAGES - 1>();               // this thread's copies of tile i have landed ...
        __syncthreads();                           // ... and everybody else's

        float acc[2][4];
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
        const __nv_bfloat16* xt = xs + s * TC * SXP;
        for (int q = 0; q < ksteps; ++q) {
            unsigned a[4];
            ldsm_x4_t(a[0], a[1], a[2], a[3], xt + (q * 16 + r + (mat >> 1) * 8) * SXP + wk * 16 + (mat & 1) * 8);
            unsigned b0, b1, b2, b3;
            ldsm_x4_t(b0, b1, b2, b3, ds + (q * 16 + r + (mat & 1) * 8) * DP + wn * 16 + (mat >> 1) * 8);
            mma_bf16(acc[0], a, b0, b1);
            mma_bf16(acc[1], a, b2, b3);
        }
        const float* pt = par + (s * 3) * SR * PP;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
                const int row = wk * 16 + g + 8 * h, col = wn * 16 + 4 * t;
                const long long o = ((long long)kt * SR + row) * N + (long long)strip * SC + col;
                const float gr[4] = {acc[0][2 * h], acc[0][2 * h + 1], acc[1][2 * h], acc[1][2 * h + 1]};
                adam4_store(*reinterpret_cast<const float4*>(pt + row * PP + col), *reinterpret_cast<const float4*>(pt + (SR + row) * PP + col),
                            *reinterpret_cast<const float4*>(pt + (2 * SR + row) * PP + col), gr, theta, mom, vel, half, dw_out, o, lr_t, omb1,
                            omb2, eps, gscale);
            }
        __syncthreads();                           // stage s has been read by everyone: refill it
        issue(i + STAGES);
    }
    cp_async_wait<0>();
}
}  // namespace

extern "C" {

int dmv_linear_wgrad_adam(const void* x_bf16, const void* dy_bf16, float* theta, float* m, float* v, void* bf16_copy, float* dw_out,
                          int M, int K, int N, const float* state4, float beta1, float beta2, float eps, float grad_scale, void* stream) {
    DMV_REQUIRE(x_bf16 && dy_bf16 && theta && m && v && state4, DMV_E_INVALID_ARG, "linear_wgrad_adam: null pointer");
    DMV_REQUIRE(M > 0 && K > 0 && N > 0, DMV_E_INVALID_ARG, "linear_wgrad_adam: bad shape");
    if ((K & 7) || (N & 7)) return dmv::fail(DMV_E_UNSUPPORTED_SHAPE, "linear_wgrad_adam: K and N must be multiples of 8");
    const uintptr_t al = (uintptr_t)x_bf16 | (uintptr_t)dy_bf16 | (uintptr_t)theta | (uintptr_t)m | (uintptr_t)v | (uintptr_t)bf16_copy |
                         (uintptr_t)dw_out;
    DMV_REQUIRE((al & 15) == 0, DMV_E_ALIGN, "linear_wgrad_adam: pointers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
    // DMV_FC_ADAM_VARIANT: 0 / 1 / 4 = generic kernel with that register prefetch depth, 9 = streaming kernel where the shape
    // allows (A/B measurements, profiles/r02_fc_adam.txt); default: see below
    static int variant = -2, sms = 0, ctas_env = 0, stages = 3;
    if (variant == -2) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const char* e = getenv("DMV_FC_ADAM_CTAS");
        ctas_env = e ? atoi(e) : 0;
        const char* s2 = getenv("DMV_FC_ADAM_STAGES");
        stages = (s2 && atoi(s2) == 2) ? 2 : 3;
        const char* g = getenv("DMV_FC_ADAM_VARIANT");
        variant = g ? atoi(g) : 9;
        if (variant == 9 &&
            (cudaFuncSetAttribute(fc_wgrad_adam_stream_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stream_smem(3)) != cudaSuccess ||
             cudaFuncSetAttribute(fc_wgrad_adam_stream_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stream_smem(2)) != cudaSuccess)) {
            cudaGetLastError();
            variant = 1;
        }
    }
    if (variant == 9 && M <= TC && K % SR == 0 && N % SC == 0) {
        const long long tiles = (long long)(K / SR) * (N / SC);
        long long grid = ctas_env > 0 ? ctas_env : sms;
        if (grid > tiles) grid = tiles;
#define DMV_FC_STREAM(ST)                                                                                                               \
    fc_wgrad_adam_stream_kernel<ST><<<(unsigned)grid, STREAM_THREADS, stream_smem(ST), st>>>(                                            \
        (const __nv_bfloat16*)x_bf16, (const __nv_bfloat16*)dy_bf16, theta, m, v, (__nv_bfloat16*)bf16_copy, dw_out, M, K, N, state4, omb1, omb2, \
        eps, grad_scale)
        if (stages == 2) DMV_FC_STREAM(2);
        else DMV_FC_STREAM(3);
#undef DMV_FC_STREAM
        return dmv::check_launch("linear_wgrad_adam (stream)");
    }
    dim3 grid((unsigned)ceil_div(N, TN), (unsigned)ceil_div(K, TK));
    DMV_REQUIRE(grid.y <= 65535u, DMV_E_UNSUPPORTED_SHAPE, "linear_wgrad_adam: K too large");
#define DMV_FC_LAUNCH(PFV)                                                                                                              \
    fc_wgrad_adam_kernel<PFV><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x_bf16, (const __nv_bfloat16*)dy_bf16, theta, m, v,             \
                                                   (__nv_bfloat16*)bf16_copy, dw_out, M, K, N, state4, omb1, omb2, eps, grad_scale)
    if (variant == 0) DMV_FC_LAUNCH(0);
    else if (variant == 4) DMV_FC_LAUNCH(4);
    else DMV_FC_LAUNCH(1);            // the best generic form (profiles/r02_fc_adam.txt): shapes the streaming kernel does not take
#undef DMV_FC_LAUNCH
    return dmv::check_launch("linear_wgrad_adam");
}
}
