// Fused reconstruction loss, its gradient, and the multi-view confidence-weighted fusion
//  in one HBM-bound pass: 12C bytes per pixel for V = 1.
//   fused = sum_v softmax_v(logits)_v * gen_v            (V > 1 only)
//   d_c   = (fused_c - target_c) * mask
//   loss  = inv_count * sum_pixels sum_c w_c * (d_c^2  |  |d_c|)
//   dL/dfused_c = inv_count * w_c * mask * (2 d_c | sign(d_c))
//   dL/dgen_v   = softmax_v * dL/dfused ;  dL/dlogit_v = softmax_v * sum_c dL/dfused_c (gen_v,c - fused_c)
// The scalar is reduced deterministically: per-CTA partials in double, summed in index
// order by the last CTA to finish.
#include "common.cuh"

namespace {
using namespace dmv;

constexpr int kThreads = 256;
constexpr int kMaxC = 8;
constexpr int kMaxV = 8;

struct LossParams {
    const float* gen;
    const float* logits;
    const float* target;
    const float* mask;
    float* grad_gen;
    float* grad_logits;
    float* fused_out;
    float* loss_out;
    double* partials;
    unsigned* counter;
    long long pixels;
    int C, V, mode;
    float inv_count;
    float w[kMaxC];
};

// Last CTA: ordered sum of the per-CTA partials by all its threads (thread t takes partials t, t + 256, ... in order,
// then a fixed tree) -- deterministic, and not one thread walking a thousand dependent loads.
__device__ __forceinline__ void final_sum(const LossParams& p, double* s_red) {
    __threadfence();
    double t = 0.0;
    for (unsigned i = threadIdx.x; i < gridDim.x; i += kThreads) t += __ldcg(p.partials + i);
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) tot += s_red[w];
        *p.loss_out = (float)(tot * (double)p.inv_count);
        *p.counter = 0;  // ready for the next launch (stream-ordered)
    }
}

template <int CT>
__global__ void __launch_bounds__(kThreads) loss_kernel(LossParams p) {
    const int C = CT ? CT : p.C;
    const int V = p.V;
    double local = 0.0;
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long px = (long long)blockIdx.x * kThreads + threadIdx.x; px < p.pixels; px += stride) {
        float sm[kMaxV];
        if (V > 1) {
            float mx = -INFINITY;
            for (int v = 0; v < V; ++v) {
                sm[v] = __ldg(p.logits + v * p.pixels + px);
                mx = fmaxf(mx, sm[v]);
            }
            float den = 0.f;
            for (int v = 0; v < V; ++v) {
                sm[v] = __expf(sm[v] - mx);
                den += sm[v];
            }
            const float inv = 1.0f / den;
            for (int v = 0; v < V; ++v) sm[v] *= inv;
        } else {
            sm[0] = 1.0f;
        }
        const float mk = p.mask ? __ldg(p.mask + px) : 1.0f;
        float gl[kMaxV];
        for (int v = 0; v < V; ++v) gl[v] = 0.f;
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < (CT ? CT : kMaxC); ++c) {
            if (!CT && c >= C) break;
            float gv[kMaxV];
            float fused = 0.f;
            for (int v = 0; v < V; ++v) {
                gv[v] = __ldg(p.gen + ((long long)v * p.pixels + px) * C + c);
                fused += sm[v] * gv[v];
            }
            if (V == 1) fused = gv[0];
            if (p.fused_out) p.fused_out[px * C + c] = fused;
            const float d = (fused - __ldg(p.target + px * C + c)) * mk;
            float g;
            if (p.mode == DMV_LOSS_L2) {
                acc += p.w[c] * d * d;
                g = 2.0f * d;
            } else {
                acc += p.w[c] * fabsf(d);
                g = (d > 0.f) ? 1.0f : (d < 0.f ? -1.0f : 0.0f);
            }
            g *= p.w[c] * mk * p.inv_count;
            for (int v = 0; v < V; ++v) {
                if (p.grad_gen) p.grad_gen[((long long)v * p.pixels + px) * C + c] = sm[v] * g;
                gl[v] += g * (gv[v] - fused);
            }
        }
        if (p.grad_logits && V > 1)
            for (int v = 0; v < V; ++v) p.grad_logits[v * p.pixels + px] = sm[v] * gl[v];
        local += (double)acc;
    }
    // block reduction (fixed tree), then ordered final sum by the last CTA
    __shared__ double s_red[kThreads / 32];
    __shared__ bool s_last;
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) t += s_red[w];
        p.partials[blockIdx.x] = t;
        __threadfence();
        const unsigned done = atomicAdd(p.counter, 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) final_sum(p, s_red);
}

// Plain (single view, unmasked) loss over the flat [pixels*C] arrays: float4 loads/stores, channel = index % C.
__global__ void __launch_bounds__(kThreads) loss_flat_kernel(LossParams p) {
    const long long n4 = (p.pixels * p.C) >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(p.gen);
    const float4* t4 = reinterpret_cast<const float4*>(p.target);
    float4* o4 = reinterpret_cast<float4*>(p.grad_gen);
    double local = 0.0;
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += stride) {
        const float4 a = __ldg(g4 + i), b = __ldg(t4 + i);
        int c = (int)((i * 4) % p.C);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
        float gv[4], acc = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float w = p.w[c];
            const float d = av[k] - bv[k];
            float g;
            if (p.mode == DMV_LOSS_L2) { acc += w * d * d; g = 2.0f * d; }
            else { acc += w * fabsf(d); g = (d > 0.f) ? 1.0f : (d < 0.f ? -1.0f : 0.0f); }
            gv[k] = g * w * p.inv_count;
            if (++c == p.C) c = 0;
        }
        if (o4) o4[i] = make_float4(gv[0], gv[1], gv[2], gv[3]);
        local += (double)acc;
    }
    __shared__ double s_red[kThreads / 32];
    __shared__ bool s_last;
    for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) t += s_red[w];
        p.partials[blockIdx.x] = t;
        __threadfence();
        const unsigned done = atomicAdd(p.counter, 1u);
        s_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) final_sum(p, s_red);
}

__global__ void scale_kernel(float* x, const float* s, long long n) {
    const float k = __ldg(s);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] *= k;
}

constexpr int kMaxBlocks = 148 * 8;
}  // namespace

extern "C" {

size_t dmv_loss_workspace_size(long long pixels) {
    (void)pixels;
    return (size_t)kMaxBlocks * sizeof(double) + 16;
}

int dmv_loss_fused_fwd_bwd(const float* gen, const float* logits, int V, const float* target, const float* mask,
                           const float* chan_weight, int mode, float inv_count, float* loss_out, float* grad_gen,
                           float* grad_logits, float* fused_out, long long pixels, int C, void* workspace,
                           size_t workspace_bytes, void* stream) {
    DMV_REQUIRE(gen && target && loss_out && chan_weight, DMV_E_INVALID_ARG, "loss: null pointer");
    DMV_REQUIRE(pixels > 0 && C > 0 && C <= kMaxC, DMV_E_UNSUPPORTED_SHAPE, "loss: need 1 <= C <= 8");
    DMV_REQUIRE(V >= 1 && V <= kMaxV, DMV_E_UNSUPPORTED_SHAPE, "loss: need 1 <= V <= 8");
    DMV_REQUIRE(V == 1 || logits, DMV_E_INVALID_ARG, "loss: V > 1 needs confidence logits");
    DMV_REQUIRE(mode == DMV_LOSS_L2 || mode == DMV_LOSS_L1, DMV_E_INVALID_ARG, "loss: unknown mode");
    DMV_REQUIRE(workspace && workspace_bytes >= dmv_loss_workspace_size(pixels), DMV_E_WORKSPACE, "loss: workspace too small");
    DMV_REQUIRE(((uintptr_t)workspace & 7) == 0, DMV_E_ALIGN, "loss: workspace must be 8-byte aligned");
    LossParams p;
    p.gen = gen; p.logits = logits; p.target = target; p.mask = mask;
    p.grad_gen = grad_gen; p.grad_logits = grad_logits; p.fused_out = fused_out; p.loss_out = loss_out;
    p.partials = reinterpret_cast<double*>(workspace);
    p.counter = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(workspace) + kMaxBlocks * sizeof(double));
    p.pixels = pixels; p.C = C; p.V = V; p.mode = mode; p.inv_count = inv_count;
    for (int c = 0; c < kMaxC; ++c) p.w[c] = c < C ? chan_weight[c] : 0.f;
    long long blocks = dmv::ceil_div_ll(pixels, kThreads);
    if (blocks > kMaxBlocks) blocks = kMaxBlocks;
    cudaStream_t st = (cudaStream_t)stream;
    const bool flat = V == 1 && !mask && !fused_out && ((pixels * C) % 4 == 0) &&
                      ((((uintptr_t)gen | (uintptr_t)target | (uintptr_t)grad_gen) & 15) == 0);
    if (flat) {
        long long fb = dmv::ceil_div_ll(pixels * C / 4, kThreads * 4);
        if (fb > kMaxBlocks) fb = kMaxBlocks;
        if (fb < 1) fb = 1;
        loss_flat_kernel<<<(int)fb, kThreads, 0, st>>>(p);
        return dmv::check_launch("loss_flat");
    }
    switch (C) {
        case 1: loss_kernel<1><<<(int)blocks, kThreads, 0, st>>>(p); break;
        case 3: loss_kernel<3><<<(int)blocks, kThreads, 0, st>>>(p); break;
        case 4: loss_kernel<4><<<(int)blocks, kThreads, 0, st>>>(p); break;
        default: loss_kernel<0><<<(int)blocks, kThreads, 0, st>>>(p); break;
    }
    return dmv::check_launch("loss_fused");
}

int dmv_scale_by_device_scalar(float* x, const float* scalar, long long n, void* stream) {
    DMV_REQUIRE(x && scalar && n >= 0, DMV_E_INVALID_ARG, "scale: bad argument");
    if (n == 0) return DMV_OK;
    long long blocks = dmv::ceil_div_ll(n, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    scale_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, scalar, n);
    return dmv::check_launch("scale");
}
}
