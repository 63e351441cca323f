// TF-flavoured Adam as a multi-tensor HBM-bound kernel, plus the
// elementwise helpers (activation forward/backward, casts).
//   ApplyAdam: m += (g - m)(1-b1); v += (g*g - v)(1-b2); theta -= (m*lr_t)/(sqrt(v)+eps)
// Bytes per parameter: read theta,g,m,v (16) + write theta,m,v (12) [+ 2 for the bf16 copy].
// The step-dependent scalars live on the device (state4 = {b1^t, b2^t, lr_t, t}) and are
// advanced by dmv_adam_tick, so a captured CUDA graph replays without host updates.
#include <stdlib.h>

#include "common.cuh"

namespace {
using namespace dmv;

constexpr int kMaxTensors = 24;
struct AdamTable {
    float* p[kMaxTensors];
    const float* g[kMaxTensors];
    float* m[kMaxTensors];
    float* v[kMaxTensors];
    __nv_bfloat16* h[kMaxTensors];
    long long start[kMaxTensors + 1];  // prefix sums in units of 4-element chunks
    int count;
};

__global__ void set_flag_kernel(int* flag, int value) { *flag = value; }

__global__ void adam_tick_kernel(float* st, float lr, float b1, float b2) {
    // float32 running powers, as TF keeps beta1_power / beta2_power variables
    const float p1 = st[0] * b1, p2 = st[1] * b2;
    st[0] = p1;
    st[1] = p2;
    st[2] = lr * sqrtf(1.0f - p2) / (1.0f - p1);
    st[3] = st[3] + 1.0f;
}

__global__ void __launch_bounds__(256) adam_multi_kernel(AdamTable t, const float* __restrict__ state, float omb1, float omb2,
                                                          float eps, float gscale, const int* __restrict__ gate) {
    if (gate && *gate == 0) return;      // deferred update with nothing pending (dmv_adam_multi_gated)
    const float lr_t = __ldg(state + 2);
    const long long total = t.start[t.count];
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long ch = (long long)blockIdx.x * blockDim.x + threadIdx.x; ch < total; ch += stride) {
        int k = 0;
        while (ch >= t.start[k + 1]) ++k;  // count <= 24: a short uniform-ish scan
        const long long base = (ch - t.start[k]) * 4;
        float* P = t.p[k] + base;
        const float* G = t.g[k] + base;
        float* M = t.m[k] + base;
        float* V = t.v[k] + base;
        float4 p4 = *reinterpret_cast<float4*>(P);
        float4 g4 = __ldg(reinterpret_cast<const float4*>(G));
        float4 m4 = *reinterpret_cast<float4*>(M);
        float4 v4 = *reinterpret_cast<float4*>(V);
        g4.x = __fmul_rn(g4.x, gscale); g4.y = __fmul_rn(g4.y, gscale); g4.z = __fmul_rn(g4.z, gscale); g4.w = __fmul_rn(g4.w, gscale);
        adam_one(p4.x, g4.x, m4.x, v4.x, lr_t, omb1, omb2, eps);
        adam_one(p4.y, g4.y, m4.y, v4.y, lr_t, omb1, omb2, eps);
        adam_one(p4.z, g4.z, m4.z, v4.z, lr_t, omb1, omb2, eps);
        adam_one(p4.w, g4.w, m4.w, v4.w, lr_t, omb1, omb2, eps);
        *reinterpret_cast<float4*>(P) = p4;
        *reinterpret_cast<float4*>(M) = m4;
        *reinterpret_cast<float4*>(V) = v4;
        if (t.h[k]) {
            __nv_bfloat162 lo = __floats2bfloat162_rn(p4.x, p4.y), hi = __floats2bfloat162_rn(p4.z, p4.w);
            uint2 pk;
            pk.x = *reinterpret_cast<unsigned*>(&lo);
            pk.y = *reinterpret_cast<unsigned*>(&hi);
            *reinterpret_cast<uint2*>(t.h[k] + base) = pk;
        }
    }
}

// scalar tail / unaligned tensors
__global__ void adam_scalar_kernel(float* P, const float* G, float* M, float* V, __nv_bfloat16* Hc, long long n,
                                   const float* __restrict__ state, float omb1, float omb2, float eps, float gscale,
                                   const int* __restrict__ gate) {
    if (gate && *gate == 0) return;
    const float lr_t = __ldg(state + 2);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float p = P[i], m = M[i], v = V[i];
        adam_one(p, __fmul_rn(G[i], gscale), m, v, lr_t, omb1, omb2, eps);
        P[i] = p; M[i] = m; V[i] = v;
        if (Hc) Hc[i] = __float2bfloat16_rn(p);
    }
}

template <typename T>
__global__ void act_fwd_kernel(const T* x, T* y, long long n, int act) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        store_from_float(y + i, apply_act(load_as_float(x + i), act));
}
template <typename T>
__global__ void act_bwd_kernel(const T* dy, const T* y, T* dpre, long long n, int act) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        store_from_float(dpre + i, load_as_float(dy + i) * act_grad_from_output(load_as_float(y + i), act));
}
// 8 bf16 per thread
__global__ void act_bwd_bf16x8_kernel(const uint4* dy, const uint4* y, uint4* dpre, long long n8, int act) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
        uint4 a = __ldg(dy + i), b = __ldg(y + i), o;
        const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
        const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
        __nv_bfloat162* po = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float2 fa = __bfloat1622float2(pa[k]), fb = __bfloat1622float2(pb[k]);
            po[k] = __floats2bfloat162_rn(fa.x * act_grad_from_output(fb.x, act), fa.y * act_grad_from_output(fb.y, act));
        }
        dpre[i] = o;
    }
}
__global__ void cast_f2b_kernel(const float* s, __nv_bfloat16* d, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) d[i] = __float2bfloat16_rn(s[i]);
}
__global__ void cast_b2f_kernel(const __nv_bfloat16* s, float* d, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) d[i] = __bfloat162float(s[i]);
}

// uint8 pixels -> float32 / 255 (read_tf_records.py:111), four pixels per thread
__global__ void u8_to_f32_kernel(const uchar4* __restrict__ s, float4* __restrict__ d, long long n4, float scale) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const uchar4 v = __ldg(s + i);
        d[i] = make_float4(__fdiv_rn((float)v.x, scale), __fdiv_rn((float)v.y, scale), __fdiv_rn((float)v.z, scale), __fdiv_rn((float)v.w, scale));
    }
}

inline int grid_for(long long n, int per_block) {
    long long b = dmv::ceil_div_ll(n, per_block);
    if (b > 148 * 16) b = 148 * 16;
    if (b < 1) b = 1;
    return (int)b;
}
}  // namespace

extern "C" {

int dmv_adam_tick(float* state4, float lr, float beta1, float beta2, void* stream) {
    DMV_REQUIRE(state4, DMV_E_INVALID_ARG, "adam_tick: null state");
    adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state4, lr, beta1, beta2);
    return dmv::check_launch("adam_tick");
}

static int adam_multi_impl(float* const* params, const float* const* grads, float* const* m, float* const* v,
                           void* const* bf16_copy, const long long* n, int count, const float* state4, float beta1, float beta2,
                           float eps, float grad_scale, const int* gate, void* stream) {
    DMV_REQUIRE(params && grads && m && v && n && state4 && count >= 0, DMV_E_INVALID_ARG, "adam: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
    int i = 0;
    while (i < count) {
        AdamTable t;
        t.count = 0;
        t.start[0] = 0;
        while (i < count && t.count < kMaxTensors) {
            const long long ni = n[i];
            DMV_REQUIRE(ni >= 0 && params[i] && grads[i] && m[i] && v[i], DMV_E_INVALID_ARG, "adam: null tensor");
            __nv_bfloat16* h = bf16_copy ? reinterpret_cast<__nv_bfloat16*>(bf16_copy[i]) : nullptr;
            const bool aligned = (((uintptr_t)params[i] | (uintptr_t)grads[i] | (uintptr_t)m[i] | (uintptr_t)v[i]) & 15) == 0 &&
                                 (!h || ((uintptr_t)h & 7) == 0);
            const long long vec = aligned ? (ni / 4) : 0;
            if (vec > 0) {
                const int k = t.count++;
                t.p[k] = params[i]; t.g[k] = grads[i]; t.m[k] = m[i]; t.v[k] = v[i]; t.h[k] = h;
                t.start[k + 1] = t.start[k] + vec;
            }
            const long long tail = ni - vec * 4;
            if (tail > 0) {
                const long long off = vec * 4;
                adam_scalar_kernel<<<grid_for(tail, 256), 256, 0, st>>>(params[i] + off, grads[i] + off, m[i] + off, v[i] + off,
                                                                         h ? h + off : nullptr, tail, state4, omb1, omb2, eps, grad_scale, gate);
                int rc = dmv::check_launch("adam_scalar");
                if (rc) return rc;
            }
            ++i;
        }
        if (t.count > 0) {
            static int cap = -1;      // DMV_ADAM_CTAS: cap of the grid (A/B: room for co-resident tensor-core CTAs)
            if (cap < 0) {
                const char* e = getenv("DMV_ADAM_CTAS");
                cap = e ? atoi(e) : 0;
            }
            int grid = grid_for(t.start[t.count], 256);
            if (cap > 0 && grid > cap) grid = cap;
            adam_multi_kernel<<<grid, 256, 0, st>>>(t, state4, omb1, omb2, eps, grad_scale, gate);
            int rc = dmv::check_launch("adam_multi");
            if (rc) return rc;
        }
    }
    return DMV_OK;
}

int dmv_adam_multi(float* const* params, const float* const* grads, float* const* m, float* const* v,
                   void* const* bf16_copy, const long long* n, int count, const float* state4, float beta1, float beta2,
                   float eps, float grad_scale, void* stream) {
    return adam_multi_impl(params, grads, m, v, bf16_copy, n, count, state4, beta1, beta2, eps, grad_scale, nullptr, stream);
}

int dmv_adam_multi_gated(float* const* params, const float* const* grads, float* const* m, float* const* v,
                         void* const* bf16_copy, const long long* n, int count, const float* state4, float beta1, float beta2,
                         float eps, float grad_scale, const int* gate, void* stream) {
    DMV_REQUIRE(gate, DMV_E_INVALID_ARG, "adam_multi_gated: null gate");
    return adam_multi_impl(params, grads, m, v, bf16_copy, n, count, state4, beta1, beta2, eps, grad_scale, gate, stream);
}

int dmv_set_flag(int* flag, int value, void* stream) {
    DMV_REQUIRE(flag, DMV_E_INVALID_ARG, "set_flag: null pointer");
    set_flag_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag, value);
    return dmv::check_launch("set_flag");
}

int dmv_act_fwd(const void* x, void* y, int dtype, long long n, int act, void* stream) {
    DMV_REQUIRE(x && y && n >= 0, DMV_E_INVALID_ARG, "act_fwd: bad argument");
    if (n == 0) return DMV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DMV_DT_F32) act_fwd_kernel<float><<<grid_for(n, 256), 256, 0, st>>>((const float*)x, (float*)y, n, act);
    else act_fwd_kernel<__nv_bfloat16><<<grid_for(n, 256), 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n, act);
    return dmv::check_launch("act_fwd");
}

int dmv_act_bwd(const void* dy, const void* y, void* dpre, int dtype, long long n, int act, void* stream) {
    DMV_REQUIRE(dy && y && dpre && n >= 0, DMV_E_INVALID_ARG, "act_bwd: bad argument");
    if (n == 0) return DMV_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DMV_DT_F32) {
        act_bwd_kernel<float><<<grid_for(n, 256), 256, 0, st>>>((const float*)dy, (const float*)y, (float*)dpre, n, act);
    } else if ((n & 7) == 0 && ((((uintptr_t)dy | (uintptr_t)y | (uintptr_t)dpre) & 15) == 0)) {
        act_bwd_bf16x8_kernel<<<grid_for(n / 8, 256), 256, 0, st>>>((const uint4*)dy, (const uint4*)y, (uint4*)dpre, n / 8, act);
    } else {
        act_bwd_kernel<__nv_bfloat16><<<grid_for(n, 256), 256, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)y,
                                                                           (__nv_bfloat16*)dpre, n, act);
    }
    return dmv::check_launch("act_bwd");
}

int dmv_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream) {
    DMV_REQUIRE(src && dst && n >= 0, DMV_E_INVALID_ARG, "cast: bad argument");
    if (n == 0) return DMV_OK;
    cast_f2b_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(src, (__nv_bfloat16*)dst, n);
    return dmv::check_launch("cast_f32_to_bf16");
}
int dmv_u8_to_f32(const unsigned char* src, float* dst, long long n, float divisor, void* stream) {
    DMV_REQUIRE(src && dst && n >= 0 && divisor != 0.f, DMV_E_INVALID_ARG, "u8_to_f32: bad argument");
    DMV_REQUIRE((n & 3) == 0 && ((uintptr_t)src & 3) == 0 && ((uintptr_t)dst & 15) == 0, DMV_E_ALIGN, "u8_to_f32: n % 4 and 4/16-byte alignment");
    if (n == 0) return DMV_OK;
    u8_to_f32_kernel<<<grid_for(n / 4, 256), 256, 0, (cudaStream_t)stream>>>((const uchar4*)src, (float4*)dst, n / 4, divisor);
    return dmv::check_launch("u8_to_f32");
}
int dmv_cast_bf16_to_f32(const void* src, float* dst, long long n, void* stream) {
    DMV_REQUIRE(src && dst && n >= 0, DMV_E_INVALID_ARG, "cast: bad argument");
    if (n == 0) return DMV_OK;
    cast_b2f_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, dst, n);
    return dmv::check_launch("cast_bf16_to_f32");
}
}
