// Data-parallel exchange of one chunk of the flat parameter storage as ONE kernel over peer-mapped memory:
// reduce-scatter (peer loads in fixed rank order, or multimem.ld_reduce through the NVSwitch) -> TF-Adam on the owned
// slice -> all-gather of the bf16 compute copy (peer stores, or multimem.st).  Contract: include/dmv3d.h.
// HBM/NVLink-bound: per owned element 4 B x world of gradient reads (4 B with multimem), 28 B of local Adam traffic,
// 2 B x world of bf16 writes (2 B with multimem).
#include "common.cuh"

namespace {
using namespace dmv;

constexpr int kMaxWorld = 8;
constexpr int kThreads = 256;

struct ExchangeArgs {
    const float* grad[kMaxWorld];
    __nv_bfloat16* half[kMaxWorld];
    unsigned* sig[kMaxWorld];
    const float* grad_mc;
    __nv_bfloat16* half_mc;
    float *master, *m, *v;
    unsigned* local;            // [2 * slots]: epoch, ticket of every slot
    const float* state;
    long long start, n_slice;
    int rank, world, slot, replicated;
    float omb1, omb2, eps, gscale;
};

__device__ __forceinline__ float4 ld_sys_v4(const float* p) {          // peer data must not be served from this SM's L1
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 mm_ld_reduce_v4(const float* p) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mm_st_v2(void* p, uint2 v) {
    asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ void st_sys_v2(void* p, uint2 v) {
    asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ unsigned ld_acq_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_rel_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_epoch(const unsigned* p, unsigned ep) {
    while ((int)(ld_acq_sys(p) - ep) < 0) __nanosleep(64);
}
__device__ __forceinline__ int sig_index(int slot, int phase, int r) { return (slot * 2 + phase) * kMaxWorld + r; }

__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// U = float4 groups a thread has in flight per peer: W * U remote-or-local gradient loads (16 B each) plus 3 U local
// loads are issued before anything is consumed -- the loop is bound by NVLink / HBM latency x bytes in flight
template <int W>
struct Unroll { static constexpr int value = W >= 8 ? 1 : (W >= 4 ? 2 : 4); };

template <int W>
__global__ void __launch_bounds__(kThreads, 2) exchange_kernel(ExchangeArgs a) {
    constexpr int U = Unroll<W>::value;
    __shared__ int s_last;
    unsigned* epoch = a.local + 2 * a.slot;
    unsigned* ticket = epoch + 1;
    // the epoch word is rewritten only by the last CTA to take a ticket, i.e. after every CTA has read it here
    const unsigned ep = *reinterpret_cast<volatile unsigned*>(epoch) + 1u;
    // ---- my gradients of this chunk are complete (stream order): tell every peer, then wait for all of them
    if (blockIdx.x == 0 && threadIdx.x < W) {
        __threadfence_system();
        st_rel_sys(a.sig[threadIdx.x] + sig_index(a.slot, 0, a.rank), ep);
    }
    if (threadIdx.x < W) wait_epoch(a.sig[a.rank] + sig_index(a.slot, 0, threadIdx.x), ep);
    __syncthreads();

    const float lr_t = __ldg(a.state + 2);
    const float gs = a.gscale;
    const long long base = a.start + (a.replicated ? 0 : (long long)a.rank * a.n_slice);
    const long long n4 = a.n_slice >> 2;
    const int lane = threadIdx.x & 31;
    const long long warp_id = ((long long)blockIdx.x * kThreads + threadIdx.x) >> 5, warps = ((long long)gridDim.x * kThreads) >> 5;
    // a warp owns blocks of U x 32 consecutive float4; within a block, load u of lane l is float4 (u * 32 + l): every
    // instruction of the warp touches 512 contiguous bytes
    for (long long blk = warp_id * (U * 32); blk < n4; blk += warps * (U * 32)) {
        float4 g[U], p[U], m[U], v[U];
        bool live[U];
#pragma unroll
        for (int u = 0; u < U; ++u) live[u] = blk + u * 32 + lane < n4;
        if (a.grad_mc) {                                   // in-switch reduction: one load returns the sum over ranks
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (live[u]) g[u] = mm_ld_reduce_v4(a.grad_mc + base + (blk + u * 32 + lane) * 4);
        } else {                                           // all loads in flight first, then a fixed-order sum r = 0..W-1
            float4 x[W][U];
#pragma unroll
            for (int r = 0; r < W; ++r)
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (live[u]) x[r][u] = ld_sys_v4(a.grad[r] + base + (blk + u * 32 + lane) * 4);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                g[u] = x[0][u];
#pragma unroll
                for (int r = 1; r < W; ++r) g[u] = add4(g[u], x[r][u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!live[u]) continue;
            const long long off = base + (blk + u * 32 + lane) * 4;
            p[u] = *reinterpret_cast<const float4*>(a.master + off);
            m[u] = *reinterpret_cast<const float4*>(a.m + off);
            v[u] = *reinterpret_cast<const float4*>(a.v + off);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!live[u]) continue;
            const long long off = base + (blk + u * 32 + lane) * 4;
            adam_one(p[u].x, __fmul_rn(g[u].x, gs), m[u].x, v[u].x, lr_t, a.omb1, a.omb2, a.eps);
            adam_one(p[u].y, __fmul_rn(g[u].y, gs), m[u].y, v[u].y, lr_t, a.omb1, a.omb2, a.eps);
            adam_one(p[u].z, __fmul_rn(g[u].z, gs), m[u].z, v[u].z, lr_t, a.omb1, a.omb2, a.eps);
            adam_one(p[u].w, __fmul_rn(g[u].w, gs), m[u].w, v[u].w, lr_t, a.omb1, a.omb2, a.eps);
            *reinterpret_cast<float4*>(a.master + off) = p[u];
            *reinterpret_cast<float4*>(a.m + off) = m[u];
            *reinterpret_cast<float4*>(a.v + off) = v[u];
            __nv_bfloat162 lo = __floats2bfloat162_rn(p[u].x, p[u].y), hi = __floats2bfloat162_rn(p[u].z, p[u].w);
            uint2 pk;
            pk.x = *reinterpret_cast<unsigned*>(&lo);
            pk.y = *reinterpret_cast<unsigned*>(&hi);
            if (a.replicated) {
                *reinterpret_cast<uint2*>(a.half[a.rank] + off) = pk;
            } else if (a.half_mc) {
                mm_st_v2(a.half_mc + off, pk);
            } else {
#pragma unroll
                for (int r = 0; r < W; ++r) st_sys_v2(a.half[r] + off, pk);
            }
        }
    }
    // ---- everything this rank writes to its peers is out: last CTA releases "done" and waits for every peer's, so that
    //      kernel completion here means (a) my bf16 chunk is complete, (b) nobody still reads my gradients
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x < W) {
        __threadfence_system();
        st_rel_sys(a.sig[threadIdx.x] + sig_index(a.slot, 1, a.rank), ep);
        wait_epoch(a.sig[a.rank] + sig_index(a.slot, 1, threadIdx.x), ep);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *ticket = 0u;
        *reinterpret_cast<volatile unsigned*>(epoch) = ep;
    }
}
}  // namespace

extern "C" {

int dmv_dp_signal_words(int slots) { return slots > 0 ? slots * 2 * kMaxWorld : 0; }

int dmv_dp_exchange_chunk(const void* const* grad_peers, void* const* half_peers, void* const* signal_peers, const void* grad_mc,
                          void* half_mc, float* master, float* m, float* v, void* local_state, long long start, long long n_slice,
                          int rank, int world, int slot, int replicated, const float* state4, float beta1, float beta2, float eps,
                          float grad_scale, int ctas, void* stream) {
    DMV_REQUIRE(grad_peers && half_peers && signal_peers && master && m && v && local_state && state4, DMV_E_INVALID_ARG,
                "dp_exchange: null argument");
    DMV_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world && slot >= 0, DMV_E_INVALID_ARG,
                "dp_exchange: world must be 1..8, 0 <= rank < world");
    DMV_REQUIRE(world == 1 || world == 2 || world == 4 || world == 8, DMV_E_UNSUPPORTED_SHAPE, "dp_exchange: world must be 1, 2, 4 or 8");
    DMV_REQUIRE(start >= 0 && n_slice >= 0 && (n_slice & 7) == 0 && (start & 3) == 0, DMV_E_ALIGN, "dp_exchange: n_slice % 8, start % 4");
    if (n_slice == 0) return DMV_OK;
    ExchangeArgs a;
    memset(&a, 0, sizeof(a));
    for (int r = 0; r < world; ++r) {
        DMV_REQUIRE(grad_peers[r] && half_peers[r] && signal_peers[r], DMV_E_INVALID_ARG, "dp_exchange: null peer pointer");
        DMV_REQUIRE((((uintptr_t)grad_peers[r] | (uintptr_t)half_peers[r]) & 15) == 0, DMV_E_ALIGN, "dp_exchange: peer buffers must be 16-byte aligned");
        a.grad[r] = (const float*)grad_peers[r];
        a.half[r] = (__nv_bfloat16*)half_peers[r];
        a.sig[r] = (unsigned*)signal_peers[r];
    }
    DMV_REQUIRE((((uintptr_t)master | (uintptr_t)m | (uintptr_t)v | (uintptr_t)grad_mc | (uintptr_t)half_mc) & 15) == 0, DMV_E_ALIGN,
                "dp_exchange: buffers must be 16-byte aligned");
    a.grad_mc = replicated ? nullptr : (const float*)grad_mc;
    a.half_mc = replicated ? nullptr : (__nv_bfloat16*)half_mc;
    a.master = master; a.m = m; a.v = v;
    a.local = (unsigned*)local_state;
    a.state = state4;
    a.start = start; a.n_slice = n_slice;
    a.rank = rank; a.world = world; a.slot = slot; a.replicated = replicated ? 1 : 0;
    a.omb1 = 1.0f - beta1; a.omb2 = 1.0f - beta2; a.eps = eps; a.gscale = grad_scale;
    // default grid: four 256-thread CTAs per SM at most (the kernel also carries the owned slice's Adam traffic, 28 B per element,
    // which needs the whole memory system at small world sizes), at least ~4 blocks of work per warp
    const int U = world >= 8 ? 1 : (world >= 4 ? 2 : 4);
    long long want = dmv::ceil_div_ll(n_slice / 4, (long long)kThreads * U * 4);
    int grid = ctas > 0 ? ctas : 4 * 148;
    if (grid > want) grid = (int)want;
    if (grid < 1) grid = 1;
    cudaStream_t st = (cudaStream_t)stream;
    switch (world) {
        case 1: exchange_kernel<1><<<grid, kThreads, 0, st>>>(a); break;
        case 2: exchange_kernel<2><<<grid, kThreads, 0, st>>>(a); break;
        case 4: exchange_kernel<4><<<grid, kThreads, 0, st>>>(a); break;
        default: exchange_kernel<8><<<grid, kThreads, 0, st>>>(a); break;
    }
    return dmv::check_launch("dp_exchange_chunk");
}
}
