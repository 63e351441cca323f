// Hardware probe (not part of the product): can a tcgen05 K-major swizzled A operand start at an
// arbitrary pixel-row offset inside a TMA-written halo tile (row pitch 16 pixels), so that one halo
// load serves all kh x kw taps of a convolution?  D = A_shifted * I, compared on the host.
#include "../../dynamic_multiview_3d_b200/csrc/tc_common.cuh"

using namespace dmv;
using namespace dmv::tc;

struct ProbeParams {
    int C, row_off, base_offset, sbo_bytes;
    float* out;
};

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_i,
                                                        const __grid_constant__ ProbeParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int row_bytes = p.C * 2;
    const int tile_bytes = 20 * 16 * row_bytes;            // 20 halo rows x 16 pixels
    uint8_t* sb = smem + ((tile_bytes + 1023) & ~1023);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sb + p.C * row_bytes);
    uint64_t* done = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(slot, 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, (uint32_t)(tile_bytes + p.C * row_bytes));
        tma_load_3d(smem, &map_x, bar, 0, 0, 0);
        tma_load_3d(sb, &map_i, bar, 0, 0, 0);
        mbar_wait(bar, 0);
        tc_fence_after();
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.C >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t a_addr = smem_u32(smem) + (uint32_t)(p.row_off * row_bytes);
        uint64_t adesc = 0;
        adesc |= (uint64_t)((a_addr & 0x3FFFF) >> 4);
        adesc |= (uint64_t)1 << 16;
        adesc |= (uint64_t)((uint32_t)p.sbo_bytes >> 4) << 32;
        adesc |= (uint64_t)1 << 46;
        adesc |= (uint64_t)(p.base_offset & 7) << 49;
        adesc |= (uint64_t)(row_bytes == 128 ? 2 : 4) << 61;
        const uint64_t bdesc = make_kmajor_desc(smem_u32(sb), row_bytes);
        for (int k = 0; k < p.C / 16; ++k) tc_mma_bf16(tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k ? 1u : 0u);
        tc_commit(done);
    }
    mbar_wait(done, 0);
    tc_fence_after();
    const int m = warp * 32 + lane;
    for (int c0 = 0; c0 < p.C; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int k = 0; k < 16; ++k) p.out[m * p.C + c0 + k] = __uint_as_float(v[k]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 64);
    }
}

extern "C" int probe_halo(const void* x, const void* ident, float* out, int C, int row_off, int base_offset, int sbo_bytes, void* stream) {
    const int row_bytes = C * 2;
    CUtensorMap mx, mi;
    {
        cuuint64_t dims[3] = {(cuuint64_t)C, 16, 24};
        cuuint64_t strides[2] = {(cuuint64_t)row_bytes, (cuuint64_t)16 * row_bytes};
        cuuint32_t box[3] = {(cuuint32_t)C, 16, 20};
        int rc = encode_map(&mx, x, 3, dims, strides, box, row_bytes);
        if (rc) return rc;
    }
    {
        cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)C, 1};
        cuuint64_t strides[2] = {(cuuint64_t)row_bytes, (cuuint64_t)C * row_bytes};
        cuuint32_t box[3] = {(cuuint32_t)C, (cuuint32_t)C, 1};
        int rc = encode_map(&mi, ident, 3, dims, strides, box, row_bytes);
        if (rc) return rc;
    }
    ProbeParams p;
    p.C = C; p.row_off = row_off; p.base_offset = base_offset; p.sbo_bytes = sbo_bytes; p.out = out;
    const size_t smem = 20 * 16 * row_bytes + 1024 + C * row_bytes + 64 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(mx, mi, p);
    return check_launch("probe");
}
