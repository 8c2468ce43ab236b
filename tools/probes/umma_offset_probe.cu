// Experiment (not product code): can a K-major UMMA A-operand start at an arbitrary ROW offset inside a TMA-written
// swizzled tile?  D[128 x 32] = A[off : off+128, :K] * B[32, :K]^T with the descriptor start address advanced by
// off * rowbytes and the descriptor base_offset field set to `base_off`.
#include "../../pnpflow_b200/csrc/pnpf_ptx.cuh"
using namespace pnpf;

template <int ROWB>   // 128 (SW128, K=64) or 64 (SW64, K=32)
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int off, int base_off) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                       // 144 rows x ROWB
    uint8_t* sB = smem + 144 * 128;           // 32 rows x ROWB (1024-aligned: 144*128 = 18432 = 18*1024)
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 32 * 128);
    uint64_t* bar2 = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar2, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<32>(slot);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, 144 * ROWB + 32 * ROWB);
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(sA)), "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(smem_u32(sB)), "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(smem_u32(bar)), "r"(0), "r"(0) : "memory");
        mbar_wait(bar, 0);
        tc_fence_after();
        uint64_t ad = make_smem_desc<ROWB>(smem_u32(sA) + off * ROWB) | (static_cast<uint64_t>(base_off & 7) << 49);
        uint64_t bd = make_smem_desc<ROWB>(smem_u32(sB));
        constexpr uint32_t idesc = make_idesc_bf16(128, 32);
        for (int kk = 0; kk < ROWB / 32; ++kk) umma_bf16(tmem, ad + 2 * kk, bd + 2 * kk, idesc, kk ? 1u : 0u);
        umma_commit(bar2);
    }
    __syncwarp();
    mbar_wait(bar2, 0);
    tc_fence_after();
    const int m = warp * 32 + lane;
    for (int c0 = 0; c0 < 32; c0 += 16) {
        uint32_t r[16];
        tmem_ld_x16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
        for (int j = 0; j < 16; ++j) out[m * 32 + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc<32>(tmem);
}

#include <cudaTypedefs.h>
#include <cstdio>
extern "C" int run_probe(const void* A /*[256][K] bf16*/, const void* B /*[32][K] bf16*/, float* out, int rowb, int off, int base_off) {
    void* fnp = nullptr; cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q) != cudaSuccess) return 10;
    auto fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fnp);
    const int K = rowb / 2;
    CUtensorMap ma, mb;
    cuuint64_t da[2] = {(cuuint64_t)K, 256}, sa[1] = {(cuuint64_t)rowb};
    cuuint32_t ba[2] = {(cuuint32_t)K, 144}, es[2] = {1, 1};
    cuuint64_t db[2] = {(cuuint64_t)K, 32};
    cuuint32_t bb[2] = {(cuuint32_t)K, 32};
    auto sw = rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    if (fn(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(A), da, sa, ba, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
           CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 11;
    if (fn(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(B), db, sa, bb, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
           CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return 12;
    const int smem = 144 * 128 + 32 * 128 + 64 + 1024;
    if (rowb == 128) {
        cudaFuncSetAttribute(probe_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        probe_kernel<128><<<1, 128, smem>>>(ma, mb, out, off, base_off);
    } else {
        cudaFuncSetAttribute(probe_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        probe_kernel<64><<<1, 128, smem>>>(ma, mb, out, off, base_off);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 13; }
    return 0;
}
