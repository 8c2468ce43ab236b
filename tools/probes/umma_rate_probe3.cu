// Experiment (not product code): true tcgen05.mma SS-mode rate with a LEAN issue stream (fully unrolled, descriptors in registers).
#include "../../pnpflow_b200/csrc/pnpf_ptx.cuh"
using namespace pnpf;

__device__ __forceinline__ void mma_p(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc) : "memory");
}

__global__ void __launch_bounds__(128, 1) rate3_kernel(long long* cycles, int N, int iters, int a_shift, int nacc) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc<512>(&slot);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint64_t a0 = make_smem_desc<64>(smem_u32(smem) + a_shift * 64);
        const uint64_t a1 = a0 + 2;
        const uint64_t b0 = make_smem_desc<64>(smem_u32(smem) + 32768), b1 = b0 + 2, b2 = b0 + 384, b3 = b2 + 2;
        const uint32_t d0 = tmem, d1 = tmem + (nacc > 1 ? N : 0);
        const long long t0 = clock64();
        for (int i = 0; i < iters; i += 8) {
            mma_p(d0, a0, b0, idesc); mma_p(d1, a1, b1, idesc); mma_p(d0, a0, b2, idesc); mma_p(d1, a1, b3, idesc);
            mma_p(d0, a1, b0, idesc); mma_p(d1, a0, b1, idesc); mma_p(d0, a1, b2, idesc); mma_p(d1, a0, b3, idesc);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        cycles[blockIdx.x] = clock64() - t0;
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}
#include <cstdio>
extern "C" int run_rate3(long long* host_cycles, int nblocks, int N, int iters, int a_shift, int nacc) {
    long long* d;
    cudaMalloc(&d, nblocks * sizeof(long long));
    cudaFuncSetAttribute(rate3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    rate3_kernel<<<nblocks, 128, 80 * 1024>>>(d, N, iters, a_shift, nacc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(host_cycles, d, nblocks * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return 0;
}
