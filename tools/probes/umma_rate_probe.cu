// Experiment (not product code): tcgen05.mma issue/throughput vs N and vs accumulator reuse pattern (SS mode, M=128, K=16).
//   mode 0: every MMA accumulates into the SAME TMEM accumulator
//   mode 1: MMAs rotate over NROT accumulators
// Each CTA (one per SM) issues `iters` MMAs on garbage smem operands and reports clock64 cycles.
#include "../../pnpflow_b200/csrc/pnpf_ptx.cuh"
using namespace pnpf;

__global__ void __launch_bounds__(128, 1) rate_kernel(long long* cycles, int N, int nrot, int iters, int rowb) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // small bf16 values
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc<512>(&slot);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_bf16(128, N);
        const uint64_t ad = rowb == 128 ? make_smem_desc<128>(smem_u32(smem)) : make_smem_desc<64>(smem_u32(smem));
        const uint64_t bd = rowb == 128 ? make_smem_desc<128>(smem_u32(smem) + 16384) : make_smem_desc<64>(smem_u32(smem) + 16384);
        const long long t0 = clock64();
        int rot = 0;
        for (int i = 0; i < iters; ++i) {
            umma_bf16(tmem + rot * N, ad + 2 * (i & 1), bd + 2 * (i & 1), idesc, i >= nrot ? 1u : 0u);
            if (++rot == nrot) rot = 0;
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

#include <cstdio>
extern "C" int run_rate(long long* host_cycles, int nblocks, int N, int nrot, int iters, int rowb) {
    long long* d;
    cudaMalloc(&d, nblocks * sizeof(long long));
    cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    rate_kernel<<<nblocks, 128, 64 * 1024>>>(d, N, nrot, iters, rowb);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(host_cycles, d, nblocks * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return 0;
}
