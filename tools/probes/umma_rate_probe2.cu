// Experiment (not product code): what slows a single MMA-issuing thread?  Variants of the SS-mode issue loop:
//   a_off   : A descriptor starts a_off rows into the swizzled tile (row-shifted halo view)
//   nbt     : the B operand rotates over nbt different weight tiles
//   spin    : number of extra warps that busy-poll an mbarrier (all 32 lanes, or lane 0 only if lane0only)
//   commit_every : tcgen05.commit to a scratch barrier every k MMAs
#include "../../pnpflow_b200/csrc/pnpf_ptx.cuh"
using namespace pnpf;

__global__ void __launch_bounds__(352, 1) rate2_kernel(long long* cycles, int N, int iters, int a_off, int nbt, int spin, int lane0only,
                                                       int commit_every) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar, bar2, bar3;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); mbar_init(&bar3, 1); fence_barrier_init(); }
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc<512>(&slot);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 0) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc_bf16(128, N);
            const uint64_t ad = make_smem_desc<64>(smem_u32(smem) + a_off * 64);
            const uint64_t bd0 = make_smem_desc<64>(smem_u32(smem) + 32768);
            const long long t0 = clock64();
            int bt = 0, cc = 0;
            uint32_t ph = 0;
            for (int i = 0; i < iters; ++i) {
                umma_bf16(tmem + (i % 3) * N, ad + 2 * (i & 1), bd0 + bt * (6144 / 16) + 2 * (i & 1), idesc, i >= 3 ? 1u : 0u);
                if (++bt == nbt) bt = 0;
                if (commit_every && ++cc == commit_every) { cc = 0; umma_commit(&bar3); (void)ph; }
            }
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            const long long t1 = clock64();
            cycles[blockIdx.x] = t1 - t0;
            mbar_arrive(&bar2);                      // release the spinners
        }
        __syncwarp();
    } else if (warp <= spin) {
        if (!lane0only || lane == 0) mbar_wait(&bar2, 0);
        __syncwarp();
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem);
}

#include <cstdio>
extern "C" int run_rate2(long long* host_cycles, int nblocks, int N, int iters, int a_off, int nbt, int spin, int lane0only, int commit_every) {
    long long* d;
    cudaMalloc(&d, nblocks * sizeof(long long));
    cudaFuncSetAttribute(rate2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    rate2_kernel<<<nblocks, 352, 100 * 1024>>>(d, N, iters, a_off, nbt, spin, lane0only, commit_every);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(host_cycles, d, nblocks * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return 0;
}
