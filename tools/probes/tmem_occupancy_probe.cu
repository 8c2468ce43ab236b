// Hardware/driver experiment: how many CTAs of a kernel that allocates tensor memory (tcgen05.alloc) can be resident on one SM?
// Prints cudaOccupancyMaxActiveBlocksPerMultiprocessor for kernels with / without tcgen05 and the co-residency actually observed
// (every CTA records its SM id and start / end clock; two CTAs of one SM overlapping in time = co-resident).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>
#include <algorithm>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int COLS, bool USE_TMEM, bool MMA>
__global__ void __launch_bounds__(320, 2) probe(long long* out, int spin) {
    __shared__ uint32_t slot;
    __shared__ uint64_t bar;
    const long long t0 = clock64();
    uint32_t base = 0;
    if (USE_TMEM) {
        if (threadIdx.x < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "n"(COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        base = slot;
    }
    if (MMA && threadIdx.x == 0) {          // makes ptxas emit the "tcgen05 used" kernel attributes
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    const long long t1 = clock64();
    while (clock64() - t1 < spin) {}
    __syncthreads();
    if (USE_TMEM && threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS));
    if (threadIdx.x == 0) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        out[3 * blockIdx.x] = smid;
        out[3 * blockIdx.x + 1] = t0;
        out[3 * blockIdx.x + 2] = clock64();
    }
}

template <typename K>
static void run(const char* name, K kern, int dyn_smem) {
    int occ = -1, sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_smem);
    cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 320, dyn_smem);
    const int grid = 2 * sms;
    long long* d;
    cudaMalloc(&d, grid * 3 * sizeof(long long));
    kern<<<grid, 320, dyn_smem>>>(d, 2000000);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<long long> h(grid * 3);
    cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    // globaltimer-free overlap test: per SM, sort CTAs by start clock (same SM clock domain) and count starts before the previous end
    int overlapping = 0;
    for (int s = 0; s < 1024; ++s) {
        std::vector<std::pair<long long, long long>> v;
        for (int b = 0; b < grid; ++b) if (h[3 * b] == s) v.push_back({h[3 * b + 1], h[3 * b + 2]});
        std::sort(v.begin(), v.end());
        for (size_t i = 1; i < v.size(); ++i) if (v[i].first < v[i - 1].second) ++overlapping;
    }
    printf("%-34s dyn smem %6d: occupancy API %d CTAs/SM, observed %d of %d CTAs co-resident with another (%s)\n", name, dyn_smem, occ, overlapping,
           grid, cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run("no tcgen05", probe<256, false, false>, 100 * 1024);
    run("tcgen05.alloc 256 cols", probe<256, true, false>, 100 * 1024);
    run("tcgen05.alloc 256 cols + commit", probe<256, true, true>, 100 * 1024);
    run("tcgen05.alloc 128 cols + commit", probe<128, true, true>, 100 * 1024);
    run("tcgen05.alloc 256 cols + commit", probe<256, true, true>, 0);
    run("tcgen05.alloc 512 cols + commit", probe<512, true, true>, 0);
    return 0;
}
