// Host helpers: error string, tensor-map encoding, GEMM kernel dispatch.
#include "pnpf_host.h"

#include <cstdlib>

#include <cudaTypedefs.h>
#include <mutex>
#include <vector>

namespace pnpf {

static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}
const char* get_error() { return g_err.c_str(); }

static std::mutex g_dev_mutex;
static int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < DeviceCache::MAX_DEV) ? dev : 0;
}
bool DeviceCache::lookup(int* value) {
    const int dev = current_device();
    std::lock_guard<std::mutex> g(g_dev_mutex);
    if (!set_[dev]) return false;
    *value = val_[dev];
    return true;
}
void DeviceCache::store(int value) {
    const int dev = current_device();
    std::lock_guard<std::mutex> g(g_dev_mutex);
    val_[dev] = value;
    set_[dev] = true;
}

int num_sms() {
    static DeviceCache cache;
    int n = 0;
    if (!cache.lookup(&n)) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, current_device());
        if (n <= 0) n = 148;
        cache.store(n);
    }
    return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

static int encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                  const cuuint32_t* box, const cuuint32_t* estr, int bk) {
    EncodeTiledFn fn = encode_fn();
    PNPF_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    PNPF_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map base %p not 16-byte aligned", base);
    const CUtensorMapSwizzle sw = (bk == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PNPF_REQUIRE(r == CUDA_SUCCESS,
                 "cuTensorMapEncodeTiled failed (CUresult %d) rank %d dims [%llu %llu %llu %llu] box [%u %u %u %u] estr [%u %u %u %u]",
                 (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
                 (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], box[2], rank > 3 ? box[3] : 0, estr[0], estr[1],
                 estr[2], rank > 3 ? estr[3] : 0);
    return 0;
}

int make_act_tmap(CUtensorMap* m, const void* base, int C, long long pitch, int W, int H, int B, int bk, int tw, int th,
                  int estride) {
    PNPF_REQUIRE(bk == 32 || bk == 64, "bad bk %d", bk);
    PNPF_REQUIRE((pitch * 2) % 16 == 0, "channel pitch %lld not a multiple of 8 elements", pitch);
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)pitch * 2, (cuuint64_t)pitch * 2 * W, (cuuint64_t)pitch * 2 * W * H};
    cuuint32_t box[4] = {(cuuint32_t)bk, (cuuint32_t)(tw * estride), (cuuint32_t)(th * estride), 1};
    cuuint32_t estr[4] = {1, (cuuint32_t)estride, (cuuint32_t)estride, 1};
    PNPF_REQUIRE(box[1] <= 256 && box[2] <= 256, "TMA box too large (%u x %u)", box[1], box[2]);
    return encode(m, base, 4, dims, strides, box, estr, bk);
}

int make_b_tmap(CUtensorMap* m, const void* base, long long K, long long ldk, int N, int batch, long long bstride, int bk,
                int bn) {
    PNPF_REQUIRE(bk == 32 || bk == 64, "bad bk %d", bk);
    PNPF_REQUIRE((ldk * 2) % 16 == 0, "weight row pitch %lld not a multiple of 8 elements", ldk);
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)N, (cuuint64_t)batch};
    cuuint64_t strides[2] = {(cuuint64_t)ldk * 2, (cuuint64_t)(batch > 1 ? bstride : (long long)N * ldk) * 2};
    cuuint32_t box[3] = {(cuuint32_t)bk, (cuuint32_t)bn, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return encode(m, base, 3, dims, strides, box, estr, bk);
}

template <int BK, int BN, bool PAIR>
static int launch_t(const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB, const GemmParams& p,
                    cudaStream_t stream) {
    using Cfg = GemmCfg<BK, BN, PAIR>;
    static DeviceCache cache;
    int max_clusters = 0;
    if (!cache.lookup(&max_clusters)) {
        PNPF_CHECK_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BK, BN, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::SMEM_BYTES));
        if (PAIR) {                                       // how many CTA pairs (one per TPC) can be resident at once
            cudaLaunchConfig_t qc = {};
            qc.gridDim = dim3(num_sms() & ~1);
            qc.blockDim = dim3(Cfg::THREADS);
            qc.dynamicSmemBytes = Cfg::SMEM_BYTES;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            qc.attrs = qa; qc.numAttrs = 1;
            PNPF_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, conv_gemm_kernel<BK, BN, PAIR>, &qc));
            PNPF_REQUIRE(max_clusters >= 1, "no CTA pair of conv_gemm_kernel<%d,%d> fits on this device", BK, BN);
        }
        cache.store(max_clusters);
    }
    const long long tiles = (long long)p.n_img * p.tiles_h * p.tiles_w * p.n_tiles_n;
    if (tiles < 1) return 0;
    if (PAIR) {
        const long long units = tiles / 2;
        const int clusters = (int)(units < max_clusters ? units : max_clusters);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * clusters);
        cfg.blockDim = dim3(Cfg::THREADS);
        cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
        cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        PNPF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, conv_gemm_kernel<BK, BN, PAIR>, tmA, tmA2, tmB, p));
        return 0;
    }
    const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
    conv_gemm_kernel<BK, BN, PAIR><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, stream>>>(tmA, tmA2, tmB, p);
    PNPF_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_conv_gemm(int BK, int BN, const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB, const CUtensorMap& tmB_half,
                     const GemmParams& p, cudaStream_t stream) {
    PNPF_REQUIRE(p.TH * p.TW == 128, "tile %dx%d is not 128 pixels", p.TH, p.TW);
    PNPF_REQUIRE(p.epi.out_mode == 2 || p.epi.n_valid % 16 == 0, "n_valid %d must be a multiple of 16 for row-major output", p.epi.n_valid);
    // CTA pairs (cta_group::2, M = 256 per MMA) for the wide tiles whenever the M tiles pair up
    static const bool no_pair = getenv("PNPF_NO_PAIR") != nullptr;          // A/B switch (tools/ab_env.py)
    // (a pair shares ONE weight tile: with a per-image B operand both M tiles must belong to the same image)
    const long long m_tiles = (long long)p.n_img * p.tiles_h * p.tiles_w;
    const bool pairable = p.b_batched ? (p.tiles_h * p.tiles_w) % 2 == 0 : m_tiles % 2 == 0;
    // Measured (profiles/r01_ab_experiments.md): with pairs the MMA issue runs at the tensor rate and the kernel becomes bound by
    // the L2 -> shared-memory fill (every input pixel is fetched once per tap): BN = 256 gains 5-6 %, BN = 128 (activation
    // traffic dominates, not halved by pairing) loses 4 % -> pairs for BN = 256 only.
    if (!no_pair && BK == 64 && pairable) {
        if (BN == 256) return launch_t<64, 256, true>(tmA, tmA2, tmB_half, p, stream);
    }
#define PNPF_CASE(bk, bn) \
    if (BK == bk && BN == bn) return launch_t<bk, bn, false>(tmA, tmA2, tmB, p, stream);
    PNPF_CASE(32, 16) PNPF_CASE(32, 32) PNPF_CASE(32, 64) PNPF_CASE(32, 128) PNPF_CASE(32, 256)
    PNPF_CASE(64, 16) PNPF_CASE(64, 32) PNPF_CASE(64, 64) PNPF_CASE(64, 128) PNPF_CASE(64, 256)
#undef PNPF_CASE
    set_error("no conv_gemm instantiation for BK=%d BN=%d", BK, BN);
    return 2;
}

}  // namespace pnpf
