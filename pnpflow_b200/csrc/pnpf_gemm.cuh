// Implicit-GEMM convolution / batched GEMM on the 5th-gen tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   D[128 pixels x BN] (+)= sum_{tap, cin-chunk} A_tap[128 x BK] * W[BN x BK]^T
//
// * A (activations) lives in HBM as NHWC fp16 and is fetched with a 4-D TILED tensor map (C, W, H, B) whose box is
//   {BK channels, TW, TH, 1}: one TMA per (tap, channel chunk) with the start coordinate shifted by (dw, dh).
//   Out-of-bounds coordinates (including negative ones) are zero-filled by the TMA unit, which IS the conv's
//   zero padding; the box lands in shared memory as 128 rows of BK*2 bytes with the hardware 128B/64B swizzle,
//   i.e. exactly the canonical K-major UMMA operand layout.  Stride-2 convs use elementStrides = 2 in W and H.
// * W (weights) is a [N_total, K_total] K-major fp16 matrix (K ordered tap-major then channel) fetched with a 3-D map.
// * One elected thread issues tcgen05.mma (M=128, N=BN, K=16) into a double-buffered TMEM accumulator; 4 epilogue
//   warps drain it with tcgen05.ld while the next tile's main loop runs (persistent CTAs, static tile striding).
// * The same kernel runs the attention GEMMs (H=1, W=M rows, per-image B operand) and the fused 1x1 shortcut
//   (extra K chunks read from a second activation map).
#pragma once
#include "pnpf_ptx.cuh"

namespace pnpf {

// Epilogue description shared by the tensor-core kernels (bias / time-embedding / residual / store mode).
struct EpiParams {
    void* out;
    int out_mode;          // 0: fp16 [img][pix][col]   1: f32 [img][pix][col]   2: f32 [img][col][pix] (NCHW)
    long long out_img_stride, out_row_stride, out_col_stride;   // elements
    int n_valid;           // columns >= n_valid are not stored
    const float* bias;     // [N_total] or nullptr
    const float* bias_img; // [img][bias_img_stride] + col, or nullptr   (time-embedding projection)
    long long bias_img_stride;
    const act16* residual;   // [img][pix][col] fp16 or nullptr
    long long res_img_stride, res_row_stride;
    double* stats;         // optional GroupNorm statistics of the OUTPUT: [img][n_valid][2] (sum, sumsq), fp64 atomics
};

// One thread = one output pixel; v[16] = accumulator columns [col0, col0+16).  Adds bias / temb / residual and stores.
__device__ __forceinline__ void epilogue_apply16(const EpiParams& p, int img, long long pix, int col0, float (&v)[16]) {
    if (p.bias) {
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
        }
    }
    if (p.bias_img) {
        const float* bi = p.bias_img + img * p.bias_img_stride + col0;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bi + j));
            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
        }
    }
    if (p.residual) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + img * p.res_img_stride + pix * p.res_row_stride + col0);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const uint4 u = __ldg(rp + q);
            const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 t = unpack2(uu[j]);
                v[q * 8 + 2 * j] += t.x;
                v[q * 8 + 2 * j + 1] += t.y;
            }
        }
    }
}
__device__ __forceinline__ void epilogue_store16(const EpiParams& p, int img, long long pix, int col0, const float (&v)[16]) {
    if (p.out_mode == 0) {
        act16* op = reinterpret_cast<act16*>(p.out) + img * p.out_img_stride + pix * p.out_row_stride + col0;
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) pk[j] = pack2(v[2 * j], v[2 * j + 1]);
        uint4* o4 = reinterpret_cast<uint4*>(op);
        o4[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        o4[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    } else if (p.out_mode == 1) {
        float4* o4 = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + img * p.out_img_stride + pix * p.out_row_stride + col0);
#pragma unroll
        for (int j = 0; j < 4; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
        float* op = reinterpret_cast<float*>(p.out) + img * p.out_img_stride + pix;
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (col0 + j < p.n_valid) op[(col0 + j) * p.out_col_stride] = v[j];
    }
}
// Column sums over the 32 rows of a warp (row = lane) of a 16-column chunk, for GroupNorm statistics.
// Butterfly: 31 shuffles.  On return even lanes hold sum(col l>>1 ... see col_of_lane), odd lanes hold the sum of squares.
__device__ __forceinline__ float warp_colsum16(const float (&v)[16], bool valid, int lane) {
    float s[16], q[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        s[j] = valid ? v[j] : 0.f;
        q[j] = s[j] * s[j];
    }
#pragma unroll
    for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
        const bool up = (lane & bit) != 0;
#pragma unroll
        for (int i = 0; i < half; ++i) {
            const float ss = up ? s[i] : s[i + half];
            const float ks = up ? s[i + half] : s[i];
            s[i] = ks + __shfl_xor_sync(0xffffffffu, ss, bit);
            const float sq = up ? q[i] : q[i + half];
            const float kq = up ? q[i + half] : q[i];
            q[i] = kq + __shfl_xor_sync(0xffffffffu, sq, bit);
        }
    }
    // now s[0]/q[0] = partial of column ((lane>>1)&15 bit-composed below) over the 16 lanes sharing lane&1
    const bool odd = lane & 1;
    const float send = odd ? s[0] : q[0];
    const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
    return odd ? q[0] + recv : s[0] + recv;
}
// column (within the 16-chunk) whose statistic lane `lane` holds after warp_colsum16
__device__ __forceinline__ int colsum16_col(int lane) {
    return ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
}

// Two adjacent 16-column chunks of one accumulator row at once: every global load (bias, time-embedding row, residual) of
// both chunks is issued BEFORE the accumulator wait, invalid lanes load from pixel 0 and are masked afterwards (no divergent
// branch), and the two statistics butterflies are independent instruction streams the scheduler interleaves — the epilogue
// is a per-warp latency chain, so halving the number of chains per tile is what shortens it.
// values of the two chunks: accumulator + bias + time-embedding row + residual (invalid lanes read pixel 0 and are masked later)
__device__ __forceinline__ void epilogue_values32(const EpiParams& p, uint32_t t_addr, int img, long long pix, bool valid, int col0, bool on0, bool on1,
                                                  float (&v)[2][16]) {
    const long long pix_s = valid ? pix : 0;
    float4 b[2][4], bi[2][4];
    uint4 rs[2][2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            b[h][j] = make_float4(0.f, 0.f, 0.f, 0.f);
            bi[h][j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        rs[h][0] = rs[h][1] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (p.bias) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
            if (h == 0 ? on0 : on1)
#pragma unroll
                for (int j = 0; j < 4; ++j) b[h][j] = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + 16 * h) + j);
    }
    if (p.bias_img) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
            if (h == 0 ? on0 : on1)
#pragma unroll
                for (int j = 0; j < 4; ++j) bi[h][j] = __ldg(reinterpret_cast<const float4*>(p.bias_img + img * p.bias_img_stride + col0 + 16 * h) + j);
    }
    if (p.residual) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
            if (h == 0 ? on0 : on1) {
                const uint4* rp = reinterpret_cast<const uint4*>(p.residual + img * p.res_img_stride + pix_s * p.res_row_stride + col0 + 16 * h);
                rs[h][0] = __ldg(rp);
                rs[h][1] = __ldg(rp + 1);
            }
    }
    uint32_t r[2][16];
    tmem_ld_x16(t_addr + col0, r[0]);
    tmem_ld_x16(t_addr + col0 + 16, r[1]);
    tmem_ld_wait();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            v[h][4 * j] = __uint_as_float(r[h][4 * j]) + b[h][j].x + bi[h][j].x;
            v[h][4 * j + 1] = __uint_as_float(r[h][4 * j + 1]) + b[h][j].y + bi[h][j].y;
            v[h][4 * j + 2] = __uint_as_float(r[h][4 * j + 2]) + b[h][j].z + bi[h][j].z;
            v[h][4 * j + 3] = __uint_as_float(r[h][4 * j + 3]) + b[h][j].w + bi[h][j].w;
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const uint32_t uu[4] = {rs[h][q].x, rs[h][q].y, rs[h][q].z, rs[h][q].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 t = unpack2(uu[j]);
                v[h][q * 8 + 2 * j] += t.x;
                v[h][q * 8 + 2 * j + 1] += t.y;
            }
        }
    }
}
__device__ __forceinline__ void epilogue_chunk32(const EpiParams& p, uint32_t t_addr, int img, long long pix, bool valid, int col0, int lane) {
    const bool on0 = col0 < p.n_valid, on1 = col0 + 16 < p.n_valid;             // warp-uniform
    float v[2][16];
    epilogue_values32(p, t_addr, img, pix, valid, col0, on0, on1, v);
    if (p.stats) {
        const float t0 = warp_colsum16(v[0], valid && on0, lane);
        const float t1 = warp_colsum16(v[1], valid && on1, lane);
        const int cs = col0 + colsum16_col(lane);
        double* sp = p.stats + (static_cast<long long>(img) * p.n_valid + cs) * 2 + (lane & 1);
        if (on0) atomicAdd(sp, static_cast<double>(t0));
        if (on1) atomicAdd(sp + 32, static_cast<double>(t1));
    }
    if (valid && on0) epilogue_store16(p, img, pix, col0, v[0]);
    if (valid && on1) epilogue_store16(p, img, pix, col0 + 16, v[1]);
}

// The same 32 columns for fp16 NHWC outputs, through a per-warp 2 KB staging tile (32 pixels x 64 bytes, 16-byte chunks XOR-swizzled
// by (pixel >> 1) & 3: conflict-free in both phases): a lane writes its own pixel's 64 bytes, then reads 16 bytes of pixels
// rd_pix, 8 + rd_pix, 16 + rd_pix, 24 + rd_pix — so every global store instruction writes 8 whole 64-byte pixel rows (32 LSU
// transactions per chunk instead of 128), and the GroupNorm statistics are summed in the read phase: a lane holds 8 channels of 4
// pixels (on the fp16 values the consumer will normalise, like the row kernel), 48 shuffles instead of the 124 + 248 selects of the
// two butterflies.  pix_rd / vmask_rd: output pixel index and validity of the four pixels this lane stores.
__device__ __forceinline__ void epilogue_chunk32_staged(const EpiParams& p, uint32_t t_addr, int img, long long pix, bool valid, int col0, int lane,
                                                        uint32_t stage_w, const int (&pix_rd)[4], uint32_t vmask_rd) {
    float v[2][16];
    epilogue_values32(p, t_addr, img, pix, valid, col0, true, true, v);
    const uint32_t st_wr = stage_w + lane * 64, sw = static_cast<uint32_t>((lane >> 1) & 3);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float* f = &v[k >> 1][(k & 1) * 8];
        const uint4 u = make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(st_wr + ((static_cast<uint32_t>(k) ^ sw) << 4)), "r"(u.x), "r"(u.y), "r"(u.z), "r"(u.w) : "memory");
    }
    __syncwarp();
    const int rd_chunk = lane & 3, rd_pix = lane >> 2;
    uint4 o4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int pr = i * 8 + rd_pix;
        const uint32_t a = stage_w + pr * 64 + ((static_cast<uint32_t>(rd_chunk) ^ static_cast<uint32_t>((pr >> 1) & 3)) << 4);
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(o4[i].x), "=r"(o4[i].y), "=r"(o4[i].z), "=r"(o4[i].w) : "r"(a) : "memory");
    }
    act16* obase = reinterpret_cast<act16*>(p.out) + img * p.out_img_stride + col0 + rd_chunk * 8;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (vmask_rd & (1u << i)) *reinterpret_cast<uint4*>(obase + static_cast<long long>(pix_rd[i]) * p.out_row_stride) = o4[i];
    if (p.stats) {
        float ssum[8], ssq[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) ssum[e] = ssq[e] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool ok = (vmask_rd >> i) & 1u;
            const uint32_t ww[4] = {o4[i].x, o4[i].y, o4[i].z, o4[i].w};
#pragma unroll
            for (int e2 = 0; e2 < 4; ++e2) {
                const float2 t = unpack2(ww[e2]);
                const float f0 = ok ? t.x : 0.f, f1 = ok ? t.y : 0.f;
                ssum[2 * e2] += f0;
                ssq[2 * e2] = fmaf(f0, f0, ssq[2 * e2]);
                ssum[2 * e2 + 1] += f1;
                ssq[2 * e2 + 1] = fmaf(f1, f1, ssq[2 * e2 + 1]);
            }
        }
        // the 8 lanes with the same rd_chunk hold partial sums of the same 8 channels: xor-reduce over lane bits 2..4
#pragma unroll
        for (int bit = 4; bit <= 16; bit <<= 1)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                ssum[e] += __shfl_xor_sync(0xffffffffu, ssum[e], bit);
                ssq[e] += __shfl_xor_sync(0xffffffffu, ssq[e], bit);
            }
        float ts = 0.f, tq = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e)
            if (rd_pix == e) { ts = ssum[e]; tq = ssq[e]; }
        double* sp = p.stats + (static_cast<long long>(img) * p.n_valid + col0 + rd_chunk * 8 + rd_pix) * 2;    // lane -> channel 8 rd_chunk + rd_pix
        atomicAdd(sp, static_cast<double>(ts));
        atomicAdd(sp + 1, static_cast<double>(tq));
    }
    __syncwarp();                                      // the tile is rewritten by the next chunk
}

struct GemmParams {
    int H, W;              // output extent in the A-box coordinate space (plain GEMM: H=1, W=M)
    int TH, TW;            // tile box, TH*TW == 128
    int tiles_h, tiles_w;
    int n_img;
    int n_tiles_n;         // N_total = n_tiles_n * BN
    int a_batched, b_batched;
    int in_stride;         // input coordinate = out * in_stride + d
    int ntaps;
    int dh[9], dw[9];
    int kchunks;           // channel chunks per tap read through tmA
    int c_base;            // first channel of tmA to read
    int kchunks2;          // trailing 1x1 chunks read through tmA2 (0 = none)
    long long* dbg;        // optional [16] cycle counters of CTA 0 (-DPNPF_ROWCONV_CLOCKS builds), else nullptr
    EpiParams epi;
};

// cycle counters of CTA 0 (tools/rowconv_dbg.py): compiled in only with -DPNPF_ROWCONV_CLOCKS
#ifdef PNPF_ROWCONV_CLOCKS
#define PNPF_CLK() clock64()
#else
#define PNPF_CLK() 0ll
#endif
#define PNPF_TIMED_WAIT(bar, par, ctr)        \
    do {                                      \
        const long long _t0 = PNPF_CLK();      \
        mbar_wait(bar, par);                  \
        ctr += PNPF_CLK() - _t0;               \
    } while (0)

// PAIR: two CTAs of a cluster run one M = 256 tcgen05.mma (cta_group::2): each stages its own 128-pixel A tile and HALF of the
// weight tile (BN/2 rows), which removes a quarter (BN = 128) to a third (BN = 256) of the shared-memory traffic per MMA — the
// resource that bounds this kernel (an SS-mode MMA reads A and B from shared memory and the TMA writes both there first).
template <int BK, int BN, bool PAIR = false>
struct GemmCfg {
    static constexpr int kRowBytes = BK * 2;
    static constexpr int A_BYTES = 128 * BK * 2;
    static constexpr int B_ROWS = PAIR ? BN / 2 : BN;        // weight rows staged by this CTA
    static constexpr int B_BYTES_RAW = B_ROWS * BK * 2;
    static constexpr int B_BYTES = (B_BYTES_RAW + 1023) / 1024 * 1024;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES_FIT = (196 * 1024) / STAGE_BYTES;
    static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
    static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    // warps: 0 TMA producer, 1 MMA issuer, 2..5 epilogue set 0, 6..9 epilogue set 1 (BN >= 64: each set drains BN / 2 columns —
    // the epilogue is a per-warp latency chain and a K = 256 1x1 conv is only 16 MMAs per tile, so one set was the bottleneck 8:1)
    static constexpr int EPI_SETS = BN >= 64 ? 2 : 1;
    static constexpr int THREADS = (2 + 4 * EPI_SETS) * 32;
    static_assert(BK == 32 || BK == 64, "BK");
    static_assert(BN == 16 || BN == 32 || BN == 64 || BN == 128 || BN == 256, "BN");
    static_assert(!PAIR || BN >= 128, "CTA pairs are used for the wide tiles only");
};

__device__ __forceinline__ void decode_tile(const GemmParams& p, int t, int& img, int& h0, int& w0, int& nt) {
    nt = t % p.n_tiles_n;
    int r = t / p.n_tiles_n;
    int twi = r % p.tiles_w;
    r /= p.tiles_w;
    int thi = r % p.tiles_h;
    img = r / p.tiles_h;
    h0 = thi * p.TH;
    w0 = twi * p.TW;
}

template <int BK, int BN, bool PAIR>
__global__ void __launch_bounds__(GemmCfg<BK, BN, PAIR>::THREADS, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ GemmParams p) {
    using Cfg = GemmCfg<BK, BN, PAIR>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tfull_bar = empty_bar + STAGES;     // [2]
    uint64_t* tempty_bar = tfull_bar + 2;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nk_main = p.ntaps * p.kchunks;
    const int nk = nk_main + p.kchunks2;
    // Work units.  Single CTA: one (M tile, N tile) per unit.  Pair: the two CTAs take M tiles 2u and 2u+1 of the same N tile
    // (the host only pairs when the M-tile count is even), the pair index strides over the units.
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const int total_units = (p.n_img * p.tiles_h * p.tiles_w / (PAIR ? 2 : 1)) * p.n_tiles_n;
    const int unit0 = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int unit_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    auto tile_of = [&](int u) {            // flat (M tile, N tile) index of this CTA's tile in unit u (decode_tile order)
        if (!PAIR) return u;
        const int nt = u % p.n_tiles_n, mp = u / p.n_tiles_n;
        return (2 * mp + static_cast<int>(rank)) * p.n_tiles_n + nt;
    };

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (p.kchunks2) tma_prefetch_desc(&tmA2);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], (PAIR ? 8 : 4) * Cfg::EPI_SETS);          // pair: the epilogue warps of BOTH CTAs release the leader's accumulator
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        if constexpr (PAIR) tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
        else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all();                   // both CTAs' barriers exist before any remote arrive / TMA signal
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer: converged warp, one ELECTED lane issues (see elect_one_sync) =====================
        {
            int stage = 0;
            uint32_t phase = 0;
            long long c_wait = 0, c_tiles = 0;
            const long long c_start = PNPF_CLK();
            for (int u = unit0; u < total_units; u += unit_step) {
                int img, h0, w0, nt;
                decode_tile(p, tile_of(u), img, h0, w0, nt);
                const int ab = p.a_batched ? img : 0;
                const int bb = p.b_batched ? img : 0;
                ++c_tiles;
                for (int i = 0; i < nk; ++i) {
                    PNPF_TIMED_WAIT(&empty_bar[stage], phase ^ 1, c_wait);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + Cfg::A_BYTES;
                    if (elect_one_sync()) {
                        if constexpr (PAIR) {
                            // both CTAs' tiles complete the LEADER's barrier; only the leader arms it (with both byte counts)
                            const uint32_t fb = mapa_u32(smem_u32(&full_bar[stage]), 0);
                            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (Cfg::A_BYTES + Cfg::B_BYTES_RAW));
                            if (i < nk_main) {
                                const int tap = i / p.kchunks;
                                const int ch = i - tap * p.kchunks;
                                tma_load_4d_pair(sa, &tmA, fb, p.c_base + ch * BK, w0 * p.in_stride + p.dw[tap], h0 * p.in_stride + p.dh[tap], ab);
                            } else {
                                tma_load_4d_pair(sa, &tmA2, fb, (i - nk_main) * BK, w0, h0, ab);
                            }
                            tma_load_3d_pair(sb, &tmB, fb, i * BK, nt * BN + static_cast<int>(rank) * Cfg::B_ROWS, bb);
                        } else {
                            mbar_arrive_expect_tx(&full_bar[stage], Cfg::A_BYTES + Cfg::B_BYTES_RAW);
                            if (i < nk_main) {
                                const int tap = i / p.kchunks;
                                const int ch = i - tap * p.kchunks;
                                tma_load_4d(sa, &tmA, &full_bar[stage], p.c_base + ch * BK, w0 * p.in_stride + p.dw[tap],
                                            h0 * p.in_stride + p.dh[tap], ab);
                            } else {
                                tma_load_4d(sa, &tmA2, &full_bar[stage], (i - nk_main) * BK, w0, h0, ab);
                            }
                            tma_load_3d(sb, &tmB, &full_bar[stage], i * BK, nt * BN, bb);
                        }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
            if (p.dbg && blockIdx.x == 0 && lane == 0) { p.dbg[0] = PNPF_CLK() - c_start; p.dbg[1] = c_wait; p.dbg[2] = c_tiles; }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer: converged warp, one ELECTED lane issues (pair: the leader CTA only) =====================
        if (!PAIR || rank == 0) {
            constexpr uint32_t idesc = make_idesc_act16(PAIR ? 256 : 128, BN);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            long long c_full = 0, c_tempty = 0;
            const long long c_start = PNPF_CLK();
            for (int u = unit0; u < total_units; u += unit_step) {
                PNPF_TIMED_WAIT(&tempty_bar[acc], acc_phase ^ 1, c_tempty);      // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int i = 0; i < nk; ++i) {
                    PNPF_TIMED_WAIT(&full_bar[stage], phase, c_full);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t adesc = make_smem_desc<Cfg::kRowBytes>(sa);
                    const uint64_t bdesc = make_smem_desc<Cfg::kRowBytes>(sa + Cfg::A_BYTES);
                    if (elect_one_sync()) {
#pragma unroll
                        for (int kk = 0; kk < BK / 16; ++kk) {
                            // advance 16 elements (32 B) along K inside the swizzle span: +2 in the (addr>>4) field
                            if constexpr (PAIR) umma_f16_pair(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (i | kk) ? 1u : 0u);
                            else umma_f16(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (i | kk) ? 1u : 0u);
                        }
                        if constexpr (PAIR) {
                            umma_commit_pair(&empty_bar[stage]);  // frees this stage in BOTH CTAs
                            if (i == nk - 1) umma_commit_pair(&tfull_bar[acc]);
                        } else {
                            umma_commit(&empty_bar[stage]);       // frees the smem slot when these MMAs retire
                            if (i == nk - 1) umma_commit(&tfull_bar[acc]);   // accumulator complete -> epilogue
                        }
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            if (p.dbg && blockIdx.x == 0 && lane == 0) { p.dbg[4] = PNPF_CLK() - c_start; p.dbg[5] = c_full; p.dbg[6] = c_tempty; }
        }
        __syncwarp();
    } else {
        // ===================== epilogue warps 2..5 (and 6..9: second half of the columns) =====================
        const int quarter = warp & 3;                             // TMEM lane quarter this warp may access
        const int m = quarter * 32 + lane;                        // accumulator row == pixel within the tile
        const int th = m / p.TW, tw = m - th * p.TW;
        int acc = 0;
        uint32_t acc_phase = 0;
        long long c_tfull = 0;
        const long long c_start = PNPF_CLK();
        const uint32_t tempty_remote = PAIR ? mapa_u32(smem_u32(&tempty_bar[0]), 0) : 0u;      // the leader's tempty_bar[0]
        for (int u = unit0; u < total_units; u += unit_step) {
            int img, h0, w0, nt;
            decode_tile(p, tile_of(u), img, h0, w0, nt);
            const int h = h0 + th, w = w0 + tw;
            const bool valid = (h < p.H) && (w < p.W);
            const long long pix = static_cast<long long>(h) * p.W + w;
            {
                const long long _t0 = PNPF_CLK();
                if (lane == 0) mbar_wait(&tfull_bar[acc], acc_phase);   // one lane polls, the warp follows
                __syncwarp();
                c_tfull += PNPF_CLK() - _t0;
            }
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
            if constexpr (Cfg::EPI_SETS == 2) {
                const int set = (warp - 2) >> 2;
#pragma unroll 1
                for (int c0 = set * (BN / 2); c0 < (set + 1) * (BN / 2); c0 += 32)      // t_addr - nt * BN: chunk32 adds the GLOBAL column
                    epilogue_chunk32(p.epi, t_addr - static_cast<uint32_t>(nt * BN), img, pix, valid, nt * BN + c0, lane);
            } else {
#pragma unroll 1
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t r[16];
                tmem_ld_x16(t_addr + c0, r);
                tmem_ld_wait();
                const int col0 = nt * BN + c0;
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
                const bool act = valid && col0 < p.epi.n_valid;
                if (act) epilogue_apply16(p.epi, img, pix, col0, v);
                if (p.epi.stats && col0 < p.epi.n_valid) {       // warp-uniform branch
                    const float tot = warp_colsum16(v, act, lane);
                    const int c = col0 + colsum16_col(lane);
                    atomicAdd(p.epi.stats + (static_cast<long long>(img) * p.epi.n_valid + c) * 2 + (lane & 1), static_cast<double>(tot));
                }
                if (act) epilogue_store16(p.epi, img, pix, col0, v);
            }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR) mbar_arrive_cluster(tempty_remote + acc * 8);     // relaxed: see pnpf_ptx.cuh
                else mbar_arrive(&tempty_bar[acc]);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (p.dbg && blockIdx.x == 0 && threadIdx.x == 64) { p.dbg[8] = PNPF_CLK() - c_start; p.dbg[9] = c_tfull; }
    }
    tc_fence_before();
    if constexpr (PAIR) {
        cluster_sync_all();               // the peer's shared memory / TMEM and the leader's barriers stay alive until both are done
        if (warp == 2) tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
    } else {
        __syncthreads();
        if (warp == 2) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

}  // namespace pnpf
