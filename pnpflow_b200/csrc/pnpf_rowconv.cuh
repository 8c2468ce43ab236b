// Row-streaming 3x3 (stride 1) implicit-GEMM convolution for wide feature maps (W % 128 == 0) and thin outputs
// (C_out <= 64) — the U-Net's two highest resolutions, where the per-tap kernel of pnpf_gemm.cuh is L2-bound because it
// fetches every input pixel nine times.
//
// A CTA owns a column strip of 128 output pixels and walks down a segment of rows.  Every INPUT row is fetched ONCE by TMA
// as a 130-pixel halo tile (w0-1 .. w0+128, out-of-image pixels zero-filled by the TMA unit = the conv padding) and feeds
// nine tensor-core taps:
//   * the three vertical taps: the row is used for output rows j-1, j, j+1 (three TMEM accumulators are live at once);
//   * the three horizontal taps: the UMMA A-descriptor simply starts 0, 1 or 2 pixel-rows into the swizzled halo tile
//     (the 128B/64B swizzle is a function of the shared-memory address bits, so a row-shifted start address reads the
//     right data with base_offset = 0 — verified on hardware by tools/probes/umma_offset_probe.cu).
// The 3x3 weights (and the optional fused 1x1 shortcut weights) stay resident in shared memory for the CTA's lifetime.
// Accumulators form a ring of NACC TMEM buffers, one per output row.
//
// These layers have only 16 tensor-core clocks of work per MMA instruction (N = 32), so instruction issue matters:
//   * the single MMA-issuing thread keeps descriptor low words in registers and only adds immediates per instruction;
//   * 8 epilogue warps: for C_out <= 32 two warp sets take alternate rows, for C_out = 64 they split the columns, so a
//     thread never handles more than 32 columns; bias (+ time-embedding row) is staged in shared memory once per item;
//   * GroupNorm statistics of the OUTPUT (sum, sum of squares per channel, for the next layer's GroupNorm) are
//     accumulated per thread in registers across the rows of an item and reduced with warp shuffles + one fp64 atomic
//     per lane once per item.
#pragma once
#include "pnpf_gemm.cuh"

namespace pnpf {

struct RowConvParams {
    int H, W, n_img;
    int strips;            // W / 128
    int seg_rows, segs;    // rows per work item, ceil(H / seg_rows)
    int kchunks;           // Cin / BK
    int kchunks2;          // C2 / BK of the fused 1x1 source (0 = none)
    int nslot;             // depth of the input-row ring
    int slot_bytes;
    EpiParams epi;
};

template <int BK, int BN>
struct RowCfg {
    static constexpr int kRowBytes = BK * 2;
    static constexpr int HALO_ROWS = 130;
    static constexpr int HALO_TILE = (136 * kRowBytes + 1023) / 1024 * 1024;
    static constexpr int X2_TILE = 128 * kRowBytes;
    static constexpr int W_TILE_RAW = BN * BK * 2;
    static constexpr int W_TILE = (W_TILE_RAW + 1023) / 1024 * 1024;
    static constexpr int NACC = (512 / BN) > 16 ? 16 : (512 / BN);
    static constexpr int TMEM_COLS = NACC * BN;              // 256 (BN=16) or 512
    static constexpr int MAX_SLOTS = 8;
    static constexpr int THREADS = 64 + 256;                 // producer, MMA issuer, 8 epilogue warps
    static constexpr int CPT = BN > 32 ? 32 : BN;            // columns per epilogue thread
    static constexpr bool ROW_SPLIT = BN <= 32;              // the two epilogue warp sets alternate rows (else: split columns)
    static constexpr int BAR_BYTES = 1024;                   // barriers + tmem slot + bias staging (2 x 64 floats)
    static_assert(BN == 16 || BN == 32 || BN == 64, "row conv is for thin outputs");
};

// tcgen05.mma with descriptors given as (low word, shared high word)
__device__ __forceinline__ void umma_bf16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <int BK, int BN>
__global__ void __launch_bounds__(RowCfg<BK, BN>::THREADS, 1)
rowconv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ RowConvParams p) {
    using Cfg = RowCfg<BK, BN>;
    constexpr int NACC = Cfg::NACC;
    constexpr int CPT = Cfg::CPT;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int n_wtiles = 9 * p.kchunks + p.kchunks2;
    uint8_t* wsm = smem;
    uint8_t* slots = smem + n_wtiles * Cfg::W_TILE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(slots + p.nslot * p.slot_bytes);
    uint64_t* wbar = bars;
    uint64_t* full_bar = bars + 1;
    uint64_t* empty_bar = full_bar + Cfg::MAX_SLOTS;
    uint64_t* tfull_bar = empty_bar + Cfg::MAX_SLOTS;
    uint64_t* tempty_bar = tfull_bar + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 16);
    float* bias_sm = reinterpret_cast<float*>(bars) + 128;          // [2 sets][64] floats at byte offset 512

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int items = p.n_img * p.segs * p.strips;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (p.kchunks2) tma_prefetch_desc(&tmA2);
        mbar_init(wbar, 1);
        for (int s = 0; s < Cfg::MAX_SLOTS; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 16; ++a) {
            mbar_init(&tfull_bar[a], 1);
            mbar_init(&tempty_bar[a], Cfg::ROW_SPLIT ? 4 : 8);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    auto decode = [&](int it, int& img, int& hb, int& he, int& w0) {
        const int strip = it % p.strips;
        int r = it / p.strips;
        const int seg = r % p.segs;
        img = r / p.segs;
        hb = seg * p.seg_rows;
        he = min(hb + p.seg_rows, p.H);
        w0 = strip * 128;
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            mbar_arrive_expect_tx(wbar, n_wtiles * Cfg::W_TILE_RAW);
            for (int i = 0; i < n_wtiles; ++i) tma_load_3d(wsm + i * Cfg::W_TILE, &tmB, wbar, i * BK, 0, 0);
            int slot = 0;
            uint32_t phase = 0;
            for (int it = blockIdx.x; it < items; it += gridDim.x) {
                int img, hb, he, w0;
                decode(it, img, hb, he, w0);
                const int j0 = max(hb - 1, 0), j1 = min(he, p.H - 1);
                for (int j = j0; j <= j1; ++j) {
                    mbar_wait(&empty_bar[slot], phase ^ 1);
                    uint8_t* sp = slots + slot * p.slot_bytes;
                    const bool centre = (j >= hb) && (j < he) && p.kchunks2;
                    mbar_arrive_expect_tx(&full_bar[slot], p.kchunks * Cfg::HALO_ROWS * Cfg::kRowBytes +
                                                               (centre ? p.kchunks2 * Cfg::X2_TILE : 0));
                    for (int c = 0; c < p.kchunks; ++c)
                        tma_load_4d(sp + c * Cfg::HALO_TILE, &tmA, &full_bar[slot], c * BK, w0 - 1, j, img);
                    if (centre)
                        for (int c = 0; c < p.kchunks2; ++c)
                            tma_load_4d(sp + p.kchunks * Cfg::HALO_TILE + c * Cfg::X2_TILE, &tmA2, &full_bar[slot], c * BK, w0, j, img);
                    if (++slot == p.nslot) { slot = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(128, BN);
            // descriptor words: hi is shared by every operand tile; lo = (addr >> 4) | LBO bit
            const uint64_t proto = make_smem_desc<Cfg::kRowBytes>(0);
            const uint32_t desc_hi = static_cast<uint32_t>(proto >> 32);
            const uint32_t lo_flags = static_cast<uint32_t>(proto);
            constexpr uint32_t ROW16 = Cfg::kRowBytes / 16, HALO16 = Cfg::HALO_TILE / 16, X216 = Cfg::X2_TILE / 16, WT16 = Cfg::W_TILE / 16;
            mbar_wait(wbar, 0);
            tc_fence_after();
            const uint32_t w_lo0 = (smem_u32(wsm) >> 4) | lo_flags;
            const uint32_t kch = p.kchunks, kch2 = p.kchunks2;
            int slot = 0;
            uint32_t phase = 0;
            uint32_t g0 = 0;                           // running output-row counter (selects the accumulator)
            for (int it = blockIdx.x; it < items; it += gridDim.x) {
                int img, hb, he, w0;
                decode(it, img, hb, he, w0);
                const int j0 = max(hb - 1, 0), j1 = min(he, p.H - 1);
                for (int j = j0; j <= j1; ++j) {
                    mbar_wait(&full_bar[slot], phase);
                    tc_fence_after();
                    const uint32_t s_lo0 = (smem_u32(slots + slot * p.slot_bytes) >> 4) | lo_flags;
#pragma unroll 1
                    for (int dh = 1; dh >= -1; --dh) {
                        const int r = j - dh;          // output row fed by input row j through vertical tap kh = dh + 1
                        if (r < hb || r >= he) continue;
                        const uint32_t g = g0 + static_cast<uint32_t>(r - hb);
                        const uint32_t acc = g % NACC;
                        const bool first = (j == max(r - 1, 0));
                        if (first) {
                            mbar_wait(&tempty_bar[acc], ((g / NACC) & 1) ^ 1);
                            tc_fence_after();
                        }
                        const uint32_t d_tmem = tmem_base + acc * BN;
                        uint32_t accum = first ? 0u : 1u;
                        uint32_t w_lo = w_lo0 + static_cast<uint32_t>(dh + 1) * 3u * kch * WT16;
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            uint32_t a_lo = s_lo0 + kw * ROW16;
                            for (uint32_t c = 0; c < kch; ++c) {
#pragma unroll
                                for (int kk = 0; kk < BK / 16; ++kk) {
                                    umma_bf16_lo(d_tmem, a_lo + 2 * kk, w_lo + 2 * kk, desc_hi, idesc, accum);
                                    accum = 1u;
                                }
                                a_lo += HALO16;
                                w_lo += WT16;
                            }
                        }
                        if (dh == 0 && kch2) {
                            uint32_t a_lo = s_lo0 + kch * HALO16;
                            uint32_t w2 = w_lo0 + 9u * kch * WT16;
                            for (uint32_t c = 0; c < kch2; ++c) {
#pragma unroll
                                for (int kk = 0; kk < BK / 16; ++kk) umma_bf16_lo(d_tmem, a_lo + 2 * kk, w2 + 2 * kk, desc_hi, idesc, 1u);
                                a_lo += X216;
                                w2 += WT16;
                            }
                        }
                    }
                    umma_commit(&empty_bar[slot]);     // the row slot can be refilled once these MMAs retire
                    if (j - 1 >= hb && j - 1 < he) umma_commit(&tfull_bar[(g0 + static_cast<uint32_t>(j - 1 - hb)) % NACC]);
                    if (j == p.H - 1 && j >= hb && j < he) umma_commit(&tfull_bar[(g0 + static_cast<uint32_t>(j - hb)) % NACC]);
                    if (++slot == p.nslot) { slot = 0; phase ^= 1; }
                }
                g0 += static_cast<uint32_t>(he - hb);
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue: warps 2..5 = set 0, warps 6..9 = set 1 =====================
        const int set = (warp - 2) >> 2;
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
        const int m = quarter * 32 + lane;            // pixel within the strip
        const int ethread = threadIdx.x - 64;         // 0..255
        const int colbase = Cfg::ROW_SPLIT ? 0 : set * 32;
        float* bsm = bias_sm + set * 64;
        const uint32_t set_bar = 1 + set;             // named barrier id of this warp set (128 threads)
        uint32_t g0 = 0;
        for (int it = blockIdx.x; it < items; it += gridDim.x) {
            int img, hb, he, w0;
            decode(it, img, hb, he, w0);
            // stage bias (+ per-image time-embedding row) of this item once
            asm volatile("bar.sync %0, 128;" ::"r"(set_bar));          // previous item's readers are done
            {
                const int c = (ethread & 127);
                if (c < BN) {
                    float b = 0.f;
                    if (c < p.epi.n_valid) {
                        if (p.epi.bias) b += __ldg(p.epi.bias + c);
                        if (p.epi.bias_img) b += __ldg(p.epi.bias_img + img * p.epi.bias_img_stride + c);
                    }
                    bsm[c] = b;
                }
            }
            asm volatile("bar.sync %0, 128;" ::"r"(set_bar));
            float ssum[CPT], ssq[CPT];
#pragma unroll
            for (int q = 0; q < CPT; ++q) ssum[q] = ssq[q] = 0.f;
            for (int r = hb + (Cfg::ROW_SPLIT ? set : 0); r < he; r += (Cfg::ROW_SPLIT ? 2 : 1)) {
                const uint32_t g = g0 + static_cast<uint32_t>(r - hb);
                const uint32_t acc = g % NACC;
                mbar_wait(&tfull_bar[acc], (g / NACC) & 1);
                tc_fence_after();
                const long long pix = static_cast<long long>(r) * p.W + w0 + m;
                const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + colbase;
                uint32_t rr[CPT / 16][16];
#pragma unroll
                for (int q = 0; q < CPT / 16; ++q) tmem_ld_x16(t_addr + q * 16, rr[q]);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);          // accumulator is in registers: release it early
#pragma unroll
                for (int q = 0; q < CPT / 16; ++q) {
                    const int col0 = colbase + q * 16;
                    if (col0 >= p.epi.n_valid) continue;
                    float v[16];
#pragma unroll
                    for (int jj = 0; jj < 16; jj += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(bsm + col0 + jj);   // bsm holds all BN columns
                        v[jj] = __uint_as_float(rr[q][jj]) + b4.x;
                        v[jj + 1] = __uint_as_float(rr[q][jj + 1]) + b4.y;
                        v[jj + 2] = __uint_as_float(rr[q][jj + 2]) + b4.z;
                        v[jj + 3] = __uint_as_float(rr[q][jj + 3]) + b4.w;
                    }
                    if (p.epi.residual) {
                        const uint4* rp = reinterpret_cast<const uint4*>(p.epi.residual + img * p.epi.res_img_stride + pix * p.epi.res_row_stride + col0);
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            const uint4 u = __ldg(rp + h2);
                            const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                v[h2 * 8 + 2 * jj] += __uint_as_float(uu[jj] << 16);
                                v[h2 * 8 + 2 * jj + 1] += __uint_as_float(uu[jj] & 0xFFFF0000u);
                            }
                        }
                    }
                    if (p.epi.stats) {
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) {
                            ssum[q * 16 + jj] += v[jj];
                            ssq[q * 16 + jj] = fmaf(v[jj], v[jj], ssq[q * 16 + jj]);
                        }
                    }
                    epilogue_store16(p.epi, img, pix, col0, v);
                }
            }
            if (p.epi.stats) {
#pragma unroll
                for (int q = 0; q < CPT / 16; ++q) {
                    const int col0 = colbase + q * 16;
                    // butterfly over the 32 lanes (= 32 pixels): afterwards even lanes hold a channel sum, odd lanes a sum of squares
                    float s[16], qq[16];
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) { s[jj] = ssum[q * 16 + jj]; qq[jj] = ssq[q * 16 + jj]; }
#pragma unroll
                    for (int half = 8, bit = 16; half >= 1; half >>= 1, bit >>= 1) {
                        const bool up = (lane & bit) != 0;
#pragma unroll
                        for (int i = 0; i < half; ++i) {
                            const float s_send = up ? s[i] : s[i + half];
                            const float s_keep = up ? s[i + half] : s[i];
                            s[i] = s_keep + __shfl_xor_sync(0xffffffffu, s_send, bit);
                            const float q_send = up ? qq[i] : qq[i + half];
                            const float q_keep = up ? qq[i + half] : qq[i];
                            qq[i] = q_keep + __shfl_xor_sync(0xffffffffu, q_send, bit);
                        }
                    }
                    const bool odd = lane & 1;
                    const float recv = __shfl_xor_sync(0xffffffffu, odd ? s[0] : qq[0], 1);
                    const float tot = odd ? qq[0] + recv : s[0] + recv;
                    const int c = col0 + colsum16_col(lane);
                    if (c < p.epi.n_valid)
                        atomicAdd(p.epi.stats + (static_cast<long long>(img) * p.epi.n_valid + c) * 2 + (lane & 1), static_cast<double>(tot));
                }
            }
            g0 += static_cast<uint32_t>(he - hb);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

}  // namespace pnpf
