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
// Accumulators form a ring of NACC TMEM buffers, one per output row; 4 epilogue warps drain finished rows while the next
// rows are being accumulated.  Optional epilogue side output: per-channel sum / sum-of-squares of the produced tensor
// (GroupNorm statistics of the NEXT layer), reduced with warp shuffles and flushed with one fp64 atomic per lane per item.
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
    static constexpr int THREADS = 192;
    static constexpr int BAR_BYTES = 512;
    static_assert(BN == 16 || BN == 32 || BN == 64, "row conv is for thin outputs");
};

template <int BK, int BN>
__global__ void __launch_bounds__(RowCfg<BK, BN>::THREADS, 1)
rowconv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ RowConvParams p) {
    using Cfg = RowCfg<BK, BN>;
    constexpr int NACC = Cfg::NACC;
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
            mbar_init(&tempty_bar[a], 4);
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
            mbar_wait(wbar, 0);
            tc_fence_after();
            const uint32_t w_addr = smem_u32(wsm);
            int slot = 0;
            uint32_t phase = 0;
            long long g0 = 0;                          // running output-row counter (selects the accumulator)
            for (int it = blockIdx.x; it < items; it += gridDim.x) {
                int img, hb, he, w0;
                decode(it, img, hb, he, w0);
                const int j0 = max(hb - 1, 0), j1 = min(he, p.H - 1);
                for (int j = j0; j <= j1; ++j) {
                    mbar_wait(&full_bar[slot], phase);
                    tc_fence_after();
                    const uint32_t s_addr = smem_u32(slots + slot * p.slot_bytes);
#pragma unroll 1
                    for (int dh = 1; dh >= -1; --dh) {
                        const int r = j - dh;          // output row fed by input row j through vertical tap kh = dh + 1
                        if (r < hb || r >= he) continue;
                        const long long g = g0 + (r - hb);
                        const int acc = static_cast<int>(g % NACC);
                        const uint32_t aphase = static_cast<uint32_t>((g / NACC) & 1);
                        const bool first = (j == max(r - 1, 0));
                        if (first) {
                            mbar_wait(&tempty_bar[acc], aphase ^ 1);
                            tc_fence_after();
                        }
                        const uint32_t d_tmem = tmem_base + acc * BN;
                        uint32_t accum = first ? 0u : 1u;
                        const int kh = dh + 1;
#pragma unroll 1
                        for (int kw = 0; kw < 3; ++kw) {
                            for (int c = 0; c < p.kchunks; ++c) {
                                const uint64_t adesc = make_smem_desc<Cfg::kRowBytes>(s_addr + c * Cfg::HALO_TILE + kw * Cfg::kRowBytes);
                                const uint64_t bdesc = make_smem_desc<Cfg::kRowBytes>(w_addr + ((kh * 3 + kw) * p.kchunks + c) * Cfg::W_TILE);
#pragma unroll
                                for (int kk = 0; kk < BK / 16; ++kk) {
                                    umma_bf16(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, accum);
                                    accum = 1u;
                                }
                            }
                        }
                        if (dh == 0 && p.kchunks2) {
                            for (int c = 0; c < p.kchunks2; ++c) {
                                const uint64_t adesc = make_smem_desc<Cfg::kRowBytes>(s_addr + p.kchunks * Cfg::HALO_TILE + c * Cfg::X2_TILE);
                                const uint64_t bdesc = make_smem_desc<Cfg::kRowBytes>(w_addr + (9 * p.kchunks + c) * Cfg::W_TILE);
#pragma unroll
                                for (int kk = 0; kk < BK / 16; ++kk) umma_bf16(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, 1u);
                            }
                        }
                    }
                    umma_commit(&empty_bar[slot]);     // the row slot can be refilled once these MMAs retire
                    if (j - 1 >= hb && j - 1 < he) umma_commit(&tfull_bar[(g0 + (j - 1 - hb)) % NACC]);   // row j-1 complete
                    if (j == p.H - 1 && j >= hb && j < he) umma_commit(&tfull_bar[(g0 + (j - hb)) % NACC]);  // bottom edge
                    if (++slot == p.nslot) { slot = 0; phase ^= 1; }
                }
                g0 += he - hb;
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue warps 2..5 =====================
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;
        long long g0 = 0;
        for (int it = blockIdx.x; it < items; it += gridDim.x) {
            int img, hb, he, w0;
            decode(it, img, hb, he, w0);
            float st[BN / 16];
#pragma unroll
            for (int q = 0; q < BN / 16; ++q) st[q] = 0.f;
            for (int r = hb; r < he; ++r) {
                const long long g = g0 + (r - hb);
                const int acc = static_cast<int>(g % NACC);
                const uint32_t aphase = static_cast<uint32_t>((g / NACC) & 1);
                mbar_wait(&tfull_bar[acc], aphase);
                tc_fence_after();
                const long long pix = static_cast<long long>(r) * p.W + w0 + m;
                const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
#pragma unroll
                for (int q = 0; q < BN / 16; ++q) {
                    uint32_t rr[16];
                    tmem_ld_x16(t_addr + q * 16, rr);
                    tmem_ld_wait();
                    float v[16];
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) v[jj] = __uint_as_float(rr[jj]);
                    const bool act = q * 16 < p.epi.n_valid;
                    if (act) epilogue_apply16(p.epi, img, pix, q * 16, v);
                    if (p.epi.stats && act) st[q] += warp_colsum16(v, true, lane);
                    if (act) epilogue_store16(p.epi, img, pix, q * 16, v);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);
            }
            if (p.epi.stats) {
#pragma unroll
                for (int q = 0; q < BN / 16; ++q) {
                    const int c = q * 16 + colsum16_col(lane);
                    if (c < p.epi.n_valid)
                        atomicAdd(p.epi.stats + (static_cast<long long>(img) * p.epi.n_valid + c) * 2 + (lane & 1), static_cast<double>(st[q]));
                }
            }
            g0 += he - hb;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

}  // namespace pnpf
