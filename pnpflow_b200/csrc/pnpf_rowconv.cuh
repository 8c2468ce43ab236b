// Row-streaming 3x3 (stride 1) implicit-GEMM convolution for wide feature maps (W % 128 == 0) and thin outputs
// (C_out <= 64) — the U-Net's two highest resolutions, where the per-tap kernel of pnpf_gemm.cuh is L2-bound because it
// fetches every input pixel nine times.
//
// A CTA owns a column strip of 128 output pixels and walks down a segment of rows.  Every INPUT row is fetched ONCE by TMA
// as a 130-pixel halo tile (w0-1 .. w0+128, out-of-image pixels zero-filled by the TMA unit = the conv padding) and feeds
// nine tensor-core taps:
//   * horizontal taps: the UMMA A-descriptor starts 0, 1 or 2 pixel-rows into the swizzled halo tile (the 128B/64B
//     swizzle is a function of the shared-memory address bits, so a row-shifted start address reads the right data with
//     base_offset = 0 — verified on hardware by tools/probes/umma_offset_probe.cu);
//   * vertical taps: input row j feeds output rows j+1, j, j-1 with the SAME A operand, so the three taps are stacked
//     along N: the resident weight tile of (kw, chunk) holds [kh][cout] rows and the accumulators of consecutive output
//     rows occupy DEcreasing TMEM column blocks; one tcgen05.mma with N = 3*C_out updates three output rows.
// With N = 32 an MMA is only 16 tensor clocks, so the kernel is organised around the MMA-issuing threads:
//   * every MMA accumulates (the epilogue re-zeroes an accumulator block with tcgen05.st after draining it), which
//     removes all per-instruction special cases; descriptor words live in registers, the tap loops are fully unrolled
//     (template KCH = channel chunks) and only add immediates;
//   * two issuer warps take alternate input rows: the MMAs are issued in row order (mbarrier hand-over), the barrier waits,
//     commits and bookkeeping around them overlap (see the MMA role below);
//   * waits are issued by one lane per warp; two epilogue warp sets (alternate rows, or the two 32-column halves of C_out = 64),
//     4 or 8 GroupNorm-transform warps (RowCfg);
//   * bias (+ the per-image time-embedding row) is staged in shared memory once per work item;
//   * GroupNorm statistics of the OUTPUT tensor are accumulated per thread across the rows of an item and reduced with
//     warp shuffles + one fp64 atomic per lane per item.
#pragma once
#include "pnpf_gemm.cuh"
#include <cuda_fp16.h>
#include <type_traits>

namespace pnpf {

struct RowConvParams {
    int H, W, n_img;
    int strips;            // W / 128
    int nsplit;            // 1, or 2: CTAs 2g / 2g+1 compute output channels [0,BN) / [BN,2BN) of the same rows (C_out = 2*BN)
    int kchunks;           // Cin / BK (== template KCH)
    int kchunks2;          // C2 / BK of the fused 1x1 source (0 = none)
    int nslot;             // depth of the input-row ring
    int slot_bytes;
    // The main input (and the fused 1x1 input) may be a channel concat [a | b] of two tensors (skip connections):
    // the first kch_a (kch2_a) chunks are fetched through the "a" tensor map, the rest through the "b" map.
    int kch_a, kch2_a;
    // Fused GroupNorm(+SiLU) of the main input: the halo tile is normalised IN PLACE in shared memory by two transform
    // warps between the TMA and the MMAs (per-channel statistics come from the producing conv's epilogue).
    int gn;                // 0 = input is used as is
    int gn_silu, gn_gs, gn_Ca, gn_Cb;    // activation flag, channels per group, channels of source a / b
    float gn_eps;
    const float* gn_gamma; // [Ca+Cb]
    const float* gn_beta;
    const double* gn_st_a; // [img][Ca][2]
    const double* gn_st_b; // [img][Cb][2]
    int mma2;              // 1: two MMA-issuing warps on alternate input rows (RowCfg::NMMA == 2 only)
    int staged_store;      // fp16 NHWC output through the per-warp staging tiles (coalesced); else per-thread global stores
    long long* dbg;        // optional [32] cycle counters of CTA 0 (profiling experiments), else nullptr
    EpiParams epi;
};

template <int BK, int BN, int NEW_>
struct RowCfg {
    static constexpr int kRowBytes = BK * 2;
    static constexpr int HALO_ROWS = 130;
    static constexpr int HALO_TILE = (136 * kRowBytes + 1023) / 1024 * 1024;
    static constexpr int X2_TILE = 128 * kRowBytes;
    static constexpr int W_TILE_RAW = BN * BK * 2;
    static constexpr int W_TILE = (W_TILE_RAW + 1023) / 1024 * 1024;      // fused 1x1 shortcut tile [BN][BK]
    static constexpr int KH_BYTES = BN * kRowBytes;                         // one vertical tap inside a stacked tile
    static constexpr int W_STACK = 3 * KH_BYTES;                            // [kh][BN][BK] weights of one (kw, chunk)
    static_assert(KH_BYTES % 1024 == 0, "stacked tap tiles must keep the swizzle phase");
    // Warp-role configurations (third template argument):
    //   8  : 8 epilogue warps (two sets) + 4 GroupNorm-transform warps, 128 registers            (no fused GroupNorm: transform idle)
    //   12 : WIDE = 8 epilogue + 8 transform warps, 96 registers — for the fused-GroupNorm layers, where the 4-warp transform
    //        (one latency chain per scheduler: ld.shared -> unpack -> fma -> tanh -> pack -> st.shared, issuing 26 % of the time)
    //        is as slow as the MMA issuer
    // (a two-CTA-per-SM configuration with 4 + 4 worker warps was measured slower in round 1 and removed:
    //  profiles/r01_ab_experiments.md)
    static constexpr bool WIDE = NEW_ == 12;
    // Below 128 registers the per-thread GroupNorm statistics of the output (2 x 32 accumulators) would spill, so they are
    // accumulated in the READ phase of the staged store instead (a lane sees 8 channels of 4 pixels per row: 16 accumulators), on
    // the fp16 values that are actually stored — exactly the tensor the next GroupNorm normalises.  The host selects this
    // configuration only together with the staged store.
    static constexpr bool STAGED_STATS = WIDE;
    static constexpr int NACC = (512 / BN) > 16 ? 16 : (512 / BN);
    static constexpr int TMEM_COLS = NACC * BN;              // 256 (BN = 16) or 512
    static constexpr int MAX_SLOTS = 8;
    // Worker warps besides the producer and the MMA issuer: NEW epilogue warps (sets of 4, one warp per TMEM lane quarter) and
    // 4 GroupNorm-transform warps.  Both are per-warp LATENCY chains (barrier wait, tcgen05.ld/st round trips or ld/st.shared +
    // fence.proxy.async).  NEW = 8: two epilogue sets (alternate rows for BN <= 32, split columns for BN = 64).  One set + 8
    // transform warps in ONE CTA measured the same or slower: the transform is limited by shared-memory bandwidth, not by its
    // warp count (profiles/r01_ab_experiments.md).
    static constexpr int NEW = WIDE ? 8 : NEW_;
    static_assert(NEW == 4 || NEW == 8, "epilogue warps");
    static_assert(BN <= 32 || NEW == 8, "C_out = 64 needs two epilogue sets");
    static_assert(NACC >= 8, "accumulator ring too shallow");
    static constexpr bool ROW_SPLIT = BN <= 32 && NEW == 8;  // the two sets alternate rows (else set s owns columns [32 s, 32 s + 32))
    static constexpr int NTW = WIDE ? 8 : 4;                 // transform warps
    // fp16 NHWC outputs are transposed through a per-warp staging tile (32 pixels x 64 bytes, 64B-swizzled): a lane packs its
    // pixel's 32 channels with four conflict-free st.shared.v4, then the warp stores eight whole 64-byte pixel rows per
    // instruction (per-lane 16-byte stores at a 64/128-byte stride cost one L1 wavefront per lane).  Warp-local: no barrier.
    static constexpr int STAGE_WARP = 32 * 64;
    static constexpr int STAGE_BYTES = BN >= 32 ? NEW * STAGE_WARP : 0;
    // A SECOND MMA-issuing warp (the last warp of the CTA): the two issuers take alternate input rows, which overlaps the
    // barrier waits, commits and bookkeeping of one row (~ 600 of the ~ 1000 clocks a single issuer spends per row: the limit of
    // the 32 -> 32 layers once the transform has 8 warps) with the MMA issue of the next.  RowConvParams::mma2 = 0 leaves the
    // second warp idle (A/B switch).
    static constexpr int NMMA = 2;
    static constexpr int THREADS = 64 + (NEW + NTW) * 32 + (NMMA - 1) * 32;
    // Register budget of WIDE: the register file is split over the four SM sub-partitions (16 K registers each); 18 warps are
    // allocated as 5 per sub-partition -> 5 x 32 x regs <= 16384 -> 96 registers (= the bound of 640 threads).
    static constexpr int BOUND_THREADS = WIDE ? 640 : THREADS;
    static constexpr int CPT = BN > 32 ? 32 : BN;            // columns per epilogue thread
    static constexpr int BAR_BYTES = 2048;                   // barriers + tmem slot | bias staging (2 x 64 floats) | GN scale/shift (2 x 128)
    static_assert(BN == 16 || BN == 32 || BN == 64, "row conv is for thin outputs");
};

// tcgen05.mma, always accumulating, descriptors given as (low word, shared high word)
__device__ __forceinline__ void umma_acc_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc)
        : "memory");
}
// registers -> TMEM, 32 lanes x 16 columns of zeros (re-arms an accumulator block)
__device__ __forceinline__ void tmem_zero_x16(uint32_t taddr) {
    const uint32_t z = 0;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(z)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t saddr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// ld/st.shared of the transform stage: volatile keeps them (and their relative order), but without a memory clobber, so the
// arithmetic of later chunks can be scheduled across the stores of earlier ones
__device__ __forceinline__ uint4 lds128_nc(uint32_t saddr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts128_nc(uint32_t saddr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}
// named barrier of epilogue warp set 0 / 1 (128 threads).  The id must be an immediate: with a register id ptxas reserves all
// 16 hardware barriers for the CTA and a second CTA can never become resident on the SM.
__device__ __forceinline__ void bar_sync_set(int set) {
    if (set == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
    else asm volatile("bar.sync 2, 128;" ::: "memory");
}

template <int BK, int BN, int KCH, int NEW_>
__global__ void __launch_bounds__(RowCfg<BK, BN, NEW_>::BOUND_THREADS, 1)
rowconv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAb,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmA2b,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ RowConvParams p) {
    using Cfg = RowCfg<BK, BN, NEW_>;
    constexpr int NACC = Cfg::NACC;
    constexpr int CPT = Cfg::CPT;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* wsm = smem;                                  // [kw][chunk] stacked tiles, then the shortcut tiles
    const int w_bytes = 3 * KCH * Cfg::W_STACK + p.kchunks2 * Cfg::W_TILE;
    uint8_t* slots = smem + w_bytes;
    uint8_t* stage = slots + p.nslot * p.slot_bytes;                // [epilogue warp] output staging tiles
    uint64_t* bars = reinterpret_cast<uint64_t*>(stage + (p.staged_store ? Cfg::STAGE_BYTES : 0));
    uint64_t* wbar = bars;
    uint64_t* full_bar = bars + 1;
    uint64_t* empty_bar = full_bar + Cfg::MAX_SLOTS;
    uint64_t* tfull_bar = empty_bar + Cfg::MAX_SLOTS;
    uint64_t* tempty_bar = tfull_bar + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 16);
    uint64_t* ready_bar = tempty_bar + 17;                          // [MAX_SLOTS] transform -> MMA (GroupNorm fusion)
    uint64_t* issued_bar = ready_bar + Cfg::MAX_SLOTS;              // [2] MMA issuer k has issued the MMAs of its n-th row (bytes 464..479)
    float* bias_sm = reinterpret_cast<float*>(bars) + 128;          // [2 sets][64] floats at byte offset 512
    float* gn_tab = reinterpret_cast<float*>(bars) + 256;           // scale[128] | shift[128] at byte offset 1024

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_mma = (Cfg::NMMA == 2 && p.mma2) ? 2 : 1;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (p.kchunks2) tma_prefetch_desc(&tmA2);
        if (p.kch_a < KCH) tma_prefetch_desc(&tmAb);
        if (p.kch2_a < p.kchunks2) tma_prefetch_desc(&tmA2b);
        mbar_init(wbar, 1);
        for (int s = 0; s < Cfg::MAX_SLOTS; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
            mbar_init(&ready_bar[s], Cfg::NTW);
        }
        mbar_init(&issued_bar[0], 1);
        mbar_init(&issued_bar[1], 1);
        for (int a = 0; a < 16; ++a) {
            mbar_init(&tfull_bar[a], n_mma);   // one tcgen05.commit per contributing issuer
            mbar_init(&tempty_bar[a], Cfg::ROW_SPLIT ? 4 : Cfg::NEW);
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (warp >= 2 && warp < 6) {                      // zero every accumulator block once (all MMAs accumulate)
        const uint32_t t0 = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16);
        for (int c = 0; c < Cfg::TMEM_COLS; c += 16) tmem_zero_x16(t0 + c);
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // Work distribution: the rows of all (image, strip) columns form one flattened row space of R = n_img * strips * H
    // rows; CTA group g of G owns the contiguous range [g*R/G, (g+1)*R/G) and walks it as (image, strip, [hb, he))
    // pieces, so every CTA streams the same number of rows (+-1) and pays the two halo rows once per piece.
    const int n_groups = gridDim.x / p.nsplit;
    const int grp = blockIdx.x / p.nsplit;
    const int n_off = (blockIdx.x - grp * p.nsplit) * BN;          // first output channel of this CTA
    const long long R_total = static_cast<long long>(p.n_img) * p.strips * p.H;         // < 2^31 (checked by the host)
    const int row_begin = static_cast<int>(R_total * grp / n_groups), row_end = static_cast<int>(R_total * (grp + 1) / n_groups);
    auto decode = [&](int row, int& img, int& hb, int& he, int& w0) {
        const int is = row / p.H;
        hb = row - is * p.H;
        he = min(p.H, hb + (row_end - row));
        img = is / p.strips;
        w0 = (is - img * p.strips) * 128;
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        // The whole warp stays converged and computes warp-uniform values; only the async instructions are issued by
        // one elected lane (elect.sync lets the compiler keep descriptors in uniform registers — a divergent `lane == 0`
        // branch forces R2UR moves and a uniformisation loop around every TMA / MMA instruction).
        {
            if (elect_one_sync()) {
                mbar_arrive_expect_tx(wbar, (9 * KCH + p.kchunks2) * Cfg::W_TILE_RAW);
                for (int kw = 0; kw < 3; ++kw)
                    for (int c = 0; c < KCH; ++c)
                        for (int kh = 0; kh < 3; ++kh)      // packed K order is (kh, kw, cin): see pack_conv_weight
                            tma_load_3d(wsm + (kw * KCH + c) * Cfg::W_STACK + kh * Cfg::KH_BYTES, &tmB, wbar,
                                        ((kh * 3 + kw) * KCH + c) * BK, n_off, 0);
                for (int c = 0; c < p.kchunks2; ++c)
                    tma_load_3d(wsm + 3 * KCH * Cfg::W_STACK + c * Cfg::W_TILE, &tmB, wbar, (9 * KCH + c) * BK, n_off, 0);
            }
            __syncwarp();
            int slot = 0;
            uint32_t phase = 0;
            long long c_wait = 0, c_rows = 0;
            const long long c_start = PNPF_CLK();
            for (int row = row_begin; row < row_end;) {
                int img, hb, he, w0;
                decode(row, img, hb, he, w0);
                row += he - hb;
                const int j0 = max(hb - 1, 0), j1 = min(he, p.H - 1);
                for (int j = j0; j <= j1; ++j) {
                    {   // the producer runs nslot rows ahead: back off instead of spinning (see mbar_wait_sleep)
                        const long long _t0 = PNPF_CLK();
                        mbar_wait_sleep(&empty_bar[slot], phase ^ 1, 200);
                        c_wait += PNPF_CLK() - _t0;
                    }
                    ++c_rows;
                    uint8_t* sp = slots + slot * p.slot_bytes;
                    const bool centre = (j >= hb) && (j < he) && p.kchunks2;
                    if (elect_one_sync()) {
                        mbar_arrive_expect_tx(&full_bar[slot], KCH * Cfg::HALO_ROWS * Cfg::kRowBytes + (centre ? p.kchunks2 * Cfg::X2_TILE : 0));
#pragma unroll
                        for (int c = 0; c < KCH; ++c)
                            tma_load_4d(sp + c * Cfg::HALO_TILE, c < p.kch_a ? &tmA : &tmAb, &full_bar[slot],
                                        (c < p.kch_a ? c : c - p.kch_a) * BK, w0 - 1, j, img);
                        if (centre)
                            for (int c = 0; c < p.kchunks2; ++c)
                                tma_load_4d(sp + KCH * Cfg::HALO_TILE + c * Cfg::X2_TILE, c < p.kch2_a ? &tmA2 : &tmA2b, &full_bar[slot],
                                            (c < p.kch2_a ? c : c - p.kch2_a) * BK, w0, j, img);
                    }
                    __syncwarp();
                    if (++slot == p.nslot) { slot = 0; phase ^= 1; }
                }
            }
            if (p.dbg && blockIdx.x == 0 && lane == 0) { p.dbg[0] = PNPF_CLK() - c_start; p.dbg[1] = c_wait; p.dbg[2] = c_rows; }
        }
        __syncwarp();
    } else if (warp == 1 || warp == 2 + Cfg::NEW + Cfg::NTW) {
        // ===================== MMA issuer(s) (converged warp, elected lane issues) =====================
        // With two issuers (n_mma == 2) warp 1 takes the even and the last warp the odd input rows of the running row sequence.
        // The MMAs themselves are still ISSUED in row order (issuer of row q waits on issued_bar until row q-1 has been issued,
        // with tcgen05 fences on both sides), so every accumulator receives its taps in the same order as with one issuer and
        // the results stay bit-reproducible; what overlaps is everything around the issue: barrier waits, commits, bookkeeping
        // (~ 600 of the ~ 1000 clocks one issuer spends per row).  Who signals what:
        //   * output row r receives taps from input rows r-1 (kh = 0), r (kh = 1), r+1 (kh = 2); rows r-1 and r+1 belong to one
        //     issuer, row r to the other, and a tcgen05.commit only covers the MMAs of the committing thread.  tfull[r] therefore
        //     expects two arrivals: the issuer of row r commits to it after its centre tap, the other one after row r+1 (its last
        //     tap; at the bottom image edge, where row r+1 does not exist, after row r-1);
        //   * each issuer waits for the drained (re-zeroed) accumulator before ITS first tap into it: rows j+1 and j every row,
        //     row j-1 only at the top of an item (otherwise it touched that accumulator two rows ago).
        const int mw = warp == 1 ? 0 : 1;
        if (mw < n_mma) {
            // descriptor words: hi is shared by every operand tile; lo = (addr >> 4) | LBO bit
            const uint64_t proto = make_smem_desc<Cfg::kRowBytes>(0);
            const uint32_t desc_hi = static_cast<uint32_t>(proto >> 32);
            const uint32_t lo_flags = static_cast<uint32_t>(proto);
            constexpr uint32_t ROW16 = Cfg::kRowBytes / 16, HALO16 = Cfg::HALO_TILE / 16, X216 = Cfg::X2_TILE / 16, WT16 = Cfg::W_TILE / 16,
                               WS16 = Cfg::W_STACK / 16, KH16 = Cfg::KH_BYTES / 16;
            constexpr uint32_t idesc1 = make_idesc_act16(128, BN), idesc2 = make_idesc_act16(128, 2 * BN), idesc3 = make_idesc_act16(128, 3 * BN);
            mbar_wait(wbar, 0);
            tc_fence_after();
            const uint32_t w_lo0 = (smem_u32(wsm) >> 4) | lo_flags;
            const uint32_t s_base_lo = (smem_u32(slots) >> 4) | lo_flags;
            const uint32_t slot16 = static_cast<uint32_t>(p.slot_bytes) >> 4;
            const uint32_t kch2 = p.kchunks2;
            const bool two = n_mma == 2;
            uint32_t slot = 0, phase = 0;
            uint32_t g0 = 0;                           // running output-row counter (selects the accumulator)
            uint32_t q = 0;                            // running input-row counter (selects the issuer)
            long long c_full = 0, c_tempty = 0, c_issue = 0, c_commit = 0;
            const long long c_start = PNPF_CLK();
            for (int row = row_begin; row < row_end;) {
                int img, hb, he, w0;
                decode(row, img, hb, he, w0);
                row += he - hb;
                const int j0 = max(hb - 1, 0), j1 = min(he, p.H - 1);
                for (int j = j0; j <= j1; ++j, ++q) {
                    if (!two || (q & 1u) == static_cast<uint32_t>(mw)) {
                    // taps kh = 0,1,2 feed output rows j+1, j, j-1; the valid ones are contiguous in kh
                    const int k0 = (j + 1 < he) ? 0 : ((j < he) ? 1 : 2);
                    const int k1 = (j - 1 >= hb) ? 2 : ((j >= hb) ? 1 : 0);
                    // accumulator of output row (j + 1 - kh): ring index acc = g % NACC, column block = NACC-1-acc
                    const uint32_t g_top = g0 + static_cast<uint32_t>(j + 1 - k0 - hb);           // row of tap k0 (largest row)
                    const uint32_t acc_top = g_top % NACC;
                    const uint32_t blk0 = (NACC - 1) - acc_top;                                     // block of tap k0
                    const uint32_t ntap = static_cast<uint32_t>(k1 - k0 + 1);
                    // taps k0.. occupy blocks blk0, blk0+1, ... until the ring wraps at NACC
                    const uint32_t n0 = min(ntap, static_cast<uint32_t>(NACC) - blk0);
                    const uint32_t n1 = ntap - n0;                                                  // wrapped part starts at block 0
                    if (!two) {
                        if (k0 == 0)                   // tap 0 opens a fresh accumulator (row j+1): it must have been drained
                            PNPF_TIMED_WAIT(&tempty_bar[acc_top], ((g_top / NACC) & 1) ^ 1, c_tempty);
                        if (j == 0)                    // top image row (hb == 0): row 0 is opened by its centre tap
                            PNPF_TIMED_WAIT(&tempty_bar[g0 % NACC], ((g0 / NACC) & 1) ^ 1, c_tempty);
                    } else {
                        // first tap of THIS issuer into the accumulators of rows j+1 and j (and j-1 at the top of the item)
                        const int r_lo = (j - 2 >= j0) ? j : j - 1;
                        for (int r = j + 1; r >= r_lo; --r)
                            if (r >= hb && r < he) {
                                const uint32_t g = g0 + static_cast<uint32_t>(r - hb);
                                PNPF_TIMED_WAIT(&tempty_bar[g % NACC], ((g / NACC) & 1) ^ 1, c_tempty);
                            }
                    }
                    tc_fence_after();
                    PNPF_TIMED_WAIT(p.gn ? &ready_bar[slot] : &full_bar[slot], phase, c_full);
                    if (two && q > 0) {                // row q-1 (other issuer, its row number (q-1)/2) has been issued
                        mbar_wait(&issued_bar[mw ^ 1], ((q - 1) >> 1) & 1);
                    }
                    tc_fence_after();
                    const long long c_i0 = PNPF_CLK();
                    const uint32_t s_lo0 = s_base_lo + slot * slot16;
                    const uint32_t d0 = tmem_base + blk0 * BN, d1 = tmem_base;
                    const uint32_t i0 = n0 == 3 ? idesc3 : (n0 == 2 ? idesc2 : idesc1);
                    const uint32_t i1 = n1 == 2 ? idesc2 : idesc1;
                    const uint32_t wk0 = w_lo0 + static_cast<uint32_t>(k0) * KH16;
                    const uint32_t wk1 = wk0 + n0 * KH16;
                    if (elect_one_sync()) {
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
#pragma unroll
                            for (int c = 0; c < KCH; ++c) {
#pragma unroll
                                for (int kk = 0; kk < BK / 16; ++kk) {
                                    const uint32_t a_lo = s_lo0 + (kw * ROW16 + c * HALO16 + 2 * kk);
                                    const uint32_t w_off = (kw * KCH + c) * WS16 + 2 * kk;
                                    umma_acc_lo(d0, a_lo, wk0 + w_off, desc_hi, i0);
                                    if (n1) umma_acc_lo(d1, a_lo, wk1 + w_off, desc_hi, i1);
                                }
                            }
                        }
                        if (kch2 && j >= hb && j < he) {   // fused 1x1 shortcut: centre row only -> accumulator of row j
                            const uint32_t gc = g0 + static_cast<uint32_t>(j - hb);
                            const uint32_t dc = tmem_base + ((NACC - 1) - (gc % NACC)) * BN;
                            uint32_t a_lo = s_lo0 + KCH * HALO16;
                            uint32_t w2 = w_lo0 + 3u * KCH * WS16;
                            for (uint32_t c = 0; c < kch2; ++c) {
#pragma unroll
                                for (int kk = 0; kk < BK / 16; ++kk) umma_acc_lo(dc, a_lo + 2 * kk, w2 + 2 * kk, desc_hi, idesc1);
                                a_lo += X216;
                                w2 += WT16;
                            }
                        }
                        if (two) {                         // hand the issue order to the other warp before the (slow) commits
                            tc_fence_before();
                            mbar_arrive(&issued_bar[mw]);
                        }
                        const long long c_i1 = PNPF_CLK();
                        c_issue += c_i1 - c_i0;
                        umma_commit(&empty_bar[slot]);     // the row slot can be refilled once these MMAs retire
                        const bool out_c = j >= hb && j < he, out_m = j - 1 >= hb && j - 1 < he;
                        if (out_m) umma_commit(&tfull_bar[(g0 + static_cast<uint32_t>(j - 1 - hb)) % NACC]);     // last tap of row j-1
                        if (!two) {
                            if (j == p.H - 1 && out_c) umma_commit(&tfull_bar[(g0 + static_cast<uint32_t>(j - hb)) % NACC]);
                        } else {
                            if (out_c) umma_commit(&tfull_bar[(g0 + static_cast<uint32_t>(j - hb)) % NACC]);     // this issuer's only tap of row j
                            if (j + 1 == p.H - 1 && j + 1 < he)      // bottom image edge: no row j+2 will carry the second arrival of row j+1
                                umma_commit(&tfull_bar[(g0 + static_cast<uint32_t>(j + 1 - hb)) % NACC]);
                        }
                        c_commit += PNPF_CLK() - c_i1;
                    }
                    __syncwarp();
                    }
                    if (++slot == static_cast<uint32_t>(p.nslot)) { slot = 0; phase ^= 1; }
                }
                g0 += static_cast<uint32_t>(he - hb);
            }
            if (p.dbg && blockIdx.x == 0 && lane == 0 && mw == 0) { p.dbg[4] = PNPF_CLK() - c_start; p.dbg[5] = c_full; p.dbg[6] = c_tempty; p.dbg[16] = c_issue; p.dbg[17] = c_commit; }
        }
        __syncwarp();
    } else if (warp >= 2 + Cfg::NEW) {
        // ===================== GroupNorm(+SiLU) transform: warps 10..13 normalise each halo tile in place =====================
        if (p.gn) {
            constexpr int NTT = Cfg::NTW * 32;            // transform threads
            const int tt = threadIdx.x - (64 + Cfg::NEW * 32);
            const int Ctot = p.gn_Ca + p.gn_Cb;
            int slot = 0;
            uint32_t phase = 0;
            long long c_twait = 0, c_tab = 0, c_fence = 0;
            const long long c_tstart = PNPF_CLK();
            for (int row = row_begin; row < row_end;) {
                int img, hb, he, w0;
                decode(row, img, hb, he, w0);
                row += he - hb;
                const long long c_t0 = PNPF_CLK();
                // per-image scale / shift of every input channel (a group may straddle the two concatenated sources)
                asm volatile("bar.sync 3, %0;" ::"n"(NTT));
                for (int c = tt; c < Ctot; c += NTT) {
                    const int g0c = (c / p.gn_gs) * p.gn_gs;
                    double S = 0, Q = 0;
                    for (int k = 0; k < p.gn_gs; ++k) {
                        const int cc = g0c + k;
                        const double* sp2 = (cc < p.gn_Ca) ? p.gn_st_a + (static_cast<long long>(img) * p.gn_Ca + cc) * 2
                                                           : p.gn_st_b + (static_cast<long long>(img) * p.gn_Cb + (cc - p.gn_Ca)) * 2;
                        S += sp2[0];
                        Q += sp2[1];
                    }
                    const double n = static_cast<double>(p.gn_gs) * p.H * p.W;
                    const double mean = S / n;
                    double var = Q / n - mean * mean;
                    if (var < 0) var = 0;
                    const float rstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(p.gn_eps)));
                    const float sc = rstd * __ldg(p.gn_gamma + c);
                    gn_tab[c] = sc;
                    gn_tab[128 + c] = __ldg(p.gn_beta + c) - static_cast<float>(mean) * sc;
                }
                asm volatile("bar.sync 3, %0;" ::"n"(NTT));
                // Thread -> 16-byte chunk mapping: consecutive lanes take consecutive chunks (conflict-free LDS/STS.128); the
                // 128 threads cover RPI = 2048 / rowbytes pixel rows per iteration, a multiple of 8, so a thread's swizzle
                // phase — hence the 8 channels its chunk holds — never changes: scale/shift stay in registers per item.
                constexpr int CPR = Cfg::kRowBytes / 16;           // chunks per pixel row (4 or 8)
                constexpr int RPI = NTT / CPR;                     // rows per iteration (16..64, always a multiple of 8)
                const int q = tt % CPR, row0 = tt / CPR;
                const int sw = (Cfg::kRowBytes == 128) ? (row0 & 7) : ((row0 >> 1) & 3);
                // silu(y) = h + h * tanh(h) with h = y / 2: the 1/2 is folded into scale / shift
                const float fold = p.gn_silu ? 0.5f : 1.f;
                float tsc[KCH][8], tsh[KCH][8];
#pragma unroll
                for (int c = 0; c < KCH; ++c)
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int ch = c * BK + ((q ^ sw) << 3) + e;
                        tsc[c][e] = fold * gn_tab[ch];
                        tsh[c][e] = fold * gn_tab[128 + ch];
                    }
                const int j0 = max(hb - 1, 0), j1 = min(he, p.H - 1);
                c_tab += PNPF_CLK() - c_t0;
                // One halo tile: all of a thread's 16-byte chunks are loaded first (independent ld.shared in flight), then
                // transformed as one straight-line block (the activation is a compile-time branch and invalid rows are computed
                // and discarded, so the scheduler can interleave the dependent chains) and stored with a predicate.
                // The NIT iterations cover NIT * RPI >= 130 pixel rows; only rows 128 and 129 fall into the LAST iteration, i.e. only
                // the first warp of the transform group has valid lanes there: every other warp skips it (warp-uniform branch).
                // Round-2 ncu: the kernel issues 3490 warp instructions per row on four schedulers in ~ 1140 clocks (77 % of the
                // issue slots), 1500 of them here — the rows are issue-bound, so instructions are what has to go.
                constexpr int NIT = (Cfg::HALO_ROWS + RPI - 1) / RPI;
                static_assert((NIT - 1) * RPI < Cfg::HALO_ROWS && (NIT - 1) * RPI + NTT / CPR > Cfg::HALO_ROWS - 1, "only the last iteration is partial");
                const bool last_it = ((tt & ~31) / CPR + (NIT - 1) * RPI) < Cfg::HALO_ROWS;       // warp-uniform
                auto transform_tile = [&](auto silu_tag, uint32_t sbase, int c) {
                    constexpr bool kSilu = decltype(silu_tag)::value;
                    uint4 u[NIT];
                    bool ok[NIT];
#pragma unroll
                    for (int k = 0; k < NIT; ++k) {
                        const int r = row0 + k * RPI;
                        const int wpix = w0 - 1 + r;
                        ok[k] = (r < Cfg::HALO_ROWS) && (wpix >= 0) && (wpix < p.W);       // conv zero padding stays zero
                        u[k] = make_uint4(0u, 0u, 0u, 0u);
                        if (k < NIT - 1 || last_it)
                            if (r < Cfg::HALO_ROWS) u[k] = lds128_nc(sbase + c * Cfg::HALO_TILE + k * RPI * Cfg::kRowBytes);
                    }
#pragma unroll
                    for (int k = 0; k < NIT; ++k) {
                        if (k == NIT - 1 && !last_it) break;
                        uint32_t wds[4] = {u[k].x, u[k].y, u[k].z, u[k].w};
#pragma unroll
                        for (int e2 = 0; e2 < 4; ++e2) {
                            const float2 xin = unpack2(wds[e2]);
                            float y0 = fmaf(xin.x, tsc[c][2 * e2], tsh[c][2 * e2]);
                            float y1 = fmaf(xin.y, tsc[c][2 * e2 + 1], tsh[c][2 * e2 + 1]);
                            if constexpr (kSilu) {
                                // silu(2h) = h + h tanh(h), fp32 MUFU.TANH per element.  (tanh.approx.f16x2 is split into two
                                // MUFU.TANH.F16 by ptxas on sm_100a and needs three extra conversions: 11-12 instructions per pair
                                // against 9 here.)
                                float t0, t1;
                                asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(y0));
                                asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(y1));
                                y0 = fmaf(y0, t0, y0);
                                y1 = fmaf(y1, t1, y1);
                            }
                            wds[e2] = pack2(y0, y1);
                        }
                        // stored as soon as it is ready: the fence below waits for the stores still in flight
                        if (ok[k]) sts128_nc(sbase + c * Cfg::HALO_TILE + k * RPI * Cfg::kRowBytes, make_uint4(wds[0], wds[1], wds[2], wds[3]));
                    }
                };
                for (int j = j0; j <= j1; ++j) {
                    {
                        const long long _t0 = PNPF_CLK();
                        mbar_wait_warp(&full_bar[slot], phase, lane);
                        c_twait += PNPF_CLK() - _t0;
                    }
                    const uint32_t sbase = smem_u32(slots) + slot * p.slot_bytes + row0 * Cfg::kRowBytes + q * 16;
                    if (p.gn_silu) {
#pragma unroll
                        for (int c = 0; c < KCH; ++c) transform_tile(std::true_type{}, sbase, c);
                    } else {
#pragma unroll
                        for (int c = 0; c < KCH; ++c) transform_tile(std::false_type{}, sbase, c);
                    }
                    const long long c_x0 = PNPF_CLK();
                    fence_proxy_async_smem();              // generic-proxy writes -> visible to the tensor core (async proxy)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&ready_bar[slot]);
                    c_fence += PNPF_CLK() - c_x0;
                    if (++slot == p.nslot) { slot = 0; phase ^= 1; }
                }
            }
            if (p.dbg && blockIdx.x == 0 && tt == 0) { p.dbg[3] = PNPF_CLK() - c_tstart; p.dbg[7] = c_twait; p.dbg[11] = c_tab; p.dbg[21] = c_fence; }
        }
    } else {
        // ===================== epilogue: warps 2..5 = set 0 (and warps 6..9 = set 1, columns 32..63, for C_out = 64) =====================
        const int set = (warp - 2) >> 2;
        const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
        const int m = quarter * 32 + lane;            // pixel within the strip
        const int ethread = threadIdx.x - 64;         // 0..255
        const int colbase = Cfg::ROW_SPLIT ? 0 : set * 32;   // column split: set 1 takes columns 32..63 of C_out = 64
        float* bsm = bias_sm + set * 64;
        // staging tile of this warp: write phase = own pixel row, read phase = 8 pixels x 4 chunks per instruction
        const uint32_t stage_w = smem_u32(stage) + (warp - 2) * Cfg::STAGE_WARP;
        const uint32_t st_wr = stage_w + lane * 64, st_wr_sw = static_cast<uint32_t>((lane >> 1) & 3);
        const int rd_pix = lane >> 2, rd_chunk = lane & 3;
        uint32_t g0 = 0;
        long long c_tfull = 0, c_rows = 0, c_ld = 0, c_st = 0, c_fin = 0, c_math = 0, c_out = 0;
        const long long c_start = PNPF_CLK();
        for (int row = row_begin; row < row_end;) {
            int img, hb, he, w0;
            decode(row, img, hb, he, w0);
            row += he - hb;
            // stage bias (+ per-image time-embedding row) of this item once
            bar_sync_set(set);          // previous item's readers are done
            {
                const int c = (ethread & 127);
                if (c < BN) {
                    float b = 0.f;
                    if (n_off + c < p.epi.n_valid) {
                        if (p.epi.bias) b += __ldg(p.epi.bias + n_off + c);
                        if (p.epi.bias_img) b += __ldg(p.epi.bias_img + img * p.epi.bias_img_stride + n_off + c);
                    }
                    bsm[c] = b;
                }
            }
            bar_sync_set(set);
            constexpr int NST = Cfg::STAGED_STATS ? 8 : CPT;
            float ssum[NST], ssq[NST];
#pragma unroll
            for (int q = 0; q < NST; ++q) ssum[q] = ssq[q] = 0.f;
            for (int r = hb + (Cfg::ROW_SPLIT ? set : 0); r < he; r += (Cfg::ROW_SPLIT ? 2 : 1)) {
                const uint32_t g = g0 + static_cast<uint32_t>(r - hb);
                const uint32_t acc = g % NACC;
                const long long pix = static_cast<long long>(r) * p.W + w0 + m;
                {
                    const long long _t0 = PNPF_CLK();
                    mbar_wait_warp_sleep(&tfull_bar[acc], (g / NACC) & 1, lane, 32);
                    c_tfull += PNPF_CLK() - _t0;
                }
                ++c_rows;
                tc_fence_after();
                const long long c_e0 = PNPF_CLK();
                const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + ((NACC - 1) - acc) * BN + colbase;
                uint32_t rr[CPT / 16][16];
#pragma unroll
                for (int q = 0; q < CPT / 16; ++q) tmem_ld_x16(t_addr + q * 16, rr[q]);
                tmem_ld_wait();
                const long long c_e1 = PNPF_CLK();
#pragma unroll
                for (int q = 0; q < CPT / 16; ++q) tmem_zero_x16(t_addr + q * 16);   // re-arm: every MMA accumulates
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[acc]);          // accumulator is in registers: release it early
                const long long c_e2 = PNPF_CLK();
                c_ld += c_e1 - c_e0;
                c_st += c_e2 - c_e1;
#pragma unroll
                for (int q = 0; q < CPT / 16; ++q) {
                    const int col0 = colbase + q * 16;              // column within this CTA's BN block
                    const int gcol0 = n_off + col0;                  // output channel
                    if (gcol0 >= p.epi.n_valid) continue;
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
                        // (prefetching these loads ahead of the accumulator wait was measured 4 % SLOWER: profiles/r01_ab_experiments.md)
                        const uint4* rp = reinterpret_cast<const uint4*>(p.epi.residual + img * p.epi.res_img_stride + pix * p.epi.res_row_stride + gcol0);
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            const uint4 u = __ldg(rp + h2);
                            const uint32_t uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                const float2 t = unpack2(uu[jj]);
                                v[h2 * 8 + 2 * jj] += t.x;
                                v[h2 * 8 + 2 * jj + 1] += t.y;
                            }
                        }
                    }
                    if constexpr (!Cfg::STAGED_STATS) {
                        if (p.epi.stats) {
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) {
                                ssum[q * 16 + jj] += v[jj];
                                ssq[q * 16 + jj] = fmaf(v[jj], v[jj], ssq[q * 16 + jj]);
                            }
                        }
                    }
                    if constexpr (CPT == 32) {
                        if (p.staged_store) {                          // pack to fp16 into this pixel's swizzled staging row
                            uint32_t pk[8];
#pragma unroll
                            for (int jj = 0; jj < 8; ++jj) pk[jj] = pack2(v[2 * jj], v[2 * jj + 1]);
                            sts128(st_wr + (((2 * q) ^ st_wr_sw) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
                            sts128(st_wr + (((2 * q + 1) ^ st_wr_sw) << 4), make_uint4(pk[4], pk[5], pk[6], pk[7]));
                            continue;
                        }
                    }
                    epilogue_store16(p.epi, img, pix, gcol0, v);
                }
                const long long c_e3 = PNPF_CLK();
                c_math += c_e3 - c_e2;
                if constexpr (CPT == 32) {
                    if (p.staged_store) {
                        __syncwarp();
                        // 4 instructions x (8 pixels x 64 bytes): whole 64-byte pixel rows per store
                        act16* orow = reinterpret_cast<act16*>(p.epi.out) + img * p.epi.out_img_stride +
                                              (static_cast<long long>(r) * p.W + w0 + quarter * 32) * p.epi.out_row_stride + n_off + colbase + rd_chunk * 8;
                        uint4 o4[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int pr = i * 8 + rd_pix;
                            o4[i] = lds128(stage_w + pr * 64 + ((rd_chunk ^ ((pr >> 1) & 3)) << 4));
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) *reinterpret_cast<uint4*>(orow + (i * 8 + rd_pix) * p.epi.out_row_stride) = o4[i];
                        if constexpr (Cfg::STAGED_STATS) {
                            if (p.epi.stats) {             // channels 8 rd_chunk .. + 7 of pixels rd_pix, 8 + rd_pix, ...
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const uint32_t ww[4] = {o4[i].x, o4[i].y, o4[i].z, o4[i].w};
#pragma unroll
                                    for (int e2 = 0; e2 < 4; ++e2) {
                                        const float2 t = unpack2(ww[e2]);
                                        const float f0 = t.x, f1 = t.y;
                                        ssum[2 * e2] += f0;
                                        ssq[2 * e2] = fmaf(f0, f0, ssq[2 * e2]);
                                        ssum[2 * e2 + 1] += f1;
                                        ssq[2 * e2 + 1] = fmaf(f1, f1, ssq[2 * e2 + 1]);
                                    }
                                }
                            }
                        }
                        __syncwarp();                                  // the tile is rewritten by the next row
                    }
                }
                c_out += PNPF_CLK() - c_e3;
            }
            const long long c_f0 = PNPF_CLK();
            if constexpr (Cfg::STAGED_STATS) {
                if (p.epi.stats) {
                    // the 8 lanes with the same rd_chunk hold partial sums of the same 8 channels: xor-reduce over lane bits 2..4,
                    // then lane (rd_pix = e) adds channel 8 rd_chunk + e
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
                    const int c = n_off + colbase + rd_chunk * 8 + rd_pix;
                    if (c < p.epi.n_valid) {
                        double* sp2 = p.epi.stats + (static_cast<long long>(img) * p.epi.n_valid + c) * 2;
                        atomicAdd(sp2, static_cast<double>(ts));
                        atomicAdd(sp2 + 1, static_cast<double>(tq));
                    }
                }
            } else if (p.epi.stats) {
#pragma unroll
                for (int q = 0; q < CPT / 16; ++q) {
                    const int col0 = n_off + colbase + q * 16;
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
            c_fin += PNPF_CLK() - c_f0;
        }
        if (p.dbg && blockIdx.x == 0 && ethread == 0) { p.dbg[18] = c_ld; p.dbg[19] = c_st; p.dbg[20] = c_fin; p.dbg[22] = c_math; p.dbg[23] = c_out; }
        if (p.dbg && blockIdx.x == 0 && ethread == 0) { p.dbg[8] = PNPF_CLK() - c_start; p.dbg[9] = c_tfull; p.dbg[10] = c_rows; }
        if (p.dbg && blockIdx.x == 0 && ethread == 128 && Cfg::NEW == 8) { p.dbg[12] = PNPF_CLK() - c_start; p.dbg[13] = c_tfull; p.dbg[14] = c_rows; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

}  // namespace pnpf
