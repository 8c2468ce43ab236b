// Patch-streaming 3x3 (stride 1) implicit-GEMM convolution for the NARROW, WIDE-CHANNEL levels of the U-Net (W <= 128,
// C_out 64 / 128 / 256: the 64^2 and 32^2 levels, the nearest-up convs and the level-1 blocks the row kernel cannot hold).  The per-tap kernel of pnpf_gemm.cuh fetches every
// input pixel nine times and is bound by the L2 -> shared-memory fill; the row kernel of pnpf_rowconv.cuh needs W % 128 == 0
// and resident weights.  Here the image is addressed in a PADDED-LINEAR space:
//
//   position o = h * P + wp,  P = W + 2,  wp in [0, W) a pixel, wp in {W, W+1} padding (never stored)
//
// so that all nine taps of a tile of 128 consecutive positions [o0, o0 + 128) are the SAME input patch read at nine row
// offsets:  tap (kh, kw) of output position o reads patch row  (o - h_first * P) + kh * P + kw,  h_first = o0 / P.
// The patch (input rows h_first-1 .. h_first-1+NR-1, columns -1 .. W) is ONE TMA box per 64-channel chunk — out-of-image
// rows / columns are zero-filled by the TMA unit, which is the conv padding — and lands in shared memory as NR * P pixel
// rows of 128 bytes in the canonical 128B-swizzled K-major layout; each tap is a UMMA descriptor whose start address is
// shifted by whole pixel rows (base_offset 0: tools/probes/umma_offset_probe.cu).  Per tile the activations are fetched
// ~1.3-1.7x instead of 9x; the weights stream through their own ring, one [BN x 64] tile per (chunk, tap).
// PAIR: two CTAs (cta_group::2) take the SAME tile of two consecutive images — one MMA instruction carries one A descriptor
// for both CTAs, so their patch origins must coincide — and each stages half of every weight tile (see pnpf_gemm.cuh).
//
// SUBPIX (default since round 2; PNPF_NO_SUBPIXEL=1 turns it off): one PHASE of "nearest-neighbour x2 upsampling followed by a 3x3 conv" (models.py:41-47)
// computed directly on the LOW-resolution tensor.  Output pixel (2h+a, 2w+b) only sees the 2x2 low-resolution pixels
// (h-1+a+i, w-1+b+j), i,j in {0,1}, with the 3x3 weights folded into 2x2 (fold_subpixel_weights): four launches (a,b) with
// 4 taps each replace one launch with 9 taps on a 4x larger tensor — 2.25x fewer FLOPs, and the upsampled tensor is never
// materialised.  In the padded-linear space tap (i,j) is simply patch-row offset (a+i)*P + (b+j); the epilogue scatters the
// tile to the stride-2 lattice of the high-resolution output.
//
// SUBPIX = 2 (round 2, default for C_out <= 128): BOTH column phases b = 0, 1 of an output-row parity a in one launch, as a
// 2 x 3-tap convolution with N = 2 C_out: tap (i, c) reads patch-row offset (a + i) P + c, accumulator columns [b C_out, (b+1) C_out)
// belong to output pixel (2h + a, 2w + b), and the packed weights hold W_ab[i][c - b] (zero where c - b is not 0 / 1:
// pack_subpixel_pair_weights).  The patch is fetched once for two phases (the single-phase form at C_out = 64 was bound by the
// L2 -> shared-memory patch fill: 66 KB per 16 short MMAs), every MMA is N = 128 / 256 wide instead of 64 / 128, a thread's two
// output pixels are adjacent (2 C_out contiguous values), and an up conv is two launches instead of four.
//
// TG = 3 (default since round 2; PNPF_PATCH_TG1=1 is the A/B off-switch): taps per weight-ring slot.  The MMA issuer pays a barrier wait and a tcgen05.commit per weight
// slot (~ 100 clocks), which is exposed when the four MMAs of a tap are short (N = 128: 4 x 64 clocks, measured ~ 90 per MMA;
// N = 64: 4 x 32): with TG = 3 a slot holds the three taps (kh, 0..2) of a kernel row, so the issuer waits and commits once per
// 12 MMAs.  Same arithmetic in the same order.  The group's twelve descriptors are one 32-bit low word plus compile-time immediates
// (umma_f16_lohi): 59 issuer clocks per MMA instead of 84 (profiles/r02_ab_experiments.md section 18).
//
// Epilogue: eight warps (two column halves).  BN <= 128, fp16 NHWC output: 32 columns at a time through a per-warp 2 KB staging tile
// (epilogue_chunk32_staged: whole 64-byte pixel rows per store instruction, GroupNorm statistics summed in the read phase); otherwise
// per-lane stores and the shuffle butterfly (epilogue_chunk32).  The accumulator-drained arrive of the CTA-pair form is RELAXED
// (mbar_arrive_cluster in pnpf_ptx.cuh: the release form is a GPU-scope fence per tile).
#pragma once
#include "pnpf_gemm.cuh"

namespace pnpf {

struct PatchConvParams {
    int H, W, P;           // P = W + 2
    int NR;                // patch rows
    int n_img, tiles_per_img;
    int kchunks;           // ceil(C_in / 64): a trailing half chunk (C_in % 64 == 32) is zero-filled by the TMA unit in the patch, so
                           // whatever the weight box holds beyond this tap's C_in columns (the next tap's weights) is multiplied by 0
    int cin;               // C_in: the packed K index of (tap, channel) is tap * C_in + channel
    int kchunks2;          // channel chunks of the fused 1x1 source (extra K through tmA2, centre tap only); 0 = none
    int patch_bytes;       // NR * P * 128 rounded up to 1024
    int na, nb;            // ring depths: patches, weight tiles
    int stage_bytes;       // 8 x 2 KB output staging tiles of the epilogue warps (fp16 NHWC outputs of BN <= 128), else 0
    long long* dbg;
    EpiParams epi;
    int sp_a, sp_b;        // SUBPIX: phase (row, column parity of the output pixels this launch produces)
};

template <int BN, bool PAIR>
struct PatchCfg {
    static constexpr int BK = 64;
    static constexpr int B_ROWS = PAIR ? BN / 2 : BN;
    static constexpr int B_BYTES = B_ROWS * 128;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int MAX_A = 4, MAX_B = 12;
    // warps: 0 patch producer, 1 MMA issuer, 2..5 epilogue (columns [0, BN/2)), 6 weight producer, 7..10 epilogue (columns
    // [BN/2, BN)).  The epilogue is a per-warp latency chain (tcgen05.ld, bias / residual loads, statistics, stores):
    // with one warp set a 128 x 128 tile took longer to drain than to compute.
    static constexpr int THREADS = 11 * 32;
    static_assert(BN == 64 || BN == 128 || BN == 256, "patch conv output widths");
};

template <int BN, bool PAIR, int SUBPIX = 0, int TG = 1>      // SUBPIX: 0 = 3x3 conv, 1 = one sub-pixel phase, 2 = the two column phases of a row parity
__global__ void __launch_bounds__(PatchCfg<BN, PAIR>::THREADS, 1)
patchconv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                 const __grid_constant__ CUtensorMap tmB, const __grid_constant__ PatchConvParams p) {
    using Cfg = PatchCfg<BN, PAIR>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_ring = smem;                                    // [na] patches
    uint8_t* b_ring = smem + p.na * p.patch_bytes;             // [nb] weight tiles
    uint8_t* stage = b_ring + p.nb * (TG * Cfg::B_BYTES);      // nb slots of TG weight tiles, then the epilogue's staging tiles
    uint64_t* a_full = reinterpret_cast<uint64_t*>(stage + p.stage_bytes);
    uint64_t* a_empty = a_full + Cfg::MAX_A;
    uint64_t* b_full = a_empty + Cfg::MAX_A;
    uint64_t* b_empty = b_full + Cfg::MAX_B;
    uint64_t* tfull_bar = b_empty + Cfg::MAX_B;                // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const int nch = p.kchunks + p.kchunks2;                    // patches per tile
    // (All CTAs walk the weight tiles in the same order on purpose: rotating the tap order per CTA to spread the requests over
    // different L2 lines measured 8 % SLOWER — concurrent requests for the same line are served together.)
    constexpr int tap_rot = 0;
    const int total_units = (p.n_img / (PAIR ? 2 : 1)) * p.tiles_per_img;      // pair: n_img is even (host)
    const int unit0 = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
    const int unit_step = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
    auto decode = [&](int u, int& img, int& o0, int& h_first) {
        const int ig = u / p.tiles_per_img;
        img = PAIR ? 2 * ig + static_cast<int>(rank) : ig;
        o0 = (u - ig * p.tiles_per_img) * 128;
        h_first = o0 / p.P;
    };

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (p.kchunks2) tma_prefetch_desc(&tmA2);
        for (int s = 0; s < Cfg::MAX_A; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < Cfg::MAX_B; ++s) { mbar_init(&b_full[s], 1); mbar_init(&b_empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], PAIR ? 16 : 8); }
        fence_barrier_init();
    }
    if (warp == 2) {
        if constexpr (PAIR) tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
        else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    if constexpr (PAIR) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== patch producer: one TMA box per (tile, 64-channel chunk) =====================
        int slot = 0;
        uint32_t phase = 0;
        long long c_wait = 0, c_tiles = 0;
        const long long c_start = PNPF_CLK();
        for (int u = unit0; u < total_units; u += unit_step) {
            int img, o0, h_first;
            decode(u, img, o0, h_first);
            ++c_tiles;
            for (int c = 0; c < nch; ++c) {
                PNPF_TIMED_WAIT(&a_empty[slot], phase ^ 1, c_wait);
                uint8_t* dst = a_ring + slot * p.patch_bytes;
                if (elect_one_sync()) {
                    const uint32_t bytes = static_cast<uint32_t>(p.NR * p.P * 128);
                    const CUtensorMap* tm = c < p.kchunks ? &tmA : &tmA2;
                    const int cc = (c < p.kchunks ? c : c - p.kchunks) * 64;
                    if constexpr (PAIR) {
                        const uint32_t fb = mapa_u32(smem_u32(&a_full[slot]), 0);
                        if (rank == 0) mbar_arrive_expect_tx(&a_full[slot], 2 * bytes);
                        tma_load_4d_pair(dst, tm, fb, cc, -1, h_first - 1, img);
                    } else {
                        mbar_arrive_expect_tx(&a_full[slot], bytes);
                        tma_load_4d(dst, tm, &a_full[slot], cc, -1, h_first - 1, img);
                    }
                }
                __syncwarp();
                if (++slot == p.na) { slot = 0; phase ^= 1; }
            }
        }
        if (p.dbg && blockIdx.x == 0 && lane == 0) { p.dbg[0] = PNPF_CLK() - c_start; p.dbg[1] = c_wait; p.dbg[2] = c_tiles; }
        __syncwarp();
    } else if (warp == 6) {
        // ===================== weight producer: one [B_ROWS x 64] tile per (chunk, tap) =====================
        int slot = 0;
        uint32_t phase = 0;
        long long c_wait = 0;
        const long long c_start = PNPF_CLK();
        for (int u = unit0; u < total_units; u += unit_step) {
            for (int c = 0; c < nch; ++c) {
                const int ntap = c < p.kchunks ? (SUBPIX == 1 ? 4 : (SUBPIX == 2 ? 6 : 9)) : 1;
                if constexpr (TG == 1) {
                    for (int t = 0; t < ntap; ++t) {
                        PNPF_TIMED_WAIT(&b_empty[slot], phase ^ 1, c_wait);
                        uint8_t* dst = b_ring + slot * Cfg::B_BYTES;
                        if (elect_one_sync()) {
                            // packed K order (pack_conv_weight): (kh, kw, cin) for the 3x3 part, then the 1x1 source channels
                            const int tap = SUBPIX ? t : (t + tap_rot) % 9;      // SUBPIX: packed K order (i, j, cin), 4 or 6 taps
                            const int k0 = c < p.kchunks ? tap * p.cin + c * 64 : 9 * p.cin + (c - p.kchunks) * 64;
                            if constexpr (PAIR) {
                                const uint32_t fb = mapa_u32(smem_u32(&b_full[slot]), 0);
                                if (rank == 0) mbar_arrive_expect_tx(&b_full[slot], 2 * Cfg::B_BYTES);
                                tma_load_3d_pair(dst, &tmB, fb, k0, static_cast<int>(rank) * Cfg::B_ROWS, 0);
                            } else {
                                mbar_arrive_expect_tx(&b_full[slot], Cfg::B_BYTES);
                                tma_load_3d(dst, &tmB, &b_full[slot], k0, 0, 0);
                            }
                        }
                        __syncwarp();
                        if (++slot == p.nb) { slot = 0; phase ^= 1; }
                    }
                } else {
                    static_assert(TG == 1 || SUBPIX != 1, "tap groups need three taps per kernel row");
                    for (int t0 = 0; t0 < ntap; t0 += TG) {
                        const int ng = min(TG, ntap - t0);           // the three taps of a kernel row, or the single 1x1 tap
                        PNPF_TIMED_WAIT(&b_empty[slot], phase ^ 1, c_wait);
                        uint8_t* dst = b_ring + slot * (TG * Cfg::B_BYTES);
                        if (elect_one_sync()) {
                            if constexpr (PAIR) {
                                const uint32_t fb = mapa_u32(smem_u32(&b_full[slot]), 0);
                                if (rank == 0) mbar_arrive_expect_tx(&b_full[slot], 2 * ng * Cfg::B_BYTES);
                                for (int j = 0; j < ng; ++j) {
                                    const int k0 = c < p.kchunks ? (t0 + j) * p.cin + c * 64 : 9 * p.cin + (c - p.kchunks) * 64;
                                    tma_load_3d_pair(dst + j * Cfg::B_BYTES, &tmB, fb, k0, static_cast<int>(rank) * Cfg::B_ROWS, 0);
                                }
                            } else {
                                mbar_arrive_expect_tx(&b_full[slot], ng * Cfg::B_BYTES);
                                for (int j = 0; j < ng; ++j) {
                                    const int k0 = c < p.kchunks ? (t0 + j) * p.cin + c * 64 : 9 * p.cin + (c - p.kchunks) * 64;
                                    tma_load_3d(dst + j * Cfg::B_BYTES, &tmB, &b_full[slot], k0, 0, 0);
                                }
                            }
                        }
                        __syncwarp();
                        if (++slot == p.nb) { slot = 0; phase ^= 1; }
                    }
                }
            }
        }
        if (p.dbg && blockIdx.x == 0 && lane == 0) { p.dbg[12] = PNPF_CLK() - c_start; p.dbg[13] = c_wait; }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (pair: the leader CTA only) =====================
        if (!PAIR || rank == 0) {
            constexpr uint32_t idesc = make_idesc_act16(PAIR ? 256 : 128, BN);
            int aslot = 0, bslot = 0;
            uint32_t aphase = 0, bphase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            long long c_afull = 0, c_bfull = 0, c_tempty = 0;
            const long long c_start = PNPF_CLK();
            for (int u = unit0; u < total_units; u += unit_step) {
                int img, o0, h_first;
                decode(u, img, o0, h_first);
                const int a_shift = o0 - h_first * p.P;           // first output position inside its row, in pixel rows
                PNPF_TIMED_WAIT(&tempty_bar[acc], acc_phase ^ 1, c_tempty);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int c = 0; c < nch; ++c) {
                    PNPF_TIMED_WAIT(&a_full[aslot], aphase, c_afull);
                    tc_fence_after();
                    const uint32_t pa = smem_u32(a_ring + aslot * p.patch_bytes);
                    const int ntap = c < p.kchunks ? (SUBPIX == 1 ? 4 : (SUBPIX == 2 ? 6 : 9)) : 1;
                    if constexpr (TG == 1) {
                        for (int t = 0; t < ntap; ++t) {
                            PNPF_TIMED_WAIT(&b_full[bslot], bphase, c_bfull);
                            tc_fence_after();
                            int kh, kw;
                            if constexpr (SUBPIX == 1) {
                                kh = p.sp_a + (t >> 1);
                                kw = p.sp_b + (t & 1);
                            } else if constexpr (SUBPIX == 2) {
                                kh = p.sp_a + t / 3;
                                kw = t - 3 * (t / 3);
                            } else {
                                const int tap = ntap == 9 ? (t + tap_rot) % 9 : 4;
                                kh = tap / 3;
                                kw = tap - 3 * kh;
                            }
                            const uint64_t adesc = make_smem_desc<128>(pa + static_cast<uint32_t>((kh * p.P + kw + a_shift) * 128));
                            const uint64_t bdesc = make_smem_desc<128>(smem_u32(b_ring + bslot * Cfg::B_BYTES));
                            if (elect_one_sync()) {
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) {
                                    if constexpr (PAIR) umma_f16_pair(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (c | t | kk) ? 1u : 0u);
                                    else umma_f16(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (c | t | kk) ? 1u : 0u);
                                }
                                if constexpr (PAIR) {
                                    umma_commit_pair(&b_empty[bslot]);
                                    if (t == ntap - 1) umma_commit_pair(&a_empty[aslot]);
                                    if (t == ntap - 1 && c == nch - 1) umma_commit_pair(&tfull_bar[acc]);
                                } else {
                                    umma_commit(&b_empty[bslot]);
                                    if (t == ntap - 1) umma_commit(&a_empty[aslot]);
                                    if (t == ntap - 1 && c == nch - 1) umma_commit(&tfull_bar[acc]);
                                }
                            }
                            __syncwarp();
                            if (++bslot == p.nb) { bslot = 0; bphase ^= 1; }
                        }
                    } else {
                        static_assert(TG == 3, "a weight-ring slot holds one kernel row");
                        constexpr uint32_t dhi = smem_desc_hi<128>();
                        for (int g = 0, t0 = 0; t0 < ntap; ++g, t0 += TG) {
                            const int ng = min(TG, ntap - t0);
                            PNPF_TIMED_WAIT(&b_full[bslot], bphase, c_bfull);
                            tc_fence_after();
                            if (elect_one_sync()) {
                                // first tap of the group: kernel row g (SUBPIX 2: sp_a + g), column 0; the fused 1x1 chunk has the
                                // centre tap only.  Descriptor low words advance by 8 per pixel row of the patch (128 B), by
                                // B_BYTES / 16 per weight tile and by 2 per 16-element K step: compile-time immediates below.
                                const int kh0 = ntap == 1 ? 1 : g + (SUBPIX == 2 ? p.sp_a : 0);
                                const int kw0 = ntap == 1 ? 1 : 0;
                                const uint32_t a_lo = smem_desc_lo<128>(pa) + static_cast<uint32_t>((kh0 * p.P + kw0 + a_shift) * 8);
                                const uint32_t b_lo = smem_desc_lo<128>(smem_u32(b_ring + bslot * (TG * Cfg::B_BYTES)));
                                const uint32_t accum0 = (c | g) ? 1u : 0u;
                                if (ng == TG) {
#pragma unroll
                                    for (int j = 0; j < TG; ++j)
#pragma unroll
                                        for (int kk = 0; kk < 4; ++kk)
                                            umma_f16_lohi<PAIR>(d_tmem, a_lo + 8 * j + 2 * kk, b_lo + j * (Cfg::B_BYTES >> 4) + 2 * kk, dhi, idesc,
                                                                (j | kk) ? 1u : accum0);
                                } else {
                                    for (int j = 0; j < ng; ++j)
#pragma unroll
                                        for (int kk = 0; kk < 4; ++kk)
                                            umma_f16_lohi<PAIR>(d_tmem, a_lo + 8 * j + 2 * kk, b_lo + j * (Cfg::B_BYTES >> 4) + 2 * kk, dhi, idesc,
                                                                (j | kk) ? 1u : accum0);
                                }
                                const bool last = t0 + ng == ntap;
                                if constexpr (PAIR) {
                                    umma_commit_pair(&b_empty[bslot]);
                                    if (last) umma_commit_pair(&a_empty[aslot]);
                                    if (last && c == nch - 1) umma_commit_pair(&tfull_bar[acc]);
                                } else {
                                    umma_commit(&b_empty[bslot]);
                                    if (last) umma_commit(&a_empty[aslot]);
                                    if (last && c == nch - 1) umma_commit(&tfull_bar[acc]);
                                }
                            }
                            __syncwarp();
                            if (++bslot == p.nb) { bslot = 0; bphase ^= 1; }
                        }
                    }
                    if (++aslot == p.na) { aslot = 0; aphase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            if (p.dbg && blockIdx.x == 0 && lane == 0) {
                p.dbg[4] = PNPF_CLK() - c_start; p.dbg[5] = c_afull; p.dbg[6] = c_tempty; p.dbg[7] = c_bfull;
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue warps 2..5 (first half of the columns) and 7..10 (second half) =====================
        const int quarter = warp & 3;                             // TMEM lane quarter this warp may access
        const int col_lo = warp >= 7 ? BN / 2 : 0;
        const int m = quarter * 32 + lane;                        // accumulator row == position within the tile
        int acc = 0;
        uint32_t acc_phase = 0;
        long long c_tfull = 0;
        [[maybe_unused]] long long c_top = 0, c_body = 0, c_tail = 0;     // cycle counters (-DPNPF_ROWCONV_CLOCKS builds only)
        const long long c_start = PNPF_CLK();
        const uint32_t tempty_remote = PAIR ? mapa_u32(smem_u32(&tempty_bar[0]), 0) : 0u;
        for (int u = unit0; u < total_units; u += unit_step) {
            const long long kt0 = PNPF_CLK();
            int img, o0, h_first;
            decode(u, img, o0, h_first);
            const int o = o0 + m;
            const int h = o / p.P, wp = o - h * p.P;
            const bool valid = (h < p.H) && (wp < p.W);
            // SUBPIX 1: stride-2 lattice of phase (a, b); SUBPIX 2: this warp set's column half IS column phase b
            const int sp_b = SUBPIX == 2 ? (warp >= 7 ? 1 : 0) : p.sp_b;
            const long long pix = SUBPIX ? static_cast<long long>(2 * h + p.sp_a) * (2 * p.W) + 2 * wp + sp_b
                                         : static_cast<long long>(h) * p.W + wp;
            long long kt1;
            {
                const long long _t0 = PNPF_CLK();
                c_top += _t0 - kt0;
                if (lane == 0) mbar_wait(&tfull_bar[acc], acc_phase);
                __syncwarp();
                kt1 = PNPF_CLK();
                c_tfull += kt1 - _t0;
            }
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
            bool staged = false;
            if constexpr (SUBPIX == 0 && BN <= 128) staged = p.stage_bytes != 0;
            if (staged) {
                // the four pixels this lane stores in the read phase of the staging tile: positions quarter * 32 + 8 i + (lane >> 2)
                int pix_rd[4];
                uint32_t vmask_rd = 0;
                const int o_rd = o0 + quarter * 32 + (lane >> 2);
                int h_rd = o_rd / p.P, wp_rd = o_rd - h_rd * p.P;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    pix_rd[i] = h_rd * p.W + wp_rd;
                    if (h_rd < p.H && wp_rd < p.W) vmask_rd |= 1u << i;
                    wp_rd += 8;                                     // P >= 10: at most one row wrap per step
                    if (wp_rd >= p.P) { wp_rd -= p.P; ++h_rd; }
                }
                const uint32_t stage_w = smem_u32(stage) + static_cast<uint32_t>(warp < 6 ? warp - 2 : warp - 3) * 2048u;
#pragma unroll 1
                for (int c0 = col_lo; c0 < col_lo + BN / 2; c0 += 32)
                    epilogue_chunk32_staged(p.epi, t_addr, img, pix, valid, c0, lane, stage_w, pix_rd, vmask_rd);
            } else {
#pragma unroll 1
                for (int c0 = col_lo; c0 < col_lo + BN / 2; c0 += 32) {
                    if constexpr (SUBPIX == 2) epilogue_chunk32(p.epi, t_addr + col_lo, img, pix, valid, c0 - col_lo, lane);   // channel = column - b C_out
                    else epilogue_chunk32(p.epi, t_addr, img, pix, valid, c0, lane);
                }
            }
            const long long kt2 = PNPF_CLK();
            c_body += kt2 - kt1;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR) mbar_arrive_cluster(tempty_remote + acc * 8);     // relaxed: see pnpf_ptx.cuh
                else mbar_arrive(&tempty_bar[acc]);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            c_tail += PNPF_CLK() - kt2;
        }
        if (p.dbg && blockIdx.x == 0 && threadIdx.x == 64) {
            p.dbg[8] = PNPF_CLK() - c_start; p.dbg[9] = c_tfull;
            p.dbg[24] = c_top; p.dbg[25] = c_body; p.dbg[26] = c_tail;
        }
    }
    tc_fence_before();
    if constexpr (PAIR) {
        cluster_sync_all();
        if (warp == 2) tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
    } else {
        __syncthreads();
        if (warp == 2) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

}  // namespace pnpf
