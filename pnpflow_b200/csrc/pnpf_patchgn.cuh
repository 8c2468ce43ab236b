// Patch-streaming 3x3 convolution with the GroupNorm(+SiLU) of its input FUSED (opt-in: PNPF_PATCH_GN=1; written at the end of
// round 1 after the GPU budget was spent, NOT yet run on hardware) — the patch kernel of pnpf_patchconv.cuh plus what the row
// kernel already does for the two high-resolution levels:
//   * every main-input patch chunk (NR * P pixel rows x 64 channels) is normalised IN PLACE in shared memory by four transform
//     warps between its TMA arrival and the MMAs (scale / shift per channel from the producer epilogues' statistics; the conv
//     zero padding — patch positions outside the image — stays zero), once per patch and not once per tap;
//   * the main input and the fused 1x1 shortcut input may be channel concats [a | b] read from their two source tensors through
//     two tensor maps each (chunks below kch_a / kch2_a come from a, the rest from b), so neither the normalised nor the raw
//     concatenated tensor is ever written to HBM (gn_apply + its raw concat copy disappear for these layers).
// Barriers per patch slot: a_full (TMA -> transform, LOCAL to each CTA also for CTA pairs: each CTA's transform warps must see
// their own patch), a_ready (transform -> MMA, in the leader: 4 warp arrivals per CTA, remote for the second CTA of a pair),
// a_empty (MMA commit -> producer, multicast for pairs).  Raw (shortcut) chunks pass through the transform warps untouched so
// that one protocol covers every chunk.
#pragma once
#include "pnpf_patchconv.cuh"
#include "pnpf_rowconv.cuh"
#include <cuda_fp16.h>

namespace pnpf {

struct PatchGnParams {
    int H, W, P;           // P = W + 2
    int NR;                // patch rows
    int n_img, tiles_per_img;
    int kchunks;           // C_in / 64 (concat total)
    int kchunks2;          // channel chunks of the fused 1x1 source (concat total); 0 = none
    int kch_a, kch2_a;     // chunks that come from source a (the rest from source b)
    int patch_bytes;       // NR * P * 128 rounded up to 1024
    int na, nb;            // ring depths: patches, weight tiles
    int gn_silu, gn_gs, gn_Ca, gn_Cb;    // activation flag, channels per group, channels of source a / b
    float gn_eps;
    const float* gn_gamma; // [Ca+Cb]
    const float* gn_beta;
    const double* gn_st_a; // [img][Ca][2]
    const double* gn_st_b; // [img][Cb][2]
    EpiParams epi;
};

template <int BN, bool PAIR>
struct PatchGnCfg {
    static constexpr int B_ROWS = PAIR ? BN / 2 : BN;
    static constexpr int B_BYTES = B_ROWS * 128;
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int MAX_A = 4, MAX_B = 12;
    static constexpr int MAX_C = 512;                  // channels of the (concatenated) main input: scale / shift table
    static constexpr int TAB_BYTES = 2 * MAX_C * 4;
    static constexpr int BAR_BYTES = 512;
    // warps: 0 patch producer, 1 MMA issuer, 2..5 epilogue (columns [0, BN/2)), 6 weight producer, 7..10 epilogue (columns
    // [BN/2, BN)), 11..14 GroupNorm transform
    static constexpr int NTW = 4;
    static constexpr int THREADS = (11 + NTW) * 32;
    static_assert(BN == 64 || BN == 128 || BN == 256, "patch conv output widths");
};

template <int BN, bool PAIR, int TG = 1>      // TG: taps per weight-ring slot (1, or 3 = a kernel row: see pnpf_patchconv.cuh)
__global__ void __launch_bounds__(PatchGnCfg<BN, PAIR>::THREADS, 1)
patchgn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAb,
               const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmA2b,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ PatchGnParams p) {
    using Cfg = PatchGnCfg<BN, PAIR>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* a_ring = smem;                                    // [na] patches
    uint8_t* b_ring = smem + p.na * p.patch_bytes;             // [nb] weight tiles
    uint64_t* a_full = reinterpret_cast<uint64_t*>(b_ring + p.nb * (TG * Cfg::B_BYTES));      // nb slots of TG weight tiles
    uint64_t* a_empty = a_full + Cfg::MAX_A;
    uint64_t* a_ready = a_empty + Cfg::MAX_A;
    uint64_t* b_full = a_ready + Cfg::MAX_A;
    uint64_t* b_empty = b_full + Cfg::MAX_B;
    uint64_t* tfull_bar = b_empty + Cfg::MAX_B;                // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* gn_tab = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(a_full) + Cfg::BAR_BYTES);   // scale[MAX_C] | shift[MAX_C]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const int nch = p.kchunks + p.kchunks2;                    // patches per tile
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
        if (p.kch_a < p.kchunks) tma_prefetch_desc(&tmAb);
        if (p.kchunks2) tma_prefetch_desc(&tmA2);
        if (p.kch2_a < p.kchunks2) tma_prefetch_desc(&tmA2b);
        for (int s = 0; s < Cfg::MAX_A; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
            mbar_init(&a_ready[s], Cfg::NTW * (PAIR ? 2 : 1));
        }
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
        // ===================== patch producer: one TMA box per (tile, 64-channel chunk), signalling the LOCAL a_full =====================
        int slot = 0;
        uint32_t phase = 0;
        for (int u = unit0; u < total_units; u += unit_step) {
            int img, o0, h_first;
            decode(u, img, o0, h_first);
            for (int c = 0; c < nch; ++c) {
                mbar_wait(&a_empty[slot], phase ^ 1);
                uint8_t* dst = a_ring + slot * p.patch_bytes;
                if (elect_one_sync()) {
                    const uint32_t bytes = static_cast<uint32_t>(p.NR * p.P * 128);
                    const CUtensorMap* tm;
                    int cc;
                    if (c < p.kchunks) {
                        tm = c < p.kch_a ? &tmA : &tmAb;
                        cc = (c < p.kch_a ? c : c - p.kch_a) * 64;
                    } else {
                        const int c2 = c - p.kchunks;
                        tm = c2 < p.kch2_a ? &tmA2 : &tmA2b;
                        cc = (c2 < p.kch2_a ? c2 : c2 - p.kch2_a) * 64;
                    }
                    mbar_arrive_expect_tx(&a_full[slot], bytes);
                    tma_load_4d(dst, tm, &a_full[slot], cc, -1, h_first - 1, img);
                }
                __syncwarp();
                if (++slot == p.na) { slot = 0; phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 6) {
        // ===================== weight producer: one [B_ROWS x 64] tile per (chunk, tap) =====================
        int slot = 0;
        uint32_t phase = 0;
        for (int u = unit0; u < total_units; u += unit_step) {
            for (int c = 0; c < nch; ++c) {
                const int ntap = c < p.kchunks ? 9 : 1;
                for (int t0 = 0; t0 < ntap; t0 += TG) {
                    const int ng = min(TG, ntap - t0);               // TG = 3: the three taps of a kernel row share a slot
                    mbar_wait(&b_empty[slot], phase ^ 1);
                    uint8_t* dst = b_ring + slot * (TG * Cfg::B_BYTES);
                    if (elect_one_sync()) {
                        if constexpr (PAIR) {
                            if (rank == 0) mbar_arrive_expect_tx(&b_full[slot], 2 * ng * Cfg::B_BYTES);
                        } else {
                            mbar_arrive_expect_tx(&b_full[slot], ng * Cfg::B_BYTES);
                        }
                        for (int j = 0; j < ng; ++j) {
                            // packed K order (pack_conv_weight): (kh, kw, cin) for the 3x3 part, then the 1x1 source channels
                            const int k0 = c < p.kchunks ? ((t0 + j) * p.kchunks + c) * 64 : (9 * p.kchunks + (c - p.kchunks)) * 64;
                            if constexpr (PAIR) {
                                const uint32_t fb = mapa_u32(smem_u32(&b_full[slot]), 0);
                                tma_load_3d_pair(dst + j * Cfg::B_BYTES, &tmB, fb, k0, static_cast<int>(rank) * Cfg::B_ROWS, 0);
                            } else {
                                tma_load_3d(dst + j * Cfg::B_BYTES, &tmB, &b_full[slot], k0, 0, 0);
                            }
                        }
                    }
                    __syncwarp();
                    if (++slot == p.nb) { slot = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (pair: the leader CTA only); waits for the TRANSFORMED patch =====================
        if (!PAIR || rank == 0) {
            constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, BN);
            int aslot = 0, bslot = 0;
            uint32_t aphase = 0, bphase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int u = unit0; u < total_units; u += unit_step) {
                int img, o0, h_first;
                decode(u, img, o0, h_first);
                const int a_shift = o0 - h_first * p.P;           // first output position inside its row, in pixel rows
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int c = 0; c < nch; ++c) {
                    mbar_wait(&a_ready[aslot], aphase);
                    tc_fence_after();
                    const uint32_t pa = smem_u32(a_ring + aslot * p.patch_bytes);
                    const int ntap = c < p.kchunks ? 9 : 1;
                    for (int t0 = 0; t0 < ntap; t0 += TG) {
                        const int ng = min(TG, ntap - t0);
                        mbar_wait(&b_full[bslot], bphase);
                        tc_fence_after();
                        const uint32_t pb = smem_u32(b_ring + bslot * (TG * Cfg::B_BYTES));
                        if (elect_one_sync()) {
                            for (int j = 0; j < ng; ++j) {
                                const int t = t0 + j;
                                const int tap = ntap == 9 ? t : 4;
                                const int kh = tap / 3, kw = tap - 3 * kh;
                                const uint64_t adesc = make_smem_desc<128>(pa + static_cast<uint32_t>((kh * p.P + kw + a_shift) * 128));
                                const uint64_t bdesc = make_smem_desc<128>(pb + static_cast<uint32_t>(j * Cfg::B_BYTES));
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) {
                                    if constexpr (PAIR) umma_bf16_pair(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (c | t | kk) ? 1u : 0u);
                                    else umma_bf16(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (c | t | kk) ? 1u : 0u);
                                }
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
                    if (++aslot == p.na) { aslot = 0; aphase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp >= 11) {
        // ===================== GroupNorm(+SiLU) transform: warps 11..14 normalise each main-input patch chunk in place =====================
        constexpr int NTT = Cfg::NTW * 32;                    // 128 transform threads
        const int tt = threadIdx.x - 11 * 32;
        const int Ctot = p.gn_Ca + p.gn_Cb;
        const uint32_t ready_remote = PAIR ? mapa_u32(smem_u32(&a_ready[0]), 0) : 0u;
        // thread -> 16-byte piece q of pixel rows row0, row0 + 16, ...: 16 is a multiple of 8, so the 128B-swizzle phase (row & 7)
        // and hence the 8 channels of the thread's piece never change
        const int q = tt & 7, row0 = tt >> 3;
        const int lp = q ^ (row0 & 7);                        // logical 16-byte piece = channels [8 lp, 8 lp + 8) of the chunk
        const int rows = p.NR * p.P;
        const float fold = p.gn_silu ? 0.5f : 1.f;            // silu(y) = h + h * tanh(h) with h = y / 2
        int slot = 0;
        uint32_t phase = 0;
        int cur_img = -1;
        for (int u = unit0; u < total_units; u += unit_step) {
            int img, o0, h_first;
            decode(u, img, o0, h_first);
            if (img != cur_img) {
                // per-image scale / shift of every input channel (a group may straddle the two concatenated sources)
                asm volatile("bar.sync 1, %0;" ::"n"(NTT) : "memory");      // nobody still reads the previous image's table
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
                    gn_tab[c] = fold * sc;
                    gn_tab[Cfg::MAX_C + c] = fold * (__ldg(p.gn_beta + c) - static_cast<float>(mean) * sc);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(NTT) : "memory");
                cur_img = img;
            }
            for (int c = 0; c < nch; ++c) {
                mbar_wait_warp(&a_full[slot], phase, lane);
                if (c < p.kchunks) {
                    float tsc[8], tsh[8];
                    {
                        const float4* ts = reinterpret_cast<const float4*>(gn_tab + c * 64 + lp * 8);
                        const float4* th = reinterpret_cast<const float4*>(gn_tab + Cfg::MAX_C + c * 64 + lp * 8);
                        const float4 s0 = ts[0], s1 = ts[1], h0 = th[0], h1 = th[1];
                        tsc[0] = s0.x; tsc[1] = s0.y; tsc[2] = s0.z; tsc[3] = s0.w; tsc[4] = s1.x; tsc[5] = s1.y; tsc[6] = s1.z; tsc[7] = s1.w;
                        tsh[0] = h0.x; tsh[1] = h0.y; tsh[2] = h0.z; tsh[3] = h0.w; tsh[4] = h1.x; tsh[5] = h1.y; tsh[6] = h1.z; tsh[7] = h1.w;
                    }
                    const uint32_t sbase = smem_u32(a_ring + slot * p.patch_bytes) + q * 16;
                    // patch row r <-> image position (h_first - 1 + r / P, r % P - 1): tracked incrementally (P > 16)
                    int hh = h_first - 1, ww = row0 - 1;
                    for (int r = row0; r < rows; r += 64) {
                        uint4 u4[4];
                        bool ok[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int rk = r + 16 * k;
                            ok[k] = (rk < rows) && (hh >= 0) && (hh < p.H) && (ww >= 0) && (ww < p.W);    // the conv zero padding stays zero
                            u4[k] = make_uint4(0u, 0u, 0u, 0u);
                            if (rk < rows) u4[k] = lds128_nc(sbase + rk * 128);
                            ww += 16;
                            if (ww >= p.P - 1) { ww -= p.P; ++hh; }
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            uint32_t wds[4] = {u4[k].x, u4[k].y, u4[k].z, u4[k].w};
#pragma unroll
                            for (int e2 = 0; e2 < 4; ++e2) {
                                float y0 = fmaf(__uint_as_float(wds[e2] << 16), tsc[2 * e2], tsh[2 * e2]);
                                float y1 = fmaf(__uint_as_float(wds[e2] & 0xFFFF0000u), tsc[2 * e2 + 1], tsh[2 * e2 + 1]);
                                if (p.gn_silu) {
                                    const __half2 hx = __floats2half2_rn(y0, y1);
                                    uint32_t hb2 = *reinterpret_cast<const uint32_t*>(&hx), tb;
                                    asm("tanh.approx.f16x2 %0, %1;" : "=r"(tb) : "r"(hb2));
                                    const __half2 th2 = *reinterpret_cast<const __half2*>(&tb);
                                    const float2 o = __half22float2(__hfma2(hx, th2, hx));
                                    y0 = o.x;
                                    y1 = o.y;
                                }
                                __nv_bfloat162 b2 = __floats2bfloat162_rn(y0, y1);
                                wds[e2] = *reinterpret_cast<uint32_t*>(&b2);
                            }
                            if (ok[k]) sts128_nc(sbase + (r + 16 * k) * 128, make_uint4(wds[0], wds[1], wds[2], wds[3]));
                        }
                    }
                }
                fence_proxy_async_smem();                      // generic-proxy writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) {
                    if constexpr (PAIR) mbar_arrive_cluster(ready_remote + slot * 8);
                    else mbar_arrive(&a_ready[slot]);
                }
                if (++slot == p.na) { slot = 0; phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue warps 2..5 (first half of the columns) and 7..10 (second half) =====================
        const int quarter = warp & 3;                             // TMEM lane quarter this warp may access
        const int col_lo = warp >= 7 ? BN / 2 : 0;
        const int m = quarter * 32 + lane;                        // accumulator row == position within the tile
        int acc = 0;
        uint32_t acc_phase = 0;
        const uint32_t tempty_remote = PAIR ? mapa_u32(smem_u32(&tempty_bar[0]), 0) : 0u;
        for (int u = unit0; u < total_units; u += unit_step) {
            int img, o0, h_first;
            decode(u, img, o0, h_first);
            const int o = o0 + m;
            const int h = o / p.P, wp = o - h * p.P;
            const bool valid = (h < p.H) && (wp < p.W);
            const long long pix = static_cast<long long>(h) * p.W + wp;
            if (lane == 0) mbar_wait(&tfull_bar[acc], acc_phase);
            __syncwarp();
            tc_fence_after();
            const uint32_t t_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int c0 = col_lo; c0 < col_lo + BN / 2; c0 += 32) epilogue_chunk32(p.epi, t_addr, img, pix, valid, c0, lane);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR) mbar_arrive_cluster(tempty_remote + acc * 8);
                else mbar_arrive(&tempty_bar[acc]);
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
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
