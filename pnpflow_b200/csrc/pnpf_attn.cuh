// Fused self-attention core for the U-Net's 16x16 attention blocks (models.py:145-162: L = 256 tokens, C = 256 channels):
//
//     S = (q * C^-1/2) k^T   ->   P = softmax(S)   ->   O = P v   ->   y = x + O Wo^T + b        per image
//
// one CTA per (image, 128-query tile), everything between the q/k/v projections and the block output on chip:
// S lives in TMEM (128 lanes x 256 fp32 columns), the softmax runs in registers (one thread = one query row), P goes to shared
// memory as the fp16 A operand of the second GEMM, O comes back through TMEM, is normalised by the row sums, goes to shared
// memory as the A operand of the output projection, and the projection's epilogue adds bias + residual, accumulates the
// GroupNorm statistics of the block output and stores fp16 NHWC.  The fp32 logits (256 KB per image) and the fp16 probabilities
// never touch HBM; five launches (qk^T, softmax, pv, proj + the S/P round trips) become one.
//
// Operands (all fp16, produced by the existing kernels of the plan):
//   tmQ : [img][L][2C] "qk" tensor of the fused q|k projection, channels [0, C) (q already carries C^-1/2)   A operand, box {64, 128}
//   tmK : same tensor, channels [C, 2C): k as [N = L keys][K = C]                                             B operand, box {64, 256}
//   tmV : V^T [img][C][L] (the vT GEMM of the plan): [N = C][K = L keys]                                        B operand, box {64, 256}
//   tmW : packed projection weights [N = C_out][K = C]                                                          B operand, box {64, 256}
// Shared memory (192 KB): R1 = 64 KB (Q, then P, then O: four 128 x 64 K-major SW128 tiles), R2 = 128 KB (K, then V^T, then Wo:
// four 256 x 64 tiles).  TMEM (512 columns): S in [0, 256), O and then Y in [256, 512).
// Warps: 0 = TMA producer, 1 = MMA issuer, 2..9 = softmax / epilogue in two sets of four (warp w owns TMEM lanes [32 (w & 3), +32);
// set 0 takes keys / columns [0, 128), set 1 [128, 256); the row max and row sum are exchanged through shared memory).
// Per unit every mbarrier completes exactly once, so its wait parity is the unit counter's low bit.
#pragma once
#include "pnpf_gemm.cuh"

namespace pnpf {

struct AttnParams {
    int n_img;
    int L, C;              // must be 256, 256 (checked by the host)
    EpiParams epi;         // bias (folded Wo bv + bo), residual x, output y (fp16 NHWC = [img][token][C]), statistics
};

struct AttnCfg {
    static constexpr int L = 256, C = 256;
    static constexpr int A_TILE = 128 * 128;            // 128 rows x 64 fp16, SW128
    static constexpr int B_TILE = 256 * 128;            // 256 rows x 64 fp16, SW128
    static constexpr int R1_BYTES = 4 * A_TILE;         // 64 KB
    static constexpr int R2_BYTES = 4 * B_TILE;         // 128 KB
    static constexpr int XCH_BYTES = 2 * 2 * 128 * 4;   // row max / row sum of the two warp sets
    static constexpr int SMEM_BYTES = R1_BYTES + R2_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/ + XCH_BYTES;
    static constexpr int THREADS = 10 * 32;
    static constexpr int TMEM_COLS = 512;
};

// byte offset of the 16-byte unit holding elements [8u, 8u + 8) of row r inside a 128 x 64 (or 256 x 64) K-major SW128 tile
__device__ __forceinline__ uint32_t sw128_off(int r, int u) { return static_cast<uint32_t>(r * 128 + ((u ^ (r & 7)) << 4)); }

template <int kL, int kC>      // (a template so that the header can be included by several translation units)
__global__ void __launch_bounds__(AttnCfg::THREADS, 1)
attn_core_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                 const __grid_constant__ CUtensorMap tmW, const __grid_constant__ AttnParams p) {
    using Cfg = AttnCfg;
    static_assert(kL == Cfg::L && kC == Cfg::C, "the tiling below is written for 256 tokens x 256 channels");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* r1 = smem;
    uint8_t* r2 = smem + Cfg::R1_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(r2 + Cfg::R2_BYTES);
    uint64_t* qk_full = bars + 0;       // TMA: Q + K landed
    uint64_t* s_full = bars + 1;        // MMA: S complete (Q, K no longer read)
    uint64_t* v_full = bars + 2;        // TMA: V^T landed
    uint64_t* p_ready = bars + 3;       // softmax warps: P written
    uint64_t* o_full = bars + 4;        // MMA: O complete (P, V^T no longer read)
    uint64_t* w_full = bars + 5;        // TMA: Wo landed
    uint64_t* o_ready = bars + 6;       // epilogue warps: normalised O written
    uint64_t* y_full = bars + 7;        // MMA: Y complete (O, Wo no longer read)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    float* xch = reinterpret_cast<float*>(bars) + 64;   // [2 quantities][2 sets][128 rows] at byte offset 256

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tiles_per_img = Cfg::L / 128;
    const int total_units = p.n_img * tiles_per_img;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        tma_prefetch_desc(&tmW);
        mbar_init(qk_full, 1);
        mbar_init(s_full, 1);
        mbar_init(v_full, 1);
        mbar_init(p_ready, 8);
        mbar_init(o_full, 1);
        mbar_init(w_full, 1);
        mbar_init(o_ready, 8);
        mbar_init(y_full, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        uint32_t par = 0;
        bool first = true;
        for (int u = blockIdx.x; u < total_units; u += gridDim.x, par ^= 1) {
            const int img = u / tiles_per_img, q0 = (u - img * tiles_per_img) * 128;
            if (!first) mbar_wait(y_full, par ^ 1);                   // previous unit's projection no longer reads R1 / R2
            first = false;
            if (elect_one_sync()) {
                mbar_arrive_expect_tx(qk_full, Cfg::R1_BYTES + Cfg::R2_BYTES);
#pragma unroll
                for (int c = 0; c < 4; ++c) tma_load_4d(r1 + c * Cfg::A_TILE, &tmQ, qk_full, c * 64, q0, 0, img);
#pragma unroll
                for (int c = 0; c < 4; ++c) tma_load_3d(r2 + c * Cfg::B_TILE, &tmK, qk_full, c * 64, 0, img);
            }
            __syncwarp();
            mbar_wait(s_full, par);                                    // K consumed -> R2 takes V^T
            if (elect_one_sync()) {
                mbar_arrive_expect_tx(v_full, Cfg::R2_BYTES);
#pragma unroll
                for (int c = 0; c < 4; ++c) tma_load_3d(r2 + c * Cfg::B_TILE, &tmV, v_full, c * 64, 0, img);
            }
            __syncwarp();
            mbar_wait(o_full, par);                                    // V^T consumed -> R2 takes Wo
            if (elect_one_sync()) {
                mbar_arrive_expect_tx(w_full, Cfg::R2_BYTES);
#pragma unroll
                for (int c = 0; c < 4; ++c) tma_load_3d(r2 + c * Cfg::B_TILE, &tmW, w_full, c * 64, 0, 0);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: three 128 x 256 x 256 GEMMs per unit =====================
        constexpr uint32_t idesc = make_idesc_act16(128, 256);
        const uint32_t a_base = smem_u32(r1), b_base = smem_u32(r2);
        uint32_t par = 0;
        auto gemm = [&](uint32_t d_tmem, uint64_t* done_bar) {     // MMAs and their commit by the SAME elected thread
            if (elect_one_sync()) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint64_t adesc = make_smem_desc<128>(a_base + c * Cfg::A_TILE);
                    const uint64_t bdesc = make_smem_desc<128>(b_base + c * Cfg::B_TILE);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) umma_f16(d_tmem, adesc + 2 * kk, bdesc + 2 * kk, idesc, (c | kk) ? 1u : 0u);
                }
                umma_commit(done_bar);
            }
            __syncwarp();
        };
        for (int u = blockIdx.x; u < total_units; u += gridDim.x, par ^= 1) {
            mbar_wait(qk_full, par);
            tc_fence_after();
            gemm(tmem_base, s_full);                                   // S = Q K^T
            mbar_wait(v_full, par);
            mbar_wait(p_ready, par);
            tc_fence_after();
            gemm(tmem_base + 256, o_full);                             // O = P V   (P unnormalised)
            mbar_wait(w_full, par);
            mbar_wait(o_ready, par);
            tc_fence_after();
            gemm(tmem_base + 256, y_full);                             // Y = O Wo^T  (O has been drained to shared memory)
        }
    } else {
        // ===================== softmax / epilogue warps 2..9: two sets, each half of the keys / output columns =====================
        const int set = (warp - 2) >> 2;
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;                             // query row of this thread inside the tile
        const int cbase = set * 128;                                   // first key / column of this set
        const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        const uint32_t r1_addr = smem_u32(r1);
        float* xmax = xch + set * 128;                                 // [set][row]
        float* xsum = xch + 256 + set * 128;
        const float* omax = xch + (set ^ 1) * 128;
        const float* osum = xch + 256 + (set ^ 1) * 128;
        uint32_t par = 0;
        for (int u = blockIdx.x; u < total_units; u += gridDim.x, par ^= 1) {
            const int img = u / tiles_per_img, q0 = (u - img * tiles_per_img) * 128;
            // ---- row softmax of S (fp32 in TMEM): max, then p = exp(s - max) rounded to fp16, row sum of the ROUNDED values
            mbar_wait_warp(s_full, par, lane);
            tc_fence_after();
            float mx = -INFINITY;
#pragma unroll 1
            for (int c0 = cbase; c0 < cbase + 128; c0 += 32) {
                uint32_t r[2][16];
                tmem_ld_x16(t_row + c0, r[0]);
                tmem_ld_x16(t_row + c0 + 16, r[1]);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) mx = fmaxf(mx, fmaxf(__uint_as_float(r[0][j]), __uint_as_float(r[1][j])));
            }
            xmax[m] = mx;
            asm volatile("bar.sync 1, 256;" ::: "memory");             // both sets: row maxima exchanged
            mx = fmaxf(mx, omax[m]);
            const float mxl = mx * 1.4426950408889634f;                // exp(s - mx) = exp2(s * log2e - mx * log2e)
            float sum = 0.f;
#pragma unroll 1
            for (int c0 = cbase; c0 < cbase + 128; c0 += 32) {
                uint32_t r[2][16];
                tmem_ld_x16(t_row + c0, r[0]);
                tmem_ld_x16(t_row + c0 + 16, r[1]);
                tmem_ld_wait();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t pk[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float e0 = exp2f(fmaf(__uint_as_float(r[h][2 * j]), 1.4426950408889634f, -mxl));
                        const float e1 = exp2f(fmaf(__uint_as_float(r[h][2 * j + 1]), 1.4426950408889634f, -mxl));
                        pk[j] = pack2(e0, e1);
                        const float2 pr = unpack2(pk[j]);
                        sum += pr.x + pr.y;
                    }
                    // keys [c0 + 16 h, + 16) -> key chunk (c0 + 16 h) / 64, 16-byte units 2 ((c0 + 16 h) % 64) / 16 and + 1
                    const int key0 = c0 + 16 * h;
                    const uint32_t tile = r1_addr + (key0 >> 6) * Cfg::A_TILE;
                    const int u0 = (key0 & 63) >> 3;
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(tile + sw128_off(m, u0)), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(tile + sw128_off(m, u0 + 1)), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
                }
            }
            xsum[m] = sum;
            fence_proxy_async_smem();                                  // generic-proxy writes of P -> visible to the tensor core
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_ready);
            asm volatile("bar.sync 1, 256;" ::: "memory");             // row sums exchanged (also orders the xmax reuse of the next unit)
            const float inv_sum = 1.f / (sum + osum[m]);
            // ---- O = P V (unnormalised) -> scale rows by 1 / sum -> fp16 A operand of the projection (R1; P is consumed)
            mbar_wait_warp(o_full, par, lane);
            tc_fence_after();
#pragma unroll 1
            for (int c0 = cbase; c0 < cbase + 128; c0 += 32) {
                uint32_t r[2][16];
                tmem_ld_x16(t_row + 256 + c0, r[0]);
                tmem_ld_x16(t_row + 256 + c0 + 16, r[1]);
                tmem_ld_wait();
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t pk[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        pk[j] = pack2(__uint_as_float(r[h][2 * j]) * inv_sum, __uint_as_float(r[h][2 * j + 1]) * inv_sum);
                    }
                    const int ch0 = c0 + 16 * h;
                    const uint32_t tile = r1_addr + (ch0 >> 6) * Cfg::A_TILE;
                    const int u0 = (ch0 & 63) >> 3;
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(tile + sw128_off(m, u0)), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]) : "memory");
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(tile + sw128_off(m, u0 + 1)), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();                                         // O has been read: the projection may overwrite its TMEM columns
            __syncwarp();
            if (lane == 0) mbar_arrive(o_ready);
            // ---- y = x + O Wo^T + bias: bias / residual / GroupNorm statistics / fp16 NHWC store (shared epilogue)
            mbar_wait_warp(y_full, par, lane);
            tc_fence_after();
            const long long pix = q0 + m;
#pragma unroll 1
            for (int c0 = cbase; c0 < cbase + 128; c0 += 32) epilogue_chunk32(p.epi, t_row + 256, img, pix, true, c0, lane);
            tc_fence_before();                                         // ordered before this warp's next p_ready arrival, which gates the next O
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

}  // namespace pnpf
