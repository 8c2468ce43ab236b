// Internal op layer: prepared (tensor maps encoded once) tensor-core ops and launchers of the SIMT kernels.
#pragma once
#include "pnpf_host.h"
#include "pnpf_rowconv.cuh"
#include "pnpf_patchconv.cuh"
#include "pnpf_attn.cuh"

namespace pnpf {


// ------------------------------------------------------------------ tensor-core ops
struct TcOp {               // a prepared conv_gemm launch
    CUtensorMap tmA, tmAb, tmA2, tmA2b, tmB;   // *b: second source of a channel concat (row conv only)
    CUtensorMap tmBh;                           // weight map with a box of BN/2 rows (CTA-pair launches of conv_gemm_kernel)
    GemmParams p;
    RowConvParams rp;       // kind == 1: row-streaming conv (pnpf_rowconv.cuh)
    PatchConvParams pp;     // kind == 2: patch-streaming conv (pnpf_patchconv.cuh)
    int patch_nb_pair = 0;  // weight-ring depth of the CTA-pair launch (half tiles)
    int patch_subpix = 0;   // patch conv computes one phase (1) / the two column phases of a row parity (2) of a sub-pixel convolution
    int kind = 0;           // 0: conv_gemm_kernel, 1: rowconv_kernel, 2: patchconv_kernel
    int BK = 0, BN = 0;
    int n_epi = 8;          // row conv: epilogue warps (RowCfg::NEW), the other worker warps run the GroupNorm transform
    double flops = 0;       // algorithmic 2*M*N*K (for reporting)
};

struct ConvDesc {
    // input: fp16 NHWC [B][Hin][Win][x_pitch], channels [c_base, c_base+Cin) are read
    const act16* x = nullptr;
    int B = 0, Hin = 0, Win = 0, Cin = 0;
    long long x_pitch = 0;
    int c_base = 0;
    int x_cvalid = 0;               // > 0: the buffer holds only this many channels (pitch >= x_cvalid); channels [x_cvalid, Cin) of the
                                    // K chunk are zero-filled by the TMA unit (out-of-bounds fill) — the 3-channel network input
    // optional second source for a fused 1x1 (extra K) at OUTPUT resolution: fp16 NHWC [B][Hout][Wout][x2_pitch]
    const act16* x2 = nullptr;
    int C2 = 0;
    long long x2_pitch = 0;
    int x2_identity = 0;            // the fused 1x1 is the identity (residual add done by the tensor core): not counted as FLOPs
    // weights: packed fp16 [N_pad][Ktot], Ktot = ksize*ksize*Cin + C2 (pack_conv_weight)
    const act16* w = nullptr;
    int N_pad = 0;
    int ksize = 3, stride = 1;
    int Hout = 0, Wout = 0;
    // epilogue
    void* out = nullptr;
    int out_mode = 0;
    long long out_img_stride = 0, out_row_stride = 0, out_col_stride = 1;
    int n_valid = 0;
    const float* bias = nullptr;
    const float* bias_img = nullptr;
    long long bias_img_stride = 0;
    const act16* residual = nullptr;
    long long res_img_stride = 0, res_row_stride = 0;
    double* stats_out = nullptr;    // optional [B][n_valid][2] GroupNorm statistics of the output (must be zeroed)
    int allow_rowconv = 1;          // use the row-streaming kernel when the shape qualifies
    // ---- row-streaming kernel only (rowconv_eligible(d) must hold, else prepare_conv fails loudly) ----
    // channel concat [x | xb] as main input, [x2 | x2b] as fused 1x1 input: Cin / C2 are the TOTAL channel counts
    const act16* xb = nullptr;
    int Cb = 0;
    long long xb_pitch = 0;
    const act16* x2b = nullptr;
    int C2b = 0;
    long long x2b_pitch = 0;
    // fused GroupNorm(+SiLU) of the main input with statistics from the producers' epilogues
    const float* gn_gamma = nullptr;
    const float* gn_beta = nullptr;
    const double* gn_stats_a = nullptr;
    const double* gn_stats_b = nullptr;
    int gn_groups = 32, gn_silu = 1;
    float gn_eps = 1e-6f;
    // ---- patch-streaming kernel only: one phase (sp_a, sp_b) of "nearest x2 upsampling + 3x3 conv" on the LOW-resolution input.
    // Hin/Win/Hout/Wout describe the low-resolution grid, the out_* strides the high-resolution tensor [B][2H][2W][C];
    // w = folded 2x2 weights of this phase, packed [N_pad][4*Cin] in (i, j, cin) order (fold_subpixel_weights + pack).
    // subpix = 2: both column phases of row parity sp_a in one launch: N_pad = 2 * C_out accumulator columns (b-major), n_valid =
    // C_out channels, bias [C_out], w = pack_subpixel_pair_weights (2 x 3 taps, [2*C_out][6*Cin]).
    int subpix = 0, sp_a = 0, sp_b = 0;
};
int prepare_conv(TcOp& op, const ConvDesc& d);
bool rowconv_eligible(const ConvDesc& d);     // would prepare_conv pick the row-streaming kernel?
bool patchconv_eligible(const ConvDesc& d);   // ... the patch-streaming kernel (3x3 stride 1, W <= 128, C_out 64 / 128 / 256)?
int rowconv_max_smem();
void describe_conv_impl(const ConvDesc& d, char* buf, size_t n);   // which kernel prepare_conv would pick (PNPF_PLAN_DUMP)

struct GemmDesc {           // out[b][m][n] = sum_k A[b|0][m][k] * Bm[b|0][n][k]  (+bias[n]) (+residual)
    const act16* A = nullptr;
    long long lda = 0, a_bstride = 0;
    int a_batched = 1;
    const act16* Bm = nullptr;
    long long ldb = 0, b_bstride = 0;
    int b_batched = 1;
    int batch = 1, M = 0, N = 0, K = 0;
    void* out = nullptr;
    int out_mode = 0;
    long long out_img_stride = 0, out_row_stride = 0;
    const float* bias = nullptr;
    const act16* residual = nullptr;
    long long res_img_stride = 0, res_row_stride = 0;
};
int prepare_gemm(TcOp& op, const GemmDesc& d);
int launch_tc(const TcOp& op, cudaStream_t s);

// ------------------------------------------------------------------ fused attention core (pnpf_attn.cuh)
struct AttnOp {
    CUtensorMap tmQ, tmK, tmV, tmW;
    AttnParams p;
    double flops = 0;
};
struct AttnDesc {
    const act16* qk = nullptr;       // [B][L][2C]: q (scaled) | k
    const act16* vT = nullptr;       // [B][C][L]
    const act16* w = nullptr;        // packed projection weights [N_pad = C][C]
    const float* bias = nullptr;    // [C]
    const act16* residual = nullptr; // [B][L][C]
    act16* out = nullptr;            // [B][L][C]
    double* stats_out = nullptr;    // optional [B][C][2]
    int B = 0, L = 0, C = 0;
};
bool attn_core_eligible(int L, int C);
int prepare_attn(AttnOp& op, const AttnDesc& d);
int launch_attn(const AttnOp& op, int n_img, cudaStream_t s);

// host-side weight repack: OIHW fp32 (reference layout) -> [N_pad][k*k*Cin_pad + C2] fp16, K ordered (kh, kw, cin),
// optionally followed by the 1x1 shortcut weights w2 [O][C2]; rows >= O and channels >= Cin are zero.
void pack_conv_weight(act16* dst, const float* w, int O, int Cin, int ks, int N_pad, int Cin_pad, const float* w2, int C2,
                      float scale);
// 3x3 weights [O][Cin][3][3] -> the 2x2 weights [O][Cin][2][2] of phase (a, b) of the equivalent sub-pixel convolution:
// out[o][c][i][j] = sum_{kh in R(a,i)} sum_{kw in R(b,j)} w[o][c][kh][kw],  R(0,0)={0}, R(0,1)={1,2}, R(1,0)={0,1}, R(1,1)={2}
void fold_subpixel_weights(const float* w, int O, int Cin, int a, int b, float* out);
// 3x3 weights [O][Cin][3][3] -> packed [2*O][6*Cin] operand of the SUBPIX = 2 launch for output-row parity a: row b*O + o, K index
// (i*3 + c)*Cin + ch holds W_ab[o][ch][i][c - b] (fold_subpixel_weights) when c - b is 0 or 1, else 0.
void pack_subpixel_pair_weights(act16* dst, const float* w, int O, int Cin, int a);

}  // namespace pnpf
