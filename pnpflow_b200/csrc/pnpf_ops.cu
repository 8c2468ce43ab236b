// Preparation (tensor maps + GemmParams) of the tensor-core ops.
#include "pnpf_ops.h"

#include <cstring>

namespace pnpf {

static int pick_bn(int N_pad, int& BN, int& n_tiles) {
    if (N_pad <= 256) {
        PNPF_REQUIRE(N_pad == 16 || N_pad == 32 || N_pad == 64 || N_pad == 128 || N_pad == 256,
                     "padded output width %d must be 16/32/64/128/256 or a multiple of 256", N_pad);
        BN = N_pad;
        n_tiles = 1;
    } else {
        PNPF_REQUIRE(N_pad % 256 == 0, "padded output width %d must be a multiple of 256", N_pad);
        BN = 256;
        n_tiles = N_pad / 256;
    }
    return 0;
}

int prepare_conv(TcOp& op, const ConvDesc& d) {
    PNPF_REQUIRE(d.ksize == 1 || d.ksize == 3, "conv kernel size %d unsupported (1 or 3)", d.ksize);
    PNPF_REQUIRE(d.stride == 1 || d.stride == 2, "conv stride %d unsupported (1 or 2)", d.stride);
    PNPF_REQUIRE(d.Cin % 32 == 0 && d.C2 % 32 == 0, "conv channels (%d,%d) must be multiples of 32", d.Cin, d.C2);
    const int BK = (d.Cin % 64 == 0 && d.C2 % 64 == 0) ? 64 : 32;
    int BN, n_tiles;
    if (int e = pick_bn(d.N_pad, BN, n_tiles)) return e;
    GemmParams& p = op.p;
    memset(&p, 0, sizeof(p));
    p.H = d.Hout;
    p.W = d.Wout;
    pick_tile(d.Wout, p.TH, p.TW);
    p.tiles_h = (d.Hout + p.TH - 1) / p.TH;
    p.tiles_w = (d.Wout + p.TW - 1) / p.TW;
    p.n_img = d.B;
    p.n_tiles_n = n_tiles;
    p.a_batched = 1;
    p.b_batched = 0;
    p.in_stride = d.stride;
    p.ntaps = d.ksize * d.ksize;
    for (int kh = 0; kh < d.ksize; ++kh)
        for (int kw = 0; kw < d.ksize; ++kw) {
            p.dh[kh * d.ksize + kw] = kh - d.ksize / 2;
            p.dw[kh * d.ksize + kw] = kw - d.ksize / 2;
        }
    p.kchunks = d.Cin / BK;
    p.c_base = d.c_base;
    p.kchunks2 = d.x2 ? d.C2 / BK : 0;
    p.out = d.out;
    p.out_mode = d.out_mode;
    p.out_img_stride = d.out_img_stride;
    p.out_row_stride = d.out_row_stride;
    p.out_col_stride = d.out_col_stride;
    p.n_valid = d.n_valid;
    p.bias = d.bias;
    p.bias_img = d.bias_img;
    p.bias_img_stride = d.bias_img_stride;
    p.residual = d.residual;
    p.res_img_stride = d.res_img_stride;
    p.res_row_stride = d.res_row_stride;
    op.BK = BK;
    op.BN = BN;
    const long long Ktot = (long long)p.ntaps * d.Cin + (d.x2 ? d.C2 : 0);
    if (int e = make_act_tmap(&op.tmA, d.x, d.c_base + d.Cin, d.x_pitch, d.Win, d.Hin, d.B, BK, p.TW, p.TH, d.stride)) return e;
    if (d.x2) {
        if (int e = make_act_tmap(&op.tmA2, d.x2, d.C2, d.x2_pitch, d.Wout, d.Hout, d.B, BK, p.TW, p.TH, 1)) return e;
    } else {
        op.tmA2 = op.tmA;
    }
    if (int e = make_b_tmap(&op.tmB, d.w, Ktot, Ktot, d.N_pad, 1, 0, BK, BN)) return e;
    op.flops = 2.0 * d.B * d.Hout * d.Wout * (double)d.n_valid * (double)Ktot;
    return 0;
}

int prepare_gemm(TcOp& op, const GemmDesc& d) {
    PNPF_REQUIRE(d.K % 64 == 0, "gemm K=%d must be a multiple of 64", d.K);
    PNPF_REQUIRE(d.M % 8 == 0, "gemm M=%d must be a multiple of 8", d.M);
    const int BK = 64;
    int BN, n_tiles;
    if (int e = pick_bn(d.N, BN, n_tiles)) return e;
    GemmParams& p = op.p;
    memset(&p, 0, sizeof(p));
    p.H = 1;
    p.W = d.M;
    p.TH = 1;
    p.TW = 128;
    p.tiles_h = 1;
    p.tiles_w = (d.M + 127) / 128;
    p.n_img = d.batch;
    p.n_tiles_n = n_tiles;
    p.a_batched = d.a_batched;
    p.b_batched = d.b_batched;
    p.in_stride = 1;
    p.ntaps = 1;
    p.kchunks = d.K / BK;
    p.out = d.out;
    p.out_mode = d.out_mode;
    p.out_img_stride = d.out_img_stride;
    p.out_row_stride = d.out_row_stride;
    p.out_col_stride = 1;
    p.n_valid = d.N;
    p.bias = d.bias;
    p.residual = d.residual;
    p.res_img_stride = d.res_img_stride;
    p.res_row_stride = d.res_row_stride;
    op.BK = BK;
    op.BN = BN;
    // A viewed as (K, M, 1, batch): W extent = M rows, "image" pitch = a_bstride
    {
        const int nb = d.a_batched ? d.batch : 1;
        PNPF_REQUIRE(nb == 1 || d.a_bstride == (long long)d.M * d.lda, "gemm A batch stride must be M*lda");
        if (int e = make_act_tmap(&op.tmA, d.A, d.K, d.lda, d.M, 1, nb, BK, 128, 1, 1)) return e;
    }
    op.tmA2 = op.tmA;
    if (int e = make_b_tmap(&op.tmB, d.Bm, d.K, d.ldb, d.N, d.b_batched ? d.batch : 1, d.b_bstride, BK, BN)) return e;
    op.flops = 2.0 * d.batch * (double)d.M * d.N * d.K;
    return 0;
}

int launch_tc(const TcOp& op, cudaStream_t s) { return launch_conv_gemm(op.BK, op.BN, op.tmA, op.tmA2, op.tmB, op.p, s); }

static inline bf16 f2bf(float f) { return __float2bfloat16_rn(f); }

void pack_conv_weight(bf16* dst, const float* w, int O, int Cin, int ks, int N_pad, int Cin_pad, const float* w2, int C2,
                      float scale) {
    const long long Ktot = (long long)ks * ks * Cin_pad + C2;
    for (long long i = 0; i < (long long)N_pad * Ktot; ++i) dst[i] = f2bf(0.f);
    for (int o = 0; o < O; ++o) {
        bf16* row = dst + (long long)o * Ktot;
        for (int kh = 0; kh < ks; ++kh)
            for (int kw = 0; kw < ks; ++kw)
                for (int c = 0; c < Cin; ++c)
                    row[(long long)(kh * ks + kw) * Cin_pad + c] = f2bf(scale * w[(((long long)o * Cin + c) * ks + kh) * ks + kw]);
        if (w2)
            for (int c = 0; c < C2; ++c) row[(long long)ks * ks * Cin_pad + c] = f2bf(w2[(long long)o * C2 + c]);
    }
}

}  // namespace pnpf
