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

static void fill_epi(EpiParams& e, const ConvDesc& d) {
    e.out = d.out;
    e.out_mode = d.out_mode;
    e.out_img_stride = d.out_img_stride;
    e.out_row_stride = d.out_row_stride;
    e.out_col_stride = d.out_col_stride;
    e.n_valid = d.n_valid;
    e.bias = d.bias;
    e.bias_img = d.bias_img;
    e.bias_img_stride = d.bias_img_stride;
    e.residual = d.residual;
    e.res_img_stride = d.res_img_stride;
    e.res_row_stride = d.res_row_stride;
    e.stats = d.stats_out;
}

constexpr int ROWCONV_SMEM_BUDGET = 200 * 1024;
int rowconv_max_smem() { return ROWCONV_SMEM_BUDGET + 1024 + 512; }

// Row-streaming kernel eligibility + preparation; returns -1 when the shape does not qualify (caller falls back).
static int try_prepare_rowconv(TcOp& op, const ConvDesc& d) {
    if (!d.allow_rowconv || d.ksize != 3 || d.stride != 1 || d.Wout % 128 != 0 || d.Wout != d.Win || d.Hout != d.Hin) return -1;
    if (!(d.N_pad == 16 || d.N_pad == 32 || d.N_pad == 64) || d.c_base != 0) return -1;
    if (d.Cin % 32 != 0 || d.C2 % 32 != 0) return -1;
    const int BK = (d.Cin % 64 == 0 && d.C2 % 64 == 0) ? 64 : 32;
    const int BN = d.N_pad;
    const int rowb = BK * 2;
    const int halo_tile = (136 * rowb + 1023) / 1024 * 1024, x2_tile = 128 * rowb;
    const int w_tile = (BN * BK * 2 + 1023) / 1024 * 1024;
    const int kch = d.Cin / BK, kch2 = d.x2 ? d.C2 / BK : 0;
    if (kch < 1 || kch > 3) return -1;
    if ((BN * rowb) % 1024 != 0) return -1;       // stacked vertical-tap tiles must keep the swizzle phase
    const int w_bytes = 3 * kch * (3 * BN * rowb) + kch2 * w_tile;
    const int slot_bytes = kch * halo_tile + kch2 * x2_tile;
    int nslot = (ROWCONV_SMEM_BUDGET - w_bytes) / slot_bytes;
    if (nslot > 8) nslot = 8;
    if (nslot < 4) return -1;
    RowConvParams& r = op.rp;
    memset(&r, 0, sizeof(r));
    r.H = d.Hout; r.W = d.Wout; r.n_img = d.B;
    r.strips = d.Wout / 128;
    int seg = 32;
    while (seg > 8 && (long long)d.B * r.strips * ((d.Hout + seg - 1) / seg) < 3LL * num_sms()) seg >>= 1;
    r.seg_rows = seg;
    r.segs = (d.Hout + seg - 1) / seg;
    r.kchunks = kch; r.kchunks2 = kch2; r.nslot = nslot; r.slot_bytes = slot_bytes;
    fill_epi(r.epi, d);
    op.kind = 1; op.BK = BK; op.BN = BN;
    const long long Ktot = 9LL * d.Cin + (d.x2 ? d.C2 : 0);
    if (int e = make_act_tmap(&op.tmA, d.x, d.Cin, d.x_pitch, d.Win, d.Hin, d.B, BK, 130, 1, 1)) return e;
    if (d.x2) {
        if (int e = make_act_tmap(&op.tmA2, d.x2, d.C2, d.x2_pitch, d.Wout, d.Hout, d.B, BK, 128, 1, 1)) return e;
    } else {
        op.tmA2 = op.tmA;
    }
    if (int e = make_b_tmap(&op.tmB, d.w, Ktot, Ktot, d.N_pad, 1, 0, BK, BN)) return e;
    op.flops = 2.0 * d.B * d.Hout * d.Wout * (double)d.n_valid * (double)Ktot;
    return 0;
}

int prepare_conv(TcOp& op, const ConvDesc& d) {
    PNPF_REQUIRE(d.ksize == 1 || d.ksize == 3, "conv kernel size %d unsupported (1 or 3)", d.ksize);
    {
        const int rc = try_prepare_rowconv(op, d);
        if (rc >= 0) return rc;
    }
    op.kind = 0;
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
    fill_epi(p.epi, d);
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
    op.kind = 0;
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
    p.epi.out = d.out;
    p.epi.out_mode = d.out_mode;
    p.epi.out_img_stride = d.out_img_stride;
    p.epi.out_row_stride = d.out_row_stride;
    p.epi.out_col_stride = 1;
    p.epi.n_valid = d.N;
    p.epi.bias = d.bias;
    p.epi.residual = d.residual;
    p.epi.res_img_stride = d.res_img_stride;
    p.epi.res_row_stride = d.res_row_stride;
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

template <int BK, int BN, int KCH>
static int launch_row_t(const TcOp& op, cudaStream_t stream) {
    using Cfg = RowCfg<BK, BN>;
    const RowConvParams& r = op.rp;
    const int smem = 3 * r.kchunks * Cfg::W_STACK + r.kchunks2 * Cfg::W_TILE + r.nslot * r.slot_bytes + Cfg::BAR_BYTES + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        PNPF_CHECK_CUDA(cudaFuncSetAttribute(rowconv_kernel<BK, BN, KCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, rowconv_max_smem()));
        attr_set = true;
    }
    PNPF_REQUIRE(smem <= rowconv_max_smem(), "row conv shared memory %d exceeds the budget", smem);
    const long long items = (long long)r.n_img * r.segs * r.strips;
    const int grid = (int)(items < num_sms() ? items : num_sms());
    if (grid < 1) return 0;
    rowconv_kernel<BK, BN, KCH><<<grid, Cfg::THREADS, smem, stream>>>(op.tmA, op.tmA2, op.tmB, r);
    PNPF_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_tc(const TcOp& op, cudaStream_t s) {
    if (op.kind == 1) {
#define PNPF_RCASE(bk, bn)                                                                \
    if (op.BK == bk && op.BN == bn) {                                                     \
        if (op.rp.kchunks == 1) return launch_row_t<bk, bn, 1>(op, s);                    \
        if (op.rp.kchunks == 2) return launch_row_t<bk, bn, 2>(op, s);                    \
        if (op.rp.kchunks == 3) return launch_row_t<bk, bn, 3>(op, s);                    \
    }
        PNPF_RCASE(32, 16) PNPF_RCASE(32, 32) PNPF_RCASE(32, 64) PNPF_RCASE(64, 16) PNPF_RCASE(64, 32) PNPF_RCASE(64, 64)
#undef PNPF_RCASE
        set_error("no rowconv instantiation for BK=%d BN=%d", op.BK, op.BN);
        return 2;
    }
    return launch_conv_gemm(op.BK, op.BN, op.tmA, op.tmA2, op.tmB, op.p, s);
}

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
