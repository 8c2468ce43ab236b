// Preparation (tensor maps + GemmParams) of the tensor-core ops.
#include "pnpf_ops.h"

#include <cstdlib>
#include <cstring>
#include <vector>

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

// Shared memory of the row kernel: 227 KB per CTA on sm_100 minus barriers/tables (RowCfg::BAR_BYTES) and alignment slack.
constexpr int ROWCONV_SMEM_MAX = 227 * 1024;
constexpr int ROWCONV_SMEM_BUDGET = ROWCONV_SMEM_MAX - 2048 - 1024;
int rowconv_max_smem() { return ROWCONV_SMEM_MAX; }

// Row-streaming kernel: shape analysis shared by rowconv_eligible() and the preparation.
struct RowShape { int BK, BN, nsplit, kch, kch2, kch_a, kch2_a, w_bytes, slot_bytes, nslot, stage_bytes, n_epi; };
static bool rowconv_shape(const ConvDesc& d, RowShape& r) {
    if (!d.allow_rowconv || d.ksize != 3 || d.stride != 1 || d.Wout % 128 != 0 || d.Wout != d.Win || d.Hout != d.Hin) return false;
    if (d.subpix) return false;                       // sub-pixel phases exist in the patch kernel only
    if (!(d.N_pad == 16 || d.N_pad == 32 || d.N_pad == 64) || d.c_base != 0) return false;
    const int Ca = d.Cin - d.Cb, C2a = d.C2 - d.C2b;
    if (d.Cin % 32 || d.C2 % 32 || Ca % 32 || d.Cb % 32 || C2a % 32 || d.C2b % 32 || Ca <= 0) return false;
    r.BK = (Ca % 64 == 0 && d.Cb % 64 == 0 && C2a % 64 == 0 && d.C2b % 64 == 0) ? 64 : 32;
    const int rowb = r.BK * 2;
    const int halo_tile = (136 * rowb + 1023) / 1024 * 1024, x2_tile = 128 * rowb;
    r.kch = d.Cin / r.BK;
    r.kch2 = d.C2 > 0 ? d.C2 / r.BK : 0;        // by channel count, not pointer: the size-query plan has no pointers
    r.kch_a = Ca / r.BK;
    r.kch2_a = d.C2 > 0 ? C2a / r.BK : 0;
    if (r.kch < 1 || r.kch > 3) return false;
    if (d.gn_gamma && d.Cin > 128) return false;     // scale/shift table
    r.slot_bytes = r.kch * halo_tile + r.kch2 * x2_tile;
    // All nine taps of the weights stay resident next to a ring of input-row slots.  When the full C_out does not leave
    // four slots, the output channels are split over two CTAs (nsplit = 2: each keeps half of the weights and both stream
    // the same rows; the second read of a row hits L2) as long as three slots remain.
    static const int max_split = getenv("PNPF_NO_NSPLIT") ? 1 : 2;     // A/B switch (tools/ab_env.py)
    for (int nsplit = 1; nsplit <= max_split; ++nsplit) {
        const int BN = d.N_pad / nsplit;
        if (BN < 16 || (BN * rowb) % 1024 != 0) continue;        // stacked vertical-tap tiles must keep the swizzle phase
        const int w_tile = (BN * r.BK * 2 + 1023) / 1024 * 1024;
        const int w_bytes = 3 * r.kch * (3 * BN * rowb) + r.kch2 * w_tile;
        // fp16 NHWC outputs with full 32-channel blocks are transposed through per-warp staging tiles (RowCfg::STAGE_BYTES)
        const bool staged = d.out_mode == 0 && BN >= 32 && d.n_valid == d.N_pad && d.out_col_stride == 1;
        // Warp-role configuration (RowCfg): WIDE (n_epi = 12: 8 epilogue + 8 transform warps) for the fused-GroupNorm layers.
        // It accumulates the output statistics in the staged store, so it needs it (the only unstaged outputs without
        // statistics are fp32 NCHW).
        static const bool no_wide = getenv("PNPF_NO_WIDE") != nullptr;    // A/B switch (tools/ab_env.py)
        const bool wide = !no_wide && staged && d.gn_gamma != nullptr;
        const int n_epi = wide ? 12 : 8;
        const int stage = staged ? 8 * 32 * 64 : 0;
        int nslot = (ROWCONV_SMEM_BUDGET - w_bytes - stage) / r.slot_bytes;
        if (nslot > 8) nslot = 8;
        // three row slots are enough for either layout: the unsplit one keeps the level-1 conv2 + shortcut layers on N = 192
        // MMAs with one read of every row instead of two half-width CTAs reading every row twice (r02 A/B: equal or faster)
        if (nslot >= 3) {
            r.BN = BN; r.nsplit = nsplit; r.w_bytes = w_bytes; r.nslot = nslot; r.stage_bytes = stage; r.n_epi = n_epi;
            return true;
        }
    }
    return false;
}
bool rowconv_eligible(const ConvDesc& d) {
    RowShape r;
    return rowconv_shape(d, r);
}
// Patch-streaming kernel: shape analysis.
struct PatchShape { int P, NR, patch_bytes, na, nb, nb_pair, tiles_per_img, stage_bytes; };
constexpr int PATCH_SMEM_MAX = 227 * 1024;
static bool patchconv_shape(const ConvDesc& d, PatchShape& r) {
    static const bool off = getenv("PNPF_NO_PATCH") != nullptr;          // A/B switch (tools/ab_env.py)
    if (off || !d.allow_rowconv || d.ksize != 3 || d.stride != 1 || d.Wout != d.Win || d.Hout != d.Hin || d.Wout > 128) return false;
    if (!(d.N_pad == 64 || d.N_pad == 128 || d.N_pad == 256) || d.c_base != 0 || d.xb || d.x2b || d.gn_gamma) return false;
    // channel counts that are odd multiples of 32 ride in 64-channel chunks whose upper half the TMA unit zero-fills (no sub-pixel form)
    if (d.Cin % 32 || d.C2 % 32 || d.Cin < 32 || d.x_cvalid) return false;
    if (d.subpix && d.Cin % 64) return false;
    if (d.subpix && (d.x2 || d.C2 || d.residual || d.out_mode != 0)) return false;
    if (d.subpix == 2 && (d.N_pad != 2 * d.n_valid || d.N_pad > 256)) return false;       // two column phases: 2 * C_out accumulator columns
    r.P = d.Wout + 2;
    r.NR = (r.P - 1 + 127) / r.P + 1 + 2;          // rows a tile of 128 positions can touch, plus the two halo rows
    r.patch_bytes = (r.NR * r.P * 128 + 1023) / 1024 * 1024;
    r.tiles_per_img = (d.Hout * r.P + 127) / 128;
    // rings: weight tiles of the non-pair variant (the larger) must fit: >= 2 patches + >= 4 weight tiles
    const int b_bytes = d.N_pad * 128;
    // fp16 NHWC outputs of the narrower tiles leave through 8 x 2 KB staging tiles (coalesced stores + cheaper statistics: the
    // BN <= 128 layers are shared-memory / epilogue bound; BN = 256 is tensor-bound and keeps its ring depth)
    static const bool no_stage = getenv("PNPF_NO_PATCH_STAGE") != nullptr;       // A/B switch (tools/ab_env.py)
    r.stage_bytes = (!no_stage && d.N_pad <= 128 && !d.subpix && d.out_mode == 0 && d.n_valid == d.N_pad && d.out_col_stride <= 1 && r.P >= 10) ? 8 * 2048 : 0;
    const int budget = PATCH_SMEM_MAX - 1024 - 512 - r.stage_bytes;
    r.na = 3;
    while (r.na > 2 && budget - r.na * r.patch_bytes < 4 * b_bytes) --r.na;
    r.nb = (budget - r.na * r.patch_bytes) / b_bytes;
    if (r.nb > 12) r.nb = 12;
    r.nb_pair = (budget - r.na * r.patch_bytes) / (b_bytes / 2);      // CTA pairs stage half tiles: twice the depth
    if (r.nb_pair > 12) r.nb_pair = 12;
    return r.nb >= 4;
}
bool patchconv_eligible(const ConvDesc& d) {
    PatchShape r;
    return patchconv_shape(d, r);
}
void describe_conv_impl(const ConvDesc& d, char* buf, size_t n) {
    RowShape r;
    if (rowconv_shape(d, r))
        snprintf(buf, n, "rowconv<%d,%d,%d> nsplit=%d nslot=%d kch2=%d w=%dKB slot=%dKB stage=%dKB epi_warps=%d gn=%d", r.BK, r.BN, r.kch, r.nsplit,
                 r.nslot, r.kch2, r.w_bytes / 1024, r.slot_bytes / 1024, r.stage_bytes / 1024, r.n_epi, d.gn_gamma ? 1 : 0);
    else if (PatchShape ps; patchconv_shape(d, ps))
        snprintf(buf, n, "patchconv<%d> P=%d NR=%d patch=%dKB na=%d nb=%d tiles/img=%d%s", d.N_pad, ps.P, ps.NR, ps.patch_bytes / 1024, ps.na, ps.nb,
                 ps.tiles_per_img, d.subpix == 2 ? " subpix2" : (d.subpix ? " subpix" : ""));
    else
        snprintf(buf, n, "conv_gemm<%d,%d> k=%d s=%d", (d.Cin % 64 == 0 && d.C2 % 64 == 0) ? 64 : 32, d.N_pad > 256 ? 256 : d.N_pad, d.ksize, d.stride);
}

// returns -1 when the shape does not qualify (caller falls back to the per-tap kernel)
static int try_prepare_rowconv(TcOp& op, const ConvDesc& d) {
    RowShape sh;
    if (!rowconv_shape(d, sh)) return -1;
    const int BK = sh.BK, BN = sh.BN;
    RowConvParams& r = op.rp;
    memset(&r, 0, sizeof(r));
    r.H = d.Hout; r.W = d.Wout; r.n_img = d.B;
    r.strips = d.Wout / 128;
    r.nsplit = sh.nsplit;
    r.staged_store = sh.stage_bytes > 0;
    op.n_epi = sh.n_epi;
    static const bool no_mma2 = getenv("PNPF_NO_MMA2") != nullptr;       // A/B switch (tools/ab_env.py)
    r.mma2 = (!no_mma2 && d.Hout >= 2) ? 1 : 0;         // second MMA-issuing warp (RowCfg::NMMA)
    PNPF_REQUIRE(sh.n_epi == 8 || r.staged_store || !d.stats_out, "row conv (WIDE): output statistics need the staged store");
    PNPF_REQUIRE((long long)d.B * r.strips * d.Hout < (1LL << 31) / 256, "row conv: batch * rows too large for 32-bit row indices");
    r.kchunks = sh.kch; r.kchunks2 = sh.kch2; r.nslot = sh.nslot; r.slot_bytes = sh.slot_bytes;
    r.kch_a = sh.kch_a; r.kch2_a = sh.kch2_a;
    if (d.gn_gamma) {
        PNPF_REQUIRE(d.gn_beta && d.gn_stats_a && (d.Cb == 0 || d.gn_stats_b), "fused GroupNorm needs beta and statistics");
        PNPF_REQUIRE(d.Cin % d.gn_groups == 0, "GroupNorm: %d channels not divisible into %d groups", d.Cin, d.gn_groups);
        r.gn = 1; r.gn_silu = d.gn_silu; r.gn_gs = d.Cin / d.gn_groups; r.gn_Ca = d.Cin - d.Cb; r.gn_Cb = d.Cb; r.gn_eps = d.gn_eps;
        r.gn_gamma = d.gn_gamma; r.gn_beta = d.gn_beta; r.gn_st_a = d.gn_stats_a; r.gn_st_b = d.gn_stats_b;
    }
    fill_epi(r.epi, d);
    op.kind = 1; op.BK = BK; op.BN = BN;
    const long long Ktot = 9LL * d.Cin + (d.x2 ? d.C2 : 0);
    if (int e = make_act_tmap(&op.tmA, d.x, d.x_cvalid > 0 ? d.x_cvalid : d.Cin - d.Cb, d.x_pitch, d.Win, d.Hin, d.B, BK, 130, 1, 1)) return e;
    op.tmAb = op.tmA;
    if (d.Cb) { if (int e = make_act_tmap(&op.tmAb, d.xb, d.Cb, d.xb_pitch, d.Win, d.Hin, d.B, BK, 130, 1, 1)) return e; }
    op.tmA2 = op.tmA;
    op.tmA2b = op.tmA;
    if (d.x2) {
        if (int e = make_act_tmap(&op.tmA2, d.x2, d.C2 - d.C2b, d.x2_pitch, d.Wout, d.Hout, d.B, BK, 128, 1, 1)) return e;
        op.tmA2b = op.tmA2;
        if (d.C2b) { if (int e = make_act_tmap(&op.tmA2b, d.x2b, d.C2b, d.x2b_pitch, d.Wout, d.Hout, d.B, BK, 128, 1, 1)) return e; }
    }
    if (int e = make_b_tmap(&op.tmB, d.w, Ktot, Ktot, d.N_pad, 1, 0, BK, BN)) return e;
    if (r.staged_store)
        PNPF_REQUIRE(d.out_row_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(d.out) & 15) == 0, "row conv: staged store needs 16-byte aligned output rows");
    op.flops = 2.0 * d.B * d.Hout * d.Wout * (double)d.n_valid * (double)(Ktot - (d.x2_identity ? d.C2 : 0));
    return 0;
}

static int try_prepare_patchconv(TcOp& op, const ConvDesc& d) {
    PatchShape sh;
    if (!patchconv_shape(d, sh)) return -1;
    PatchConvParams& q = op.pp;
    memset(&q, 0, sizeof(q));
    q.H = d.Hout; q.W = d.Wout; q.P = sh.P; q.NR = sh.NR; q.n_img = d.B; q.tiles_per_img = sh.tiles_per_img;
    q.kchunks = (d.Cin + 63) / 64; q.kchunks2 = d.x2 ? (d.C2 + 63) / 64 : 0; q.cin = d.Cin;
    q.patch_bytes = sh.patch_bytes; q.na = sh.na; q.nb = sh.nb; q.stage_bytes = sh.stage_bytes;
    op.patch_nb_pair = sh.nb_pair;
    op.patch_subpix = d.subpix;
    q.sp_a = d.sp_a; q.sp_b = d.sp_b;
    PNPF_REQUIRE(!d.subpix || ((d.sp_a | d.sp_b) & ~1) == 0, "sub-pixel phase (%d,%d) must be 0/1", d.sp_a, d.sp_b);
    fill_epi(q.epi, d);
    op.kind = 2; op.BK = 64; op.BN = d.N_pad;
    const long long Ktot = d.subpix == 2 ? 6LL * d.Cin : (d.subpix ? 4LL * d.Cin : 9LL * d.Cin + (d.x2 ? d.C2 : 0));
    if (int e = make_act_tmap(&op.tmA, d.x, d.Cin, d.x_pitch, d.Win, d.Hin, d.B, 64, sh.P, sh.NR, 1)) return e;
    op.tmA2 = op.tmA;
    if (d.x2) { if (int e = make_act_tmap(&op.tmA2, d.x2, d.C2, d.x2_pitch, d.Wout, d.Hout, d.B, 64, sh.P, sh.NR, 1)) return e; }
    if (int e = make_b_tmap(&op.tmB, d.w, Ktot, Ktot, d.N_pad, 1, 0, 64, d.N_pad)) return e;
    if (int e = make_b_tmap(&op.tmBh, d.w, Ktot, Ktot, d.N_pad, 1, 0, 64, d.N_pad / 2)) return e;
    op.flops = 2.0 * d.B * d.Hout * d.Wout * (double)d.n_valid * (double)(d.subpix == 2 ? 8LL * d.Cin : Ktot);     // useful MACs (two phases x 4 taps)
    return 0;
}

int prepare_conv(TcOp& op, const ConvDesc& d) {
    PNPF_REQUIRE(d.ksize == 1 || d.ksize == 3, "conv kernel size %d unsupported (1 or 3)", d.ksize);
    {
        const int rc = try_prepare_rowconv(op, d);
        if (rc >= 0) return rc;
    }
    {
        const int rc = try_prepare_patchconv(op, d);
        if (rc >= 0) return rc;
    }
    PNPF_REQUIRE(!d.subpix, "sub-pixel convolution needs the patch-streaming kernel: check patchconv_eligible() first");
    PNPF_REQUIRE(!d.xb && !d.x2b && !d.gn_gamma, "two-source / fused-GroupNorm convolution needs the row-streaming kernel "
                 "(3x3 stride 1, W %% 128 == 0, C_out <= 64): check rowconv_eligible() first");
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
    if (int e = make_act_tmap(&op.tmA, d.x, d.x_cvalid > 0 ? d.x_cvalid : d.c_base + d.Cin, d.x_pitch, d.Win, d.Hin, d.B, BK, p.TW, p.TH, d.stride)) return e;
    if (d.x2) {
        if (int e = make_act_tmap(&op.tmA2, d.x2, d.C2, d.x2_pitch, d.Wout, d.Hout, d.B, BK, p.TW, p.TH, 1)) return e;
    } else {
        op.tmA2 = op.tmA;
    }
    if (int e = make_b_tmap(&op.tmB, d.w, Ktot, Ktot, d.N_pad, 1, 0, BK, BN)) return e;
    op.tmBh = op.tmB;                 // CTA-pair launches stage half of the weight tile per CTA (box of BN/2 rows)
    if (BK == 64 && BN >= 128) { if (int e = make_b_tmap(&op.tmBh, d.w, Ktot, Ktot, d.N_pad, 1, 0, BK, BN / 2)) return e; }
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
    op.tmBh = op.tmB;
    if (BN >= 128) { if (int e = make_b_tmap(&op.tmBh, d.Bm, d.K, d.ldb, d.N, d.b_batched ? d.batch : 1, d.b_bstride, BK, BN / 2)) return e; }
    op.flops = 2.0 * d.batch * (double)d.M * d.N * d.K;
    return 0;
}

template <int BK, int BN, int KCH, int NEW>
static int launch_row_t(const TcOp& op, cudaStream_t stream) {
    using Cfg = RowCfg<BK, BN, NEW>;
    const RowConvParams& r = op.rp;
    const int smem = 3 * r.kchunks * Cfg::W_STACK + r.kchunks2 * Cfg::W_TILE + r.nslot * r.slot_bytes + (r.staged_store ? Cfg::STAGE_BYTES : 0) +
                     Cfg::BAR_BYTES + 1024;
    static DeviceCache cache;
    int dummy = 0;
    if (!cache.lookup(&dummy)) {
        PNPF_CHECK_CUDA(cudaFuncSetAttribute(rowconv_kernel<BK, BN, KCH, NEW>, cudaFuncAttributeMaxDynamicSharedMemorySize, rowconv_max_smem()));
        cache.store(1);
    }
    PNPF_REQUIRE(smem <= rowconv_max_smem(), "row conv shared memory %d exceeds the budget", smem);
    const int ctas_per_sm = 1;      // (a two-CTA-per-SM variant was measured slower in round 1: profiles/r01_ab_experiments.md)
    // one CTA group (nsplit CTAs) per contiguous range of the flattened row space; at least 8 rows per range
    const long long rows = (long long)r.n_img * r.strips * r.H;
    long long groups = (long long)num_sms() * ctas_per_sm / r.nsplit;
    if (groups > (rows + 7) / 8) groups = (rows + 7) / 8;
    const int grid = (int)groups * r.nsplit;
    if (grid < 1) return 0;
    rowconv_kernel<BK, BN, KCH, NEW><<<grid, Cfg::THREADS, smem, stream>>>(op.tmA, op.tmAb, op.tmA2, op.tmA2b, op.tmB, r);
    PNPF_CHECK_CUDA(cudaGetLastError());
    return 0;
}

template <int BN, bool PAIR, int SUBPIX = 0, int TG = 1>
static int launch_patch_t(const TcOp& op, cudaStream_t stream) {
    using Cfg = PatchCfg<BN, PAIR>;
    PatchConvParams q = op.pp;
    if (PAIR) q.nb = op.patch_nb_pair;
    q.nb /= TG;                                       // ring depth in slots of TG weight tiles
    const int smem = q.na * q.patch_bytes + q.nb * TG * Cfg::B_BYTES + q.stage_bytes + 512 + 1024;
    static DeviceCache cache;
    int max_clusters = 0;
    if (!cache.lookup(&max_clusters)) {
        PNPF_CHECK_CUDA(cudaFuncSetAttribute(patchconv_kernel<BN, PAIR, SUBPIX, TG>, cudaFuncAttributeMaxDynamicSharedMemorySize, PATCH_SMEM_MAX));
        if (PAIR) {
            cudaLaunchConfig_t qc = {};
            qc.gridDim = dim3(num_sms() & ~1);
            qc.blockDim = dim3(Cfg::THREADS);
            qc.dynamicSmemBytes = PATCH_SMEM_MAX;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            qc.attrs = qa; qc.numAttrs = 1;
            PNPF_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, patchconv_kernel<BN, PAIR, SUBPIX, TG>, &qc));
            PNPF_REQUIRE(max_clusters >= 1, "no CTA pair of patchconv_kernel<%d> fits on this device", BN);
        }
        cache.store(max_clusters);
    }
    PNPF_REQUIRE(smem <= PATCH_SMEM_MAX, "patch conv shared memory %d exceeds the budget", smem);
    const long long units = (long long)(q.n_img / (PAIR ? 2 : 1)) * q.tiles_per_img;
    if (units < 1) return 0;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    if (PAIR) {
        const int clusters = (int)(units < max_clusters ? units : max_clusters);
        cfg.gridDim = dim3(2 * clusters);
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
    } else {
        cfg.gridDim = dim3((unsigned)(units < num_sms() ? units : num_sms()));
    }
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    PNPF_CHECK_CUDA(cudaLaunchKernelEx(&cfg, patchconv_kernel<BN, PAIR, SUBPIX, TG>, op.tmA, op.tmA2, PAIR ? op.tmBh : op.tmB, q));
    return 0;
}

int launch_tc(const TcOp& op, cudaStream_t s) {
    if (op.kind == 2) {
        static const bool no_pair = getenv("PNPF_NO_PAIR") != nullptr;
        const bool pair = !no_pair && op.pp.n_img % 2 == 0;
        if (op.patch_subpix == 2) {                   // two column phases per launch: 2 x 3 taps, tap groups of three
            const bool tg = (pair ? op.patch_nb_pair : op.pp.nb) >= 6;
            if (op.BN == 128) {
                if (tg) return pair ? launch_patch_t<128, true, 2, 3>(op, s) : launch_patch_t<128, false, 2, 3>(op, s);
                return pair ? launch_patch_t<128, true, 2>(op, s) : launch_patch_t<128, false, 2>(op, s);
            }
            if (op.BN == 256) {
                if (tg) return pair ? launch_patch_t<256, true, 2, 3>(op, s) : launch_patch_t<256, false, 2, 3>(op, s);
                return pair ? launch_patch_t<256, true, 2>(op, s) : launch_patch_t<256, false, 2>(op, s);
            }
            set_error("no two-phase sub-pixel instantiation for BN=%d", op.BN);
            return 2;
        }
        if (op.patch_subpix) {
            if (op.BN == 64) return pair ? launch_patch_t<64, true, 1>(op, s) : launch_patch_t<64, false, 1>(op, s);
            if (op.BN == 128) return pair ? launch_patch_t<128, true, 1>(op, s) : launch_patch_t<128, false, 1>(op, s);
            if (op.BN == 256) return pair ? launch_patch_t<256, true, 1>(op, s) : launch_patch_t<256, false, 1>(op, s);
        }
        // three taps (one kernel row) per weight-ring slot: the issuer waits and commits once per 12 MMAs instead of once per 4
        // (same arithmetic in the same order; r02 A/B: profiles/r02_optin_validation.log).  Needs >= 2 slots of 3 tiles.
        static const bool tg1 = getenv("PNPF_PATCH_TG1") != nullptr;        // A/B switch (tools/ab_env.py)
        if (!tg1 && (pair ? op.patch_nb_pair : op.pp.nb) >= 6) {
            if (op.BN == 64) return pair ? launch_patch_t<64, true, 0, 3>(op, s) : launch_patch_t<64, false, 0, 3>(op, s);
            if (op.BN == 128) return pair ? launch_patch_t<128, true, 0, 3>(op, s) : launch_patch_t<128, false, 0, 3>(op, s);
            if (op.BN == 256) return pair ? launch_patch_t<256, true, 0, 3>(op, s) : launch_patch_t<256, false, 0, 3>(op, s);
        }
        if (op.BN == 64) return pair ? launch_patch_t<64, true>(op, s) : launch_patch_t<64, false>(op, s);
        if (op.BN == 128) return pair ? launch_patch_t<128, true>(op, s) : launch_patch_t<128, false>(op, s);
        if (op.BN == 256) return pair ? launch_patch_t<256, true>(op, s) : launch_patch_t<256, false>(op, s);
        set_error("no patchconv instantiation for BN=%d", op.BN);
        return 2;
    }
    if (op.kind == 1) {
#define PNPF_RCASE(bk, bn, ne)                                                            \
    if (op.BK == bk && op.BN == bn && op.n_epi == ne) {                                   \
        if (op.rp.kchunks == 1) return launch_row_t<bk, bn, 1, ne>(op, s);                \
        if (op.rp.kchunks == 2) return launch_row_t<bk, bn, 2, ne>(op, s);                \
        if (op.rp.kchunks == 3) return launch_row_t<bk, bn, 3, ne>(op, s);                \
    }
        PNPF_RCASE(32, 32, 12) PNPF_RCASE(32, 64, 12) PNPF_RCASE(64, 32, 12) PNPF_RCASE(64, 64, 12)
        PNPF_RCASE(32, 16, 8) PNPF_RCASE(32, 32, 8) PNPF_RCASE(32, 64, 8) PNPF_RCASE(64, 16, 8) PNPF_RCASE(64, 32, 8) PNPF_RCASE(64, 64, 8)
#undef PNPF_RCASE
        set_error("no rowconv instantiation for BK=%d BN=%d epilogue warps %d", op.BK, op.BN, op.n_epi);
        return 2;
    }
    return launch_conv_gemm(op.BK, op.BN, op.tmA, op.tmA2, op.tmB, op.tmBh, op.p, s);
}

// ---- fused attention core ---------------------------------------------------------------------------------------------------
bool attn_core_eligible(int L, int C) {
    static const bool off = getenv("PNPF_NO_FUSED_ATTN") != nullptr;     // A/B switch (tools/ab_env.py)
    return !off && L == AttnCfg::L && C == AttnCfg::C;
}
int prepare_attn(AttnOp& op, const AttnDesc& d) {
    PNPF_REQUIRE(d.L == AttnCfg::L && d.C == AttnCfg::C, "fused attention core is built for L = %d tokens, C = %d channels (got %d, %d)", AttnCfg::L,
                 AttnCfg::C, d.L, d.C);
    PNPF_REQUIRE(d.qk && d.vT && d.w && d.out, "fused attention: null operand");
    memset(&op.p, 0, sizeof(op.p));
    op.p.n_img = d.B; op.p.L = d.L; op.p.C = d.C;
    EpiParams& e = op.p.epi;
    e.out = d.out; e.out_mode = 0; e.out_img_stride = (long long)d.L * d.C; e.out_row_stride = d.C; e.out_col_stride = 1; e.n_valid = d.C;
    e.bias = d.bias;
    e.residual = d.residual; e.res_img_stride = (long long)d.L * d.C; e.res_row_stride = d.C;
    e.stats = d.stats_out;
    if (int rc = make_act_tmap(&op.tmQ, d.qk, d.C, 2LL * d.C, d.L, 1, d.B, 64, 128, 1, 1)) return rc;
    if (int rc = make_b_tmap(&op.tmK, d.qk + d.C, d.C, 2LL * d.C, d.L, d.B, (long long)d.L * 2 * d.C, 64, 256)) return rc;
    if (int rc = make_b_tmap(&op.tmV, d.vT, d.L, d.L, d.C, d.B, (long long)d.C * d.L, 64, 256)) return rc;
    if (int rc = make_b_tmap(&op.tmW, d.w, d.C, d.C, d.C, 1, 0, 64, 256)) return rc;
    op.flops = (double)d.B * (2.0 * d.L * d.L * d.C * 2 + 2.0 * d.L * d.C * d.C);
    return 0;
}
int launch_attn(const AttnOp& op, int n_img, cudaStream_t s) {
    static DeviceCache cache;
    int dummy = 0;
    if (!cache.lookup(&dummy)) {
        PNPF_CHECK_CUDA(cudaFuncSetAttribute(attn_core_kernel<AttnCfg::L, AttnCfg::C>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg::SMEM_BYTES));
        cache.store(1);
    }
    AttnParams q = op.p;
    q.n_img = n_img;
    const int units = n_img * (AttnCfg::L / 128);
    if (units < 1) return 0;
    const int grid = units < num_sms() ? units : num_sms();
    attn_core_kernel<AttnCfg::L, AttnCfg::C><<<grid, AttnCfg::THREADS, AttnCfg::SMEM_BYTES, s>>>(op.tmQ, op.tmK, op.tmV, op.tmW, q);
    PNPF_CHECK_CUDA(cudaGetLastError());
    return 0;
}

static inline act16 f2bf(float f) { return to_act16(f); }

void pack_conv_weight(act16* dst, const float* w, int O, int Cin, int ks, int N_pad, int Cin_pad, const float* w2, int C2,
                      float scale) {
    const long long Ktot = (long long)ks * ks * Cin_pad + C2;
    for (long long i = 0; i < (long long)N_pad * Ktot; ++i) dst[i] = f2bf(0.f);
    for (int o = 0; o < O; ++o) {
        act16* row = dst + (long long)o * Ktot;
        for (int kh = 0; kh < ks; ++kh)
            for (int kw = 0; kw < ks; ++kw)
                for (int c = 0; c < Cin; ++c)
                    row[(long long)(kh * ks + kw) * Cin_pad + c] = f2bf(scale * w[(((long long)o * Cin + c) * ks + kh) * ks + kw]);
        if (w2)
            for (int c = 0; c < C2; ++c) row[(long long)ks * ks * Cin_pad + c] = f2bf(w2[(long long)o * C2 + c]);
    }
}

void pack_subpixel_pair_weights(act16* dst, const float* w, int O, int Cin, int a) {
    const long long Ktot = 6LL * Cin;
    std::vector<float> f((size_t)O * Cin * 4);
    for (long long i = 0; i < 2LL * O * Ktot; ++i) dst[i] = f2bf(0.f);
    for (int b = 0; b < 2; ++b) {
        fold_subpixel_weights(w, O, Cin, a, b, f.data());                 // [O][Cin][2][2]
        for (int o = 0; o < O; ++o) {
            act16* row = dst + (long long)(b * O + o) * Ktot;
            for (int i = 0; i < 2; ++i)
                for (int j = 0; j < 2; ++j)
                    for (int ch = 0; ch < Cin; ++ch) row[(long long)(i * 3 + (b + j)) * Cin + ch] = f2bf(f[(((size_t)o * Cin + ch) * 2 + i) * 2 + j]);
        }
    }
}

void fold_subpixel_weights(const float* w, int O, int Cin, int a, int b, float* out) {
    // rows of the 3x3 kernel that land on low-resolution row h-1+a+i for an output row 2h+a (same for columns): see pnpf_ops.h
    auto lo = [](int ph, int i) { return ph == 0 ? (i == 0 ? 0 : 1) : (i == 0 ? 0 : 2); };
    auto hi = [](int ph, int i) { return ph == 0 ? (i == 0 ? 0 : 2) : (i == 0 ? 1 : 2); };
    for (long long oc = 0; oc < (long long)O * Cin; ++oc)
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) {
                float acc = 0.f;
                for (int kh = lo(a, i); kh <= hi(a, i); ++kh)
                    for (int kw = lo(b, j); kw <= hi(b, j); ++kw) acc += w[oc * 9 + kh * 3 + kw];
                out[oc * 4 + i * 2 + j] = acc;
            }
}

}  // namespace pnpf
