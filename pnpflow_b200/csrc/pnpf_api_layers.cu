// C-ABI: layer-level entry points (used by the parity tests; they drive the very same kernels as the U-Net plan).
#include "../../include/pnpflow_b200.h"
#include "pnpf_ops.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace pnpf;

extern "C" int pnpf_abi_version(void) { return PNPF_ABI_VERSION; }
extern "C" const char* pnpf_last_error(void) { return get_error(); }

static int round_up_n(int cout) {
    if (cout <= 16) return 16;
    if (cout <= 32) return 32;
    if (cout <= 64) return 64;
    if (cout <= 128) return 128;
    return (cout + 255) / 256 * 256;
}

extern "C" int pnpf_conv2d_nhwc(const void* x, int B, int Hin, int Win, int Cin, const float* host_w, const float* host_bias,
                                int Cout, int ksize, int stride, const void* x2, int C2, const float* host_w2,
                                const void* residual, void* out, int out_f32, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PNPF_REQUIRE(x && host_w && out, "null pointer");
    PNPF_REQUIRE(Cout % 16 == 0, "layer-level conv needs Cout %% 16 == 0 (got %d)", Cout);
    const int N_pad = round_up_n(Cout);
    const int Hout = (Hin + 2 * (ksize / 2) - ksize) / stride + 1;
    const int Wout = (Win + 2 * (ksize / 2) - ksize) / stride + 1;
    const long long Ktot = (long long)ksize * ksize * Cin + (x2 ? C2 : 0);
    std::vector<act16> wp((size_t)N_pad * Ktot);
    pack_conv_weight(wp.data(), host_w, Cout, Cin, ksize, N_pad, Cin, x2 ? host_w2 : nullptr, x2 ? C2 : 0, 1.0f);
    std::vector<float> bp(N_pad, 0.f);
    if (host_bias)
        for (int i = 0; i < Cout; ++i) bp[i] = host_bias[i];
    act16* dw = nullptr;
    float* db = nullptr;
    PNPF_CHECK_CUDA(cudaMalloc(&dw, wp.size() * sizeof(act16)));
    PNPF_CHECK_CUDA(cudaMalloc(&db, bp.size() * sizeof(float)));
    PNPF_CHECK_CUDA(cudaMemcpyAsync(dw, wp.data(), wp.size() * sizeof(act16), cudaMemcpyHostToDevice, s));
    PNPF_CHECK_CUDA(cudaMemcpyAsync(db, bp.data(), bp.size() * sizeof(float), cudaMemcpyHostToDevice, s));
    ConvDesc d;
    d.x = static_cast<const act16*>(x);
    d.B = B; d.Hin = Hin; d.Win = Win; d.Cin = Cin; d.x_pitch = Cin;
    d.x2 = static_cast<const act16*>(x2); d.C2 = C2; d.x2_pitch = C2;
    d.w = dw; d.N_pad = N_pad; d.ksize = ksize; d.stride = stride; d.Hout = Hout; d.Wout = Wout;
    d.out = out; d.out_mode = out_f32 ? 1 : 0;
    d.out_img_stride = (long long)Hout * Wout * Cout; d.out_row_stride = Cout; d.n_valid = Cout;
    d.bias = db;
    d.residual = static_cast<const act16*>(residual);
    d.res_img_stride = (long long)Hout * Wout * Cout; d.res_row_stride = Cout;
    TcOp op;
    int rc = prepare_conv(op, d);
    long long* dbg = nullptr;
    if (!rc && getenv("PNPF_ROWCONV_DBG")) {          // profiling experiment: cycle counters of CTA 0
        cudaMalloc(&dbg, 32 * sizeof(long long));
        cudaMemset(dbg, 0, 32 * sizeof(long long));
        op.rp.dbg = dbg;
        op.p.dbg = dbg;
        op.pp.dbg = dbg;
    }
    if (!rc) rc = launch_tc(op, s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (dbg) {
        long long h[32];
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%s producer: total %lld wait_empty %lld rows/tiles %lld | mma: total %lld wait_full %lld wait_tempty %lld | "
               "epi0: total %lld wait_tfull %lld rows %lld | epi1: total %lld wait_tfull %lld rows %lld\n", op.kind == 1 ? "ROWCONV_DBG" : "GEMM_DBG",
               h[0], h[1], h[2], h[4], h[5], h[6], h[8], h[9], h[10], h[12], h[13], h[14]);
        if (op.kind == 2)
            printf("PATCH_DBG tiles %lld | patch producer: total %lld wait_empty %lld | weight producer: total %lld wait_empty %lld | mma: total %lld "
                   "wait_patch %lld wait_weights %lld wait_tempty %lld | epi: total %lld wait_tfull %lld\n", h[2], h[0], h[1], h[12], h[13], h[4], h[5], h[7], h[6],
                   h[8], h[9]);
        if (op.kind == 2)
            printf("   epi: tile decode %lld chunks (tcgen05.ld, bias / residual, statistics, stores) %lld fence + arrive %lld\n", h[24], h[25], h[26]);
        if (op.kind == 1)
            printf("   mma: issue %lld commit %lld | epi0: tmem_ld %lld tmem_st+arrive %lld math+stage %lld store %lld stats_flush %lld\n", h[16], h[17], h[18],
                   h[19], h[22], h[23], h[20]);
        fflush(stdout);
        cudaFree(dbg);
    }
    cudaFree(dw);
    cudaFree(db);
    if (rc) return rc;
    PNPF_CHECK_CUDA(e);
    return 0;
}

extern "C" int pnpf_fold_subpixel_weights(const float* host_w, int Cout, int Cin, int a, int b, float* host_out) {
    PNPF_REQUIRE(host_w && host_out && Cout > 0 && Cin > 0 && (a | b) >= 0 && (a | b) <= 1, "fold_subpixel_weights: bad arguments");
    fold_subpixel_weights(host_w, Cout, Cin, a, b, host_out);
    return 0;
}

extern "C" int pnpf_pack_subpixel_pair_weights(const float* host_w, int Cout, int Cin, int a, float* host_out) {
    PNPF_REQUIRE(host_w && host_out && Cout > 0 && Cin > 0 && (a == 0 || a == 1), "pack_subpixel_pair_weights: bad arguments");
    std::vector<act16> packed((size_t)2 * Cout * 6 * Cin);
    pack_subpixel_pair_weights(packed.data(), host_w, Cout, Cin, a);
    for (size_t i = 0; i < packed.size(); ++i) host_out[i] = __half2float(packed[i]);
    return 0;
}

extern "C" int pnpf_upconv2x_nhwc(const void* x, int B, int H, int W, int Cin, const float* host_w, const float* host_bias, int Cout,
                                  void* out, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PNPF_REQUIRE(x && host_w && out, "null pointer");
    const int N_pad = round_up_n(Cout);
    PNPF_REQUIRE(Cout == N_pad, "sub-pixel up conv needs Cout in {64, 128, 256} (got %d)", Cout);
    std::vector<float> bp(N_pad, 0.f);
    if (host_bias)
        for (int i = 0; i < Cout; ++i) bp[i] = host_bias[i];
    // same choice as the U-Net plan: two column phases per launch (2 launches) when 2 * Cout <= 256, else four single phases
    const bool pairph = 2 * N_pad <= 256 && getenv("PNPF_NO_SUBPIX2") == nullptr;
    const size_t wn = pairph ? (size_t)2 * N_pad * 6 * Cin : (size_t)N_pad * 4 * Cin;
    const int nlaunch = pairph ? 2 : 4;
    std::vector<act16> wp(nlaunch * wn);
    std::vector<float> f((size_t)Cout * Cin * 4);
    for (int ph = 0; ph < nlaunch; ++ph) {
        if (pairph) {
            pack_subpixel_pair_weights(wp.data() + ph * wn, host_w, Cout, Cin, ph);
        } else {
            fold_subpixel_weights(host_w, Cout, Cin, ph >> 1, ph & 1, f.data());
            pack_conv_weight(wp.data() + ph * wn, f.data(), Cout, Cin, 2, N_pad, Cin, nullptr, 0, 1.0f);
        }
    }
    act16* dw = nullptr;
    float* db = nullptr;
    PNPF_CHECK_CUDA(cudaMalloc(&dw, wp.size() * sizeof(act16)));
    PNPF_CHECK_CUDA(cudaMalloc(&db, bp.size() * sizeof(float)));
    PNPF_CHECK_CUDA(cudaMemcpyAsync(dw, wp.data(), wp.size() * sizeof(act16), cudaMemcpyHostToDevice, s));
    PNPF_CHECK_CUDA(cudaMemcpyAsync(db, bp.data(), bp.size() * sizeof(float), cudaMemcpyHostToDevice, s));
    int rc = 0;
    for (int ph = 0; ph < nlaunch && !rc; ++ph) {
        ConvDesc d;
        d.x = static_cast<const act16*>(x);
        d.B = B; d.Hin = d.Hout = H; d.Win = d.Wout = W; d.Cin = Cin; d.x_pitch = Cin;
        d.w = dw + ph * wn; d.ksize = 3; d.stride = 1;
        if (pairph) { d.N_pad = 2 * N_pad; d.subpix = 2; d.sp_a = ph; d.sp_b = 0; }
        else { d.N_pad = N_pad; d.subpix = 1; d.sp_a = ph >> 1; d.sp_b = ph & 1; }
        d.out = out; d.out_mode = 0;
        d.out_img_stride = 4LL * H * W * Cout; d.out_row_stride = Cout; d.n_valid = Cout;
        d.bias = db;
        TcOp op;
        if (!patchconv_eligible(d)) {
            set_error("sub-pixel up conv: shape %dx%d Cin=%d Cout=%d is not a patch-kernel shape", H, W, Cin, Cout);
            rc = 2;
            break;
        }
        rc = prepare_conv(op, d);
        if (!rc) rc = launch_tc(op, s);
    }
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(dw);
    cudaFree(db);
    if (rc) return rc;
    PNPF_CHECK_CUDA(e);
    return 0;
}

extern "C" int pnpf_gemm_nt(const void* A, const void* Bm, void* out, int batch, int M, int N, int K, int out_f32, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    GemmDesc d;
    d.A = static_cast<const act16*>(A); d.lda = K; d.a_bstride = (long long)M * K; d.a_batched = 1;
    d.Bm = static_cast<const act16*>(Bm); d.ldb = K; d.b_bstride = (long long)N * K; d.b_batched = 1;
    d.batch = batch; d.M = M; d.N = N; d.K = K;
    d.out = out; d.out_mode = out_f32 ? 1 : 0; d.out_img_stride = (long long)M * N; d.out_row_stride = N;
    TcOp op;
    if (int rc = prepare_gemm(op, d)) return rc;
    if (int rc = launch_tc(op, s)) return rc;
    PNPF_CHECK_CUDA(cudaStreamSynchronize(s));
    return 0;
}

#include "pnpf_kernels.cuh"

extern "C" int pnpf_gn_conv2d_nhwc(const void* xa, int Ca, const void* xb, int Cb, int B, int H, int W, const float* host_gamma,
                                   const float* host_beta, const float* host_w, const float* host_bias, int Cout, int silu,
                                   void* out, int out_f32, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PNPF_REQUIRE(xa && host_gamma && host_beta && host_w && out, "null pointer");
    PNPF_REQUIRE(Cout % 16 == 0, "Cout %% 16");
    const int Cin = Ca + Cb;
    const int N_pad = round_up_n(Cout);
    const long long Ktot = 9LL * Cin;
    std::vector<act16> wp((size_t)N_pad * Ktot);
    pack_conv_weight(wp.data(), host_w, Cout, Cin, 3, N_pad, Cin, nullptr, 0, 1.0f);
    std::vector<float> bp(N_pad, 0.f);
    if (host_bias)
        for (int i = 0; i < Cout; ++i) bp[i] = host_bias[i];
    act16* dw = nullptr;
    float *db = nullptr, *dg = nullptr, *dbeta = nullptr;
    double *sta = nullptr, *stb = nullptr;
    PNPF_CHECK_CUDA(cudaMalloc(&dw, wp.size() * sizeof(act16)));
    PNPF_CHECK_CUDA(cudaMalloc(&db, bp.size() * sizeof(float)));
    PNPF_CHECK_CUDA(cudaMalloc(&dg, Cin * sizeof(float)));
    PNPF_CHECK_CUDA(cudaMalloc(&dbeta, Cin * sizeof(float)));
    PNPF_CHECK_CUDA(cudaMalloc(&sta, (size_t)B * Ca * 2 * sizeof(double)));
    PNPF_CHECK_CUDA(cudaMalloc(&stb, (size_t)B * (Cb ? Cb : 1) * 2 * sizeof(double)));
    PNPF_CHECK_CUDA(cudaMemcpyAsync(dw, wp.data(), wp.size() * sizeof(act16), cudaMemcpyHostToDevice, s));
    PNPF_CHECK_CUDA(cudaMemcpyAsync(db, bp.data(), bp.size() * sizeof(float), cudaMemcpyHostToDevice, s));
    PNPF_CHECK_CUDA(cudaMemcpyAsync(dg, host_gamma, Cin * sizeof(float), cudaMemcpyHostToDevice, s));
    PNPF_CHECK_CUDA(cudaMemcpyAsync(dbeta, host_beta, Cin * sizeof(float), cudaMemcpyHostToDevice, s));
    PNPF_CHECK_CUDA(cudaMemsetAsync(sta, 0, (size_t)B * Ca * 2 * sizeof(double), s));
    PNPF_CHECK_CUDA(cudaMemsetAsync(stb, 0, (size_t)B * (Cb ? Cb : 1) * 2 * sizeof(double), s));
    int rc = launch_gn_stats(GnSrc{static_cast<const act16*>(xa), Ca, Ca, nullptr, 0, 0, nullptr, nullptr}, B, H * W, sta, s);
    if (!rc && Cb) rc = launch_gn_stats(GnSrc{static_cast<const act16*>(xb), Cb, Cb, nullptr, 0, 0, nullptr, nullptr}, B, H * W, stb, s);
    ConvDesc d;
    d.x = static_cast<const act16*>(xa); d.x_pitch = Ca; d.Cin = Cin;
    d.xb = static_cast<const act16*>(xb); d.Cb = Cb; d.xb_pitch = Cb;
    d.B = B; d.Hin = d.Hout = H; d.Win = d.Wout = W;
    d.w = dw; d.N_pad = N_pad; d.ksize = 3; d.stride = 1;
    d.out = out; d.out_mode = out_f32 ? 1 : 0;
    d.out_img_stride = (long long)H * W * Cout; d.out_row_stride = Cout; d.n_valid = Cout;
    d.bias = db;
    d.gn_gamma = dg; d.gn_beta = dbeta; d.gn_stats_a = sta; d.gn_stats_b = Cb ? stb : nullptr; d.gn_silu = silu;
    TcOp op;
    if (!rc) rc = prepare_conv(op, d);
    long long* dbg = nullptr;
    if (!rc && op.kind == 1 && getenv("PNPF_ROWCONV_DBG")) {
        cudaMalloc(&dbg, 32 * sizeof(long long));
        cudaMemset(dbg, 0, 32 * sizeof(long long));
        op.rp.dbg = dbg;
    }
    if (!rc) rc = launch_tc(op, s);
    cudaError_t e = cudaStreamSynchronize(s);
    if (dbg) {
        long long h[32];
        cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
        printf("ROWCONV_GN_DBG producer: total %lld wait_empty %lld rows %lld | transform: total %lld wait_full %lld table %lld | mma: total %lld "
               "wait_ready %lld wait_tempty %lld | epi0: total %lld wait_tfull %lld\n", h[0], h[1], h[2], h[3], h[7], h[11], h[4], h[5], h[6], h[8], h[9]);
        printf("   epi0: math+stage %lld store %lld\n", h[22], h[23]);
        printf("   mma: issue %lld commit %lld | epi0: tmem_ld %lld tmem_st+arrive %lld stats_flush %lld | transform: fence+arrive %lld\n", h[16], h[17], h[18],
               h[19], h[20], h[21]);
        fflush(stdout);
        cudaFree(dbg);
    }
    cudaFree(dw); cudaFree(db); cudaFree(dg); cudaFree(dbeta); cudaFree(sta); cudaFree(stb);
    if (rc) return rc;
    PNPF_CHECK_CUDA(e);
    return 0;
}

// Fused attention core on device tensors (parity test of pnpf_attn.cuh): qk fp16 [B][L][2C] (q scaled | k), vT fp16 [B][C][L],
// host_wo fp32 [C][C] (proj_out weight, OI), host_bias fp32 [C] or NULL, residual fp16 [B][L][C] or NULL -> out fp16 [B][L][C].
extern "C" int pnpf_attn_core_nhwc(const void* qk, const void* vT, const float* host_wo, const float* host_bias, const void* residual, void* out,
                                   int B, int L, int C, void* stream) {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    PNPF_REQUIRE(qk && vT && host_wo && out, "null pointer");
    PNPF_REQUIRE(attn_core_eligible(L, C), "fused attention core handles L = 256 tokens, C = 256 channels (got %d, %d)", L, C);
    std::vector<act16> wp((size_t)C * C);
    pack_conv_weight(wp.data(), host_wo, C, C, 1, C, C, nullptr, 0, 1.0f);
    std::vector<float> bp(C, 0.f);
    if (host_bias)
        for (int i = 0; i < C; ++i) bp[i] = host_bias[i];
    act16* dw = nullptr;
    float* db = nullptr;
    PNPF_CHECK_CUDA(cudaMalloc(&dw, wp.size() * sizeof(act16)));
    PNPF_CHECK_CUDA(cudaMalloc(&db, bp.size() * sizeof(float)));
    PNPF_CHECK_CUDA(cudaMemcpyAsync(dw, wp.data(), wp.size() * sizeof(act16), cudaMemcpyHostToDevice, s));
    PNPF_CHECK_CUDA(cudaMemcpyAsync(db, bp.data(), bp.size() * sizeof(float), cudaMemcpyHostToDevice, s));
    AttnDesc d;
    d.qk = static_cast<const act16*>(qk); d.vT = static_cast<const act16*>(vT); d.w = dw; d.bias = db;
    d.residual = static_cast<const act16*>(residual); d.out = static_cast<act16*>(out); d.B = B; d.L = L; d.C = C;
    AttnOp op;
    int rc = prepare_attn(op, d);
    if (!rc) rc = launch_attn(op, B, s);
    cudaError_t e = cudaStreamSynchronize(s);
    cudaFree(dw);
    cudaFree(db);
    if (rc) return rc;
    PNPF_CHECK_CUDA(e);
    return 0;
}
