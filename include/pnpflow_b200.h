/* pnpflow_b200 — C ABI of libpnpflow_sm100a.so
 *
 * B200-native (sm_100a) engine for the PnP-Flow restoration hot path.  The reference
 * (annegnx/PnP-Flow) has NO native interface on this path — it is pure Python/PyTorch — so every
 * entry point below cites the reference *Python* symbol it replaces (file:line under /root/reference).
 * The Python binding a maintainer would add is shown in INTEGRATION.md (ctypes, mirrors
 * pnpflow_b200/_lib.py).
 *
 * Conventions
 *  - all pointers are raw device pointers unless a parameter is named host_*; no torch types;
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream and
 *    CUDA-graph capturable unless stated otherwise; the library never calls cudaMalloc on the step path;
 *  - return value 0 = success; otherwise pnpf_last_error() (thread-local) describes the failure.
 *    There is NO CPU fallback: unsupported shapes are errors.
 *  - image tensors at the boundary are fp32 NCHW contiguous exactly like the reference's
 *    (SURVEY.md §8a); inside the engine activations and weights are IEEE fp16 (NHWC; 11-bit significand like
 *    the TF32 operands of the reference's cuDNN path), accumulated in fp32.
 */
#ifndef PNPFLOW_B200_H
#define PNPFLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNPF_ABI_VERSION 1

int pnpf_abi_version(void);
const char* pnpf_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * Velocity U-Net  v_theta(x, t)        replaces pnpflow/models.py:302-495 (UNet) as called from
 *                                      pnpflow/methods/pnp_flow.py:19-21 (PNP_FLOW.model_forward)
 * ---------------------------------------------------------------------------------------------- */
typedef struct pnpf_engine pnpf_engine; /* opaque */

typedef struct {
    int input_channels;        /* UNet(input_channels=...)      pnpflow/utils.py:172 */
    int input_height;          /* UNet(input_height=...)        pnpflow/utils.py:173 */
    int ch;                    /* 32                            pnpflow/utils.py:174 */
    int num_levels;            /* len(ch_mult)                  */
    int ch_mult[8];            /* (1,2,4,8)                     pnpflow/utils.py:175 */
    int num_res_blocks;        /* 6                             pnpflow/utils.py:176 */
    int num_attn_resolutions;  /* len(attn_resolutions)         */
    int attn_resolutions[8];   /* (16,8)                        pnpflow/utils.py:177 */
} pnpf_unet_config;

/* models.py:302-440 (UNet.__init__): builds the layer plan; weights are loaded afterwards. */
int pnpf_create(const pnpf_unet_config* cfg, pnpf_engine** out);
void pnpf_destroy(pnpf_engine* e);

/* pnpflow/utils.py:225 (model.load_state_dict): one call per state_dict entry, fp32 HOST data in the
 * reference layout (conv OIHW, linear [out,in], vectors).  Unknown names / wrong shapes are errors. */
int pnpf_load_weight(pnpf_engine* e, const char* name, const float* host_data, const int64_t* shape, int ndim);
/* Number of state_dict entries the plan expects, and the i-th expected name (for the loader / tests). */
int pnpf_num_weights(pnpf_engine* e);
const char* pnpf_weight_name(pnpf_engine* e, int i);
int pnpf_weight_shape(pnpf_engine* e, int i, int64_t shape[4], int* ndim);
/* Repack all loaded weights to the engine layout (fp16 K-major GEMM operands, folded attention scale/bias)
 * and upload them.  Fails if any expected entry is missing.  Synchronous (allocates device memory). */
int pnpf_finalize_weights(pnpf_engine* e);

/* Workspace: caller-owned device buffer (e.g. torch.empty) holding all activations for up to max_batch
 * images; the engine owns only packed weights and TMA descriptors.  Binding (re)builds the launch plan. */
size_t pnpf_workspace_bytes(pnpf_engine* e, int max_batch);
int pnpf_bind_workspace(pnpf_engine* e, void* workspace, size_t bytes, int max_batch);

/* models.py:442-495 (UNet.forward): x fp32 [batch,C,H,W], t fp32 [batch] (raw, in [0,1]),
 * v fp32 [batch,C,H,W].  batch <= max_batch of the bound workspace. */
int pnpf_unet_forward(pnpf_engine* e, const float* x, const float* t, float* v, int batch, void* stream);

/* Debug/parity taps: number of ops in the plan, their names, and "run the first n_ops ops only" so that a
 * test can read an intermediate activation (fp16 NHWC) back through pnpf_debug_read_op_output. */
int pnpf_debug_num_ops(pnpf_engine* e);
const char* pnpf_debug_op_name(pnpf_engine* e, int i);
int pnpf_debug_forward_partial(pnpf_engine* e, const float* x, const float* t, int batch, int n_ops, void* stream);
/* copies op i's output as fp32 NCHW [batch,C,H,W] into dst (device); writes C,H,W to dims[3] */
int pnpf_debug_read_op_output(pnpf_engine* e, int i, int batch, float* dst, size_t dst_elems, int dims[3], void* stream);
/* per-op accounting (per image): kind 1 = tensor-core conv/GEMM op, 0 = SIMT; algorithmic FLOPs and HBM bytes */
int pnpf_debug_op_info(pnpf_engine* e, int i, int* kind, double* flops, double* bytes);
/* which kernel runs op i: "rowconv<BK,BN,KCH> ...", "patchconv<BN> ...", "conv_gemm<BK,BN> ...", "gn_apply", ... (reports) */
const char* pnpf_debug_op_impl(pnpf_engine* e, int i);
/* one forward with CUDA events around every op: host_ms[n], n == pnpf_debug_num_ops(). Synchronous (bench/roofline). */
int pnpf_profile_forward(pnpf_engine* e, const float* x, const float* t, float* v, int batch, float* host_ms, int n,
                         void* stream);
double pnpf_unet_flops_per_image(pnpf_engine* e);  /* algorithmic 2*MAC of all tensor-core ops, per image */
int pnpf_unet_num_launches(pnpf_engine* e);        /* kernels launched by one pnpf_unet_forward */

/* ------------------------------------------------------------------------------------------------
 * PnP-Flow per-pixel kernels (fp32 NCHW)
 * ---------------------------------------------------------------------------------------------- */
/* Operator descriptor for the data-fidelity step.  Replaces the Degradation.H / H_adj pair of
 * pnpflow/degradations.py:6-127 inside grad_datafit (pnp_flow.py:39-41). */
enum {
    PNPF_OP_IDENTITY = 0, /* Denoising            degradations.py:15-20 */
    PNPF_OP_BOX = 1,      /* BoxInpainting        degradations.py:23-32, utils.py:327-336 */
    PNPF_OP_MASK = 2,     /* Random/Paintbrush    degradations.py:35-52 (cached uint8 keep-mask [B,H,W]) */
    PNPF_OP_SR = 3,       /* Superresolution      degradations.py:92-127 mode None */
    PNPF_OP_BLUR = 4,     /* GaussianDeblurring   degradations.py:55-89 (separable circular Gaussian) */
    PNPF_OP_SR_BICUBIC = 5 /* Superresolution(mode='bicubic')  degradations.py:97-109,117-127 + utils.py:365-396 (separable circular
                              4*sf-tap bicubic filter + decimation; adjoint = zero-fill + correlation) */
};
typedef struct {
    int kind;
    int half_size;        /* BOX: half_size_mask */
    const uint8_t* mask;  /* MASK: device [B,H,W] keep mask (1 = observed) */
    int sf;               /* SR, SR_BICUBIC: scale factor */
    const float* taps;    /* BLUR: device 1-D normalised Gaussian, ksize taps (kernel = outer(taps,taps));
                             SR_BICUBIC: device 1-D normalised bicubic weights, ksize = 4*sf taps */
    int ksize;            /* BLUR: 61; SR_BICUBIC: 4*sf */
    float* scratch;       /* BLUR: device scratch, B*C*H*W floats; SR_BICUBIC: B*C*(H/sf)*(W/sf) floats */
} pnpf_operator;

/* y = A x.                       Degradation.H      (degradations.py)      x [B,C,H,W] -> y [B,C,Hy,Wy] */
int pnpf_apply_H(const pnpf_operator* op, const float* x, float* y, int B, int C, int H, int W, void* stream);
/* x = A^T y.                     Degradation.H_adj  (degradations.py) */
int pnpf_apply_H_adj(const pnpf_operator* op, const float* y, float* x, int B, int C, int H, int W, void* stream);
/* z = x - gamma * A^T(A x - y)   pnp_flow.py:39-41 + :111-112 with gamma = lr_t (sigma^2 cancels, :60-62) */
int pnpf_datafit_step(const pnpf_operator* op, const float* x, const float* y, float* z, float gamma, int B, int C, int H,
                      int W, void* stream);
/* Laplace noise model: z = x - gamma * A^T(2*heaviside(Ax - y, 0) - 1)   pnp_flow.py:42-43 + :111-112, gamma = lr_t / sigma
 * (lr = sigma * lr_pnp, :64-66, so sigma cancels like sigma^2 does in the gaussian case). */
int pnpf_datafit_step_laplace(const pnpf_operator* op, const float* x, const float* y, float* z, float gamma, int B, int C, int H,
                              int W, void* stream);
/* zt[s] = t*z + (1-t)*eps[s], s < S      pnp_flow.py:47-48 (interpolation_step) for the S Monte-Carlo draws of one
 * step (:115-117); n = B*C*H*W, eps and zt are [S][n] (draw-major), z is [n]. */
int pnpf_interp(const float* z, const float* eps, float t, float* zt, long long n, int S, void* stream);
/* x_new = (sum_{s<S} (zt_s + (1-t) v_s)) / S        pnp_flow.py:50-52,114-121.
 * zt, v: [S][n] contiguous (draw-major), x_new: [n]. Summation order s = 0..S-1 like the reference. */
int pnpf_push_accum(const float* zt, const float* v, float t, int S, float* x_new, long long n, void* stream);

/* ONE whole PnP-Flow iteration (pnp_flow.py:107-121) on the S*B Monte-Carlo batch, for hosts that do not want to sequence
 * the four calls above themselves:
 *     z    = x - gamma * A^T(Ax - y)            (laplace != 0: A^T(2*heaviside(Ax - y, 0) - 1))       :39-45,111-112
 *     zt_s = t z + (1-t) eps_s,  s < S          eps: [S][B,C,H,W] noise the caller drew (torch Philox for seed parity)   :47-48
 *     v_s  = v_theta(zt_s, t)                   one U-Net evaluation at batch S*B (workspace bound for >= S*B)           :19-21
 *     x_new = (sum_s (zt_s + (1-t) v_s)) / S                                                                              :50-52,114-121
 * t = fp32(fp32(1/steps) * it) and gamma = lr_pnp * g(t) are computed by the host (pnp_flow.py:29-37,107-108).
 * Caller-owned scratch: z [B,C,H,W], zt and v [S*B,C,H,W], t_dev [S*B]; x_new [B,C,H,W] (may alias x: x is last read by the first kernel).  Asynchronous on
 * `stream`, CUDA-graph capturable, no allocation. */
int pnpf_step(pnpf_engine* e, const pnpf_operator* op, int laplace, const float* x, const float* y, const float* eps, float t,
              float gamma, int S, int B, int C, int H, int W, float* z, float* zt, float* t_dev, float* v, float* x_new,
              void* stream);

/* Forward-only sibling that shares the U-Net engine: one fixed-grid Euler step of the flow-matching ODE
 *     x_next = x + dt * v_theta(x, t0)
 * as FLOW_MATCHING.generate_samples(integration_method="euler") drives it through torchdiffeq's fixed-grid solver
 * (pnpflow/train_flow_matching.py:170-198, cnf.forward :252-262).  t_dev [B] and v [B,C,H,W] are caller-owned scratch;
 * x_next may alias x.  Asynchronous on `stream`, graph capturable. */
int pnpf_euler_step(pnpf_engine* e, const float* x, float t0, float dt, int B, int C, int H, int W, float* t_dev, float* v,
                    float* x_next, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Layer-level entry points (parity tests of the individual kernels; same kernels the plan uses)
 * ---------------------------------------------------------------------------------------------- */
/* F.conv2d(x, w, b, stride, padding=ksize//2) on fp16 NHWC activations with the tcgen05 implicit-GEMM kernel.
 * x: device fp16 [B,Hin,Win,Cin]; host_w: HOST fp32 OIHW [Cout,Cin,k,k]; host_bias: HOST fp32 [Cout] or NULL;
 * x2/host_w2: optional fused 1x1 over a second fp16 NHWC source at output resolution (ResBlock shortcut,
 * models.py:85-92,108); residual: optional fp16 NHWC [B,Hout,Wout,Cout] added in the epilogue;
 * out: device, fp16 NHWC (out_f32=0) or fp32 NHWC (out_f32=1), [B,Hout,Wout,Cout].  Synchronous. */
int pnpf_conv2d_nhwc(const void* x, int B, int Hin, int Win, int Cin, const float* host_w, const float* host_bias, int Cout,
                     int ksize, int stride, const void* x2, int C2, const float* host_w2, const void* residual, void* out,
                     int out_f32, void* stream);
/* out = conv3x3(act(GroupNorm_32(cat[xa | xb]))) + bias with the GroupNorm(+SiLU) applied INSIDE the row-streaming conv kernel
 * (shared-memory transform between TMA and MMA) and the concat read straight from its two sources (xb may be NULL).
 * models.py:94-101 (norm1 -> act -> conv1 on torch.cat([h, skip])).  Needs W % 128 == 0 and Cout <= 64.  Synchronous. */
int pnpf_gn_conv2d_nhwc(const void* xa, int Ca, const void* xb, int Cb, int B, int H, int W, const float* host_gamma,
                        const float* host_beta, const float* host_w, const float* host_bias, int Cout, int silu, void* out,
                        int out_f32, void* stream);
/* Sub-pixel form of Upsample (models.py:41-47: F.interpolate(scale 2, nearest) followed by a 3x3 conv): the 3x3 weights
 * host_w [Cout,Cin,3,3] folded into the 2x2 weights host_out [Cout,Cin,2,2] that output pixels (2h+a, 2w+b) apply to the
 * low-resolution pixels (h-1+a+i, w-1+b+j).  Host-only (no GPU needed). */
int pnpf_fold_subpixel_weights(const float* host_w, int Cout, int Cin, int a, int b, float* host_out);
/* The packed operand of the two-column-phase launch for output-row parity a (0 / 1): host_out [2*Cout][6*Cin] (fp16 values
 * widened to fp32): row b*Cout + o, K index (i*3 + c)*Cin + ch holds W_ab[o][ch][i][c - b] when c - b is 0 or 1, else 0, so that
 * out[(2h+a, 2w+b), o] = sum_{i,c,ch} host_out[b*Cout+o][(i*3+c)*Cin+ch] * x[h-1+a+i, w-1+c, ch].  Host-only (no GPU needed). */
int pnpf_pack_subpixel_pair_weights(const float* host_w, int Cout, int Cin, int a, float* host_out);
/* conv3x3(nearest_x2(x)) + bias computed as four sub-pixel phases on the low-resolution tensor (patch-streaming kernel,
 * the form the U-Net plan uses for the three up convs).  x: device fp16 [B,H,W,Cin]; out: device fp16 [B,2H,2W,Cout];
 * W <= 128, Cin % 64 == 0, Cout in {64,128,256}.  Synchronous. */
int pnpf_upconv2x_nhwc(const void* x, int B, int H, int W, int Cin, const float* host_w, const float* host_bias, int Cout,
                       void* out, void* stream);
/* out[b] = A[b] (M x K) * Bm[b]^T (N x K), fp16 row-major operands, fp32 (out_f32=1) or fp16 output. Synchronous. */
int pnpf_gemm_nt(const void* A, const void* Bm, void* out, int batch, int M, int N, int K, int out_f32, void* stream);
/* Fused attention core of the 16x16 attention blocks (models.py:145-162 after the q/k/v projections):
 * out = residual + softmax(q k^T) v Wo^T + bias, one kernel, logits and probabilities stay on chip (pnpf_attn.cuh).
 * qk: device fp16 [B,L,2C] (q already scaled by C^-1/2 | k); vT: device fp16 [B,C,L]; host_wo: HOST fp32 [C,C] (proj_out, OI);
 * host_bias: HOST fp32 [C] or NULL; residual: device fp16 [B,L,C] or NULL; out: device fp16 [B,L,C].  L = C = 256.  Synchronous. */
int pnpf_attn_core_nhwc(const void* qk, const void* vT, const float* host_wo, const float* host_bias, const void* residual, void* out,
                        int B, int L, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PNPFLOW_B200_H */
