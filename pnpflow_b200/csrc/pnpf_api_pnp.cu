// C-ABI: PnP-Flow per-pixel kernels (data-fidelity step, interpolation, Euler push + Monte-Carlo average).
#include "../../include/pnpflow_b200.h"
#include "pnpf_kernels.cuh"
#include "pnpf_ops.h"

using namespace pnpf;

static int to_desc(const pnpf_operator* op, OpDesc& d) {
    PNPF_REQUIRE(op, "null operator");
    PNPF_REQUIRE(op->kind >= PNPF_OP_IDENTITY && op->kind <= PNPF_OP_SR_BICUBIC, "unknown operator kind %d", op->kind);
    d.kind = op->kind;
    d.half_size = op->half_size;
    d.sf = op->sf;
    d.ksize = op->ksize;
    d.mask = op->mask;
    d.taps = op->taps;
    d.scratch = op->scratch;
    return 0;
}

extern "C" int pnpf_apply_H(const pnpf_operator* op, const float* x, float* y, int B, int C, int H, int W, void* stream) {
    OpDesc d;
    if (int rc = to_desc(op, d)) return rc;
    PNPF_REQUIRE(x && y, "null pointer");
    return launch_apply_H(d, x, y, B, C, H, W, false, static_cast<cudaStream_t>(stream));
}
extern "C" int pnpf_apply_H_adj(const pnpf_operator* op, const float* y, float* x, int B, int C, int H, int W, void* stream) {
    OpDesc d;
    if (int rc = to_desc(op, d)) return rc;
    PNPF_REQUIRE(x && y, "null pointer");
    return launch_apply_H(d, y, x, B, C, H, W, true, static_cast<cudaStream_t>(stream));
}
extern "C" int pnpf_datafit_step(const pnpf_operator* op, const float* x, const float* y, float* z, float gamma, int B, int C,
                                 int H, int W, void* stream) {
    OpDesc d;
    if (int rc = to_desc(op, d)) return rc;
    PNPF_REQUIRE(x && y && z, "null pointer");
    return launch_datafit(d, x, y, z, gamma, 0, B, C, H, W, static_cast<cudaStream_t>(stream));
}
extern "C" int pnpf_datafit_step_laplace(const pnpf_operator* op, const float* x, const float* y, float* z, float gamma, int B,
                                         int C, int H, int W, void* stream) {
    OpDesc d;
    if (int rc = to_desc(op, d)) return rc;
    PNPF_REQUIRE(x && y && z, "null pointer");
    return launch_datafit(d, x, y, z, gamma, 1, B, C, H, W, static_cast<cudaStream_t>(stream));
}
extern "C" int pnpf_interp(const float* z, const float* eps, float t, float* zt, long long n, int S, void* stream) {
    PNPF_REQUIRE(z && eps && zt && n >= 0 && S >= 1, "bad argument");
    return launch_interp(z, eps, t, zt, n, S, static_cast<cudaStream_t>(stream));
}
extern "C" int pnpf_push_accum(const float* zt, const float* v, float t, int S, float* x_new, long long n, void* stream) {
    PNPF_REQUIRE(zt && v && x_new && n >= 0, "bad argument");
    return launch_push_accum(zt, v, t, S, x_new, n, static_cast<cudaStream_t>(stream));
}

// ---- one whole PnP-Flow step (pnp_flow.py:107-121) --------------------------------------------------------------------------
__global__ void fill_f32_kernel(float* __restrict__ p, float v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

extern "C" int pnpf_step(pnpf_engine* e, const pnpf_operator* op, int laplace, const float* x, const float* y, const float* eps,
                         float t, float gamma, int S, int B, int C, int H, int W, float* z, float* zt, float* t_dev, float* v,
                         float* x_new, void* stream) {
    OpDesc d;
    if (int rc = to_desc(op, d)) return rc;
    PNPF_REQUIRE(e && x && y && eps && z && zt && t_dev && v && x_new, "null pointer");
    PNPF_REQUIRE(S >= 1 && B >= 1, "num_samples %d, batch %d", S, B);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long n = (long long)B * C * H * W;
    if (int rc = launch_datafit(d, x, y, z, gamma, laplace ? 1 : 0, B, C, H, W, st)) return rc;          // :111-112
    if (int rc = launch_interp(z, eps, t, zt, n, S, st)) return rc;                                      // :47-48 for the S draws
    fill_f32_kernel<<<(S * B + 255) / 256, 256, 0, st>>>(t_dev, t, S * B);                               // t1 = ones(B) * delta * it (:107-108)
    PNPF_CHECK_CUDA(cudaGetLastError());
    if (int rc = pnpf_unet_forward(e, zt, t_dev, v, S * B, stream)) return rc;                           // :19-21 on the S*B batch
    return launch_push_accum(zt, v, t, S, x_new, n, st);                                                 // :50-52,114-121
}

// ---- forward-only sibling of the path: Euler sampling of the flow-matching ODE (train_flow_matching.py:170-198) ---------------
extern "C" int pnpf_euler_step(pnpf_engine* e, const float* x, float t0, float dt, int B, int C, int H, int W, float* t_dev, float* v,
                               float* x_next, void* stream) {
    PNPF_REQUIRE(e && x && t_dev && v && x_next && B >= 1, "bad argument");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    fill_f32_kernel<<<(B + 255) / 256, 256, 0, st>>>(t_dev, t0, B);                  // cnf.forward: t.repeat(x.shape[0]) (:258-262)
    PNPF_CHECK_CUDA(cudaGetLastError());
    if (int rc = pnpf_unet_forward(e, x, t_dev, v, B, stream)) return rc;
    return launch_axpy(x, v, dt, x_next, (long long)B * C * H * W, st);              // y1 = y0 + dt * f(t0, y0)
}
