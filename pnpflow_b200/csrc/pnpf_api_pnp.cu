// C-ABI: PnP-Flow per-pixel kernels (data-fidelity step, interpolation, Euler push + Monte-Carlo average).
#include "../../include/pnpflow_b200.h"
#include "pnpf_kernels.cuh"

using namespace pnpf;

static int to_desc(const pnpf_operator* op, OpDesc& d) {
    PNPF_REQUIRE(op, "null operator");
    PNPF_REQUIRE(op->kind >= PNPF_OP_IDENTITY && op->kind <= PNPF_OP_BLUR, "unknown operator kind %d", op->kind);
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
