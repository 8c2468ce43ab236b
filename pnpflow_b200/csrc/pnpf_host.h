// Host-side helpers shared by the translation units of libpnpflow_sm100a.so:
// error reporting, TMA tensor-map encoding (driver entry point resolved at run time, no -lcuda), GEMM launch.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "pnpf_gemm.cuh"

namespace pnpf {

// ---- error plumbing: every C-ABI entry returns 0 / non-zero and leaves a message for pnpf_last_error() ----
void set_error(const char* fmt, ...);
const char* get_error();

#define PNPF_CHECK_CUDA(expr)                                                                     \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            pnpf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return 1;                                                                             \
        }                                                                                         \
    } while (0)

#define PNPF_REQUIRE(cond, ...)            \
    do {                                   \
        if (!(cond)) {                     \
            pnpf::set_error(__VA_ARGS__);  \
            return 2;                      \
        }                                  \
    } while (0)

int num_sms();                      // of the CURRENT device (cached per device)

// Kernel attributes (dynamic shared-memory opt-in) and occupancy answers are PER DEVICE: every launcher keeps one of these as a
// function-local static and initialises it the first time it runs on a device.  lookup()/store() are guarded, a racing double
// initialisation is harmless (the attribute calls are idempotent).
struct DeviceCache {
    static constexpr int MAX_DEV = 64;
    bool lookup(int* value);        // false: not initialised for the current device yet
    void store(int value);
private:
    int val_[MAX_DEV] = {};
    bool set_[MAX_DEV] = {};
};

// ---- tensor maps ----
// Activation map: fp16 NHWC buffer viewed as 4-D (C, W, H, B); `pitch` = channel pitch of the buffer in elements
// (>= C when the view is a channel slice).  box = {bk, tw, th, 1}; `estride` = traversal stride in W and H (1 or 2).
int make_act_tmap(CUtensorMap* m, const void* base, int C, long long pitch, int W, int H, int B, int bk, int tw, int th,
                  int estride);
// Weight / B-operand map: fp16 [batch][N][K] K-major viewed as 3-D (K, N, batch); row pitch `ldk` elements,
// batch stride `bstride` elements (ignored when batch == 1).  box = {bk, bn, 1}.
int make_b_tmap(CUtensorMap* m, const void* base, long long K, long long ldk, int N, int batch, long long bstride, int bk,
                int bn);

// ---- GEMM launch (dispatch on BK in {32,64} and BN in {16,32,64,128,256}) ----
int launch_conv_gemm(int BK, int BN, const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB, const CUtensorMap& tmB_half,
                     const GemmParams& p, cudaStream_t stream);

// tile box for a W-wide feature map: TW = min(pow2ceil(W),128), TH = 128/TW
inline void pick_tile(int W, int& TH, int& TW) {
    int tw = 1;
    while (tw < W && tw < 128) tw <<= 1;
    if (tw < 8) tw = 8;
    TW = tw;
    TH = 128 / tw;
}

}  // namespace pnpf
