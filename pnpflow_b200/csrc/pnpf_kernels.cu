// SIMT kernels of the engine (everything that is not a tensor-core contraction).  sm_100a.
#include "pnpf_kernels.cuh"

#include <cstdlib>

namespace pnpf {

// =================================================================================================
// helpers
// =================================================================================================
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 t = unpack2(w[j]);
        f[2 * j] = t.x;
        f[2 * j + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) w[j] = pack2(f[2 * j], f[2 * j + 1]);
    return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ const act16* src_ptr(const GnSrc& s, int img, long long pix, int HW, int c0) {
    if (c0 < s.C1) return s.p1 + ((long long)img * HW + pix) * s.pitch1 + c0;
    return s.p2 + ((long long)img * HW + pix) * s.pitch2 + (c0 - s.C1);
}

// pixels per block: the largest of {1024..64} that still gives >= 4 blocks per SM (small feature maps would
// otherwise run on a handful of SMs)
static inline int gn_pix_per_block(int HW, int B) {
    const long long want = 4LL * num_sms();
    for (int p = 1024; p > 64; p >>= 1)
        if ((long long)B * ((HW + p - 1) / p) >= want) return p;
    return 64;
}

// =================================================================================================
// GroupNorm statistics: per (image, channel) sum and sum of squares, fp32 in-block, fp64 atomics across blocks
// =================================================================================================
__global__ void gn_stats_kernel(GnSrc s, int HW, int pix_per_block, double* __restrict__ stats) {
    extern __shared__ float red[];                    // [ppb][C][2]
    const int C = s.C1 + s.C2;
    const int nvec = C >> 3;
    const int ppb = blockDim.x / nvec;
    const int v = threadIdx.x % nvec, pl = threadIdx.x / nvec;
    const int img = blockIdx.y;
    const int p0 = blockIdx.x * pix_per_block;
    const int p1 = min(p0 + pix_per_block, HW);
    float sum[8], sq[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sum[j] = sq[j] = 0.f;
    for (int p = p0 + pl; p < p1; p += ppb) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(src_ptr(s, img, p, HW, v * 8)));
        float f[8];
        unpack8(u, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sum[j] += f[j];
            sq[j] = fmaf(f[j], f[j], sq[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        red[(pl * C + v * 8 + j) * 2] = sum[j];
        red[(pl * C + v * 8 + j) * 2 + 1] = sq[j];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int q = 0; q < ppb; ++q) {
            a += red[(q * C + c) * 2];
            b += red[(q * C + c) * 2 + 1];
        }
        atomicAdd(&stats[((long long)img * C + c) * 2], (double)a);
        atomicAdd(&stats[((long long)img * C + c) * 2 + 1], (double)b);
    }
}

static inline int gn_threads(int C) {
    const int nvec = C / 8;
    int ppb = 256 / nvec;
    if (ppb < 1) ppb = 1;
    return nvec * ppb;
}

int launch_gn_stats(const GnSrc& s, int B, int HW, double* stats, cudaStream_t st) {
    const int C = s.C1 + s.C2;
    PNPF_REQUIRE(C % 32 == 0 && s.C1 % 8 == 0 && C <= 2048, "GroupNorm channels (%d+%d) unsupported", s.C1, s.C2);
    const int threads = gn_threads(C);
    const int ppb = threads / (C / 8);
    const size_t smem = (size_t)ppb * C * 2 * sizeof(float);
    const int ppb_blk = gn_pix_per_block(HW, B);
    dim3 grid((HW + ppb_blk - 1) / ppb_blk, B);
    gn_stats_kernel<<<grid, threads, smem, st>>>(s, HW, ppb_blk, stats);
    PNPF_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// =================================================================================================
// GroupNorm apply (+ SiLU) -> fp16 NHWC conv operand; optional raw concat copy
// =================================================================================================
// Prologue: one thread per GROUP reduces the fp64 channel statistics (a group may straddle the two concatenated
// sources) to mean / rstd, then one thread per channel derives scale / shift in fp32.  Main loop: a thread owns one
// 16-byte channel vector and walks the block's pixels four at a time (four independent loads in flight);
// SiLU(y) = h + h*tanh(h) with h = y/2 folded into scale/shift -> one MUFU op per element.
__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {        // read-once data: do not allocate in L1
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void stg_stream(uint4* p, const uint4& v) {  // write-once data: evict-first
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// Persistent form: the grid is exactly ONE wave (occupancy x SMs CTAs, all resident at once) and CTA b owns the contiguous
// pixel range [b T / G, (b + 1) T / G) of the flattened (image, pixel) space, T = B * HW — equal work for every CTA, no tail.
// (Round 1 launched (HW / ppb) x B blocks: 640 blocks on 444 resident slots = 1.44 waves, i.e. 72 % of the achievable rate, which
// is what the 0.70-0.74 of HBM in profiles/r01_gn_apply_ncu.json was.)  A CTA touches one image, rarely two: the per-image
// scale / shift table is rebuilt when the image changes.
template <int U, bool STREAM>
__global__ void __launch_bounds__(256, 3)
gn_apply_kernel(GnSrc s, int HW, long long total_pix, const float* __restrict__ gamma,
                const float* __restrict__ beta, float eps, int gs, int silu, act16* __restrict__ dst,
                act16* __restrict__ raw_dst) {
    extern __shared__ float ss[];                     // scale[C] | shift[C] | mean[G] | rstd[G]
    const int C = s.C1 + s.C2;
    const int G = C / gs;
    float* gmean = ss + 2 * C;
    float* grstd = gmean + G;
    const int nvec = C >> 3;
    const int ppb = blockDim.x / nvec;
    const int v = threadIdx.x % nvec, pl = threadIdx.x / nvec;
    const float half = silu ? 0.5f : 1.f;
    const long long g_begin = total_pix * blockIdx.x / gridDim.x, g_end = total_pix * (blockIdx.x + 1) / gridDim.x;
    for (long long g0 = g_begin; g0 < g_end;) {
        const int img = (int)(g0 / HW);
        const int p0 = (int)(g0 - (long long)img * HW);
        const int p1 = (int)min((long long)HW, p0 + (g_end - g0));
        g0 += p1 - p0;
        __syncthreads();                              // the previous image's table is no longer read
        for (int g = threadIdx.x; g < G; g += blockDim.x) {
            double S = 0, Q = 0;
            for (int j = 0; j < gs; ++j) {
                const int cc = g * gs + j;
                const double* sp = (cc < s.C1) ? s.st1 + ((long long)img * s.C1 + cc) * 2 : s.st2 + ((long long)img * s.C2 + (cc - s.C1)) * 2;
                S += sp[0];
                Q += sp[1];
            }
            const double n = (double)gs * HW;
            const double mean = S / n;
            double var = Q / n - mean * mean;
            if (var < 0) var = 0;
            gmean[g] = (float)mean;
            grstd[g] = (float)(1.0 / sqrt(var + (double)eps));
        }
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const int g = c / gs;
            const float sc = grstd[g] * gamma[c];
            ss[c] = half * sc;
            ss[C + c] = half * (beta[c] - gmean[g] * sc);
        }
        __syncthreads();
        float sc[8], sh[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sc[j] = ss[v * 8 + j];
            sh[j] = ss[C + v * 8 + j];
        }
        for (int p = p0 + pl; p < p1; p += U * ppb) {
            uint4 u[U];
#pragma unroll
            for (int k = 0; k < U; ++k)
                if (p + k * ppb < p1) {
                    const uint4* sp = reinterpret_cast<const uint4*>(src_ptr(s, img, p + k * ppb, HW, v * 8));
                    u[k] = STREAM ? ldg_stream(sp) : __ldg(sp);
                }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                if (p + k * ppb >= p1) break;
                const long long o = ((long long)img * HW + p + k * ppb) * C + v * 8;
                if (raw_dst) {
                    if (STREAM) stg_stream(reinterpret_cast<uint4*>(raw_dst + o), u[k]);
                    else *reinterpret_cast<uint4*>(raw_dst + o) = u[k];
                }
                float f[8];
                unpack8(u[k], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float h = fmaf(f[j], sc[j], sh[j]);
                    f[j] = silu ? fmaf(h, tanh_approx(h), h) : h;
                }
                if (STREAM) stg_stream(reinterpret_cast<uint4*>(dst + o), pack8(f));
                else *reinterpret_cast<uint4*>(dst + o) = pack8(f);
            }
        }
    }
}

int launch_gn_apply(const GnSrc& s, int B, int HW, const float* gamma, const float* beta, float eps,
                    int groups, int silu, act16* dst, act16* raw_dst, cudaStream_t st) {
    const int C = s.C1 + s.C2;
    PNPF_REQUIRE(C % groups == 0 && C % 8 == 0 && s.C1 % 8 == 0, "GroupNorm channels (%d+%d) unsupported", s.C1, s.C2);
    const int threads = gn_threads(C);
    PNPF_REQUIRE(s.st1 && (s.C2 == 0 || s.st2), "GroupNorm apply without statistics");
    const size_t smem = (2 * C + 2 * groups) * sizeof(float);
    // one wave: resident CTAs per SM for this block size (register-limited: 3 at 256 threads) x SMs
    static DeviceCache cache[9];                      // per block-size class (threads / 32)
    int per_sm = 0;
    DeviceCache& dc = cache[(threads / 32) % 9];
    if (!dc.lookup(&per_sm)) {
        PNPF_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gn_apply_kernel<8, true>, threads, smem));
        if (per_sm < 1) per_sm = 1;
        dc.store(per_sm);
    }
    const long long total_pix = (long long)B * HW;
    long long grid = (long long)per_sm * num_sms();
    const long long max_grid = (total_pix + 63) / 64;         // at least 64 pixels per CTA
    if (grid > max_grid) grid = max_grid;
    if (grid < 1) grid = 1;
    // 8 independent 16-byte loads per thread in flight + streaming cache hints: 13 % faster than 4 loads / default caching
    // (same-box sweep, profiles/r01_ab_experiments.md)
    gn_apply_kernel<8, true><<<(unsigned)grid, threads, smem, st>>>(s, HW, total_pix, gamma, beta, eps, C / groups, silu, dst, raw_dst);
    PNPF_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// =================================================================================================
// row softmax: one warp per row, fp32 in, fp16 out
// =================================================================================================
__global__ void softmax_rows_kernel(const float* __restrict__ S, act16* __restrict__ P, long long rows, int L) {
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int lane = threadIdx.x & 31;
    const float* s = S + row * L;
    float m = -INFINITY;
    for (int i = lane * 4; i < L; i += 128) {
        const float4 v = *reinterpret_cast<const float4*>(s + i);
        m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int i = lane * 4; i < L; i += 128) {
        const float4 v = *reinterpret_cast<const float4*>(s + i);
        sum += __expf(v.x - m) + __expf(v.y - m) + __expf(v.z - m) + __expf(v.w - m);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.f / sum;
    act16* p = P + row * L;
    for (int i = lane * 4; i < L; i += 128) {
        const float4 v = *reinterpret_cast<const float4*>(s + i);
        const uint32_t a = pack2(__expf(v.x - m) * inv, __expf(v.y - m) * inv);
        const uint32_t b = pack2(__expf(v.z - m) * inv, __expf(v.w - m) * inv);
        uint2 o2;
        o2.x = a;
        o2.y = b;
        *reinterpret_cast<uint2*>(p + i) = o2;
    }
}

int launch_softmax_rows(const float* S, act16* P, long long rows, int L, cudaStream_t st) {
    PNPF_REQUIRE(L % 4 == 0, "softmax length %d must be a multiple of 4", L);
    const int wpb = 8;
    softmax_rows_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, st>>>(S, P, rows, L);
    PNPF_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// =================================================================================================
// time embedding: sincos -> dense -> swish -> dense -> swish -> all temb_proj layers
// =================================================================================================
__device__ __forceinline__ float swishf(float x) { return x / (1.f + expf(-x)); }

constexpr int TEMB_IMGS = 8;                          // images per block: every weight load serves 8 images
__global__ void temb_kernel(TembWeights w, const float* __restrict__ t, int B, float* __restrict__ out) {
    // grid (group of TEMB_IMGS images, chunk of 256 projection outputs): every block recomputes the two small dense layers of
    // its images (20K MAC each); per image the arithmetic and its order are those of a one-image block
    extern __shared__ float sm[];                     // per image: emb[ch] | h[temb_ch] | s[temb_ch]
    const int per = w.ch + 2 * w.temb_ch;
    const int img0 = blockIdx.x * TEMB_IMGS;
    const int half = w.ch / 2;
    for (int i = threadIdx.x; i < TEMB_IMGS * w.ch; i += blockDim.x) {
        const int g = i / w.ch, c = i - g * w.ch;
        const float tt = t[min(img0 + g, B - 1)];
        const float a = tt * w.freqs[c % half];
        sm[g * per + c] = (c < half) ? sinf(a) : cosf(a);
    }
    __syncthreads();
    for (int o = threadIdx.x; o < w.temb_ch; o += blockDim.x) {
        float acc[TEMB_IMGS];
#pragma unroll
        for (int g = 0; g < TEMB_IMGS; ++g) acc[g] = w.b0[o];
        for (int k = 0; k < w.ch; ++k) {
            const float wv = __ldg(w.w0_t + k * w.temb_ch + o);
#pragma unroll
            for (int g = 0; g < TEMB_IMGS; ++g) acc[g] = fmaf(wv, sm[g * per + k], acc[g]);
        }
#pragma unroll
        for (int g = 0; g < TEMB_IMGS; ++g) sm[g * per + w.ch + o] = swishf(acc[g]);
    }
    __syncthreads();
    for (int o = threadIdx.x; o < w.temb_ch; o += blockDim.x) {
        float acc[TEMB_IMGS];
#pragma unroll
        for (int g = 0; g < TEMB_IMGS; ++g) acc[g] = w.b2[o];
        for (int k = 0; k < w.temb_ch; ++k) {
            const float wv = __ldg(w.w2_t + k * w.temb_ch + o);
#pragma unroll
            for (int g = 0; g < TEMB_IMGS; ++g) acc[g] = fmaf(wv, sm[g * per + w.ch + k], acc[g]);
        }
#pragma unroll
        for (int g = 0; g < TEMB_IMGS; ++g)           // ResidualBlock applies act(temb) before temb_proj (models.py:101)
            sm[g * per + w.ch + w.temb_ch + o] = swishf(acc[g]);
    }
    __syncthreads();
    const int o = blockIdx.y * blockDim.x + threadIdx.x;
    if (o < w.total_proj) {
        float acc[TEMB_IMGS];
#pragma unroll
        for (int g = 0; g < TEMB_IMGS; ++g) acc[g] = w.bp[o];
#pragma unroll 4
        for (int k = 0; k < w.temb_ch; ++k) {
            const float wv = __ldg(w.wp_t + (long long)k * w.total_proj + o);
#pragma unroll
            for (int g = 0; g < TEMB_IMGS; ++g) acc[g] = fmaf(wv, sm[g * per + w.ch + w.temb_ch + k], acc[g]);
        }
#pragma unroll
        for (int g = 0; g < TEMB_IMGS; ++g)
            if (img0 + g < B) out[(long long)(img0 + g) * w.total_proj + o] = acc[g];
    }
}

int launch_temb(const TembWeights& w, const float* t, int B, float* out, cudaStream_t st) {
    const size_t smem = (size_t)TEMB_IMGS * (w.ch + 2 * w.temb_ch) * sizeof(float);
    dim3 grid((B + TEMB_IMGS - 1) / TEMB_IMGS, (w.total_proj + 255) / 256);
    temb_kernel<<<grid, 256, smem, st>>>(w, t, B, out);
    PNPF_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// =================================================================================================
// layout shims
// =================================================================================================
__global__ void nchw_to_nhwc_pad_kernel(const float* __restrict__ x, int C, int HW, act16* __restrict__ dst, int Cpad,
                                        long long total_pix) {
    const long long gp = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gp >= total_pix) return;
    const long long img = gp / HW, p = gp - img * HW;
    act16* o = dst + gp * Cpad;
    for (int c0 = 0; c0 < Cpad; c0 += 8) {
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = (c0 + j < C) ? __ldg(x + (img * C + c0 + j) * HW + p) : 0.f;
        *reinterpret_cast<uint4*>(o + c0) = pack8(f);
    }
}
int launch_nchw_to_nhwc_pad(const float* x, int B, int C, int HW, act16* dst, int Cpad, cudaStream_t st) {
    const long long n = (long long)B * HW;
    nchw_to_nhwc_pad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, C, HW, dst, Cpad, n);
    PNPF_CHECK_CUDA(cudaGetLastError());
    return 0;
}

__global__ void nhwc_to_nchw_f32_kernel(const act16* __restrict__ src, long long pitch, int C, int HW, float* __restrict__ dst,
                                        long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over [B][C][HW]
    if (i >= total) return;
    const long long p = i % HW;
    const long long bc = i / HW;
    const long long c = bc % C, b = bc / C;
    dst[i] = act16_to_float(src[(b * HW + p) * pitch + c]);
}
int launch_nhwc_to_nchw_f32(const act16* src, long long pitch, int B, int C, int HW, float* dst, cudaStream_t st) {
    const long long n = (long long)B * C * HW;
    nhwc_to_nchw_f32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, pitch, C, HW, dst, n);
    PNPF_CHECK_CUDA(cudaGetLastError());
    return 0;
}

__global__ void upsample2x_kernel(const uint4* __restrict__ src, int H, int W, int Cv, uint4* __restrict__ dst, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over [B][2H][2W][Cv]
    if (i >= total) return;
    const int cv = (int)(i % Cv);
    long long r = i / Cv;
    const int ow = (int)(r % (2 * W));
    r /= (2 * W);
    const int oh = (int)(r % (2 * H));
    const long long b = r / (2 * H);
    dst[i] = __ldg(src + ((b * H + (oh >> 1)) * W + (ow >> 1)) * Cv + cv);
}
int launch_upsample2x(const act16* src, int B, int H, int W, int C, act16* dst, cudaStream_t st) {
    PNPF_REQUIRE(C % 8 == 0, "upsample channels %d", C);
    const long long n = (long long)B * 4 * H * W * (C / 8);
    upsample2x_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(src), H, W, C / 8,
                                                                   reinterpret_cast<uint4*>(dst), n);
    PNPF_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// =================================================================================================
// PnP-Flow per-pixel kernels
// =================================================================================================
__device__ __forceinline__ int floor_div(int a, int b) {      // b > 0
    const int q = a / b;
    return (a % b != 0 && a < 0) ? q - 1 : q;
}
__device__ __forceinline__ bool keep_pixel(const OpDesc& op, int b, int h, int w, int H, int W) {
    if (op.kind == 1) {                               // box: zero the square [d-hs, d+hs)^2, d = H//2 (utils.py:331-335)
        const int d = H / 2;
        const bool in = (h >= d - op.half_size) && (h < d + op.half_size) && (w >= d - op.half_size) && (w < d + op.half_size);
        return !in;
    }
    if (op.kind == 2) return op.mask[((long long)b * H + h) * W + w] != 0;
    return true;
}

// H / H_adj for identity, box, mask (y = m*x) and SR (gather / zero-fill scatter)
__global__ void apply_diag_kernel(OpDesc op, const float* __restrict__ x, float* __restrict__ y, int C, int H, int W,
                                  long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int w = (int)(i % W);
    long long r = i / W;
    const int h = (int)(r % H);
    r /= H;
    const int b = (int)(r / C);
    y[i] = keep_pixel(op, b, h, w, H, W) ? x[i] : 0.f * x[i];
}
__global__ void sr_down_kernel(const float* __restrict__ x, float* __restrict__ y, int sf, int H, int W, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over [BC][H/sf][W/sf]
    if (i >= total) return;
    const int Ws = W / sf, Hs = H / sf;
    const int w = (int)(i % Ws);
    long long r = i / Ws;
    const int h = (int)(r % Hs);
    const long long bc = r / Hs;
    y[i] = x[(bc * H + (long long)h * sf) * W + (long long)w * sf];
}
__global__ void sr_up_kernel(const float* __restrict__ y, float* __restrict__ x, int sf, int H, int W, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // over [BC][H][W]
    if (i >= total) return;
    const int w = (int)(i % W);
    long long r = i / W;
    const int h = (int)(r % H);
    const long long bc = r / H;
    x[i] = ((h % sf) == 0 && (w % sf) == 0) ? y[(bc * (H / sf) + h / sf) * (W / sf) + w / sf] : 0.f;
}

// Separable circular Gaussian blur of one (b,c) plane tile.  mode 0: out = G*in ; 1: out = G*in - aux ; 2: out = aux - gamma*(G*in)
constexpr int BLUR_TILE = 32;
__global__ void blur_kernel(const float* __restrict__ in, const float* __restrict__ aux, float* __restrict__ out,
                            const float* __restrict__ taps, int ksize, int H, int W, int mode, float gamma) {
    extern __shared__ float sm[];
    const int R = (ksize - 1) / 2;
    const int HT = BLUR_TILE + 2 * R;                 // halo tile side
    float* tile = sm;                                 // [HT][HT]
    float* tmp = tile + HT * HT;                      // [HT][BLUR_TILE]  (row pass result)
    float* g = tmp + HT * BLUR_TILE;                  // [ksize]
    const long long plane = blockIdx.z;
    const int h0 = blockIdx.y * BLUR_TILE, w0 = blockIdx.x * BLUR_TILE;
    const float* src = in + plane * H * W;
    for (int i = threadIdx.x; i < ksize; i += blockDim.x) g[i] = taps[i];
    for (int i = threadIdx.x; i < HT * HT; i += blockDim.x) {
        const int th = i / HT, tw = i - th * HT;
        int hh = (h0 + th - R) % H; if (hh < 0) hh += H;
        int ww = (w0 + tw - R) % W; if (ww < 0) ww += W;
        tile[i] = src[(long long)hh * W + ww];
    }
    __syncthreads();
    // out[i] = sum_k g[k] * x[i - (k - R)]  (circular; degradations.py:62-68,78-79)
    for (int i = threadIdx.x; i < HT * BLUR_TILE; i += blockDim.x) {
        const int th = i / BLUR_TILE, tw = i - th * BLUR_TILE;
        float acc = 0.f;
        for (int k = 0; k < ksize; ++k) acc = fmaf(g[k], tile[th * HT + (tw + 2 * R - k)], acc);
        tmp[i] = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < BLUR_TILE * BLUR_TILE; i += blockDim.x) {
        const int th = i / BLUR_TILE, tw = i - th * BLUR_TILE;
        const int h = h0 + th, w = w0 + tw;
        if (h >= H || w >= W) continue;
        float acc = 0.f;
        for (int k = 0; k < ksize; ++k) acc = fmaf(g[k], tmp[(th + 2 * R - k) * BLUR_TILE + tw], acc);
        const long long o = plane * H * W + (long long)h * W + w;
        if (mode == 1) acc = acc - aux[o];
        else if (mode == 2) acc = aux[o] - gamma * acc;
        else if (mode == 3) acc = (acc - aux[o] > 0.f) ? 1.f : -1.f;      // laplace: 2*heaviside(Gx - y, 0) - 1 (pnp_flow.py:43)
        out[o] = acc;
    }
}
static int launch_blur(const OpDesc& op, const float* in, const float* aux, float* out, int planes, int H, int W, int mode,
                       float gamma, cudaStream_t st) {
    PNPF_REQUIRE(op.taps && op.ksize % 2 == 1 && op.ksize <= H && op.ksize <= W, "blur kernel size %d vs image %dx%d", op.ksize, H, W);
    const int R = (op.ksize - 1) / 2, HT = BLUR_TILE + 2 * R;
    const size_t smem = ((size_t)HT * HT + (size_t)HT * BLUR_TILE + op.ksize) * sizeof(float);
    static DeviceCache cache;
    int dummy = 0;
    if (!cache.lookup(&dummy)) {
        PNPF_CHECK_CUDA(cudaFuncSetAttribute(blur_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        cache.store(1);
    }
    PNPF_REQUIRE(smem <= 160 * 1024, "blur kernel too large for shared memory");
    dim3 grid((W + BLUR_TILE - 1) / BLUR_TILE, (H + BLUR_TILE - 1) / BLUR_TILE, planes);
    blur_kernel<<<grid, 256, smem, st>>>(in, aux, out, op.taps, op.ksize, H, W, mode, gamma);
    PNPF_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// ---- Superresolution(mode='bicubic') (degradations.py:97-109,117-127; utils.py:365-396) -------------------------------------
// The reference filters with the 4 sf x 4 sf bicubic kernel through the FFT (circular, the filter image rolled by -(K-1)//2 =
// -K/2), then decimates; the adjoint zero-fills and correlates.  The kernel is separable, k = outer(g, g) with g = w / sum(w),
// so both directions are direct sums here:
//   H:      y[i, j]  = sum_{m1, m2} g[m1] g[m2] x[(sf i + K/2 - m1) mod H, (sf j + K/2 - m2) mod W]      K^2 MACs per LOW-res pixel
//   H_adj:  u[n1,n2] = sum_{i, j}   g[sf i - n1 + K/2] g[sf j - n2 + K/2] r[i mod Hs, j mod Ws]          (K/sf)^2 = 16 MACs per pixel
// mode 0: out = Hx ; 1: out = Hx - aux ; 3: out = sign(Hx - aux)   (aux = y, low resolution)
__global__ void sr_bicubic_down_kernel(const float* __restrict__ x, const float* __restrict__ aux, float* __restrict__ out,
                                       const float* __restrict__ taps, int K, int sf, int H, int W, int mode) {
    extern __shared__ float g[];
    for (int i = threadIdx.x; i < K; i += blockDim.x) g[i] = taps[i];
    __syncthreads();
    const int Hs = H / sf, Ws = W / sf;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;            // low-resolution pixel inside the plane
    if (idx >= Hs * Ws) return;
    const long long plane = blockIdx.y;
    const int i = idx / Ws, j = idx - i * Ws;
    const float* src = x + plane * H * W;
    const int o = K / 2;
    float acc = 0.f;
    for (int m1 = 0; m1 < K; ++m1) {
        int hh = (sf * i + o - m1) % H; if (hh < 0) hh += H;
        const float* row = src + (long long)hh * W;
        float racc = 0.f;
        for (int m2 = 0; m2 < K; ++m2) {
            int ww = (sf * j + o - m2) % W; if (ww < 0) ww += W;
            racc = fmaf(g[m2], __ldg(row + ww), racc);
        }
        acc = fmaf(g[m1], racc, acc);
    }
    const long long oi = plane * Hs * Ws + idx;
    if (mode == 1) acc = acc - aux[oi];
    else if (mode == 3) acc = (acc - aux[oi] > 0.f) ? 1.f : -1.f;     // laplace: 2*heaviside(Hx - y, 0) - 1 (pnp_flow.py:43)
    out[oi] = acc;
}
// mode 0: out = H^T r ; 2: out = aux - gamma * H^T r   (aux = x, full resolution)
__global__ void sr_bicubic_up_kernel(const float* __restrict__ r, const float* __restrict__ aux, float* __restrict__ out,
                                     const float* __restrict__ taps, int K, int sf, int H, int W, int mode, float gamma) {
    extern __shared__ float g[];
    for (int i = threadIdx.x; i < K; i += blockDim.x) g[i] = taps[i];
    __syncthreads();
    const int Hs = H / sf, Ws = W / sf;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;            // full-resolution pixel inside the plane
    if (idx >= H * W) return;
    const long long plane = blockIdx.y;
    const int n1 = idx / W, n2 = idx - n1 * W;
    const float* src = r + plane * Hs * Ws;
    const int o = K / 2;
    // low-resolution rows i with 0 <= sf i - n1 + o < K  (i may be negative or >= Hs: circular)
    const int i0 = -floor_div(o - n1, sf), j0 = -floor_div(o - n2, sf);      // ceil((n - o) / sf)
    float acc = 0.f;
    for (int i = i0; sf * i - n1 + o < K; ++i) {
        int ii = i % Hs; if (ii < 0) ii += Hs;
        const float* row = src + (long long)ii * Ws;
        float racc = 0.f;
        for (int j = j0; sf * j - n2 + o < K; ++j) {
            int jj = j % Ws; if (jj < 0) jj += Ws;
            racc = fmaf(g[sf * j - n2 + o], __ldg(row + jj), racc);
        }
        acc = fmaf(g[sf * i - n1 + o], racc, acc);
    }
    const long long oi = plane * H * W + idx;
    out[oi] = mode == 2 ? __fsub_rn(aux[oi], __fmul_rn(gamma, acc)) : acc;
}
static int launch_sr_bicubic(const OpDesc& op, bool up, const float* in, const float* aux, float* out, int planes, int H, int W, int mode,
                             float gamma, cudaStream_t st) {
    PNPF_REQUIRE(op.taps && op.sf >= 1 && op.ksize == 4 * op.sf && H % op.sf == 0 && W % op.sf == 0 && op.ksize <= H / 1 && op.ksize <= W,
                 "bicubic SR: factor %d, %d taps vs image %dx%d", op.sf, op.ksize, H, W);
    PNPF_REQUIRE(planes <= 65535, "bicubic SR: too many planes (%d)", planes);
    const int n = up ? H * W : (H / op.sf) * (W / op.sf);
    dim3 grid((n + 255) / 256, planes);
    if (up) sr_bicubic_up_kernel<<<grid, 256, op.ksize * sizeof(float), st>>>(in, aux, out, op.taps, op.ksize, op.sf, H, W, mode, gamma);
    else sr_bicubic_down_kernel<<<grid, 256, op.ksize * sizeof(float), st>>>(in, aux, out, op.taps, op.ksize, op.sf, H, W, mode);
    PNPF_CHECK_CUDA(cudaGetLastError());
    return 0;
}

int launch_apply_H(const OpDesc& op, const float* x, float* y, int B, int C, int H, int W, bool adjoint, cudaStream_t st) {
    const long long n = (long long)B * C * H * W;
    if (op.kind == 3) {
        PNPF_REQUIRE(op.sf >= 1 && H % op.sf == 0 && W % op.sf == 0, "SR factor %d vs %dx%d", op.sf, H, W);
        if (!adjoint) {
            const long long m = n / (op.sf * op.sf);
            sr_down_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(x, y, op.sf, H, W, m);
        } else {
            sr_up_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, y, op.sf, H, W, n);
        }
    } else if (op.kind == 4) {
        return launch_blur(op, x, nullptr, y, B * C, H, W, 0, 0.f, st);
    } else if (op.kind == 5) {
        return launch_sr_bicubic(op, adjoint, x, nullptr, y, B * C, H, W, 0, 0.f, st);
    } else {
        PNPF_REQUIRE(op.kind != 2 || op.mask, "mask operator without a device mask");
        apply_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(op, x, y, C, H, W, n);
    }
    PNPF_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// z = x - gamma * A^T(Ax - y) for the diagonal operators and SR, one pass (12 N bytes)
// residual of the data term: gaussian  r = Ax - y ;  laplace  r = 2*heaviside(Ax - y, 0) - 1  (pnp_flow.py:41,43)
__device__ __forceinline__ float datafit_residual(float ax, float yv, int laplace) {
    const float d = __fsub_rn(ax, yv);
    return laplace ? ((d > 0.f) ? 1.f : -1.f) : d;
}

__global__ void datafit_diag_kernel(OpDesc op, const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ z,
                                    float gamma, int laplace, int C, int H, int W, long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int w = (int)(i % W);
    long long r = i / W;
    const int h = (int)(r % H);
    r /= H;                                           // r = b*C + c
    const float xv = x[i];
    float out = xv;
    if (op.kind == 3) {
        if ((h % op.sf) == 0 && (w % op.sf) == 0) {
            const float yv = y[(r * (H / op.sf) + h / op.sf) * (W / op.sf) + w / op.sf];
            out = __fsub_rn(xv, __fmul_rn(gamma, datafit_residual(xv, yv, laplace)));
        }
    } else {
        const int b = (int)(r / C);
        if (keep_pixel(op, b, h, w, H, W)) out = __fsub_rn(xv, __fmul_rn(gamma, datafit_residual(xv, y[i], laplace)));
    }
    z[i] = out;
}

// 16-byte version (W % 4 == 0): a thread owns four consecutive pixels of one (b, c) plane row; grid.y = plane, so all index
// arithmetic is 32-bit.  Same separately rounded operations per element as the scalar kernel (bit-identical results).
__global__ void datafit_diag_vec4_kernel(OpDesc op, const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ z,
                                         float gamma, int laplace, int C, int H, int W) {
    const int W4 = W >> 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;            // vector index inside the plane
    if (idx >= H * W4) return;
    const int plane = blockIdx.y;
    const int h = idx / W4, w = (idx - h * W4) << 2;
    const long long base = ((long long)plane * H + h) * W + w;
    const float4 xv = __ldg(reinterpret_cast<const float4*>(x + base));
    float xs[4] = {xv.x, xv.y, xv.z, xv.w}, out[4] = {xv.x, xv.y, xv.z, xv.w};
    if (op.kind == 3) {
        if ((h % op.sf) == 0) {
            const float* yrow = y + ((long long)plane * (H / op.sf) + h / op.sf) * (W / op.sf);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (((w + e) % op.sf) == 0)
                    out[e] = __fsub_rn(xs[e], __fmul_rn(gamma, datafit_residual(xs[e], __ldg(yrow + (w + e) / op.sf), laplace)));
        }
    } else {
        const float4 yv = __ldg(reinterpret_cast<const float4*>(y + base));
        const float ys[4] = {yv.x, yv.y, yv.z, yv.w};
        bool keep[4] = {true, true, true, true};
        if (op.kind == 1) {
            const int d = H / 2;
            const bool hin = (h >= d - op.half_size) && (h < d + op.half_size);
#pragma unroll
            for (int e = 0; e < 4; ++e) keep[e] = !(hin && (w + e >= d - op.half_size) && (w + e < d + op.half_size));
        } else if (op.kind == 2) {
            const uchar4 m = __ldg(reinterpret_cast<const uchar4*>(op.mask + ((long long)(plane / C) * H + h) * W + w));
            keep[0] = m.x != 0; keep[1] = m.y != 0; keep[2] = m.z != 0; keep[3] = m.w != 0;
        }
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (keep[e]) out[e] = __fsub_rn(xs[e], __fmul_rn(gamma, datafit_residual(xs[e], ys[e], laplace)));
    }
    *reinterpret_cast<float4*>(z + base) = make_float4(out[0], out[1], out[2], out[3]);
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int launch_datafit(const OpDesc& op, const float* x, const float* y, float* z, float gamma, int laplace, int B, int C, int H, int W,
                   cudaStream_t st) {
    const long long n = (long long)B * C * H * W;
    if (op.kind == 4) {
        PNPF_REQUIRE(op.scratch, "blur data-fidelity step needs pnpf_operator.scratch");
        if (int e = launch_blur(op, x, y, op.scratch, B * C, H, W, laplace ? 3 : 1, 0.f, st)) return e;   // r = Gx - y (or its sign)
        return launch_blur(op, op.scratch, x, z, B * C, H, W, 2, gamma, st);                 // z = x - gamma G r
    }
    if (op.kind == 5) {
        PNPF_REQUIRE(op.scratch, "bicubic SR data-fidelity step needs pnpf_operator.scratch (B*C*H*W/sf^2 floats)");
        if (int e = launch_sr_bicubic(op, false, x, y, op.scratch, B * C, H, W, laplace ? 3 : 1, 0.f, st)) return e;   // r = Hx - y (or its sign)
        return launch_sr_bicubic(op, true, op.scratch, x, z, B * C, H, W, 2, gamma, st);                               // z = x - gamma H^T r
    }
    PNPF_REQUIRE(op.kind >= 0 && op.kind <= 3, "unknown operator kind %d", op.kind);
    PNPF_REQUIRE(op.kind != 2 || op.mask, "mask operator without a device mask");
    PNPF_REQUIRE(op.kind != 3 || (op.sf >= 1 && H % op.sf == 0 && W % op.sf == 0), "SR factor %d vs %dx%d", op.sf, H, W);
    const bool vec = (W % 4 == 0) && aligned16(x) && aligned16(z) && (op.kind == 3 || aligned16(y)) &&
                     (op.kind != 2 || (reinterpret_cast<uintptr_t>(op.mask) & 3) == 0) && (long long)B * C <= 65535;
    if (vec) {
        dim3 grid((unsigned)((H * (W / 4) + 255) / 256), (unsigned)(B * C));
        datafit_diag_vec4_kernel<<<grid, 256, 0, st>>>(op, x, y, z, gamma, laplace, C, H, W);
    } else {
        datafit_diag_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(op, x, y, z, gamma, laplace, C, H, W, n);
    }
    PNPF_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// zt[s] = t*z + eps[s]*(1-t): three separately rounded ops like the eager reference (pnp_flow.py:48).
// grid.y = draw s (no 64-bit modulo); the 16-byte version handles n % 4 == 0, the scalar one everything else.
__global__ void interp_kernel(const float* __restrict__ z, const float* __restrict__ eps, float t, float omt,
                              float* __restrict__ zt, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long o = (long long)blockIdx.y * n + i;
    if (i < n) zt[o] = __fadd_rn(__fmul_rn(t, __ldg(z + i)), __fmul_rn(eps[o], omt));
}
__device__ __forceinline__ float4 ldg_stream_f4(const float* p) {       // read-once data: do not allocate in L1
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__global__ void interp_vec4_kernel(const float* __restrict__ z, const float* __restrict__ eps, float t, float omt,
                                   float* __restrict__ zt, long long n4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const long long o = ((long long)blockIdx.y * n4 + i) * 4;
    const float4 a = __ldg(reinterpret_cast<const float4*>(z) + i);      // z is re-read by every draw: keep it cached
    const float4 e = ldg_stream_f4(eps + o);
    float4 r;
    r.x = __fadd_rn(__fmul_rn(t, a.x), __fmul_rn(e.x, omt));
    r.y = __fadd_rn(__fmul_rn(t, a.y), __fmul_rn(e.y, omt));
    r.z = __fadd_rn(__fmul_rn(t, a.z), __fmul_rn(e.z, omt));
    r.w = __fadd_rn(__fmul_rn(t, a.w), __fmul_rn(e.w, omt));
    *reinterpret_cast<float4*>(zt + o) = r;
}
int launch_interp(const float* z, const float* eps, float t, float* zt, long long n, int S, cudaStream_t st) {
    PNPF_REQUIRE(S >= 1 && S <= 65535, "num_samples %d", S);
    if (n % 4 == 0 && aligned16(z) && aligned16(eps) && aligned16(zt)) {
        const long long n4 = n / 4;
        dim3 grid((unsigned)((n4 + 255) / 256), (unsigned)S);
        interp_vec4_kernel<<<grid, 256, 0, st>>>(z, eps, t, 1.0f - t, zt, n4);
    } else {
        dim3 grid((unsigned)((n + 255) / 256), (unsigned)S);
        interp_kernel<<<grid, 256, 0, st>>>(z, eps, t, 1.0f - t, zt, n);
    }
    PNPF_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// x_new = (sum_s (zt_s + (1-t) v_s)) / S, summed in draw order with separately rounded ops (pnp_flow.py:114-121: x_new += ...;
// x_new /= num_samples).  The final "/= S" is a MULTIPLICATION by fp32(1/S): that is what the reference's GPU path executes —
// ATen's CUDA true-divide kernel turns division by a CPU scalar into `a * (1 / b)` (BinaryDivTrueKernel.cu), only the CPU kernel
// divides.  Verified on the B200: with __fdiv_rn the result differs from eager torch-CUDA in the last bit, with the reciprocal it
// is torch.equal (tests/test_gpu_pnp.py).
__global__ void push_accum_kernel(const float* __restrict__ zt, const float* __restrict__ v, float omt, int S, float invS,
                                  float* __restrict__ x_new, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc = __fadd_rn(acc, __fadd_rn(zt[s * n + i], __fmul_rn(omt, v[s * n + i])));
    x_new[i] = __fmul_rn(acc, invS);
}
__global__ void push_accum_vec4_kernel(const float* __restrict__ zt, const float* __restrict__ v, float omt, int S, float invS,
                                       float* __restrict__ x_new, long long n4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 5
    for (int s = 0; s < S; ++s) {
        const float4 a = ldg_stream_f4(zt + ((long long)s * n4 + i) * 4);
        const float4 b = ldg_stream_f4(v + ((long long)s * n4 + i) * 4);
        acc.x = __fadd_rn(acc.x, __fadd_rn(a.x, __fmul_rn(omt, b.x)));
        acc.y = __fadd_rn(acc.y, __fadd_rn(a.y, __fmul_rn(omt, b.y)));
        acc.z = __fadd_rn(acc.z, __fadd_rn(a.z, __fmul_rn(omt, b.z)));
        acc.w = __fadd_rn(acc.w, __fadd_rn(a.w, __fmul_rn(omt, b.w)));
    }
    *reinterpret_cast<float4*>(x_new + i * 4) = make_float4(__fmul_rn(acc.x, invS), __fmul_rn(acc.y, invS), __fmul_rn(acc.z, invS), __fmul_rn(acc.w, invS));
}
int launch_push_accum(const float* zt, const float* v, float t, int S, float* x_new, long long n, cudaStream_t st) {
    PNPF_REQUIRE(S >= 1, "num_samples %d", S);
    if (n % 4 == 0 && aligned16(zt) && aligned16(v) && aligned16(x_new)) {
        const long long n4 = n / 4;
        push_accum_vec4_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(zt, v, 1.0f - t, S, 1.0f / (float)S, x_new, n4);
    } else {
        push_accum_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(zt, v, 1.0f - t, S, 1.0f / (float)S, x_new, n);
    }
    PNPF_CHECK_CUDA(cudaGetLastError());
    return 0;
}

// out = x + a * v with separately rounded product and sum (one fixed-grid Euler step  y1 = y0 + dt * f(t0, y0))
__global__ void axpy_kernel(const float* __restrict__ x, const float* __restrict__ v, float a, float* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = __fadd_rn(x[i], __fmul_rn(a, v[i]));
}
__global__ void axpy_vec4_kernel(const float* __restrict__ x, const float* __restrict__ v, float a, float* __restrict__ out, long long n4) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    const float4 p = ldg_stream_f4(x + i * 4), q = ldg_stream_f4(v + i * 4);
    *reinterpret_cast<float4*>(out + i * 4) = make_float4(__fadd_rn(p.x, __fmul_rn(a, q.x)), __fadd_rn(p.y, __fmul_rn(a, q.y)),
                                                          __fadd_rn(p.z, __fmul_rn(a, q.z)), __fadd_rn(p.w, __fmul_rn(a, q.w)));
}
int launch_axpy(const float* x, const float* v, float a, float* out, long long n, cudaStream_t st) {
    if (n % 4 == 0 && aligned16(x) && aligned16(v) && aligned16(out)) axpy_vec4_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, st>>>(x, v, a, out, n / 4);
    else axpy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, v, a, out, n);
    PNPF_CHECK_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace pnpf
