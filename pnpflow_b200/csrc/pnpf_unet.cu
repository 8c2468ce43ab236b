// U-Net engine: layer plan (same walk as the reference UNet.__init__/forward, models.py:302-495), weight registry and
// repacking, workspace layout, and the launch sequence of one velocity evaluation.
#include "../../include/pnpflow_b200.h"
#include "pnpf_kernels.cuh"
#include "pnpf_ops.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

using namespace pnpf;

namespace {

constexpr int CIN_PAD = 32;       // K chunk of begin_conv: the image channels are zero-padded to 32 (weights) ...
// ... but the bf16/fp16 NHWC copy of the network input only STORES cin_store() channels per pixel (8: one 16-byte vector); the TMA
// unit zero-fills channels [cin_store, 32) of every box (ConvDesc::x_cvalid), so the shim writes and begin_conv reads 16 instead of
// 64 bytes per pixel (335 -> 84 MB per evaluation at 256^2, batch 80)
static int cin_store(const pnpf_unet_config& c) { return c.input_channels <= 8 ? 8 : 16; }
constexpr int GROUPS = 32;        // models.py:33-38
constexpr float GN_EPS = 1e-6f;

struct LayerSpec {
    enum Kind { CONV, RES, ATTN, DOWN, UP, END } kind;
    std::string prefix;
    int in_ch, out_ch, side, skip_ch;
    bool push;
};

struct WEntry {
    std::string name;
    std::vector<int64_t> shape;
    std::vector<float> data;
    bool loaded = false;
};

struct Act {                      // fp16 NHWC activation [B][side][side][C]
    act16* p = nullptr;
    int C = 0, side = 0;
    double* stats = nullptr;      // [B][C][2] per-channel (sum, sumsq) written by the producing conv's epilogue
    long long elems_per_img() const { return (long long)C * side * side; }
};

struct Op {
    enum Kind { MEMSET, IN_SHIM, TEMB, TC, GN_STATS, GN_APPLY, SOFTMAX, UPSAMPLE, ATTN } kind;
    std::string name;
    TcOp tc;
    AttnOp attn;                  // kind == ATTN: fused attention core (pnpf_attn.cuh)
    GnSrc gsrc{};
    int HW = 0;
    double* stats = nullptr;
    const float* gamma = nullptr;
    const float* beta = nullptr;
    int silu = 0;
    act16* dst = nullptr;
    act16* raw_dst = nullptr;
    const float* S = nullptr;     // softmax
    act16* P = nullptr;
    long long rows_per_img = 0;
    int L = 0;
    const act16* up_src = nullptr; // upsample
    int up_side = 0, up_C = 0;
    std::string impl;             // kernel that runs the op (pnpf_debug_op_impl)
    double flops = 0;             // algorithmic 2*MAC per image (tensor-core ops)
    double bytes = 0;             // algorithmic HBM bytes per image: every operand read once + every output written once
    // debug view of the output (fp16 NHWC), C == 0 -> not readable
    const act16* out_p = nullptr;
    int out_C = 0, out_side = 0;
    long long out_pitch = 0;
};

}  // namespace

struct pnpf_engine {
    pnpf_unet_config cfg;
    std::vector<LayerSpec> layers;
    std::vector<WEntry> weights;
    std::map<std::string, int> windex;
    // device weights
    uint8_t* wdev = nullptr;
    size_t wdev_bytes = 0;
    std::map<std::string, size_t> woff;      // packed item name -> byte offset in wdev
    bool finalized = false;
    int total_proj = 0;
    std::map<std::string, int> proj_off;     // resblock prefix -> offset in the temb projection table
    // plan
    int max_batch = 0;
    uint8_t* ws = nullptr;
    size_t ws_bytes = 0;
    std::vector<Op> ops;
    act16* in_nhwc = nullptr;
    float* tproj = nullptr;
    double* stats_arena = nullptr;
    size_t stats_bytes = 0;
    double flops_per_img = 0;
    float* v_out_slot = nullptr;  // the END conv writes here (patched per call)
    int end_op = -1;
};

// ------------------------------------------------------------------------------------------------
// plan construction
// ------------------------------------------------------------------------------------------------
// The nearest-x2 + 3x3 convs (models.py:41-47) run as four sub-pixel phases (folded 2x2 kernels) on the LOW-resolution tensor
// whenever the patch kernel takes the shape: 2.25x fewer MACs and no materialised upsampled tensor.
static bool subpixel_up_enabled() {
    static const bool off = getenv("PNPF_NO_SUBPIXEL") != nullptr;       // A/B switch (tools/ab_env.py)
    return !off;
}

static bool has_attn(const pnpf_unet_config& c, int side) {
    for (int i = 0; i < c.num_attn_resolutions; ++i)
        if (c.attn_resolutions[i] == side) return true;
    return false;
}

static void add_w(pnpf_engine* e, const std::string& name, std::vector<int64_t> shape) {
    WEntry w;
    w.name = name;
    w.shape = std::move(shape);
    e->windex[name] = (int)e->weights.size();
    e->weights.push_back(std::move(w));
}
static void add_conv_w(pnpf_engine* e, const std::string& p, int cin, int cout, int k) {
    add_w(e, p + ".weight", {cout, cin, k, k});
    add_w(e, p + ".bias", {cout});
}
static void add_vec2(pnpf_engine* e, const std::string& p, int c) {
    add_w(e, p + ".weight", {c});
    add_w(e, p + ".bias", {c});
}

static int build_layers(pnpf_engine* e) {
    const pnpf_unet_config& c = e->cfg;
    PNPF_REQUIRE(c.num_levels >= 1 && c.num_levels <= 8, "num_levels %d", c.num_levels);
    PNPF_REQUIRE(c.ch % 32 == 0, "ch=%d must be a multiple of 32 (GroupNorm groups / K chunk)", c.ch);
    PNPF_REQUIRE(c.input_channels >= 1 && c.input_channels <= 16, "input_channels %d unsupported (1..16)", c.input_channels);
    PNPF_REQUIRE(c.input_height % (1 << (c.num_levels - 1)) == 0, "input_height doesn't satisfy the condition");   // models.py:334
    const int temb_ch = c.ch * 4;
    add_w(e, "temb_net.main.0.weight", {temb_ch, c.ch});
    add_w(e, "temb_net.main.0.bias", {temb_ch});
    add_w(e, "temb_net.main.2.weight", {temb_ch, temb_ch});
    add_w(e, "temb_net.main.2.bias", {temb_ch});
    auto L = [&](LayerSpec::Kind k, std::string p, int ic, int oc, int side, int sk, bool push) {
        e->layers.push_back(LayerSpec{k, std::move(p), ic, oc, side, sk, push});
    };
    auto res_w = [&](const std::string& p, int ic, int oc) {
        add_w(e, p + ".temb_proj.weight", {oc, temb_ch});
        add_w(e, p + ".temb_proj.bias", {oc});
        add_vec2(e, p + ".norm1", ic);
        add_conv_w(e, p + ".conv1", ic, oc, 3);
        add_vec2(e, p + ".norm2", oc);
        add_conv_w(e, p + ".conv2", oc, oc, 3);
        if (ic != oc) add_conv_w(e, p + ".shortcut", ic, oc, 1);
        e->proj_off[p] = e->total_proj;
        e->total_proj += oc;
    };
    auto attn_w = [&](const std::string& p, int ch) {
        for (const char* n : {"attn_q", "attn_k", "attn_v", "proj_out"}) add_conv_w(e, p + "." + n, ch, ch, 1);
        add_vec2(e, p + ".norm", ch);
    };
    int side = c.input_height;
    std::vector<int> skip;
    char buf[128];
    L(LayerSpec::CONV, "begin_conv", c.input_channels, c.ch, side, 0, true);
    add_conv_w(e, "begin_conv", c.input_channels, c.ch, 3);
    skip.push_back(c.ch);
    int in_ch = c.ch;
    for (int lvl = 0; lvl < c.num_levels; ++lvl) {
        const int out_ch = c.ch * c.ch_mult[lvl];
        for (int blk = 0; blk < c.num_res_blocks; ++blk) {
            const bool at = has_attn(c, side);
            snprintf(buf, sizeof(buf), "down_modules.%d.%da_%da_block", lvl, lvl, blk);
            L(LayerSpec::RES, buf, in_ch, out_ch, side, 0, !at);
            res_w(buf, in_ch, out_ch);
            if (at) {
                snprintf(buf, sizeof(buf), "down_modules.%d.%da_%db_attn", lvl, lvl, blk);
                L(LayerSpec::ATTN, buf, out_ch, out_ch, side, 0, true);
                attn_w(buf, out_ch);
            }
            skip.push_back(out_ch);
            in_ch = out_ch;
        }
        if (lvl != c.num_levels - 1) {
            snprintf(buf, sizeof(buf), "down_modules.%d.%db_downsample", lvl, lvl);
            L(LayerSpec::DOWN, buf, in_ch, in_ch, side, 0, true);
            add_conv_w(e, buf, in_ch, in_ch, 3);
            side /= 2;
            skip.push_back(in_ch);
        }
    }
    L(LayerSpec::RES, "mid_modules.0", in_ch, in_ch, side, 0, false);
    res_w("mid_modules.0", in_ch, in_ch);
    L(LayerSpec::ATTN, "mid_modules.1", in_ch, in_ch, side, 0, false);
    attn_w("mid_modules.1", in_ch);
    L(LayerSpec::RES, "mid_modules.2", in_ch, in_ch, side, 0, false);
    res_w("mid_modules.2", in_ch, in_ch);
    for (int idx = 0; idx < c.num_levels; ++idx) {
        const int lvl = c.num_levels - 1 - idx;
        const int out_ch = c.ch * c.ch_mult[lvl];
        for (int blk = 0; blk < c.num_res_blocks + 1; ++blk) {
            const int sc = skip.back();
            skip.pop_back();
            snprintf(buf, sizeof(buf), "up_modules.%d.%da_%da_block", idx, lvl, blk);
            L(LayerSpec::RES, buf, in_ch + sc, out_ch, side, sc, false);
            res_w(buf, in_ch + sc, out_ch);
            if (has_attn(c, side)) {
                snprintf(buf, sizeof(buf), "up_modules.%d.%da_%db_attn", idx, lvl, blk);
                L(LayerSpec::ATTN, buf, out_ch, out_ch, side, 0, false);
                attn_w(buf, out_ch);
            }
            in_ch = out_ch;
        }
        if (lvl != 0) {
            snprintf(buf, sizeof(buf), "up_modules.%d.%db_upsample.up_conv", idx, lvl);
            L(LayerSpec::UP, buf, in_ch, in_ch, side, 0, false);
            add_conv_w(e, buf, in_ch, in_ch, 3);
            side *= 2;
        }
    }
    PNPF_REQUIRE(skip.empty(), "internal: skip stack not empty");
    L(LayerSpec::END, "end_conv", in_ch, c.input_channels, side, 0, false);
    add_vec2(e, "end_conv.0", in_ch);
    add_conv_w(e, "end_conv.2", in_ch, c.input_channels, 3);
    return 0;
}

// Residual blocks without a shortcut conv on the row-streaming levels add their input through the tensor core: conv2's
// packed weights get an identity 1x1 block appended (exact: fp16 x 1.0 accumulated in fp32), so the residual rides the
// kernel's TMA/MMA path instead of per-thread global loads in the epilogue.
static bool identity_shortcut(const LayerSpec& L) {
    static const bool off = getenv("PNPF_NO_IDENTITY") != nullptr;     // A/B switch (tools/ab_env.py)
    if (off) return false;
    return L.kind == LayerSpec::RES && L.in_ch == L.out_ch && L.skip_ch == 0 && L.side % 128 == 0 && L.out_ch % 32 == 0 && L.out_ch <= 64;
}

extern "C" int pnpf_create(const pnpf_unet_config* cfg, pnpf_engine** out) {
    PNPF_REQUIRE(cfg && out, "null argument");
    pnpf_engine* e = new pnpf_engine();
    e->cfg = *cfg;
    if (int rc = build_layers(e)) {
        delete e;
        return rc;
    }
    *out = e;
    return 0;
}

extern "C" void pnpf_destroy(pnpf_engine* e) {
    if (!e) return;
    if (e->wdev) cudaFree(e->wdev);
    delete e;
}

extern "C" int pnpf_num_weights(pnpf_engine* e) { return e ? (int)e->weights.size() : 0; }
extern "C" const char* pnpf_weight_name(pnpf_engine* e, int i) {
    return (e && i >= 0 && i < (int)e->weights.size()) ? e->weights[i].name.c_str() : nullptr;
}

extern "C" int pnpf_weight_shape(pnpf_engine* e, int i, int64_t shape[4], int* ndim) {
    PNPF_REQUIRE(e && shape && ndim && i >= 0 && i < (int)e->weights.size(), "bad argument");
    *ndim = (int)e->weights[i].shape.size();
    for (int k = 0; k < *ndim && k < 4; ++k) shape[k] = e->weights[i].shape[k];
    return 0;
}

extern "C" int pnpf_load_weight(pnpf_engine* e, const char* name, const float* host_data, const int64_t* shape, int ndim) {
    PNPF_REQUIRE(e && name && host_data && shape, "null argument");
    auto it = e->windex.find(name);
    PNPF_REQUIRE(it != e->windex.end(), "unexpected state_dict key '%s'", name);
    WEntry& w = e->weights[it->second];
    PNPF_REQUIRE((int)w.shape.size() == ndim, "'%s': expected %d dims, got %d", name, (int)w.shape.size(), ndim);
    size_t n = 1;
    for (int i = 0; i < ndim; ++i) {
        PNPF_REQUIRE(w.shape[i] == shape[i], "'%s': dim %d is %lld, expected %lld", name, i, (long long)shape[i], (long long)w.shape[i]);
        n *= (size_t)shape[i];
    }
    w.data.assign(host_data, host_data + n);
    w.loaded = true;
    e->finalized = false;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// weight repacking
// ------------------------------------------------------------------------------------------------
namespace {
struct Packer {
    std::vector<uint8_t> host;
    std::map<std::string, size_t>* off;
    size_t reserve(const std::string& name, size_t bytes) {
        size_t o = (host.size() + 255) / 256 * 256;
        host.resize(o + bytes, 0);
        (*off)[name] = o;
        return o;
    }
    float* f32(const std::string& name, size_t n) { size_t o = reserve(name, n * 4); return reinterpret_cast<float*>(host.data() + o); }
    act16* b16(const std::string& name, size_t n) { size_t o = reserve(name, n * 2); return reinterpret_cast<act16*>(host.data() + o); }
};
int round_n(int cout) {
    if (cout <= 16) return 16;
    if (cout <= 32) return 32;
    if (cout <= 64) return 64;
    if (cout <= 128) return 128;
    return (cout + 255) / 256 * 256;
}
}  // namespace

static const std::vector<float>& W(pnpf_engine* e, const std::string& n) { return e->weights[e->windex.at(n)].data; }

extern "C" int pnpf_finalize_weights(pnpf_engine* e) {
    PNPF_REQUIRE(e, "null engine");
    for (const WEntry& w : e->weights) PNPF_REQUIRE(w.loaded, "missing state_dict key '%s'", w.name.c_str());
    const pnpf_unet_config& c = e->cfg;
    const int temb_ch = c.ch * 4;
    Packer pk;
    pk.off = &e->woff;
    e->woff.clear();
    // --- time embedding
    {
        const int half = c.ch / 2;
        float* fr = pk.f32("temb.freqs", half);
        const float k = logf(10000.0f) / (float)(half - 1);       // models.py:270-272 (python double -> fp32 tensor math)
        const double kd = std::log(10000.0) / (half - 1);
        (void)k;
        for (int i = 0; i < half; ++i) fr[i] = expf((float)i * (float)(-kd));
        auto cp = [&](const char* dst, const std::string& src) {
            const std::vector<float>& v = W(e, src);
            float* d = pk.f32(dst, v.size());
            memcpy(d, v.data(), v.size() * 4);
        };
        auto cp_t = [&](const char* dst, const std::string& src, int rows, int cols) {      // [rows][cols] -> [cols][rows]
            const std::vector<float>& v = W(e, src);
            float* d = pk.f32(dst, v.size());
            for (int r = 0; r < rows; ++r)
                for (int q = 0; q < cols; ++q) d[(size_t)q * rows + r] = v[(size_t)r * cols + q];
        };
        cp_t("temb.w0_t", "temb_net.main.0.weight", temb_ch, c.ch);
        cp("temb.b0", "temb_net.main.0.bias");
        cp_t("temb.w2_t", "temb_net.main.2.weight", temb_ch, temb_ch);
        cp("temb.b2", "temb_net.main.2.bias");
        // NB: Packer pointers are invalidated by the next reserve() (vector growth): reserve both, then take pointers
        const size_t wp_off = pk.reserve("temb.wp_t", (size_t)temb_ch * e->total_proj * sizeof(float));
        const size_t bp_off = pk.reserve("temb.bp", (size_t)e->total_proj * sizeof(float));
        float* wp = reinterpret_cast<float*>(pk.host.data() + wp_off);
        float* bp = reinterpret_cast<float*>(pk.host.data() + bp_off);
        for (const LayerSpec& L : e->layers) {
            if (L.kind != LayerSpec::RES) continue;
            const int off = e->proj_off.at(L.prefix);
            const std::vector<float>& w = W(e, L.prefix + ".temb_proj.weight");   // [oc][temb_ch]
            const std::vector<float>& b = W(e, L.prefix + ".temb_proj.bias");
            for (int o = 0; o < L.out_ch; ++o) {
                bp[off + o] = b[o];
                for (int k2 = 0; k2 < temb_ch; ++k2) wp[(size_t)k2 * e->total_proj + off + o] = w[(size_t)o * temb_ch + k2];
            }
        }
    }
    auto pack_bias = [&](const std::string& name, int n_pad, const std::vector<float>& b, const std::vector<float>* b2) {
        float* d = pk.f32(name, n_pad);
        for (size_t i = 0; i < b.size(); ++i) d[i] = b[i] + (b2 ? (*b2)[i] : 0.f);
    };
    auto pack_gn = [&](const std::string& p) {
        const std::vector<float>& g = W(e, p + ".weight");
        const std::vector<float>& b = W(e, p + ".bias");
        memcpy(pk.f32(p + ".gamma", g.size()), g.data(), g.size() * 4);
        memcpy(pk.f32(p + ".beta", b.size()), b.data(), b.size() * 4);
    };
    for (const LayerSpec& L : e->layers) {
        const std::string& p = L.prefix;
        switch (L.kind) {
            case LayerSpec::CONV: {
                const int np = round_n(L.out_ch);
                act16* d = pk.b16(p + ".w", (size_t)np * 9 * CIN_PAD);
                pack_conv_weight(d, W(e, p + ".weight").data(), L.out_ch, L.in_ch, 3, np, CIN_PAD, nullptr, 0, 1.f);
                pack_bias(p + ".b", np, W(e, p + ".bias"), nullptr);
                break;
            }
            case LayerSpec::DOWN:
            case LayerSpec::UP: {
                const int np = round_n(L.out_ch);
                act16* d = pk.b16(p + ".w", (size_t)np * 9 * L.in_ch);
                pack_conv_weight(d, W(e, p + ".weight").data(), L.out_ch, L.in_ch, 3, np, L.in_ch, nullptr, 0, 1.f);
                pack_bias(p + ".b", np, W(e, p + ".bias"), nullptr);
                if (L.kind == LayerSpec::UP && subpixel_up_enabled()) {
                    // sub-pixel form of nearest x2 + 3x3 conv: four folded 2x2 weight sets, packed [np][4*Cin] in (i, j, cin) order
                    std::vector<float> f((size_t)L.out_ch * L.in_ch * 4);
                    for (int ph = 0; ph < 4; ++ph) {
                        fold_subpixel_weights(W(e, p + ".weight").data(), L.out_ch, L.in_ch, ph >> 1, ph & 1, f.data());
                        act16* ds = pk.b16(p + ".w_sp" + std::to_string(ph), (size_t)np * 4 * L.in_ch);
                        pack_conv_weight(ds, f.data(), L.out_ch, L.in_ch, 2, np, L.in_ch, nullptr, 0, 1.f);
                    }
                    if (2 * np <= 256 && np == L.out_ch)            // two column phases per launch (SUBPIX = 2): [2*C_out][6*Cin] per row parity
                        for (int a = 0; a < 2; ++a) {
                            act16* ds = pk.b16(p + ".w_sp2a" + std::to_string(a), (size_t)2 * np * 6 * L.in_ch);
                            pack_subpixel_pair_weights(ds, W(e, p + ".weight").data(), L.out_ch, L.in_ch, a);
                        }
                }
                break;
            }
            case LayerSpec::RES: {
                const int np = round_n(L.out_ch);
                pack_gn(p + ".norm1");
                pack_gn(p + ".norm2");
                act16* d1 = pk.b16(p + ".conv1.w", (size_t)np * 9 * L.in_ch);
                pack_conv_weight(d1, W(e, p + ".conv1.weight").data(), L.out_ch, L.in_ch, 3, np, L.in_ch, nullptr, 0, 1.f);
                pack_bias(p + ".conv1.b", np, W(e, p + ".conv1.bias"), nullptr);
                const bool sc = L.in_ch != L.out_ch;
                act16* d2 = pk.b16(p + ".conv2.w", (size_t)np * (9 * L.out_ch + (sc ? L.in_ch : 0)));
                pack_conv_weight(d2, W(e, p + ".conv2.weight").data(), L.out_ch, L.out_ch, 3, np, L.out_ch,
                                 sc ? W(e, p + ".shortcut.weight").data() : nullptr, sc ? L.in_ch : 0, 1.f);
                pack_bias(p + ".conv2.b", np, W(e, p + ".conv2.bias"), sc ? &W(e, p + ".shortcut.bias") : nullptr);
                if (identity_shortcut(L)) {
                    std::vector<float> eye((size_t)L.out_ch * L.out_ch, 0.f);
                    for (int o = 0; o < L.out_ch; ++o) eye[(size_t)o * L.out_ch + o] = 1.f;
                    act16* d3 = pk.b16(p + ".conv2.wid", (size_t)np * (9 * L.out_ch + L.out_ch));
                    pack_conv_weight(d3, W(e, p + ".conv2.weight").data(), L.out_ch, L.out_ch, 3, np, L.out_ch, eye.data(), L.out_ch, 1.f);
                }
                break;
            }
            case LayerSpec::ATTN: {
                const int C = L.in_ch;
                pack_gn(p + ".norm");
                const float scale = 1.0f / sqrtf((float)C);               // models.py:154 folded into Wq, bq
                const int np = round_n(2 * C);
                act16* dqk = pk.b16(p + ".qk.w", (size_t)np * C);
                std::vector<float> wqk((size_t)2 * C * C), bqk(2 * C);
                const std::vector<float>&wq = W(e, p + ".attn_q.weight"), &wk = W(e, p + ".attn_k.weight");
                const std::vector<float>&bq = W(e, p + ".attn_q.bias"), &bk = W(e, p + ".attn_k.bias");
                for (size_t i = 0; i < (size_t)C * C; ++i) {
                    wqk[i] = wq[i] * scale;
                    wqk[(size_t)C * C + i] = wk[i];
                }
                for (int i = 0; i < C; ++i) {
                    bqk[i] = bq[i] * scale;
                    bqk[C + i] = bk[i];
                }
                pack_conv_weight(dqk, wqk.data(), 2 * C, C, 1, np, C, nullptr, 0, 1.f);
                pack_bias(p + ".qk.b", np, bqk, nullptr);
                act16* dv = pk.b16(p + ".v.w", (size_t)C * C);              // A operand of the V^T GEMM: [C rows][C k]
                const std::vector<float>& wv = W(e, p + ".attn_v.weight");
                for (size_t i = 0; i < (size_t)C * C; ++i) dv[i] = to_act16(wv[i]);
                const int npo = round_n(C);
                act16* dpo = pk.b16(p + ".proj.w", (size_t)npo * C);
                const std::vector<float>& wo = W(e, p + ".proj_out.weight");
                pack_conv_weight(dpo, wo.data(), C, C, 1, npo, C, nullptr, 0, 1.f);
                // softmax rows sum to 1 => P(V + 1 bv^T) = PV + 1 bv^T: fold Wo*bv into the projection bias
                const std::vector<float>&bv = W(e, p + ".attn_v.bias"), &bo = W(e, p + ".proj_out.bias");
                std::vector<float> bfold(C);
                for (int o = 0; o < C; ++o) {
                    double acc = bo[o];
                    for (int k2 = 0; k2 < C; ++k2) acc += (double)wo[(size_t)o * C + k2] * bv[k2];
                    bfold[o] = (float)acc;
                }
                pack_bias(p + ".proj.b", npo, bfold, nullptr);
                break;
            }
            case LayerSpec::END: {
                pack_gn(p + ".0");
                const int np = round_n(L.out_ch);
                act16* d = pk.b16(p + ".2.w", (size_t)np * 9 * L.in_ch);
                pack_conv_weight(d, W(e, p + ".2.weight").data(), L.out_ch, L.in_ch, 3, np, L.in_ch, nullptr, 0, 1.f);
                pack_bias(p + ".2.b", np, W(e, p + ".2.bias"), nullptr);
                break;
            }
        }
    }
    if (e->wdev) cudaFree(e->wdev);
    e->wdev = nullptr;
    e->wdev_bytes = pk.host.size();
    PNPF_CHECK_CUDA(cudaMalloc(&e->wdev, e->wdev_bytes));
    PNPF_CHECK_CUDA(cudaMemcpy(e->wdev, pk.host.data(), e->wdev_bytes, cudaMemcpyHostToDevice));
    e->finalized = true;
    e->ops.clear();               // a bound plan refers to the old weight arena
    e->ws = nullptr;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// workspace layout + op list
// ------------------------------------------------------------------------------------------------
namespace {
struct Arena {
    uint8_t* base;
    size_t off = 0;
    template <typename T>
    T* take(size_t n) {
        off = (off + 1023) / 1024 * 1024;
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += n * sizeof(T);
        return p;
    }
};
}  // namespace

template <typename T>
static const T* wptr(pnpf_engine* e, const std::string& name) {
    return reinterpret_cast<const T*>(e->wdev + e->woff.at(name));
}

// Builds the op list against `base` (nullptr = size query only).  Returns bytes needed.
static int build_plan(pnpf_engine* e, uint8_t* base, int Bm, size_t* need) {
    const pnpf_unet_config& c = e->cfg;
    const bool real = base != nullptr;
    Arena A{base};
    std::vector<Op> ops;
    double flops = 0;
    const int side0 = c.input_height;
    // ---- maxima for the shared temporaries
    long long m_a1 = 0, m_h1 = 0, m_up = 0, m_h = 0, m_qk = 0, m_S = 0, m_P = 0;
    size_t stats_elems = 0;
    for (const LayerSpec& L : e->layers) {
        const long long px = (long long)L.side * L.side;
        switch (L.kind) {
            case LayerSpec::RES:
                m_a1 = std::max(m_a1, px * L.in_ch);
                m_h1 = std::max(m_h1, px * L.out_ch);
                if (!L.push) m_h = std::max(m_h, px * L.out_ch);
                stats_elems += 4 * (size_t)L.out_ch;          // conv1 output + block output
                break;
            case LayerSpec::ATTN:
                m_a1 = std::max(m_a1, px * L.in_ch);
                m_h1 = std::max(m_h1, px * L.in_ch);
                m_qk = std::max(m_qk, px * 2 * L.in_ch);
                m_S = std::max(m_S, px * px);
                m_P = std::max(m_P, px * px);
                if (!L.push) m_h = std::max(m_h, px * L.out_ch);
                stats_elems += 2 * (size_t)L.out_ch;
                break;
            case LayerSpec::UP:
                m_up = std::max(m_up, 4 * px * L.in_ch);
                m_h = std::max(m_h, 4 * px * L.out_ch);
                stats_elems += 2 * (size_t)L.out_ch;
                break;
            case LayerSpec::END:
                m_a1 = std::max(m_a1, px * L.in_ch);
                break;
            default:                                          // CONV, DOWN: one produced tensor
                stats_elems += 2 * (size_t)L.out_ch;
                break;
        }
    }
    act16* in_nhwc = A.take<act16>((size_t)Bm * side0 * side0 * cin_store(c));
    float* tproj = A.take<float>((size_t)Bm * e->total_proj);
    double* stats = A.take<double>((size_t)Bm * stats_elems);
    const size_t stats_bytes = (size_t)Bm * stats_elems * sizeof(double);
    act16* t_a1 = A.take<act16>((size_t)Bm * m_a1);       // normalised conv1 / attention input
    act16* t_xcat = A.take<act16>((size_t)Bm * m_a1);     // raw concat (shortcut operand)
    act16* t_h1 = A.take<act16>((size_t)Bm * m_h1);       // conv1 output / attention O
    act16* t_a2 = A.take<act16>((size_t)Bm * m_h1);       // normalised conv2 input / V^T
    act16* t_up = A.take<act16>((size_t)Bm * m_up);
    act16* t_h[2] = {A.take<act16>((size_t)Bm * m_h), A.take<act16>((size_t)Bm * m_h)};
    act16* t_qk = A.take<act16>((size_t)Bm * m_qk);
    float* t_S = A.take<float>((size_t)Bm * m_S);
    act16* t_P = A.take<act16>((size_t)Bm * m_P);

    size_t stats_off = 0;   // in doubles, per image block layout: [op][img][C][2] -> we give every GN op its own [Bm][C][2]
    auto new_stats = [&](int C) {
        double* p = real ? stats + stats_off : nullptr;
        stats_off += (size_t)Bm * C * 2;
        return p;
    };
    int hsel = 0;
    auto new_h = [&](int C, int side, bool push) {
        Act a;
        a.C = C;
        a.side = side;
        a.stats = new_stats(C);
        if (push) a.p = A.take<act16>((size_t)Bm * a.elems_per_img());
        else { a.p = t_h[hsel]; hsel ^= 1; }
        return a;
    };
    auto set_out = [&](Op& o, const act16* p, int C, int side) { o.out_p = p; o.out_C = C; o.out_side = side; o.out_pitch = C; };

    auto add_gn = [&](const std::string& name, const GnSrc& src, int side, const std::string& wname, int silu, act16* dst,
                      act16* raw) {
        const int C = src.C1 + src.C2;
        Op a;
        a.kind = Op::GN_APPLY; a.name = name; a.gsrc = src; a.HW = side * side; a.impl = "gn_apply";
        if (real) { a.gamma = wptr<float>(e, wname + ".gamma"); a.beta = wptr<float>(e, wname + ".beta"); }
        a.silu = silu; a.dst = dst; a.raw_dst = raw;
        a.bytes = (raw ? 6.0 : 4.0) * C * side * side;
        set_out(a, dst, C, side);
        ops.push_back(a);
    };
    auto add_conv = [&](const std::string& name, ConvDesc d, const act16* dbg_out, int dbg_C, int dbg_side) -> int {
        Op o;
        o.kind = Op::TC; o.name = name;
        d.B = Bm;
        {
            char buf[200];
            describe_conv_impl(d, buf, sizeof(buf));
            o.impl = buf;
        }
        if (!real && getenv("PNPF_PLAN_DUMP")) {
            char buf[200];
            describe_conv_impl(d, buf, sizeof(buf));
            fprintf(stderr, "plan: %-44s %4dx%-4d Cin=%-3d C2=%-3d N=%-3d %s\n", name.c_str(), d.Hout, d.Wout, d.Cin, d.C2, d.n_valid, buf);
        }
        if (real) { if (int rc = prepare_conv(o.tc, d)) return rc; }
        o.flops = 2.0 * d.Hout * d.Wout * (double)d.n_valid * ((double)(d.subpix == 2 ? 8 : (d.subpix ? 4 : d.ksize * d.ksize)) * d.Cin + ((d.x2 && !d.x2_identity) ? d.C2 : 0));
        o.bytes = 2.0 * d.Hin * d.Win * (d.x_cvalid > 0 ? d.x_cvalid : d.Cin) + (d.x2 ? 2.0 * d.Hout * d.Wout * d.C2 : 0.0) +   // Cin / C2 are concat totals
                  (d.residual ? 2.0 * d.Hout * d.Wout * d.n_valid : 0.0) +
                  (d.out_mode == 0 ? 2.0 : 4.0) * d.Hout * d.Wout * d.n_valid;
        flops += o.flops;
        if (dbg_out) set_out(o, dbg_out, dbg_C, dbg_side);
        ops.push_back(o);
        return 0;
    };
    auto add_gemm = [&](const std::string& name, GemmDesc d) -> int {
        Op o;
        o.kind = Op::TC; o.name = name;
        {
            char buf[96];
            snprintf(buf, sizeof(buf), "conv_gemm<64,%d> gemm M=%d N=%d K=%d", d.N > 256 ? 256 : d.N, d.M, d.N, d.K);
            o.impl = buf;
        }
        d.batch = Bm;
        if (real) { if (int rc = prepare_gemm(o.tc, d)) return rc; }
        o.flops = 2.0 * (double)d.M * d.N * d.K;
        o.bytes = 2.0 * ((double)d.M * d.K * (d.a_batched ? 1 : 0) + (double)d.N * d.K * (d.b_batched ? 1 : 0)) +
                  (d.out_mode == 0 ? 2.0 : 4.0) * d.M * d.N;
        flops += o.flops;
        ops.push_back(o);
        return 0;
    };

    { Op o; o.kind = Op::MEMSET; o.name = "zero_gn_stats"; ops.push_back(o); }
    { Op o; o.kind = Op::IN_SHIM; o.name = "input_nchw_to_nhwc"; o.bytes = (4.0 * c.input_channels + 2.0 * cin_store(c)) * side0 * side0;
      set_out(o, in_nhwc, cin_store(c), side0); ops.push_back(o); }
    { Op o; o.kind = Op::TEMB; o.name = "time_embedding"; ops.push_back(o); }

    std::vector<Act> hs;
    Act h;
    for (const LayerSpec& L : e->layers) {
        const std::string& p = L.prefix;
        const int side = L.side;
        const long long px = (long long)side * side;
        switch (L.kind) {
            case LayerSpec::CONV: {
                Act y = new_h(L.out_ch, side, true);
                ConvDesc d;
                d.x = in_nhwc; d.Hin = d.Win = d.Hout = d.Wout = side; d.Cin = CIN_PAD; d.x_pitch = cin_store(c); d.x_cvalid = cin_store(c);
                d.N_pad = round_n(L.out_ch); d.ksize = 3; d.stride = 1;
                if (real) { d.w = wptr<act16>(e, p + ".w"); d.bias = wptr<float>(e, p + ".b"); }
                d.stats_out = y.stats;
                d.out = y.p; d.out_mode = 0; d.out_img_stride = px * L.out_ch; d.out_row_stride = L.out_ch; d.n_valid = L.out_ch;
                if (int rc = add_conv(p, d, y.p, y.C, side)) return rc;
                h = y;
                break;
            }
            case LayerSpec::RES: {
                GnSrc src{};
                Act skip{};
                if (L.skip_ch) {
                    skip = hs.back();
                    hs.pop_back();
                    src = GnSrc{h.p, h.C, h.C, skip.p, skip.C, skip.C, h.stats, skip.stats};
                } else {
                    src = GnSrc{h.p, h.C, h.C, nullptr, 0, 0, h.stats, nullptr};
                }
                const bool sc = L.in_ch != L.out_ch;
                double* h1_stats = new_stats(L.out_ch);
                Act y = new_h(L.out_ch, side, L.push);
                // ---- conv1 (+ temb) and conv2 (+ shortcut / residual) descriptors in their FUSED form: GroupNorm+SiLU applied
                //      inside the row-streaming kernel, concat inputs read straight from the two source tensors
                ConvDesc d1;
                d1.x = h.p; d1.x_pitch = h.C; d1.Cin = L.in_ch;
                if (L.skip_ch) { d1.xb = skip.p; d1.Cb = skip.C; d1.xb_pitch = skip.C; }
                d1.Hin = d1.Win = d1.Hout = d1.Wout = side;
                d1.N_pad = round_n(L.out_ch); d1.ksize = 3; d1.stride = 1;
                if (real) {
                    d1.w = wptr<act16>(e, p + ".conv1.w"); d1.bias = wptr<float>(e, p + ".conv1.b");
                    d1.gn_gamma = wptr<float>(e, p + ".norm1.gamma"); d1.gn_beta = wptr<float>(e, p + ".norm1.beta");
                } else {
                    d1.gn_gamma = reinterpret_cast<const float*>(0x10); d1.gn_beta = d1.gn_gamma;   // shape analysis only
                }
                d1.gn_stats_a = h.stats; d1.gn_stats_b = L.skip_ch ? skip.stats : nullptr;
                if (!real) { d1.gn_stats_a = reinterpret_cast<const double*>(0x10); d1.gn_stats_b = d1.gn_stats_a; }
                d1.bias_img = real ? tproj + e->proj_off.at(p) : nullptr; d1.bias_img_stride = e->total_proj;
                d1.stats_out = h1_stats;
                d1.out = t_h1; d1.out_mode = 0; d1.out_img_stride = px * L.out_ch; d1.out_row_stride = L.out_ch; d1.n_valid = L.out_ch;
                ConvDesc d2;
                d2.x = t_h1; d2.x_pitch = L.out_ch; d2.Cin = L.out_ch;
                d2.Hin = d2.Win = d2.Hout = d2.Wout = side;
                d2.N_pad = round_n(L.out_ch); d2.ksize = 3; d2.stride = 1;
                if (real) {
                    d2.w = wptr<act16>(e, p + ".conv2.w"); d2.bias = wptr<float>(e, p + ".conv2.b");
                    d2.gn_gamma = wptr<float>(e, p + ".norm2.gamma"); d2.gn_beta = wptr<float>(e, p + ".norm2.beta");
                } else {
                    d2.gn_gamma = reinterpret_cast<const float*>(0x10); d2.gn_beta = d2.gn_gamma;
                }
                d2.gn_stats_a = real ? h1_stats : reinterpret_cast<const double*>(0x10);
                if (sc) {
                    d2.x2 = h.p; d2.x2_pitch = h.C; d2.C2 = L.in_ch;
                    if (L.skip_ch) { d2.x2b = skip.p; d2.C2b = skip.C; d2.x2b_pitch = skip.C; }
                } else {
                    PNPF_REQUIRE(!L.skip_ch, "internal: concat block without shortcut");
                    d2.residual = h.p; d2.res_img_stride = px * L.out_ch; d2.res_row_stride = L.out_ch;
                    if (identity_shortcut(L)) {                    // residual as an identity 1x1 on the tensor core
                        ConvDesc di = d2;
                        di.residual = nullptr;
                        di.x2 = h.p; di.x2_pitch = h.C; di.C2 = L.in_ch; di.x2_identity = 1;
                        if (real) di.w = wptr<act16>(e, p + ".conv2.wid");
                        if (rowconv_eligible(di)) d2 = di;
                    }
                }
                d2.stats_out = y.stats;
                d2.out = y.p; d2.out_mode = 0; d2.out_img_stride = px * L.out_ch; d2.out_row_stride = L.out_ch; d2.n_valid = L.out_ch;
                bool fuse2 = rowconv_eligible(d2);
                bool fuse1 = rowconv_eligible(d1);
                if (!fuse2 && L.skip_ch && sc) fuse1 = false;       // the unfused conv2 needs the raw concat copy made by norm1
                if (!fuse1) {                                        // separate GroupNorm pass -> normalised (concatenated) operand
                    add_gn(p + ".norm1", src, side, p + ".norm1", 1, t_a1, (L.skip_ch && sc) ? t_xcat : nullptr);
                    d1.x = t_a1; d1.x_pitch = L.in_ch; d1.xb = nullptr; d1.Cb = 0;
                    d1.gn_gamma = d1.gn_beta = nullptr; d1.gn_stats_a = d1.gn_stats_b = nullptr;
                }
                if (int rc = add_conv(p + ".conv1", d1, t_h1, L.out_ch, side)) return rc;
                if (!fuse2) {
                    add_gn(p + ".norm2", GnSrc{t_h1, L.out_ch, L.out_ch, nullptr, 0, 0, h1_stats, nullptr}, side, p + ".norm2", 1, t_a2, nullptr);
                    d2.x = t_a2;
                    d2.gn_gamma = d2.gn_beta = nullptr; d2.gn_stats_a = nullptr;
                    if (sc && L.skip_ch) { d2.x2 = t_xcat; d2.x2_pitch = L.in_ch; d2.x2b = nullptr; d2.C2b = 0; }
                }
                if (int rc = add_conv(p, d2, y.p, y.C, side)) return rc;
                h = y;
                break;
            }
            case LayerSpec::ATTN: {
                const int C = L.in_ch;
                const int Lk = (int)px;
                PNPF_REQUIRE(Lk % 16 == 0 && (Lk <= 256 ? (Lk == 16 || Lk == 32 || Lk == 64 || Lk == 128 || Lk == 256) : Lk % 256 == 0),
                             "attention over %d tokens unsupported by the tensor-core path", Lk);
                PNPF_REQUIRE(C % 64 == 0, "attention channels %d must be a multiple of 64", C);
                add_gn(p + ".norm", GnSrc{h.p, C, C, nullptr, 0, 0, h.stats, nullptr}, side, p + ".norm", 0, t_a1, nullptr);
                ConvDesc dq;                                   // [q*scale | k] = hn * [Wq*scale ; Wk]^T
                dq.x = t_a1; dq.Hin = dq.Win = dq.Hout = dq.Wout = side; dq.Cin = C; dq.x_pitch = C;
                dq.N_pad = round_n(2 * C); dq.ksize = 1; dq.stride = 1;
                if (real) { dq.w = wptr<act16>(e, p + ".qk.w"); dq.bias = wptr<float>(e, p + ".qk.b"); }
                dq.out = t_qk; dq.out_mode = 0; dq.out_img_stride = px * 2 * C; dq.out_row_stride = 2 * C; dq.n_valid = 2 * C;
                if (int rc = add_conv(p + ".qk", dq, t_qk, 2 * C, side)) return rc;
                GemmDesc gv;                                   // V^T[b] (C x L) = Wv (C x C) * hn[b]^T
                gv.A = real ? wptr<act16>(e, p + ".v.w") : nullptr; gv.lda = C; gv.a_batched = 0;
                gv.Bm = t_a1; gv.ldb = C; gv.b_bstride = px * C; gv.b_batched = 1;
                gv.M = C; gv.N = Lk; gv.K = C;
                gv.out = t_a2; gv.out_mode = 0; gv.out_img_stride = (long long)C * Lk; gv.out_row_stride = Lk;
                if (int rc = add_gemm(p + ".vT", gv)) return rc;
                if (attn_core_eligible(Lk, C)) {
                    // fused core: S = q k^T in TMEM -> row softmax in registers -> P v -> proj_out + bias + residual, one launch;
                    // the fp32 logits and the probabilities never reach HBM (pnpf_attn.cuh)
                    Act y = new_h(C, side, L.push);
                    Op o;
                    o.kind = Op::ATTN; o.name = p; o.impl = "attn_core<256,256> S->softmax->PV->proj fused";
                    AttnDesc ad;
                    ad.qk = t_qk; ad.vT = t_a2; ad.residual = h.p; ad.out = y.p; ad.stats_out = y.stats; ad.B = Bm; ad.L = Lk; ad.C = C;
                    if (real) {
                        ad.w = wptr<act16>(e, p + ".proj.w"); ad.bias = wptr<float>(e, p + ".proj.b");
                        if (int rc = prepare_attn(o.attn, ad)) return rc;
                    }
                    o.flops = 2.0 * Lk * Lk * C * 2 + 2.0 * Lk * C * C;
                    o.bytes = 2.0 * (3.0 * Lk * C + 2.0 * Lk * C);          // q, k, v^T, residual in; y out
                    flops += o.flops;
                    set_out(o, y.p, C, side);
                    ops.push_back(o);
                    h = y;
                    break;
                }
                GemmDesc gs;                                   // S[b] = q[b] k[b]^T   (fp32 logits)
                gs.A = t_qk; gs.lda = 2 * C; gs.a_bstride = px * 2 * C; gs.a_batched = 1;
                gs.Bm = t_qk + C; gs.ldb = 2 * C; gs.b_bstride = px * 2 * C; gs.b_batched = 1;
                gs.M = Lk; gs.N = Lk; gs.K = C;
                gs.out = t_S; gs.out_mode = 1; gs.out_img_stride = px * px; gs.out_row_stride = Lk;
                if (int rc = add_gemm(p + ".qkT", gs)) return rc;
                {
                    Op o;
                    o.kind = Op::SOFTMAX; o.name = p + ".softmax"; o.S = t_S; o.P = t_P; o.rows_per_img = Lk; o.L = Lk;
                    o.bytes = 6.0 * Lk * Lk;
                    ops.push_back(o);
                }
                GemmDesc go;                                   // O[b] (L x C) = P[b] (L x L) * V^T[b]^T
                go.A = t_P; go.lda = Lk; go.a_bstride = px * px; go.a_batched = 1;
                go.Bm = t_a2; go.ldb = Lk; go.b_bstride = (long long)C * Lk; go.b_batched = 1;
                go.M = Lk; go.N = C; go.K = Lk;
                go.out = t_h1; go.out_mode = 0; go.out_img_stride = px * C; go.out_row_stride = C;
                if (int rc = add_gemm(p + ".pv", go)) return rc;
                Act y = new_h(C, side, L.push);
                ConvDesc dp;                                   // y = x + proj_out(O) (+ folded V bias)
                dp.x = t_h1; dp.Hin = dp.Win = dp.Hout = dp.Wout = side; dp.Cin = C; dp.x_pitch = C;
                dp.N_pad = round_n(C); dp.ksize = 1; dp.stride = 1;
                if (real) { dp.w = wptr<act16>(e, p + ".proj.w"); dp.bias = wptr<float>(e, p + ".proj.b"); }
                dp.residual = h.p; dp.res_img_stride = px * C; dp.res_row_stride = C;
                dp.stats_out = y.stats;
                dp.out = y.p; dp.out_mode = 0; dp.out_img_stride = px * C; dp.out_row_stride = C; dp.n_valid = C;
                if (int rc = add_conv(p, dp, y.p, C, side)) return rc;
                h = y;
                break;
            }
            case LayerSpec::DOWN: {
                const int so = side / 2;
                Act y = new_h(L.out_ch, so, true);
                ConvDesc d;
                d.x = h.p; d.Hin = d.Win = side; d.Hout = d.Wout = so; d.Cin = L.in_ch; d.x_pitch = L.in_ch;
                d.N_pad = round_n(L.out_ch); d.ksize = 3; d.stride = 2;
                if (real) { d.w = wptr<act16>(e, p + ".w"); d.bias = wptr<float>(e, p + ".b"); }
                d.stats_out = y.stats;
                d.out = y.p; d.out_mode = 0; d.out_img_stride = (long long)so * so * L.out_ch; d.out_row_stride = L.out_ch; d.n_valid = L.out_ch;
                if (int rc = add_conv(p, d, y.p, y.C, so)) return rc;
                h = y;
                break;
            }
            case LayerSpec::UP: {
                const int so = side * 2;
                if (subpixel_up_enabled()) {
                    // four sub-pixel phases on the low-resolution tensor (pnpf_patchconv.cuh SUBPIX) instead of
                    // upsample2x + a 3x3 conv on the 4x larger tensor
                    ConvDesc d;
                    d.x = h.p; d.Hin = d.Win = d.Hout = d.Wout = side; d.Cin = L.in_ch; d.x_pitch = L.in_ch;
                    d.N_pad = round_n(L.out_ch); d.ksize = 3; d.stride = 1; d.subpix = 1;
                    d.out_mode = 0; d.out_img_stride = (long long)so * so * L.out_ch; d.out_row_stride = L.out_ch; d.n_valid = L.out_ch;
                    // two column phases per launch (SUBPIX = 2: 2 x 3 taps, N = 2 C_out) when 2 C_out fits one MMA / the double-buffered TMEM
                    static const bool no_pair_phase = getenv("PNPF_NO_SUBPIX2") != nullptr;      // A/B switch (tools/ab_env.py)
                    ConvDesc d2 = d;
                    d2.subpix = 2; d2.N_pad = 2 * round_n(L.out_ch);
                    if (!no_pair_phase && round_n(L.out_ch) == L.out_ch && d2.N_pad <= 256 && patchconv_eligible(d2)) {
                        Act y = new_h(L.out_ch, so, false);
                        d2.out = y.p; d2.stats_out = y.stats;
                        for (int a = 0; a < 2; ++a) {
                            d2.sp_a = a; d2.sp_b = 0;
                            if (real) { d2.w = wptr<act16>(e, p + ".w_sp2a" + std::to_string(a)); d2.bias = wptr<float>(e, p + ".b"); }
                            const std::string nm = a == 1 ? p : p + ".rows0";
                            if (int rc = add_conv(nm, d2, a == 1 ? y.p : nullptr, y.C, so)) return rc;
                        }
                        h = y;
                        break;
                    }
                    if (patchconv_eligible(d)) {
                        Act y = new_h(L.out_ch, so, false);
                        d.out = y.p; d.stats_out = y.stats;
                        for (int ph = 0; ph < 4; ++ph) {
                            d.sp_a = ph >> 1; d.sp_b = ph & 1;
                            if (real) { d.w = wptr<act16>(e, p + ".w_sp" + std::to_string(ph)); d.bias = wptr<float>(e, p + ".b"); }
                            // the op that completes the tensor carries the layer's name (debug taps compare it with the oracle)
                            const std::string nm = ph == 3 ? p : p + ".phase" + std::to_string(ph);
                            if (int rc = add_conv(nm, d, ph == 3 ? y.p : nullptr, y.C, so)) return rc;
                        }
                        h = y;
                        break;
                    }
                }
                {
                    Op o;
                    o.kind = Op::UPSAMPLE; o.name = p + ".nearest2x"; o.up_src = h.p; o.up_side = side; o.up_C = L.in_ch; o.dst = t_up;
                    o.bytes = 10.0 * L.in_ch * side * side;
                    set_out(o, t_up, L.in_ch, so);
                    ops.push_back(o);
                }
                Act y = new_h(L.out_ch, so, false);
                ConvDesc d;
                d.x = t_up; d.Hin = d.Win = d.Hout = d.Wout = so; d.Cin = L.in_ch; d.x_pitch = L.in_ch;
                d.N_pad = round_n(L.out_ch); d.ksize = 3; d.stride = 1;
                if (real) { d.w = wptr<act16>(e, p + ".w"); d.bias = wptr<float>(e, p + ".b"); }
                d.stats_out = y.stats;
                d.out = y.p; d.out_mode = 0; d.out_img_stride = (long long)so * so * L.out_ch; d.out_row_stride = L.out_ch; d.n_valid = L.out_ch;
                if (int rc = add_conv(p, d, y.p, y.C, so)) return rc;
                h = y;
                break;
            }
            case LayerSpec::END: {
                ConvDesc d;
                d.x = h.p; d.x_pitch = h.C; d.Hin = d.Win = d.Hout = d.Wout = side; d.Cin = L.in_ch;
                d.N_pad = round_n(L.out_ch); d.ksize = 3; d.stride = 1;
                if (real) {
                    d.w = wptr<act16>(e, p + ".2.w"); d.bias = wptr<float>(e, p + ".2.b");
                    d.gn_gamma = wptr<float>(e, p + ".0.gamma"); d.gn_beta = wptr<float>(e, p + ".0.beta");
                    d.gn_stats_a = h.stats;
                } else {
                    d.gn_gamma = reinterpret_cast<const float*>(0x10); d.gn_beta = d.gn_gamma;
                    d.gn_stats_a = reinterpret_cast<const double*>(0x10);
                }
                d.out = reinterpret_cast<void*>(0x10);         // patched per call with the caller's v pointer
                d.out_mode = 2; d.out_img_stride = px * L.out_ch; d.out_row_stride = 1; d.out_col_stride = px; d.n_valid = L.out_ch;
                if (!rowconv_eligible(d)) {
                    add_gn(p + ".0", GnSrc{h.p, h.C, h.C, nullptr, 0, 0, h.stats, nullptr}, side, p + ".0", 1, t_a1, nullptr);
                    d.x = t_a1; d.x_pitch = L.in_ch;
                    d.gn_gamma = d.gn_beta = nullptr; d.gn_stats_a = nullptr;
                }
                if (int rc = add_conv(p + ".2", d, nullptr, 0, 0)) return rc;
                break;
            }
        }
        if (L.push) hs.push_back(h);
    }
    PNPF_REQUIRE(hs.empty(), "internal: skip stack not empty after plan");
    *need = A.off + 1024;
    if (real) {
        e->ops = std::move(ops);
        e->in_nhwc = in_nhwc;
        e->tproj = tproj;
        e->stats_arena = stats;
        e->stats_bytes = stats_bytes;
        e->flops_per_img = flops;
        e->end_op = (int)e->ops.size() - 1;
    }
    return 0;
}

extern "C" size_t pnpf_workspace_bytes(pnpf_engine* e, int max_batch) {
    if (!e || max_batch < 1) return 0;
    size_t need = 0;
    if (build_plan(e, nullptr, max_batch, &need)) return 0;
    return need;
}

extern "C" int pnpf_bind_workspace(pnpf_engine* e, void* workspace, size_t bytes, int max_batch) {
    PNPF_REQUIRE(e && workspace && max_batch >= 1, "bad argument");
    PNPF_REQUIRE(e->finalized, "pnpf_finalize_weights must be called before pnpf_bind_workspace");
    PNPF_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, "workspace must be 1024-byte aligned");
    size_t need = 0;
    if (int rc = build_plan(e, nullptr, max_batch, &need)) return rc;
    PNPF_REQUIRE(bytes >= need, "workspace too small: %zu < %zu bytes", bytes, need);
    if (int rc = build_plan(e, static_cast<uint8_t*>(workspace), max_batch, &need)) return rc;
    e->ws = static_cast<uint8_t*>(workspace);
    e->ws_bytes = bytes;
    e->max_batch = max_batch;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// execution
// ------------------------------------------------------------------------------------------------
static int run_ops(pnpf_engine* e, const float* x, const float* t, float* v, int batch, int n_ops, cudaStream_t st) {
    PNPF_REQUIRE(e && e->ws && !e->ops.empty(), "no workspace bound");
    PNPF_REQUIRE(batch >= 1 && batch <= e->max_batch, "batch %d exceeds the bound workspace (max %d)", batch, e->max_batch);
    const pnpf_unet_config& c = e->cfg;
    const int HW0 = c.input_height * c.input_height;
    for (int i = 0; i < n_ops && i < (int)e->ops.size(); ++i) {
        Op& o = e->ops[i];
        int rc = 0;
        switch (o.kind) {
            case Op::MEMSET:
                PNPF_CHECK_CUDA(cudaMemsetAsync(e->stats_arena, 0, e->stats_bytes, st));
                break;
            case Op::IN_SHIM:
                rc = launch_nchw_to_nhwc_pad(x, batch, c.input_channels, HW0, e->in_nhwc, cin_store(c), st);
                break;
            case Op::TEMB: {
                TembWeights w;
                w.ch = c.ch; w.temb_ch = c.ch * 4; w.total_proj = e->total_proj;
                w.freqs = wptr<float>(e, "temb.freqs");
                w.w0_t = wptr<float>(e, "temb.w0_t"); w.b0 = wptr<float>(e, "temb.b0");
                w.w2_t = wptr<float>(e, "temb.w2_t"); w.b2 = wptr<float>(e, "temb.b2");
                w.wp_t = wptr<float>(e, "temb.wp_t"); w.bp = wptr<float>(e, "temb.bp");
                rc = launch_temb(w, t, batch, e->tproj, st);
                break;
            }
            case Op::TC: {
                TcOp tc = o.tc;
                tc.p.n_img = batch;
                tc.rp.n_img = batch;
                tc.pp.n_img = batch;            // launch_tc re-derives the CTA-pair decision from it
                if (i == e->end_op) {
                    PNPF_REQUIRE(v != nullptr, "null output pointer");
                    tc.p.epi.out = v;
                    tc.rp.epi.out = v;
                }
                rc = launch_tc(tc, st);
                break;
            }
            case Op::GN_STATS:
                break;
            case Op::GN_APPLY:
                rc = launch_gn_apply(o.gsrc, batch, o.HW, o.gamma, o.beta, GN_EPS, GROUPS, o.silu, o.dst, o.raw_dst, st);
                break;
            case Op::SOFTMAX:
                rc = launch_softmax_rows(o.S, o.P, o.rows_per_img * batch, o.L, st);
                break;
            case Op::ATTN:
                rc = launch_attn(o.attn, batch, st);
                break;
            case Op::UPSAMPLE:
                rc = launch_upsample2x(o.up_src, batch, o.up_side, o.up_side, o.up_C, o.dst, st);
                break;
        }
        if (rc) return rc;
    }
    return 0;
}

extern "C" int pnpf_unet_forward(pnpf_engine* e, const float* x, const float* t, float* v, int batch, void* stream) {
    PNPF_REQUIRE(e, "null engine");
    return run_ops(e, x, t, v, batch, (int)e->ops.size(), static_cast<cudaStream_t>(stream));
}

extern "C" int pnpf_debug_num_ops(pnpf_engine* e) { return e ? (int)e->ops.size() : 0; }
extern "C" const char* pnpf_debug_op_impl(pnpf_engine* e, int i) {
    if (!e || i < 0 || i >= (int)e->ops.size()) return nullptr;
    const Op& o = e->ops[i];
    if (!o.impl.empty()) return o.impl.c_str();
    switch (o.kind) {
        case Op::MEMSET: return "cudaMemsetAsync";
        case Op::IN_SHIM: return "nchw_to_nhwc_pad";
        case Op::TEMB: return "temb";
        case Op::SOFTMAX: return "softmax_rows";
        case Op::UPSAMPLE: return "upsample2x";
        default: return "?";
    }
}
extern "C" const char* pnpf_debug_op_name(pnpf_engine* e, int i) {
    return (e && i >= 0 && i < (int)e->ops.size()) ? e->ops[i].name.c_str() : nullptr;
}
extern "C" int pnpf_debug_forward_partial(pnpf_engine* e, const float* x, const float* t, int batch, int n_ops, void* stream) {
    PNPF_REQUIRE(e, "null engine");
    PNPF_REQUIRE(n_ops < (int)e->ops.size(), "partial forward must stop before the last op (use pnpf_unet_forward)");
    return run_ops(e, x, t, nullptr, batch, n_ops, static_cast<cudaStream_t>(stream));
}
extern "C" int pnpf_debug_read_op_output(pnpf_engine* e, int i, int batch, float* dst, size_t dst_elems, int dims[3], void* stream) {
    PNPF_REQUIRE(e && i >= 0 && i < (int)e->ops.size(), "bad op index");
    const Op& o = e->ops[i];
    PNPF_REQUIRE(o.out_C > 0, "op %d (%s) has no readable act16 NHWC output", i, o.name.c_str());
    const size_t n = (size_t)batch * o.out_C * o.out_side * o.out_side;
    dims[0] = o.out_C; dims[1] = o.out_side; dims[2] = o.out_side;
    PNPF_REQUIRE(dst_elems >= n, "destination too small");
    return launch_nhwc_to_nchw_f32(o.out_p, o.out_pitch, batch, o.out_C, o.out_side * o.out_side, dst, static_cast<cudaStream_t>(stream));
}
extern "C" double pnpf_unet_flops_per_image(pnpf_engine* e) { return e ? e->flops_per_img : 0.0; }
extern "C" int pnpf_unet_num_launches(pnpf_engine* e) { return e ? (int)e->ops.size() : 0; }

// Per-op device time of one forward, CUDA events on `stream` around every op (synchronous; not graph-capturable).
extern "C" int pnpf_profile_forward(pnpf_engine* e, const float* x, const float* t, float* v, int batch, float* host_ms,
                                    int n, void* stream) {
    PNPF_REQUIRE(e && host_ms && n == (int)e->ops.size(), "host_ms must hold pnpf_debug_num_ops() floats");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& q : ev) PNPF_CHECK_CUDA(cudaEventCreate(&q));
    int rc = 0;
    PNPF_CHECK_CUDA(cudaEventRecord(ev[0], st));
    // run op by op through the same dispatcher (n_ops = i+1 would re-run the prefix, so replicate the loop here)
    for (int i = 0; i < n && !rc; ++i) {
        std::vector<Op> one(1, e->ops[i]);
        std::swap(one, e->ops);
        const int end_saved = e->end_op;
        e->end_op = (i == end_saved) ? 0 : -1;
        rc = run_ops(e, x, t, v, batch, 1, st);
        e->end_op = end_saved;
        std::swap(one, e->ops);
        cudaEventRecord(ev[i + 1], st);
    }
    cudaError_t err = cudaStreamSynchronize(st);
    for (int i = 0; i < n; ++i) {
        float ms = 0.f;
        if (!rc && err == cudaSuccess) cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
        host_ms[i] = ms;
    }
    for (auto& q : ev) cudaEventDestroy(q);
    if (rc) return rc;
    PNPF_CHECK_CUDA(err);
    return 0;
}
// kind: 1 = tensor-core (conv_gemm) op, 0 = SIMT / memset; flops and algorithmic HBM bytes are PER IMAGE
extern "C" int pnpf_debug_op_info(pnpf_engine* e, int i, int* kind, double* flops, double* bytes) {
    PNPF_REQUIRE(e && i >= 0 && i < (int)e->ops.size() && kind && flops && bytes, "bad argument");
    *kind = (e->ops[i].kind == Op::TC || e->ops[i].kind == Op::ATTN) ? 1 : 0;
    *flops = e->ops[i].flops;
    *bytes = e->ops[i].bytes;
    return 0;
}
