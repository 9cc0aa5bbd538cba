// The forward-pass engine: static layer schedule of the four U-Nets, packed-weight blob
// layout, workspace plan and the launch sequence of Net.forward (model.py:59-65) -- plus the
// extern "C" surface declared in include/rrin_b200.h.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/rrin_b200.h"
#include "common.cuh"
#include "rrin_internal.h"

namespace rrin {

// ------------------------------------------------------------------ error text
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ------------------------------------------------------------------ static schedule
enum SrcKind { K_PLAIN = 0, K_CAT = 1, K_POOL = 2, K_UP = 3, K_HEAD = 4 };

struct Layer {
    std::string key;
    int unet;        // 0 Flow, 1 refine_flow, 2 Mask, 3 final   (execution order)
    int cin, cout;   // true channel counts (unet.py)
    int cin_pad;     // channels of the stored input tensor(s) (heads: 16)
    int level, src, act, out_f32, cfg;
    size_t w_off, b_off;   // byte offsets into the packed blob
};

struct UNetDef { const char* name; int cin, ncls, depth; };
// execution order (model.py:35,42,52,62); shapes from model.py:27-30
static const UNetDef kUNets[4] = {{"Flow", 6, 4, 5}, {"refine_flow", 10, 4, 4}, {"Mask", 16, 2, 4}, {"final", 9, 3, 4}};

struct Schedule {
    std::vector<Layer> layers;
    int first[5];          // first layer index of each U-Net, first[4] = total
    size_t blob_bytes;
    std::string error;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static Schedule build_schedule() {
    Schedule s;
    size_t off = 0;
    auto add = [&](int u, const std::string& key, int cin, int cout, int level, int src, int act, int out_f32) {
        Layer L;
        L.key = std::string(kUNets[u].name) + "." + key;
        L.unet = u; L.cin = cin; L.cout = cout; L.level = level; L.src = src; L.act = act; L.out_f32 = out_f32;
        L.cin_pad = (src == K_HEAD) ? 16 : cin;
        L.cfg = conv_select_config(L.cin_pad, out_f32 ? 16 : cout, out_f32);
        if (L.cfg < 0) s.error = "no conv configuration for " + L.key;
        L.w_off = off;
        off = align_up(off + (L.cfg < 0 ? 0 : conv_packed_weight_bytes(cout, L.cin_pad, L.cfg)), 256);
        L.b_off = off;
        off = align_up(off + (L.cfg < 0 ? 0 : (size_t)conv_packed_bias_count(cout, L.cfg) * 4), 256);
        s.layers.push_back(L);
    };
    for (int u = 0; u < 4; ++u) {
        s.first[u] = (int)s.layers.size();
        const int d = kUNets[u].depth;
        int prev = kUNets[u].cin;
        char k[64];
        for (int i = 0; i < d; ++i) {               // unet.py:24-28,42-46
            const int c = 32 << i;
            snprintf(k, sizeof k, "down_path.%d.block.0", i);
            add(u, k, prev, c, i, i == 0 ? K_HEAD : K_POOL, 1, 0);
            snprintf(k, sizeof k, "down_path.%d.block.2", i);
            add(u, k, c, c, i, K_PLAIN, 1, 0);
            prev = c;
        }
        add(u, "midconv", prev, prev, d - 1, K_PLAIN, 1, 0);   // unet.py:29,47
        for (int j = 0; j < d - 1; ++j) {           // unet.py:32-36,48-49,90-95
            const int lvl = d - 2 - j, c = 32 << lvl;
            snprintf(k, sizeof k, "up_path.%d.up.1", j);
            add(u, k, prev, c, lvl, K_UP, 0, 0);
            snprintf(k, sizeof k, "up_path.%d.conv_block.block.0", j);
            add(u, k, prev, c, lvl, K_CAT, 1, 0);
            snprintf(k, sizeof k, "up_path.%d.conv_block.block.2", j);
            add(u, k, c, c, lvl, K_PLAIN, 1, 0);
            prev = c;
        }
        add(u, "last", prev, kUNets[u].ncls, 0, K_PLAIN, 0, 1);   // unet.py:38,51
    }
    s.first[4] = (int)s.layers.size();
    s.blob_bytes = off;
    return s;
}

static const Schedule& schedule() {
    static const Schedule s = build_schedule();
    return s;
}

}  // namespace rrin

using namespace rrin;

// ------------------------------------------------------------------ engine object
struct rrin_engine {
    int Np, Nt, H, W, pair_mul;
    size_t ws_bytes;
    // workspace offsets (bytes)
    size_t off_tmp[3], off_skip[4], off_h16, off_flow4, off_u4, off_out4, off_xt8;
    int launches;
    std::vector<cudaEvent_t>* prof = nullptr;   // when set, an event is recorded after every launch
    int prof_n = 0;
};

static inline void mark(const rrin_engine* e, cudaStream_t st) {
    if (e->prof && e->prof_n < (int)e->prof->size()) cudaEventRecord((*e->prof)[const_cast<rrin_engine*>(e)->prof_n++], st);
}

static size_t lvl_bytes(int B, int H, int W, int lvl) {   // bf16 NHWC tensor of level lvl
    return (size_t)B * (H >> lvl) * (W >> lvl) * (32 << lvl) * 2;
}

extern "C" {

int rrin_version(void) { return 100; }
const char* rrin_last_error(void) { return g_err; }

int rrin_num_convs(void) { return (int)schedule().layers.size(); }

int rrin_conv_info(int idx, char* key, int key_cap, int* cin, int* cout, int* level, int* src_mode, int* act) {
    const Schedule& s = schedule();
    if (idx < 0 || idx >= (int)s.layers.size()) { set_error("rrin_conv_info: index %d out of range", idx); return RRIN_ERR_BAD_ARG; }
    const Layer& L = s.layers[idx];
    if (key && key_cap > 0) { strncpy(key, L.key.c_str(), key_cap - 1); key[key_cap - 1] = 0; }
    if (cin) *cin = L.cin;
    if (cout) *cout = L.cout;
    if (level) *level = L.level;
    if (src_mode) *src_mode = L.src;
    if (act) *act = L.act;
    return RRIN_OK;
}

size_t rrin_packed_weights_bytes(void) { return schedule().blob_bytes; }

int rrin_pack_conv(int idx, const float* w, const float* b, void* blob, void* stream) {
    const Schedule& s = schedule();
    if (!s.error.empty()) { set_error("%s", s.error.c_str()); return RRIN_ERR_UNSUPPORTED; }
    if (idx < 0 || idx >= (int)s.layers.size() || !w || !b || !blob) { set_error("rrin_pack_conv: bad argument"); return RRIN_ERR_BAD_ARG; }
    const Layer& L = s.layers[idx];
    uint8_t* base = static_cast<uint8_t*>(blob);
    return conv_pack_weights(w, b, L.cout, L.cin, L.cin_pad, L.cfg, base + L.w_off, reinterpret_cast<float*>(base + L.b_off),
                             static_cast<cudaStream_t>(stream));
}

int rrin_engine_create(int n_pairs, int n_samples, int H, int W, rrin_engine** out) {
    if (!out) { set_error("rrin_engine_create: null out"); return RRIN_ERR_BAD_ARG; }
    *out = nullptr;
    if (n_pairs <= 0 || n_samples <= 0 || H <= 0 || W <= 0) { set_error("rrin_engine_create: empty shape"); return RRIN_ERR_BAD_SHAPE; }
    if (H % 16 || W % 16) { set_error("H and W must be multiples of 16 (got %dx%d)", H, W); return RRIN_ERR_BAD_SHAPE; }
    if (n_pairs != n_samples && n_pairs != 1) { set_error("n_pairs must equal n_samples or be 1 (got %d, %d)", n_pairs, n_samples); return RRIN_ERR_BAD_SHAPE; }
    if (!schedule().error.empty()) { set_error("%s", schedule().error.c_str()); return RRIN_ERR_UNSUPPORTED; }
    rrin_engine* e = new rrin_engine();
    e->Np = n_pairs; e->Nt = n_samples; e->H = H; e->W = W;
    e->pair_mul = (n_pairs == n_samples) ? 1 : 0;
    const int B = n_pairs > n_samples ? n_pairs : n_samples;
    const size_t px = (size_t)H * W;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    for (int i = 0; i < 3; ++i) e->off_tmp[i] = take(lvl_bytes(B, H, W, 0));
    for (int l = 0; l < 4; ++l) e->off_skip[l] = take(lvl_bytes(B, H, W, l));
    e->off_h16 = take((size_t)B * px * 16 * 2);
    e->off_flow4 = take((size_t)n_pairs * px * 16);
    e->off_u4 = take((size_t)n_samples * px * 16);
    e->off_out4 = take((size_t)n_samples * px * 16);
    e->off_xt8 = take((size_t)n_samples * px * 32);
    e->ws_bytes = off;
    e->launches = rrin_num_convs() + 5;
    *out = e;
    return RRIN_OK;
}

void rrin_engine_destroy(rrin_engine* e) { delete e; }
size_t rrin_engine_workspace_bytes(const rrin_engine* e) { return e ? e->ws_bytes : 0; }
int rrin_engine_num_launches(const rrin_engine* e) { return e ? e->launches : 0; }

// One U-Net (unet.py:40-51) at batch B: head16 -> fp32 NHWC4 `out4`.
static int run_unet(const rrin_engine* e, int u, int B, const uint8_t* blob, uint8_t* ws, const void* head16, float* out4,
                    cudaStream_t st) {
    const Schedule& s = schedule();
    const int d = kUNets[u].depth;
    void* tmp[3] = {ws + e->off_tmp[0], ws + e->off_tmp[1], ws + e->off_tmp[2]};
    int li = s.first[u];
    auto conv = [&](const void* src0, const void* src1, int c0, int c1, int mode, int lvl, void* out) -> int {
        const Layer& L = s.layers[li++];
        ConvDesc cd;
        cd.src0 = src0; cd.src1 = src1; cd.c0 = c0; cd.c1 = c1; cd.mode = mode;
        cd.N = B; cd.H = e->H >> lvl; cd.W = e->W >> lvl;
        cd.cout = L.cout; cd.wpack = blob + L.w_off; cd.bias = reinterpret_cast<const float*>(blob + L.b_off);
        cd.out = out; cd.out_f32 = L.out_f32; cd.act = L.act; cd.cfg = L.cfg;
        const int r = conv_launch(cd, st);
        mark(e, st);
        return r;
    };
    const void* x = head16;
    int xc = 16;
    for (int i = 0; i < d; ++i) {
        const int c = 32 << i;
        if (int r = conv(x, nullptr, xc, 0, i == 0 ? SRC_PLAIN : SRC_POOL, i, tmp[0])) return r;
        void* o = (i < d - 1) ? (void*)(ws + e->off_skip[i]) : tmp[1];
        if (int r = conv(tmp[0], nullptr, c, 0, SRC_PLAIN, i, o)) return r;
        x = o; xc = c;
    }
    if (int r = conv(x, nullptr, xc, 0, SRC_PLAIN, d - 1, tmp[0])) return r;   // midconv
    int cur = 0;                                                                // x lives in tmp[cur]
    for (int j = 0; j < d - 1; ++j) {
        const int lvl = d - 2 - j, c = 32 << lvl;
        const int ui = (cur + 1) % 3, vi = (cur + 2) % 3;
        if (int r = conv(tmp[cur], nullptr, 2 * c, 0, SRC_UP, lvl, tmp[ui])) return r;              // up.1
        if (int r = conv(tmp[ui], ws + e->off_skip[lvl], c, c, SRC_CAT, lvl, tmp[vi])) return r;    // block.0 on cat(up, skip)
        if (int r = conv(tmp[vi], nullptr, c, 0, SRC_PLAIN, lvl, tmp[cur])) return r;               // block.2
    }
    return conv(tmp[cur], nullptr, 32, 0, SRC_PLAIN, 0, out4);                                      // last
}

int rrin_engine_forward(rrin_engine* e, const void* blob_, void* workspace, const float* in0, const float* in1,
                        const float* coef, float* out, void* stream) {
    if (!e || !blob_ || !workspace || !in0 || !in1 || !coef || !out) { set_error("rrin_engine_forward: null argument"); return RRIN_ERR_BAD_ARG; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint8_t* blob = static_cast<const uint8_t*>(blob_);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    void* h16 = ws + e->off_h16;
    float* flow4 = reinterpret_cast<float*>(ws + e->off_flow4);
    float* u4 = reinterpret_cast<float*>(ws + e->off_u4);
    float* out4 = reinterpret_cast<float*>(ws + e->off_out4);
    float* xt8 = reinterpret_cast<float*>(ws + e->off_xt8);
    const int H = e->H, W = e->W, Np = e->Np, Nt = e->Nt, pm = e->pair_mul;
    int r;
    mark(e, st);                                                                                  // t0
    if ((r = pack_pair(in0, in1, Np, H, W, h16, st))) return r;                                   // model.py:33
    mark(e, st);
    if ((r = run_unet(e, 0, Np, blob, ws, h16, flow4, st))) return r;                             // model.py:35
    if ((r = flow_tscale_pack(flow4, in0, in1, coef, Nt, pm, H, W, h16, st))) return r;           // model.py:37-41
    mark(e, st);
    if ((r = run_unet(e, 1, Nt, blob, ws, h16, u4, st))) return r;                                // model.py:42
    if ((r = warp_pack(flow4, u4, in0, in1, coef, Nt, pm, H, W, h16, xt8, st))) return r;         // model.py:44-50
    mark(e, st);
    if ((r = run_unet(e, 2, Nt, blob, ws, h16, u4, st))) return r;                                // model.py:52
    if ((r = blend_pack(u4, xt8, in0, in1, coef, Nt, pm, H, W, out4, h16, st))) return r;         // model.py:52-55,61
    mark(e, st);
    if ((r = run_unet(e, 3, Nt, blob, ws, h16, u4, st))) return r;                                // model.py:62
    r = residue_clamp(u4, out4, Nt, H, W, out, st);                                               // model.py:62-63
    mark(e, st);
    return r;
}

// Launch i of one forward, in stream order: kernel class name, the reference layer it computes,
// algorithmic FLOPs and algorithmic HBM bytes (each operand/result tensor crossing once).
int rrin_engine_launch_info(const rrin_engine* e, int i, char* name, int name_cap, char* layer, int layer_cap,
                            double* flops, double* bytes) {
    if (!e || i < 0 || i >= e->launches) { set_error("rrin_engine_launch_info: bad index %d", i); return RRIN_ERR_BAD_ARG; }
    const Schedule& s = schedule();
    const double px = (double)e->H * e->W;
    std::string nm, ly; double fl = 0, by = 0;
    // glue launches sit before U-Net 0, between U-Nets, and at the end
    int pos = i, li = -1;
    const char* glue_names[5] = {"pack_pair", "flow_tscale_pack", "warp_pack", "blend_pack", "residue_clamp"};
    const double glue_bytes[5] = {56, 72, 120, 120, 44};          // per pixel, see DESIGN.md
    int g = -1;
    for (int u = 0; u < 4 && g < 0 && li < 0; ++u) {
        if (pos == 0) { g = u; break; }
        pos -= 1;
        const int nl = s.first[u + 1] - s.first[u];
        if (pos < nl) { li = s.first[u] + pos; break; }
        pos -= nl;
    }
    if (g < 0 && li < 0) g = 4;
    if (g >= 0) {
        nm = glue_names[g]; ly = std::string("model.py glue: ") + glue_names[g];
        by = glue_bytes[g] * px * (g == 0 ? e->Np : e->Nt);
    } else {
        const Layer& L = s.layers[li];
        int kc, nt, msub; conv_config_info(L.cfg, &kc, &nt, &msub);
        char b[96]; snprintf(b, sizeof b, "conv3x3_umma<KC%d,NT%d,MSUB%d>", kc, nt, msub);
        nm = b; ly = L.key;
        const int B = (L.unet == 0) ? e->Np : e->Nt;
        const double lp = px * B / (double)(1 << (2 * L.level));
        fl = 2.0 * 9 * L.cin * L.cout * lp;
        double in_b = 2.0 * L.cin_pad * lp;                       // bf16 operand tensor(s) at this level
        if (L.src == K_POOL) in_b *= 4; else if (L.src == K_UP) in_b /= 4;
        const double out_b = L.out_f32 ? 16.0 * lp : 2.0 * L.cout * lp;
        by = in_b + out_b + (double)conv_packed_weight_bytes(L.cout, L.cin_pad, L.cfg);
    }
    if (name && name_cap > 0) { strncpy(name, nm.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
    if (layer && layer_cap > 0) { strncpy(layer, ly.c_str(), layer_cap - 1); layer[layer_cap - 1] = 0; }
    if (flops) *flops = fl;
    if (bytes) *bytes = by;
    return RRIN_OK;
}

// One forward with a CUDA event after every launch; ms[i] = device time of launch i.
// Synchronises the stream (profiling aid for bench.py, not part of the product path).
int rrin_engine_forward_profiled(rrin_engine* e, const void* blob, void* workspace, const float* in0, const float* in1,
                                 const float* coef, float* out, void* stream, float* ms_host) {
    if (!e || !ms_host) { set_error("rrin_engine_forward_profiled: null argument"); return RRIN_ERR_BAD_ARG; }
    std::vector<cudaEvent_t> ev(e->launches + 1);
    for (auto& x : ev) RRIN_CUDA_CHECK(cudaEventCreate(&x));
    e->prof = &ev; e->prof_n = 0;
    int r = rrin_engine_forward(e, blob, workspace, in0, in1, coef, out, stream);
    e->prof = nullptr;
    if (r == RRIN_OK) {
        cudaError_t ce = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
        if (ce != cudaSuccess) { set_error("profiled forward failed: %s", cudaGetErrorString(ce)); r = RRIN_ERR_CUDA; }
    }
    if (r == RRIN_OK)
        for (int i = 0; i < e->launches; ++i) cudaEventElapsedTime(&ms_host[i], ev[i], ev[i + 1]);
    for (auto& x : ev) cudaEventDestroy(x);
    return r;
}

int rrin_engine_tap(const rrin_engine* e, const void* workspace, int which, float* dst, void* stream) {
    if (!e || !workspace || !dst) { set_error("rrin_engine_tap: null argument"); return RRIN_ERR_BAD_ARG; }
    const uint8_t* ws = static_cast<const uint8_t*>(workspace);
    const size_t px = (size_t)e->H * e->W;
    const void* src; size_t bytes;
    if (which == 0) { src = ws + e->off_flow4; bytes = (size_t)e->Np * px * 16; }
    else if (which == 1) { src = ws + e->off_out4; bytes = (size_t)e->Nt * px * 16; }
    else { set_error("rrin_engine_tap: unknown tap %d", which); return RRIN_ERR_BAD_ARG; }
    RRIN_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
    return RRIN_OK;
}

// ------------------------------------------------------------------ unit-level wrappers
int rrin_conv_select_config(int cin, int cout, int out_f32) { return conv_select_config(cin, cout, out_f32); }
size_t rrin_conv_packed_weight_bytes(int cout, int cin_pad, int cfg) { return conv_packed_weight_bytes(cout, cin_pad, cfg); }
int rrin_conv_packed_bias_count(int cout, int cfg) { return conv_packed_bias_count(cout, cfg); }
int rrin_pack_conv_raw(const float* w, const float* b, int cout, int cin, int cin_pad, int cfg, void* wpack, float* bias_pack, void* stream) {
    return conv_pack_weights(w, b, cout, cin, cin_pad, cfg, wpack, bias_pack, static_cast<cudaStream_t>(stream));
}
int rrin_conv3x3(const void* src0, const void* src1, int c0, int c1, int src_mode, int N, int H, int W, int cout,
                 const void* wpack, const float* bias_pack, void* out, int out_f32, int act, int cfg, void* stream) {
    ConvDesc cd;
    cd.src0 = src0; cd.src1 = src1; cd.c0 = c0; cd.c1 = c1; cd.mode = src_mode; cd.N = N; cd.H = H; cd.W = W;
    cd.cout = cout; cd.wpack = wpack; cd.bias = bias_pack; cd.out = out; cd.out_f32 = out_f32; cd.act = act; cd.cfg = cfg;
    return conv_launch(cd, static_cast<cudaStream_t>(stream));
}
int rrin_pack_pair(const float* in0, const float* in1, int N, int H, int W, void* x16, void* stream) {
    return pack_pair(in0, in1, N, H, W, x16, static_cast<cudaStream_t>(stream));
}
int rrin_flow_tscale_pack(const float* flow4, const float* in0, const float* in1, const float* coef, int n, int pair_mul,
                          int H, int W, void* r16, void* stream) {
    return flow_tscale_pack(flow4, in0, in1, coef, n, pair_mul, H, W, r16, static_cast<cudaStream_t>(stream));
}
int rrin_warp_pack(const float* flow4, const float* res4, const float* in0, const float* in1, const float* coef, int n,
                   int pair_mul, int H, int W, void* m16, float* xt8, void* stream) {
    return warp_pack(flow4, res4, in0, in1, coef, n, pair_mul, H, W, m16, xt8, static_cast<cudaStream_t>(stream));
}
int rrin_blend_pack(const float* mask4, const float* xt8, const float* in0, const float* in1, const float* coef, int n,
                    int pair_mul, int H, int W, float* out4, void* f16, void* stream) {
    return blend_pack(mask4, xt8, in0, in1, coef, n, pair_mul, H, W, out4, f16, static_cast<cudaStream_t>(stream));
}
int rrin_residue_clamp(const float* res4, const float* out4, int n, int H, int W, float* out_nchw, void* stream) {
    return residue_clamp(res4, out4, n, H, W, out_nchw, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
