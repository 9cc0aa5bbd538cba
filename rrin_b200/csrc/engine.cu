// The forward-pass engine: static layer schedule of the four U-Nets, packed-weight blob
// layout, workspace plan and the launch sequence of Net.forward (model.py:59-65) -- plus the
// extern "C" surface declared in include/rrin_b200.h.
//
// Tensor layouts inside a U-Net:
//   level 0 (full resolution, 32 channels): space-to-depth on the half-res grid, i.e. bf16
//           [N, H/2, W/2, 4 phases, 32] == NHWC with 128 channels at half resolution;
//   level l >= 1 (32*2^l channels): bf16 NHWC [N, H/2^l, W/2^l, C].
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

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

bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("RRIN_PDL"); return !(e && e[0] == '0'); }();
    return on;
}

// ------------------------------------------------------------------ static schedule
enum SrcKind { K_PLAIN = 0, K_CAT = 1, K_POOL = 2, K_UP = 3, K_HEAD = 4 };

// One packed copy of a conv's weights (a layer has a second one when its upsample is folded:
// the folded weights for the interior plus the plain weights for the exact border ring).
struct Pack {
    int kind = -1, cfg = -1, n_stages = 0, sched = SCHED_TAPS9, n_cols = 0;
    size_t w_off = 0, b_off = 0;
};

struct Layer {
    std::string key;
    int unet;        // 0 Flow, 1 refine_flow, 2 Mask, 3 final   (execution order)
    int cin, cout;   // true channel counts (unet.py)
    int level, src, act, is_last;
    Pack main, fold, strip; // `fold` / `strip` only for up.1 convs writing level 0 or 1 (folded interior + exact border ring)
};

struct UNetDef { const char* name; int cin, ncls, depth; };
// execution order (model.py:35,42,52,62); shapes from model.py:27-30
static const UNetDef kUNets[4] = {{"Flow", 6, 4, 5}, {"refine_flow", 10, 4, 4}, {"Mask", 16, 2, 4}, {"final", 9, 3, 4}};

struct Schedule {
    std::vector<Layer> layers;
    int first[5];          // first layer index of each U-Net, first[4] = total
    size_t blob_bytes;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// levels >= 2 plain / cat convs: single-CTA tile (config 16).  RRIN_BIG_CFG=19 selects the CTA-pair (cta_group::2) variant, which
// is parity-tested but measured ~10 % slower at 1080p (profiles/README.md)
static int big_cfg() {
    const char* e = getenv("RRIN_BIG_CFG");
    return (e && atoi(e) == 19) ? 19 : 16;
}

// exact bilinear upsample convs of levels >= 2: TMA-staged transform (config 20); RRIN_UP_CFG=5 selects the register-computed
// producer kernel (A/B timing runs)
static int up_cfg() {
    const char* e = getenv("RRIN_UP_CFG");
    return (e && atoi(e) == 5) ? 5 : 20;
}

// level-1 block.0 conv on the pooled 32-channel level-0 tensor: TMA config 21 (RRIN_POOL1_CFG=3: the cp.async-fed v1 kernel)
static int pool1_cfg() {
    static const int c = getenv("RRIN_POOL1_CFG") ? atoi(getenv("RRIN_POOL1_CFG")) : 21;
    return (c == 3 || c == 39 || c == 45 || c == 46) ? c : 21;
}

// level-1 plain / cat convs on CTA pairs (config 22): experimental, RRIN_L1_PAIR=1
static bool l1_pair() {
    static const bool on = getenv("RRIN_L1_PAIR") && atoi(getenv("RRIN_L1_PAIR")) != 0;
    return on;
}
// level-1 64->64 / cat(64+64)->64 config ids.  64->64: two tile streams per CTA on 16 x 8 tiles (config 36; N = 64 MMAs run 48
// cycles, one thread issues one per ~39: 0.21 -> 0.18 ms against the single-issuer config 14).  cat(64+64)->64: CTA pairs with
// streamed weights (config 22: one issuer drives two SMs, half of every weight block per SM; 0.35 -> 0.31 ms against config 15).
// RRIN_L1_CFG=<a>,<b> with a in
// {14, 28, 30, 36, 43}, b in {15, 22, 29, 31, 41, 42} (22, 28-31: CTA pairs; 36, 41-43: two tile streams)
static void l1_cfgs(int& c64, int& ccat) {
    static int a = -1, b = -1;
    if (a < 0) {
        a = 36; b = 22;
        const char* e = getenv("RRIN_L1_CFG");
        int x = 0, y = 0;
        if (e && sscanf(e, "%d,%d", &x, &y) == 2) {
            if (x == 14 || x == 28 || x == 30 || x == 36 || x == 43) a = x;
            if (y == 15 || y == 22 || y == 29 || y == 31 || y == 41 || y == 42) b = y;
        }
    }
    c64 = a; ccat = b;
}

// Level-0 cat(32+32)->32 conv: CTA pairs that keep the layer's 192 KB of space-to-depth weights resident as two 96 KB halves
// (config 25; measured 0.44 -> 0.33 ms per 4-pair launch against the single-CTA config 12, which re-streams them per tile).
// Level-0 32->32 conv: its 96 KB of weights fit one SM; two tile streams per CTA (config 35: two MMA-issuing warps; 0.27 / 0.25 ->
// 0.24 / 0.21 ms against the single-issuer config 11; the pair variants 23 / 26 / 27 measured slower).
// RRIN_L0_PAIR=0 selects 11 / 12; RRIN_L0_PAIR=<a>,<b> picks explicit ids (A/B runs: a in {11,23,26,27,35}, b in {12,24,25}).
static void l0_pair_cfgs(int& c32, int& ccat) {
    static int a = -1, b = -1;
    if (a < 0) {
        a = 35; b = 25;
        const char* e = getenv("RRIN_L0_PAIR");
        if (e) {
            int x = 0, y = 0;
            const int n = sscanf(e, "%d,%d", &x, &y);
            if (n == 1 && x == 0) { a = 11; b = 12; }
            else if (n == 2) { a = (x == 23 || x == 26 || x == 27 || x == 11 || x == 35) ? x : 35; b = (y == 24 || y == 25 || y == 12 || y == 47) ? y : 25; }
        }
    }
    c32 = a; ccat = b;
}

// transposed launches for levels >= 2 (RRIN_TRANSPOSE=0: never)
constexpr long kBandRows = 16;          // conv3x3.cuh kTileH
static bool transpose_ok() {
    static const bool on = [] { const char* e = getenv("RRIN_TRANSPOSE"); return !(e && e[0] == '0'); }();
    return on;
}

// refine_flow.last + fused backward warps: config 44 stages the frames' tile windows in shared memory (RRIN_WARP_STAGE=0: gather
// from global memory like the other `last` epilogues)
static bool warp_staging() {
    static const bool on = [] { const char* e = getenv("RRIN_WARP_STAGE"); return !(e && e[0] == '0'); }();
    return on;
}

static Schedule build_schedule() {
    Schedule s;
    size_t off = 0;
    auto place = [&](Pack& p) {
        p.w_off = off;
        off = align_up(off + conv_packed_weight_bytes(p.cfg, p.n_cols, p.n_stages, p.sched), 256);
        p.b_off = off;
        off = align_up(off + (size_t)conv_packed_bias_count(p.cfg, p.n_cols) * 4, 256);
    };
    auto add = [&](int u, const std::string& key, int cin, int cout, int level, int src, int act, int is_last) {
        Layer L;
        L.key = std::string(kUNets[u].name) + "." + key;
        L.unet = u; L.cin = cin; L.cout = cout; L.level = level; L.src = src; L.act = act; L.is_last = is_last;
        Pack& m = L.main;
        // configs 0..6: transform kernel (conv3x3.cuh: pool / bilinear sources, exact border ring);
        // configs 10..16: TMA-fed kernel (conv3x3_v2.cuh: every source that is a stored tensor as-is)
        if (level == 0) {                       // space-to-depth: 4 phases x cout columns on the half-res grid
            m.n_cols = is_last ? 16 : 128;
            static const int head_cfg = [] { const char* e = getenv("RRIN_HEAD_CFG"); return (e && atoi(e) == 38) ? 38 : 10; }();
            if (src == K_HEAD) { m.kind = PACK_S2D; m.sched = SCHED_S2D16; m.cfg = head_cfg; m.n_stages = 1; }
            else if (src == K_UP) { m.kind = PACK_S2D; m.sched = SCHED_S2D16; m.cfg = 1; m.n_stages = cin / 32; }   // exact bilinear (ring / small frames)
            else {
                int c32, ccat;
                l0_pair_cfgs(c32, ccat);
                static const int last_cfg = [] { const char* e = getenv("RRIN_LAST_CFG"); const int v = e ? atoi(e) : 37; return (v == 13 || (v >= 32 && v <= 34) || v == 37 || v == 40) ? v : 37; }();
                m.kind = PACK_S2D8; m.sched = SCHED_S2D8; m.cfg = is_last ? last_cfg : (cin == 32 ? c32 : ccat); m.n_stages = 2 * (cin / 32);
                if ((m.cfg >= 23 && m.cfg <= 27) || m.cfg == 47) m.kind = PACK_S2D8_CG2;
            }
        } else {
            m.kind = PACK_NORMAL; m.sched = SCHED_TAPS9; m.n_cols = cout;
            // K_POOL sources are read from the pooled copy the previous level's block.2 epilogue wrote: plain convs
            if (level == 1) {
                if (src == K_POOL) { m.cfg = pool1_cfg(); m.n_stages = 1; }      // 32 stored channels: 64-byte TMA rows (config 21)
                else {
                    int c64, ccat;
                    l1_cfgs(c64, ccat);
                    m.cfg = (src == K_UP) ? 4 : (cin == 64 ? c64 : ccat); m.n_stages = cin / 64;
                    if (src != K_UP && l1_pair()) m.cfg = 22;
                    if (m.cfg == 22 || (m.cfg >= 28 && m.cfg <= 31)) m.kind = PACK_NORMAL_CG2;
                }
            } else {
                m.cfg = (src == K_UP) ? up_cfg() : big_cfg(); m.n_stages = cin / 64;
                if (m.cfg == 19) m.kind = PACK_NORMAL_CG2;
            }
        }
        place(m);
        if (src == K_UP && level <= 1) {        // folded bilinear x2: runs on the coarser grid with 4*cout columns
            Pack& f = L.fold;
            f.kind = PACK_FOLD; f.sched = SCHED_TAPS9; f.cfg = (level == 0) ? 18 : 17; f.n_stages = cin / 64; f.n_cols = 4 * cout;   // 17: scatter epilogue
            place(f);
            Pack& r = L.strip;                    // exact bilinear on 128-pixel border strips (conv3x3.cuh, STRIP configs)
            if (level == 0) { r.kind = PACK_S2D; r.sched = SCHED_S2D16; r.cfg = 8; r.n_stages = cin / 16; r.n_cols = 128; }
            else { r.kind = PACK_NORMAL; r.sched = SCHED_TAPS9; r.cfg = 7; r.n_stages = cin / 64; r.n_cols = cout; }
            place(r);
        }
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

// One kernel launch of a forward pass, resolved against the workspace at run time.
struct Launch {
    int glue = -1;           // >= 0: glue kernel id (0 pack_pair .. 4 residue_clamp)
    int layer = -1;          // conv: index into schedule().layers
    int use_fold = 0;        // conv: which Pack (0 main, 1 fold, 2 strip)
    int fuse_mode = 0;       // `last` convs: glue fused into the epilogue (conv3x3_v2.cuh FuseParams::mode), 0 = none
    ConvDesc cd;             // pointers hold workspace OFFSETS (+1 so that 0 stays "null") until launch
    alignas(64) unsigned char tmap[3][128];   // TMA configs: tensor maps of src0 / src1 / out, encoded for `tmap_ws`
    std::string name;
    double flops = 0, bytes = 0;
};

}  // namespace rrin

using namespace rrin;

// ------------------------------------------------------------------ engine object
struct rrin_engine {
    int Np, Nt, H, W, pair_mul;
    int f16 = 0;                                // operand / activation format (RRIN_PRECISION_*)
    size_t ws_bytes;
    size_t off_tmp[3], off_skip[4], off_pool[4], off_h16, off_flow4, off_u4, off_out4, off_xt8;
    std::vector<Launch> launches;
    const void* tmap_ws = nullptr;              // workspace base the cached tensor maps were encoded for
    std::vector<cudaEvent_t>* prof = nullptr;   // when set, an event is recorded after every launch
    int prof_n = 0;
    // rrin_engine_forward_graph: instantiated CUDA graphs of the launch sequence, keyed by every pointer baked into it
    struct GraphEntry {
        const void* key[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // blob, workspace, in0, in1, coef, out
        cudaGraphExec_t exec = nullptr;
        unsigned long long last_use = 0;
        int dev = -1;
    };
    std::vector<GraphEntry> graphs;
    cudaStream_t cap_stream = nullptr;          // capture happens on a private stream (torch's default stream is the legacy stream,
    int cap_dev = -1;                           // which cannot be captured); the instantiated graph launches into any stream
    unsigned long long tick = 0;
    int graph_launches = 0, direct_launches = 0, captures = 0;
};

static inline void mark(rrin_engine* e, cudaStream_t st) {
    if (e->prof && e->prof_n < (int)e->prof->size()) cudaEventRecord((*e->prof)[e->prof_n++], st);
}

static size_t lvl_bytes(int B, int H, int W, int lvl) {   // bf16 activation tensor of level lvl (either layout)
    return (size_t)B * (H >> lvl) * (W >> lvl) * (32 << lvl) * 2;
}

static void* enc(size_t off) { return reinterpret_cast<void*>(off + 1); }   // workspace offset as a tagged pointer

// Builds the conv launches of one U-Net at batch B: head16 -> fp32 out4.
static void plan_unet(rrin_engine* e, int u, int B, size_t head_off, size_t out_off) {
    const Schedule& s = schedule();
    const int d = kUNets[u].depth;
    const int H = e->H, W = e->W;
    int li = s.first[u];
    const bool fold_ok = (H > 64 && W > 64);     // the exact border ring needs more than 2x2 tiles of 32x32 pixels
    auto base = [&](int layer, int use_fold, int grid_lvl) {
        const Layer& L = s.layers[layer];
        const Pack& pk = use_fold == 1 ? L.fold : (use_fold == 2 ? L.strip : L.main);
        Launch ln;
        ln.layer = layer; ln.use_fold = use_fold;
        ln.cd.N = B; ln.cd.H = H >> grid_lvl; ln.cd.W = W >> grid_lvl;
        ln.cd.sched = pk.sched; ln.cd.n_cols = pk.n_cols; ln.cd.cfg = pk.cfg; ln.cd.act = L.act;
        return ln;
    };
    auto finish = [&](Launch& ln, const char* tag, bool counts_flops) {
        const Layer& L = s.layers[ln.layer];
        // Levels >= 2 (streamed 9-tap configs with TMA stores): run the kernel's 16-row bands along whichever image dimension pads
        // less (1088 = 17 * 64: the level-2..4 tensors have 136 / 68 / 34 rows but widths that are multiples of 16)
        if ((ln.cd.cfg == 16 || ln.cd.cfg == 20) && ln.cd.epi == EPI_BF16 && transpose_ok()) {
            auto padded = [](long h, long w) { return ((h + kBandRows - 1) / kBandRows * kBandRows) * ((w + 7) / 8 * 8); };
            if (padded(ln.cd.W, ln.cd.H) < padded(ln.cd.H, ln.cd.W)) ln.cd.transposed = 1;
        }
        int kcs, kb, nt, msub;
        conv_config_info(ln.cd.cfg, &kcs, &kb, &nt, &msub);
        char b[128];
        snprintf(b, sizeof b, "conv3x3_%s#%d<KCS%d,KB%d,NT%d,MSUB%d>%s", ln.cd.cfg >= 10 ? "tma" : "umma", ln.cd.cfg, kcs, kb, nt, msub, tag);
        ln.name = b;
        const double lp = (double)B * (H >> L.level) * (W >> L.level);      // output pixels of the reference conv
        if (counts_flops) {
            ln.flops = 2.0 * 9 * L.cin * L.cout * lp;
            double in_b = 2.0 * (L.src == K_HEAD ? 16 : L.cin) * lp;
            if (L.src == K_UP) in_b /= 4;
            if (ln.cd.pool_out) in_b += 2.0 * L.cout * lp / 4;                  // pooled second output
            const Pack& pk = ln.use_fold == 1 ? L.fold : (ln.use_fold == 2 ? L.strip : L.main);
            ln.bytes = in_b + (L.is_last ? 16.0 * lp : 2.0 * L.cout * lp) + (double)conv_packed_weight_bytes(pk.cfg, pk.n_cols, pk.n_stages, pk.sched);
        }
        e->launches.push_back(ln);
    };
    const size_t tmp[3] = {e->off_tmp[0], e->off_tmp[1], e->off_tmp[2]};
    // ---- encoder
    size_t x = head_off;
    for (int i = 0; i < d; ++i) {
        const int c = 32 << i;
        {   // block.0
            Launch ln = base(li, 0, i == 0 ? 1 : i);
            ln.cd.mode = SRC_PLAIN;                    // level >= 1: avg_pool2d (unet.py:46) was applied by the previous block.2 epilogue
            if (i == 0) { ln.cd.c0 = 64; ln.cd.cout_stride = 128; }
            else { ln.cd.c0 = c / 2; ln.cd.cout_stride = c; }
            ln.cd.src0 = enc(i == 0 ? x : e->off_pool[i - 1]); ln.cd.out = enc(tmp[0]); ln.cd.epi = EPI_BF16;
            finish(ln, "", true); ++li;
        }
        {   // block.2
            Launch ln = base(li, 0, i == 0 ? 1 : i);
            ln.cd.mode = SRC_PLAIN; ln.cd.c0 = (i == 0) ? 128 : c; ln.cd.cout_stride = (i == 0) ? 128 : c;
            const size_t o = (i < d - 1) ? e->off_skip[i] : tmp[1];
            ln.cd.src0 = enc(tmp[0]); ln.cd.out = enc(o); ln.cd.epi = EPI_BF16;
            if (i < d - 1) ln.cd.pool_out = enc(e->off_pool[i]);               // skip + its 2x2 mean for the next level
            finish(ln, "", true); ++li;
            x = o;
        }
    }
    {   // midconv at level d-1 (>= 3)
        Launch ln = base(li, 0, d - 1);
        ln.cd.mode = SRC_PLAIN; ln.cd.c0 = 32 << (d - 1); ln.cd.cout_stride = ln.cd.c0;
        ln.cd.src0 = enc(x); ln.cd.out = enc(tmp[0]); ln.cd.epi = EPI_BF16;
        finish(ln, "", true); ++li;
    }
    // ---- decoder
    int cur = 0;
    for (int j = 0; j < d - 1; ++j) {
        const int lvl = d - 2 - j, c = 32 << lvl;
        const int ui = (cur + 1) % 3, vi = (cur + 2) % 3;
        // up.1 (no activation): input tmp[cur] is the level lvl+1 tensor with 2c channels
        if (lvl <= 1 && fold_ok) {
            Launch f = base(li, 1, lvl + 1);          // folded: runs on the coarse grid; its outermost 2 hi-res pixels are
            f.cd.mode = SRC_PLAIN; f.cd.c0 = 2 * c;   // wrong (zero fill instead of the reference border) and rewritten by the ring
            f.cd.src0 = enc(tmp[cur]); f.cd.out = enc(tmp[ui]);
            if (lvl == 1) { f.cd.epi = EPI_SCATTER; f.cd.cout_stride = c; }
            else { f.cd.epi = EPI_BF16; f.cd.cout_stride = 128; }     // (phase, co) columns == space-to-depth pixel
            finish(f, " fold", true);
            Launch r = base(li, 2, 1);                // exact bilinear + conv on the border ring, in 128-pixel strips
            r.cd.mode = (lvl == 0) ? SRC_UP_S2D : SRC_UP; r.cd.c0 = 2 * c; r.cd.ring_only = 1;
            r.cd.src0 = enc(tmp[cur]); r.cd.out = enc(tmp[ui]); r.cd.epi = EPI_BF16; r.cd.cout_stride = (lvl == 0) ? 128 : c;
            finish(r, " ring", false);
        } else {
            Launch r = base(li, 0, lvl == 0 ? 1 : lvl);
            r.cd.mode = (lvl == 0) ? SRC_UP_S2D : SRC_UP; r.cd.c0 = 2 * c;
            r.cd.src0 = enc(tmp[cur]); r.cd.out = enc(tmp[ui]); r.cd.epi = EPI_BF16; r.cd.cout_stride = (lvl == 0) ? 128 : c;
            finish(r, " up", true);
        }
        ++li;
        {   // conv_block.block.0 on cat(up, skip)
            Launch ln = base(li, 0, lvl == 0 ? 1 : lvl);
            ln.cd.mode = SRC_CAT; ln.cd.c0 = ln.cd.c1 = (lvl == 0) ? 128 : c; ln.cd.cout_stride = (lvl == 0) ? 128 : c;
            ln.cd.src0 = enc(tmp[ui]); ln.cd.src1 = enc(e->off_skip[lvl]); ln.cd.out = enc(tmp[vi]); ln.cd.epi = EPI_BF16;
            finish(ln, "", true); ++li;
        }
        {   // conv_block.block.2
            Launch ln = base(li, 0, lvl == 0 ? 1 : lvl);
            ln.cd.mode = SRC_PLAIN; ln.cd.c0 = (lvl == 0) ? 128 : c; ln.cd.cout_stride = ln.cd.c0;
            ln.cd.src0 = enc(tmp[vi]); ln.cd.out = enc(tmp[cur]); ln.cd.epi = EPI_BF16;
            finish(ln, "", true); ++li;
        }
    }
    {   // last (level 0, fp32 space-to-depth output)
        Launch ln = base(li, 0, 1);
        ln.cd.mode = SRC_PLAIN; ln.cd.c0 = 128; ln.cd.epi = EPI_F32X16; ln.cd.cout_stride = 16;
        ln.cd.src0 = enc(tmp[cur]); ln.cd.out = enc(out_off);
        finish(ln, "", true); ++li;
    }
}

static void plan_glue(rrin_engine* e, int id) {
    static const char* names[5] = {"pack_pair", "flow_tscale_pack", "warp_pack", "blend_pack", "residue_clamp"};
    static const double bytes_px[5] = {56, 72, 120, 120, 44};          // per pixel, see DESIGN.md
    Launch ln;
    ln.glue = id; ln.name = names[id];
    ln.bytes = bytes_px[id] * (double)e->H * e->W * (id == 0 ? e->Np : e->Nt);
    e->launches.push_back(ln);
}

extern "C" {

int rrin_version(void) { return 200; }
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
    return rrin_pack_conv_ex(idx, w, b, blob, RRIN_PRECISION_BF16, stream);
}

static int check_precision(const char* who, int precision) {
    if (precision != RRIN_PRECISION_BF16 && precision != RRIN_PRECISION_FP16) { set_error("%s: unknown precision %d", who, precision); return RRIN_ERR_BAD_ARG; }
    return RRIN_OK;
}

int rrin_pack_conv_ex(int idx, const float* w, const float* b, void* blob, int precision, void* stream) {
    if (int r = check_precision("rrin_pack_conv_ex", precision)) return r;
    const int f16 = precision == RRIN_PRECISION_FP16;
    const Schedule& s = schedule();
    if (idx < 0 || idx >= (int)s.layers.size() || !w || !b || !blob) { set_error("rrin_pack_conv: bad argument"); return RRIN_ERR_BAD_ARG; }
    const Layer& L = s.layers[idx];
    uint8_t* base = static_cast<uint8_t*>(blob);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int r = conv_pack_weights(L.main.kind, w, b, L.cout, L.cin, L.main.n_stages, L.main.cfg, base + L.main.w_off,
                              reinterpret_cast<float*>(base + L.main.b_off), st, f16);
    if (r == RRIN_OK && L.fold.kind >= 0)
        r = conv_pack_weights(L.fold.kind, w, b, L.cout, L.cin, L.fold.n_stages, L.fold.cfg, base + L.fold.w_off,
                              reinterpret_cast<float*>(base + L.fold.b_off), st, f16);
    if (r == RRIN_OK && L.strip.kind >= 0)
        r = conv_pack_weights(L.strip.kind, w, b, L.cout, L.cin, L.strip.n_stages, L.strip.cfg, base + L.strip.w_off,
                              reinterpret_cast<float*>(base + L.strip.b_off), st, f16);
    return r;
}

int rrin_engine_create(int n_pairs, int n_samples, int H, int W, rrin_engine** out) {
    return rrin_engine_create_ex(n_pairs, n_samples, H, W, RRIN_PRECISION_BF16, out);
}

int rrin_engine_create_ex(int n_pairs, int n_samples, int H, int W, int precision, rrin_engine** out) {
    if (int r = check_precision("rrin_engine_create_ex", precision)) return r;
    if (!out) { set_error("rrin_engine_create: null out"); return RRIN_ERR_BAD_ARG; }
    *out = nullptr;
    if (n_pairs <= 0 || n_samples <= 0 || H <= 0 || W <= 0) { set_error("rrin_engine_create: empty shape"); return RRIN_ERR_BAD_SHAPE; }
    if (H % 16 || W % 16) { set_error("H and W must be multiples of 16 (got %dx%d)", H, W); return RRIN_ERR_BAD_SHAPE; }
    if (n_pairs != n_samples && n_pairs != 1) { set_error("n_pairs must equal n_samples or be 1 (got %d, %d)", n_pairs, n_samples); return RRIN_ERR_BAD_SHAPE; }
    rrin_engine* e = new rrin_engine();
    e->Np = n_pairs; e->Nt = n_samples; e->H = H; e->W = W; e->f16 = precision == RRIN_PRECISION_FP16;
    e->pair_mul = (n_pairs == n_samples) ? 1 : 0;
    const int B = n_pairs > n_samples ? n_pairs : n_samples;
    const size_t px = (size_t)H * W;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    for (int i = 0; i < 3; ++i) e->off_tmp[i] = take(lvl_bytes(B, H, W, 0));
    for (int l = 0; l < 4; ++l) e->off_skip[l] = take(lvl_bytes(B, H, W, l));
    for (int l = 0; l < 4; ++l) e->off_pool[l] = take(lvl_bytes(B, H, W, l) / 4);          // avg_pool2d of level l: level-(l+1) grid, level-l channels
    e->off_h16 = take((size_t)B * px * 16 * 2);
    e->off_flow4 = take((size_t)n_pairs * px * 16);
    e->off_u4 = take((size_t)n_samples * px * 16);
    e->off_out4 = take((size_t)n_samples * px * 16);
    e->off_xt8 = take((size_t)n_samples * px * 32);
    e->ws_bytes = off;
    // launch sequence of Net.forward
    // The glue between the U-Nets (K2-K5) runs fused in the epilogue of the preceding `last` conv; RRIN_FUSE=0 launches
    // the stand-alone glue kernels instead (same per-block device functions: bit-identical results).
    static const bool fuse = [] { const char* v = getenv("RRIN_FUSE"); return !(v && v[0] == '0'); }();
    static const double glue_bytes_px[5] = {56, 72, 120, 120, 44};
    plan_glue(e, 0);                                              // model.py:33
    for (int u = 0; u < 4; ++u) {
        // Flow (model.py:35) | refine_flow (:42) | Mask (:52) | final (:62)
        plan_unet(e, u, u == 0 ? n_pairs : n_samples, e->off_h16, u == 0 ? e->off_flow4 : e->off_u4);
        if (fuse) {
            Launch& last = e->launches.back();
            last.fuse_mode = u + 1;
            if (u == 1 && warp_staging()) {             // refine_flow.last runs both backward warps: frame windows staged in shared memory
                last.cd.cfg = 44;                       // (same packed weights and tensor-map geometry as the other `last` configs with 16 x 16 tiles)
                last.name = "conv3x3_tma#44<KCS64,KB32,NT16,MSUB2>";
            }
            last.name += " +glue";
            last.bytes += glue_bytes_px[u + 1] * (double)H * W * n_samples - 16.0 * H * W * n_samples;   // the fp32 hand-over tensor is not materialised
        } else {
            plan_glue(e, u + 1);                                  // model.py:37-41 | 44-50 | 52-55,61 | 62-63
        }
    }
    *out = e;
    return RRIN_OK;
}

void rrin_engine_destroy(rrin_engine* e) {
    if (!e) return;
    for (auto& g : e->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
    delete e;
}
size_t rrin_engine_workspace_bytes(const rrin_engine* e) { return e ? e->ws_bytes : 0; }
int rrin_engine_num_launches(const rrin_engine* e) { return e ? (int)e->launches.size() : 0; }

int rrin_engine_forward(rrin_engine* e, const void* blob_, void* workspace, const float* in0, const float* in1,
                        const float* coef, float* out, void* stream) {
    if (!e || !blob_ || !workspace || !in0 || !in1 || !coef || !out) { set_error("rrin_engine_forward: null argument"); return RRIN_ERR_BAD_ARG; }
    // frames are read with 8-byte loads and the result is written with 8-byte stores; the workspace and the weight blob
    // hold TMA / 256-bit-store targets.  A misaligned pointer would fault on the device (sticky): refuse it here.
    if ((reinterpret_cast<uintptr_t>(in0) | reinterpret_cast<uintptr_t>(in1) | reinterpret_cast<uintptr_t>(out)) & 15) {
        set_error("rrin_engine_forward: in0, in1 and out must be 16-byte aligned"); return RRIN_ERR_BAD_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(blob_) | reinterpret_cast<uintptr_t>(workspace)) & 255) {
        set_error("rrin_engine_forward: the weight blob and the workspace must be 256-byte aligned"); return RRIN_ERR_BAD_ARG;
    }
    if (reinterpret_cast<uintptr_t>(coef) & 3) { set_error("rrin_engine_forward: coef must be 4-byte aligned"); return RRIN_ERR_BAD_ARG; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint8_t* blob = static_cast<const uint8_t*>(blob_);
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    const Schedule& s = schedule();
    void* h16 = ws + e->off_h16;
    float* flow4 = reinterpret_cast<float*>(ws + e->off_flow4);
    float* u4 = reinterpret_cast<float*>(ws + e->off_u4);
    float* out4 = reinterpret_cast<float*>(ws + e->off_out4);
    float* xt8 = reinterpret_cast<float*>(ws + e->off_xt8);
    const int H = e->H, W = e->W, Np = e->Np, Nt = e->Nt, pm = e->pair_mul;
    auto dec = [&](const void* p) -> void* { return p ? ws + (reinterpret_cast<size_t>(p) - 1) : nullptr; };
    if (e->tmap_ws != workspace) {                                // (re-)encode the TMA tensor maps for this workspace
        for (Launch& ln : e->launches) {
            if (ln.glue >= 0 || ln.cd.cfg < 10) continue;
            const ConvDesc& c = ln.cd;
            int r = (c.mode == SRC_UP) ? conv_make_tmap(dec(c.src0), c.N, c.H / 2, c.W / 2, c.c0, c.cfg, 2, ln.tmap[0], c.transposed)   // raw coarse tile
                                       : conv_make_tmap(dec(c.src0), c.N, c.H, c.W, c.c0, c.cfg, 0, ln.tmap[0], c.transposed);
            if (r == RRIN_OK && c.mode == SRC_CAT) r = conv_make_tmap(dec(c.src1), c.N, c.H, c.W, c.c1, c.cfg, 0, ln.tmap[1], c.transposed);
            if (r == RRIN_OK && conv_config_tma_epilogue(c.cfg)) r = conv_make_tmap(dec(c.out), c.N, c.H, c.W, c.cout_stride, c.cfg, 1, ln.tmap[2], c.transposed);
            if (r != RRIN_OK) return r;
        }
        e->tmap_ws = workspace;
    }
    mark(e, st);                                                  // t0
    for (const Launch& ln : e->launches) {
        int r = RRIN_OK;
        if (ln.glue >= 0) {
            switch (ln.glue) {
                case 0: r = pack_pair(in0, in1, Np, H, W, h16, st, e->f16); break;
                case 1: r = flow_tscale_pack(flow4, in0, in1, coef, Nt, pm, H, W, h16, st, e->f16); break;
                case 2: r = warp_pack(flow4, u4, in0, in1, coef, Nt, pm, H, W, h16, xt8, st, e->f16); break;
                case 3: r = blend_pack(u4, xt8, in0, in1, coef, Nt, pm, H, W, out4, h16, st, e->f16); break;
                case 4: r = residue_clamp(u4, out4, Nt, H, W, out, st); break;
            }
        } else {
            const Layer& L = s.layers[ln.layer];
            const Pack& pk = ln.use_fold == 1 ? L.fold : (ln.use_fold == 2 ? L.strip : L.main);
            ConvDesc cd = ln.cd;
            cd.src0 = dec(cd.src0); cd.src1 = dec(cd.src1); cd.out = dec(cd.out); cd.pool_out = dec(cd.pool_out);
            cd.wpack = blob + pk.w_off;
            cd.f16 = e->f16;
            cd.bias = reinterpret_cast<const float*>(blob + pk.b_off);
            if (cd.cfg >= 10) { cd.tmap0 = ln.tmap[0]; cd.tmap1 = ln.tmap[1]; cd.tmap_out = ln.tmap[2]; }
            if (ln.fuse_mode) {
                ConvFuse& f = cd.fuse;
                f.mode = ln.fuse_mode; f.H = H; f.W = W; f.Nt = Nt; f.pair_mul = pm;
                f.in0 = in0; f.in1 = in1; f.coef = coef; f.h16 = h16;
                if (ln.fuse_mode == 2) { f.aux = flow4; f.dst = xt8; }
                else if (ln.fuse_mode == 3) { f.aux = xt8; f.dst = out4; }
                else if (ln.fuse_mode == 4) { f.aux = out4; f.dst = out; f.h16 = nullptr; }
            }
            r = conv_launch(cd, st);
        }
        if (r != RRIN_OK) return r;
        mark(e, st);
    }
    return RRIN_OK;
}

// Net.forward replayed from a CUDA graph (model.py:59-65; same launches, same programmatic-dependent-launch edges as
// rrin_engine_forward, one cudaGraphLaunch instead of ~90 kernel launches).  A graph bakes in every pointer, so graphs are
// cached per (blob, workspace, in0, in1, coef, out): the FIRST call with a new pointer set launches directly and only
// remembers the set; the second call captures and instantiates (this is the one place the library allocates -- host and
// driver memory of the graph -- after engine creation); later calls replay.  A streaming caller that cycles through a few
// staging buffers (rrin_b200.pipeline, convert.py's loop under torch's caching allocator) therefore replays every step,
// while a caller with ever-new pointers never pays for a capture.  At most kMaxGraphs graphs are kept (least recently used
// is dropped).
static constexpr int kMaxGraphs = 16;

int rrin_engine_forward_graph(rrin_engine* e, const void* blob, void* workspace, const float* in0, const float* in1,
                              const float* coef, float* out, void* stream) {
    if (!e) { set_error("rrin_engine_forward_graph: null engine"); return RRIN_ERR_BAD_ARG; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int dev = -1;
    RRIN_CUDA_CHECK(cudaGetDevice(&dev));
    const void* key[6] = {blob, workspace, in0, in1, coef, out};
    rrin_engine::GraphEntry* hit = nullptr;
    for (auto& g : e->graphs)
        if (g.dev == dev && memcmp(g.key, key, sizeof key) == 0) { hit = &g; break; }
    ++e->tick;
    if (hit && hit->exec) {
        hit->last_use = e->tick;
        RRIN_CUDA_CHECK(cudaGraphLaunch(hit->exec, st));
        ++e->graph_launches;
        return RRIN_OK;
    }
    if (!hit) {                                   // first sighting: remember the pointer set, launch directly
        if ((int)e->graphs.size() >= kMaxGraphs) {
            size_t lru = 0;
            for (size_t i = 1; i < e->graphs.size(); ++i) if (e->graphs[i].last_use < e->graphs[lru].last_use) lru = i;
            if (e->graphs[lru].exec) cudaGraphExecDestroy(e->graphs[lru].exec);
            e->graphs.erase(e->graphs.begin() + lru);
        }
        rrin_engine::GraphEntry g;
        memcpy(g.key, key, sizeof key);
        g.dev = dev; g.last_use = e->tick;
        e->graphs.push_back(g);
        ++e->direct_launches;
        return rrin_engine_forward(e, blob, workspace, in0, in1, coef, out, stream);
    }
    // second sighting: capture the launch sequence on the private stream and instantiate it -- unless this caller keeps
    // producing new pointer sets (captures that are hardly ever replayed): then stay on direct launches
    hit->last_use = e->tick;
    if (e->captures >= 32 && e->graph_launches < 4 * e->captures) {
        ++e->direct_launches;
        return rrin_engine_forward(e, blob, workspace, in0, in1, coef, out, stream);
    }
    ++e->captures;
    if (e->cap_stream && e->cap_dev != dev) { cudaStreamDestroy(e->cap_stream); e->cap_stream = nullptr; }
    if (!e->cap_stream) { RRIN_CUDA_CHECK(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking)); e->cap_dev = dev; }
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal);
    int r = RRIN_OK;
    if (ce == cudaSuccess) {
        r = rrin_engine_forward(e, blob, workspace, in0, in1, coef, out, e->cap_stream);
        ce = cudaStreamEndCapture(e->cap_stream, &graph);          // always ends the capture, also after a failed launch
        if (r == RRIN_OK && ce == cudaSuccess && graph) ce = cudaGraphInstantiate(&hit->exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
    }
    if (r != RRIN_OK || ce != cudaSuccess || !hit->exec) {         // capture unavailable: stay on direct launches for this set
        cudaGetLastError();
        hit->exec = nullptr;
        ++e->direct_launches;
        return rrin_engine_forward(e, blob, workspace, in0, in1, coef, out, stream);
    }
    RRIN_CUDA_CHECK(cudaGraphLaunch(hit->exec, st));
    ++e->graph_launches;
    return RRIN_OK;
}

// how many forwards of this engine were replayed from a graph / launched kernel by kernel (tests, bench.py)
int rrin_engine_graph_stats(const rrin_engine* e, int* graph_launches, int* direct_launches, int* graphs_alive) {
    if (!e) { set_error("rrin_engine_graph_stats: null engine"); return RRIN_ERR_BAD_ARG; }
    if (graph_launches) *graph_launches = e->graph_launches;
    if (direct_launches) *direct_launches = e->direct_launches;
    if (graphs_alive) { int n = 0; for (auto& g : e->graphs) n += g.exec != nullptr; *graphs_alive = n; }
    return RRIN_OK;
}

// Launch i of one forward, in stream order: kernel class name, the reference layer it computes,
// algorithmic FLOPs and algorithmic HBM bytes (each operand/result tensor crossing once).
int rrin_engine_launch_info(const rrin_engine* e, int i, char* name, int name_cap, char* layer, int layer_cap,
                            double* flops, double* bytes) {
    if (!e || i < 0 || i >= (int)e->launches.size()) { set_error("rrin_engine_launch_info: bad index %d", i); return RRIN_ERR_BAD_ARG; }
    const Launch& ln = e->launches[i];
    const std::string ly = ln.glue >= 0 ? std::string("model.py glue: ") + ln.name : schedule().layers[ln.layer].key;
    if (name && name_cap > 0) { strncpy(name, ln.name.c_str(), name_cap - 1); name[name_cap - 1] = 0; }
    if (layer && layer_cap > 0) { strncpy(layer, ly.c_str(), layer_cap - 1); layer[layer_cap - 1] = 0; }
    if (flops) *flops = ln.flops;
    if (bytes) *bytes = ln.bytes;
    return RRIN_OK;
}

// One forward with a CUDA event after every launch; ms[i] = device time of launch i.
// Synchronises the stream (profiling aid for bench.py, not part of the product path).
int rrin_engine_forward_profiled(rrin_engine* e, const void* blob, void* workspace, const float* in0, const float* in1,
                                 const float* coef, float* out, void* stream, float* ms_host) {
    if (!e || !ms_host) { set_error("rrin_engine_forward_profiled: null argument"); return RRIN_ERR_BAD_ARG; }
    const int nl = (int)e->launches.size();
    std::vector<cudaEvent_t> ev(nl + 1);
    for (auto& x : ev) RRIN_CUDA_CHECK(cudaEventCreate(&x));
    e->prof = &ev; e->prof_n = 0;
    int r = rrin_engine_forward(e, blob, workspace, in0, in1, coef, out, stream);
    e->prof = nullptr;
    if (r == RRIN_OK) {
        cudaError_t ce = cudaStreamSynchronize(static_cast<cudaStream_t>(stream));
        if (ce != cudaSuccess) { set_error("profiled forward failed: %s", cudaGetErrorString(ce)); r = RRIN_ERR_CUDA; }
    }
    if (r == RRIN_OK)
        for (int i = 0; i < nl; ++i) cudaEventElapsedTime(&ms_host[i], ev[i], ev[i + 1]);
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
int rrin_conv_config_info(int cfg, int* kcs, int* kb, int* nt, int* msub) { return conv_config_info(cfg, kcs, kb, nt, msub); }
size_t rrin_conv_packed_weight_bytes(int cfg, int n_cols, int n_stages, int sched) {
    if (!conv_config_valid(cfg)) return 0;
    return conv_packed_weight_bytes(cfg, n_cols, n_stages, sched);
}
int rrin_conv_packed_bias_count(int cfg, int n_cols) {
    if (!conv_config_valid(cfg)) return 0;
    return conv_packed_bias_count(cfg, n_cols);
}
int rrin_pack_conv_raw(int kind, const float* w, const float* b, int cout, int cin, int n_stages, int cfg, void* wpack,
                       float* bias_pack, void* stream) {
    return conv_pack_weights(kind, w, b, cout, cin, n_stages, cfg, wpack, bias_pack, static_cast<cudaStream_t>(stream));
}
int rrin_pack_conv_raw_ex(int kind, const float* w, const float* b, int cout, int cin, int n_stages, int cfg, void* wpack,
                          float* bias_pack, int precision, void* stream) {
    if (int r = check_precision("rrin_pack_conv_raw_ex", precision)) return r;
    return conv_pack_weights(kind, w, b, cout, cin, n_stages, cfg, wpack, bias_pack, static_cast<cudaStream_t>(stream), precision == RRIN_PRECISION_FP16);
}
int rrin_conv3x3(const void* src0, const void* src1, int c0, int c1, int src_mode, int pad_clamp, int N, int H, int W,
                 int sched, int n_cols, const void* wpack, const float* bias_pack, void* out, int epi, int cout_stride,
                 int act, int ring_only, int cfg, void* pool_out, void* stream) {
    return rrin_conv3x3_ex(src0, src1, c0, c1, src_mode, pad_clamp, N, H, W, sched, n_cols, wpack, bias_pack, out, epi, cout_stride, act,
                           ring_only, cfg, pool_out, RRIN_PRECISION_BF16, 0, stream);
}
int rrin_conv3x3_ex(const void* src0, const void* src1, int c0, int c1, int src_mode, int pad_clamp, int N, int H, int W,
                    int sched, int n_cols, const void* wpack, const float* bias_pack, void* out, int epi, int cout_stride,
                    int act, int ring_only, int cfg, void* pool_out, int precision, int transposed, void* stream) {
    if (int r = check_precision("rrin_conv3x3_ex", precision)) return r;
    ConvDesc cd;
    cd.f16 = precision == RRIN_PRECISION_FP16;
    cd.transposed = transposed ? 1 : 0;
    cd.src0 = src0; cd.src1 = src1; cd.c0 = c0; cd.c1 = c1; cd.mode = src_mode; cd.pad_clamp = pad_clamp;
    cd.N = N; cd.H = H; cd.W = W; cd.sched = sched; cd.n_cols = n_cols; cd.wpack = wpack; cd.bias = bias_pack;
    cd.out = out; cd.epi = epi; cd.cout_stride = cout_stride; cd.act = act; cd.ring_only = ring_only; cd.cfg = cfg; cd.pool_out = pool_out;
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
int rrin_frame_from_u8(const uint8_t* src_hwc, int H0, int W0, int C, int top_pad, int bottom_pad, float* dst_nchw, void* stream) {
    return frame_from_u8(src_hwc, H0, W0, C, top_pad, bottom_pad, dst_nchw, static_cast<cudaStream_t>(stream));
}
int rrin_frame_to_u8(const float* src_nchw, int H, int W, int H0, int W0, uint8_t* dst_hwc, void* stream) {
    return frame_to_u8(src_nchw, H, W, H0, W0, dst_hwc, static_cast<cudaStream_t>(stream));
}
int rrin_warp(const float* img, const float* flow, int N, int C, int H, int W, float* out, void* stream) {
    return warp_frames(img, flow, N, C, H, W, out, static_cast<cudaStream_t>(stream));
}
int rrin_residue_clamp(const float* res4, const float* out4, int n, int H, int W, float* out_nchw, void* stream) {
    return residue_clamp(res4, out4, n, H, W, out_nchw, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
