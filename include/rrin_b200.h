/* rrin_b200 -- C-ABI of the B200-native (sm_100a) RRIN forward pass.
 *
 * This is the drop-in boundary beneath the Python `Net` module.
 * The reference (Thomasedv/RRIN) has no FFI of its own: its hot path is the Python
 * `Net.forward` (/root/reference/model.py:59-65) calling PyTorch operators.  Each entry
 * point below therefore cites the reference *call site(s)* whose work it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless noted;
 *   - `stream` is a cudaStream_t passed as void* (the caller's current stream);
 *   - functions return 0 on success, non-zero otherwise (rrin_last_error() gives the text);
 *     they never throw, never allocate device memory, never synchronise the device (rrin_engine_forward_profiled, a
 *     profiling aid, synchronises; rrin_engine_forward_graph instantiates a CUDA graph on its second call per pointer set);
 *   - an engine and its workspace serve ONE forward at a time: calls on different streams must be ordered by the caller
 *     (rrin_b200.engine.Engine.run does it with an event);
 *   - frames / results: fp32 NCHW [N,3,H,W] in [0,1]   (dataloader.py:116-118, convert.py:133)
 *   - U-Net activations: bf16 NHWC (level 0: space-to-depth, see K1); U-Net outputs: fp32 [N,H/2,W/2,4,4];
 *   - H and W must be multiples of 16 (four 2x2 pools in the Flow U-Net, unet.py:46).
 */
#ifndef RRIN_B200_H_
#define RRIN_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define RRIN_API __attribute__((visibility("default")))
#else
#define RRIN_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define RRIN_OK 0
#define RRIN_ERR_BAD_SHAPE 1
#define RRIN_ERR_UNSUPPORTED 2
#define RRIN_ERR_CUDA 3
#define RRIN_ERR_BAD_ARG 4

/* 16-bit operand / activation format of the tensor-core path (accumulation is always fp32; flows, mask logits, the blend
 * and the final residue stay fp32).  BF16: fp32 range, 8-bit significand -- the default, PSNR >= 50 dB path.  FP16: 11-bit
 * significand at the same MMA rate and size -- the "fp32-accumulate path within 1e-3" precision mode; activations above 65504
 * would overflow.  The reference computes these convs in fp32 (unet.py:29,38,59,62,78). */
#define RRIN_PRECISION_BF16 0
#define RRIN_PRECISION_FP16 1

/* library version (major*10000 + minor*100 + patch) and last error text of this thread */
RRIN_API int rrin_version(void);
RRIN_API const char* rrin_last_error(void);

/* ---- static description of the 81 convolutions, in execution order ----------------------
 * Mirrors Net.__init__ (model.py:27-30) and UNet.__init__/forward (unet.py:10-51):
 * U-Nets run in the order Flow, refine_flow, Mask, final (model.py:35,42,52,62).
 * `key` receives the state_dict prefix of the conv, e.g. "Flow.down_path.0.block.0"
 * (append ".weight" / ".bias"); src_mode: 0 plain, 1 cat(up,skip), 2 avg-pool, 3 bilinear x2,
 * 4 packed head input. */
RRIN_API int rrin_num_convs(void);
RRIN_API int rrin_conv_info(int idx, char* key, int key_cap, int* cin, int* cout, int* level, int* src_mode, int* act);

/* ---- K7: one-time weight repack (replaces nothing in the reference; cost of load_state_dict,
 * convert.py:100-104).  w: fp32 OIHW [cout,cin,3,3], b: fp32 [cout] -> this conv's slice of the
 * packed blob (bf16, per-tap K-major UMMA core-matrix order).  blob must be 256-byte aligned. */
RRIN_API size_t rrin_packed_weights_bytes(void);
RRIN_API int rrin_pack_conv(int idx, const float* w, const float* b, void* blob, void* stream);               /* bf16 */
RRIN_API int rrin_pack_conv_ex(int idx, const float* w, const float* b, void* blob, int precision, void* stream);

/* ---- the engine: Net.forward (model.py:59-65) for a fixed problem shape -------------------
 * n_pairs frame pairs and n_samples interpolated frames: either n_samples == n_pairs (sample i
 * uses pair i: a batch, model.py:59) or n_pairs == 1 (all samples interpolate the same pair at
 * different t; the t-independent Flow U-Net, model.py:33-35, then runs once). */
typedef struct rrin_engine rrin_engine;
RRIN_API int rrin_engine_create(int n_pairs, int n_samples, int H, int W, rrin_engine** out);                    /* bf16 */
RRIN_API int rrin_engine_create_ex(int n_pairs, int n_samples, int H, int W, int precision, rrin_engine** out);  /* blob packed with the same precision */
RRIN_API void rrin_engine_destroy(rrin_engine* e);
RRIN_API size_t rrin_engine_workspace_bytes(const rrin_engine* e);   /* caller-provided scratch, 256-B aligned */
RRIN_API int rrin_engine_num_launches(const rrin_engine* e);         /* kernels enqueued by one forward */
/* coef: fp32 [n_samples][6] = {-(1-t)t, t*t, (1-t)(1-t), t(1-t), 1-t, t} (model.py:38-39,54).
 * in0,in1: fp32 NCHW [n_pairs,3,H,W]; out: fp32 NCHW [n_samples,3,H,W]. */
RRIN_API int rrin_engine_forward(rrin_engine* e, const void* blob, void* workspace, const float* in0, const float* in1,
                        const float* coef, float* out, void* stream);
/* The same forward replayed from a CUDA graph (one cudaGraphLaunch instead of ~90 kernel launches; replaces the ~250-300
 * eager op launches per Net.forward of the reference, model.py:59-65).  Graphs are cached per pointer set (blob, workspace,
 * in0, in1, coef, out): the first call with a new set launches directly, the second captures + instantiates (the only
 * allocation the library makes after rrin_engine_create: host/driver memory of the graph), later calls replay; at most 16
 * sets are kept.  Results are bit-identical to rrin_engine_forward.  rrin_engine_graph_stats reports how many forwards of
 * this engine were replayed / launched directly and how many graphs are alive. */
RRIN_API int rrin_engine_forward_graph(rrin_engine* e, const void* blob, void* workspace, const float* in0, const float* in1,
                              const float* coef, float* out, void* stream);
RRIN_API int rrin_engine_graph_stats(const rrin_engine* e, int* graph_launches, int* direct_launches, int* graphs_alive);
/* Profiling aids (bench.py): description of launch i of one forward (kernel class, reference layer,
 * algorithmic FLOPs and HBM bytes), and a forward that records a CUDA event after every launch and
 * returns per-launch device milliseconds in the HOST array ms_host[num_launches] (synchronises). */
RRIN_API int rrin_engine_launch_info(const rrin_engine* e, int i, char* name, int name_cap, char* layer, int layer_cap,
                            double* flops, double* bytes);
RRIN_API int rrin_engine_forward_profiled(rrin_engine* e, const void* blob, void* workspace, const float* in0, const float* in1,
                                 const float* coef, float* out, void* stream, float* ms_host);
/* debug/test taps: copy out intermediate fp32 NHWC4 tensors of the last forward.
 * which: 0 flow (n_pairs), 1 blend output (n_samples); layout fp32 [n,H/2,W/2,4 phases,4].  dst: n*H*W*4 floats. */
RRIN_API int rrin_engine_tap(const rrin_engine* e, const void* workspace, int which, float* dst, void* stream);

/* ---- unit-level entry points (used by tests and micro-benchmarks) -------------------------
 * K1: conv3x3(pad 1) + bias + optional LeakyReLU(0.1) on tcgen05 tensor cores.
 * Replaces nn.Conv2d (unet.py:29,38,59,62,78) + LeakyReLU (unet.py:47,60,63), with the input
 * transform folded into the operand loads or the weights.
 *   src_mode : 0 plain [N,H,W,c0] | 1 cat(src0,src1) (unet.py:93) | 2 avg_pool2d of [N,2H,2W,c0] (unet.py:46)
 *              | 3 bilinear x2 of [N,H/2,W/2,c0] (unet.py:77) | 4 mean over the 4 phases of a space-to-depth
 *              source [N,H,W,c0] | 5 space-to-depth grid: phase (a,b) = bilinear x2 of [N,H,W,c0] at (2y+a,2x+b)
 *   sched    : 0 nine 3x3 taps | 1 space-to-depth (16 (block shift, phase) entries; level-0 tensors are stored
 *              [N,H/2,W/2,4 phases,C] and the conv runs on the half-resolution grid with 4*Cout columns)
 *   epi      : 0 bf16 NHWC [N,H,W,cout_stride] | 1 fp32 [N,H,W,16] | 2 folded-upsample scatter into bf16
 *              [N,2H,2W,cout_stride]
 *   pack kind: 0 plain | 1 space-to-depth | 2 bilinear-x2-folded weights (4*cout columns, pad_clamp=1 input)
 *   sched 2  : half-phase schedule of the TMA-fed kernel for level-0 tensors (pack kind 3)
 *   pool_out : optional second output of the TMA-epilogue configs: F.avg_pool2d(out, 2) (unet.py:46) written by the
 *              same epilogue, bf16 NHWC [N,H/2,W/2,cout_stride] (space-to-depth grid: [N,H,W,cout_stride/4]); or NULL
 *   cfg      : tile configuration: 0..8 transform kernel (pool / bilinear sources, border strips), 10..44 TMA-fed kernel
 *              (rrin_conv_config_info gives KCS, KB, NT, MSUB; the table with the CTA-pair, two-tile-stream and
 *              frame-staging variants is in rrin_b200/csrc/conv3x3_launch.cuh).
 *   pack kind: ... 4 plain, CTA-pair halves | 5 half-phase space-to-depth, CTA-pair halves
 *   alignment: activation / output / pooled tensors 32-byte aligned (TMA maps and 256-bit stores); a misaligned
 *              pointer is rejected with RRIN_ERR_BAD_ARG. */
RRIN_API int rrin_conv_config_info(int cfg, int* kcs, int* kb, int* nt, int* msub);
RRIN_API size_t rrin_conv_packed_weight_bytes(int cfg, int n_cols, int n_stages, int sched);
RRIN_API int rrin_conv_packed_bias_count(int cfg, int n_cols);
RRIN_API int rrin_pack_conv_raw(int kind, const float* w, const float* b, int cout, int cin, int n_stages, int cfg,
                       void* wpack, float* bias_pack, void* stream);
RRIN_API int rrin_conv3x3(const void* src0, const void* src1, int c0, int c1, int src_mode, int pad_clamp, int N, int H, int W,
                 int sched, int n_cols, const void* wpack, const float* bias_pack, void* out, int epi, int cout_stride,
                 int act, int ring_only, int cfg, void* pool_out, void* stream);
/* the same two with the 16-bit format given (RRIN_PRECISION_*); the plain names are the bf16 ones */
RRIN_API int rrin_pack_conv_raw_ex(int kind, const float* w, const float* b, int cout, int cin, int n_stages, int cfg,
                          void* wpack, float* bias_pack, int precision, void* stream);
RRIN_API int rrin_conv3x3_ex(const void* src0, const void* src1, int c0, int c1, int src_mode, int pad_clamp, int N, int H, int W,
                    int sched, int n_cols, const void* wpack, const float* bias_pack, void* out, int epi, int cout_stride,
                    int act, int ring_only, int cfg, void* pool_out, int precision, int transposed, void* stream);
/* transposed != 0 (streamed 9-tap TMA configs with the TMA-store epilogue, e.g. 16 and 20): the kernel's 16-row bands run along the
 * image WIDTH (tensor maps with swapped W / H, weight taps swapped) -- same result, less padding when H is not a multiple of 16 */

/* Glue kernels.  Frames are fp32 NCHW; tensors exchanged with the U-Nets are space-to-depth on the half-res
 * grid: head inputs bf16 [N,H/2,W/2,4,16]; U-Net outputs fp32 [N,H/2,W/2,4,4]; xt8 fp32 [N,H/2,W/2,4,8].
 * The packed tensors (x16, r16, m16, f16, flow4, res4, mask4, xt8, out4) must be 32-byte aligned (256-bit stores). */
/* K6: torch.cat((x0,x1),1) (model.py:33) -> packed 16-channel bf16 NHWC head input. */
RRIN_API int rrin_pack_pair(const float* in0, const float* in1, int N, int H, int W, void* x16, void* stream);
/* K2: flow t-scaling (model.py:37-39) + cat((F_t0,F_t1,x),1) (model.py:41) -> refine head input. */
RRIN_API int rrin_flow_tscale_pack(const float* flow4, const float* in0, const float* in1, const float* coef, int n_samples,
                          int pair_mul, int H, int W, void* r16, void* stream);
/* K3: residue add (model.py:44-45) + warp x2 (model.py:8-21,47-48: grid build + F.grid_sample)
 * + cat((F_t0,F_t1,x,xt1,xt2),1) (model.py:50) -> Mask head input m16 and fp32 xt8 = [xt1,xt2,0,0]. */
RRIN_API int rrin_warp_pack(const float* flow4, const float* res4, const float* in0, const float* in1, const float* coef,
                   int n_samples, int pair_mul, int H, int W, void* m16, float* xt8, void* stream);
/* K4: sigmoid (model.py:52) + weights (model.py:54) + blend (model.py:55) + cat((in0,in1,output),1)
 * (model.py:61) -> out4 (fp32 NHWC4) and the `final` head input f16. */
RRIN_API int rrin_blend_pack(const float* mask4, const float* xt8, const float* in0, const float* in1, const float* coef,
                    int n_samples, int pair_mul, int H, int W, float* out4, void* f16, void* stream);
/* K3b: the reference's public `warp(img, flow)` (model.py:8-21) on its own: img fp32 NCHW [N,C,H,W], flow fp32 [N,2,H,W]
 * (channel 0 = x displacement, 1 = y displacement, in pixels) -> out fp32 [N,C,H,W].  Grid build in the reference's fp32
 * op order (model.py:15-18) + F.grid_sample defaults (bilinear, zeros padding, align_corners=False).  Any H, W >= 1.
 * Net.forward never calls this (its two warps run fused in K3); it exists so that `from model import warp` keeps working. */
RRIN_API int rrin_warp(const float* img, const float* flow, int N, int C, int H, int W, float* out, void* stream);
/* K5: final residue add + clamp(0,1) (model.py:62-63) -> fp32 NCHW result. */
RRIN_API int rrin_residue_clamp(const float* res4, const float* out4, int n_samples, int H, int W, float* out_nchw, void* stream);

/* ---- uint8 frame I/O for the streaming pipeline (frames cross PCIe as bytes) ----------------
 * K8: transforms.Pad((0, top_pad, 0, right_pad), 'edge') + ToTensor() + drop alpha (dataloader.py:93-118).
 *     src: uint8 HWC [H0,W0,C], C = 3 or 4 (device); dst: fp32 NCHW [3, top_pad+H0+bottom_pad, W0].
 *     torchvision's Pad order is (left, top, right, bottom): the reference's "right_pad" is a BOTTOM pad.
 * K9: to_pil_image (mul(255).byte(), truncating) + crop((0, H-H0, W0, H)) (utils.py:51-58).
 *     src: fp32 NCHW [3,H,W]; dst: uint8 HWC [H0,W0,3], rows H-H0..H-1 and columns 0..W0-1 of src. */
RRIN_API int rrin_frame_from_u8(const uint8_t* src_hwc, int H0, int W0, int C, int top_pad, int bottom_pad, float* dst_nchw, void* stream);
RRIN_API int rrin_frame_to_u8(const float* src_nchw, int H, int W, int H0, int W0, uint8_t* dst_hwc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RRIN_B200_H_ */
