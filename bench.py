#!/usr/bin/env python
"""bench.py -- 1080p interpolated frames/s of the RRIN forward pass on N B200s.

Contract: ``python bench.py --gpus N --steps K --warmup W`` (N>1: launched under torchrun,
one rank per GPU).  A *step* is one pass of the hot path (``Net.forward``,
/root/reference/model.py:59-65) over one batch of ``--batch`` (default 4) consecutive
1920x1088 frame pairs at t=0.5 -- the unit of BASELINE.json configs[2] ("1080p 2x
interpolation of a 240-frame synthetic clip, sharded over 1/2/4/8 B200").  Each rank owns a
contiguous shard of the synthetic clip (rrin_b200.sharding), device resident, and interpolates
K consecutive batches of it; there is no data-path collective (frame pairs are independent,
SURVEY.md 8(e)), torch.distributed is used only for the barrier and the max over ranks of the
device-timed region.  One JSON line is printed by rank 0; ``value`` counts interpolated frames.

The line carries: ``value`` (device-resident inputs), ``e2e`` (the streaming pipeline from pinned host buffers, copies
inside the timed region), ``roofline`` (dominant kernel class from live per-launch CUDA-event times against
MEASURED_PEAKS.json; all convs; the whole step against the layer-wise roofline; every fused warp / blend launch against the
HBM peak, the warp launch also under x200 flow weights), ``launch_gap`` (step time vs the sum of the launches), ``batch1``
(single-pair calls), ``clocks`` (nvidia-smi samples inside the timed region) and, at N = 1, ``cpu_baseline`` (the oracle
port on the host cores -- the only place this arm touches oracle/).

``--config`` selects the other named configurations of BASELINE.json through the same contract: ``720p_b8``
(configs[1]), ``1080p_t7`` (configs[3]: 7 timesteps per pair, Flow U-Net once), ``4k`` (configs[4]) and ``clip``
(configs[2] as a strong-scaling job: the whole 240-frame clip per step, 239 pairs split over the ranks, frames/s =
239 / the slowest rank's time).  ``--precision fp16`` runs the precision mode.

``--impl reference`` times the reference's own CPU path instead: the oracle port
(oracle/rrin_oracle.py: the same torch CPU operators at the same call sites as the reference,
which is pure Python and cannot travel to the GPU box) on all host threads, each step a bounded
strip of the same workload.

``--impl library`` (not part of the driver's contract; context only) times the same oracle
restatement as torch eager operators on cuda:0 -- cuDNN convolutions in fp32, TF32 and bf16
autocast/channels_last -- the "GPU library baseline" of SURVEY.md 8(d): what the unmodified
reference would execute on a B200.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 1088, 1920                       # 1080p padded to a multiple of 16 (dataloader.py:93-108)
FLOP_PER_PX = 1_736_064                 # 81 convs, true channel counts
CLIP_FRAMES = 240
METRIC = "1080p_interpolated_frames_per_sec"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def wait_ready(self, timeout: float = 8.0):
        """Block until nvidia-smi has delivered its first sample: its start-up (NVML initialisation) takes a few hundred
        milliseconds during which kernel launches of this process can stall, which must not fall into a timed region."""
        t0 = time.time()
        while self.proc is not None and not self.rows and time.time() - t0 < timeout and self.proc.poll() is None:
            time.sleep(0.02)

    def close(self):
        if self.proc is not None:
            self.proc.terminate()

    def window(self, t0: float, t1: float):
        """Median SM clock and the throttle reasons seen between the wall-clock times t0 and t1 (the sampler keeps running)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in list(self.rows):
            if ts < t0 or ts > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except Exception:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_reference_throughput(budget_s: float, threads: int, H: int = H, W: int = W):
    """Sizes a strip of the HxW workload for ~budget_s of the oracle port; returns (state_dict, frame0, frame1, strip_h)."""
    import torch
    from oracle import rrin_oracle as O
    torch.set_num_threads(threads)
    sd = O.seeded_state_dict()
    a, b = O.seeded_frames(1, 64, W, seed=1)
    O.forward(sd, a, b, 0.5)                                   # warm-up (mkldnn primitive caches)
    t0 = time.perf_counter(); O.forward(sd, a, b, 0.5); probe = time.perf_counter() - t0
    px_per_s = 64 * W / probe
    strip_h = int(min(H, max(64, (budget_s * px_per_s / W) // 16 * 16)))
    a, b = O.seeded_frames(1, strip_h, W, seed=1)
    return sd, a, b, strip_h


def run_reference(args):
    import torch
    from oracle import rrin_oracle as O
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                                  # rank 0 alone runs the CPU arm
    threads = os.cpu_count() or 1
    cfg = CONFIGS[args.config]
    H, W = cfg["h"], cfg["w"]
    B = args.batch if args.batch else cfg["batch"]
    budget_total = 150.0
    per_step = budget_total / max(1, args.steps + args.warmup)
    sd, a, b, strip_h = cpu_reference_throughput(per_step, threads, H, W)
    for _ in range(args.warmup):
        O.forward(sd, a, b, 0.5)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.forward(sd, a, b, 0.5)
    dt = time.perf_counter() - t0
    frac = strip_h / H
    fps = args.steps * frac / dt
    sample = f"{strip_h}x{W} strip of one {H}x{W} frame pair (= {frac:.4f} frame) per step, fp32, t=0.5"
    metric = METRIC if args.config in ("1080p", "clip") else f"{args.config}_interpolated_frames_per_sec"
    line = {"impl": "reference", "metric": metric, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "strong" if cfg["clip"] else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["what"].format(B=B) + ", random-init weights (torch.manual_seed(0))", "name": args.config,
                       "reference_path": "oracle port of Net.forward on torch CPU operators (the reference is Python and cannot travel to the GPU box); "
                                         "each step is a bounded sample of the workload: " + sample},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ----------------------------------------------------------------------------- GPU library arm
def run_library(args):
    """SURVEY.md 8(d) "GPU library baseline": the oracle's restatement of Net.forward issued as torch eager operators on
    cuda:0 -- cuDNN convolutions, ATen grid_sample / upsample / pooling -- i.e. what the unmodified reference executes on a
    B200 today (the reference itself cannot travel to the GPU box).  Three precisions; `value` is the fastest.  None of
    rrin_b200's kernels run here; the line is context for the headline, not a product path."""
    import torch
    from oracle import rrin_oracle as O
    if int(os.environ.get("RANK", "0")) != 0:
        return
    dev = torch.device("cuda", 0)
    torch.backends.cudnn.benchmark = True
    sd32 = {k: v.to(dev) for k, v in O.seeded_state_dict().items()}
    a, b = (x.to(dev) for x in O.seeded_frames(1, H, W, seed=1, smooth=True))
    K, Wm = max(1, min(args.steps, 20)), max(args.warmup, 3)

    def timed(fn):
        for _ in range(Wm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return K * 1e3 / e0.elapsed_time(e1)

    modes = {}
    with torch.no_grad():
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        modes["fp32"] = timed(lambda: O.forward(sd32, a, b, 0.5))
        torch.backends.cudnn.allow_tf32 = True                  # PyTorch's default for cuDNN convolutions
        modes["tf32"] = timed(lambda: O.forward(sd32, a, b, 0.5))
        sdcl = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd32.items()}
        acl, bcl = a.contiguous(memory_format=torch.channels_last), b.contiguous(memory_format=torch.channels_last)

        def bf16():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return O.forward(sdcl, acl, bcl, 0.5)
        modes["bf16_autocast_channels_last"] = timed(bf16)
    best = max(modes, key=modes.get)
    line = {"impl": "library", "metric": METRIC, "value": modes[best], "unit": "frames/s", "n_gpus": 1, "steps": K, "warmup": Wm,
            "ms_per_step": 1e3 / modes[best], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": best,
            "data": "synthetic",
            "config": {"workload": "1080p (1920x1088) 2x interpolation, one frame pair per step, t=0.5, random-init weights",
                       "path": "oracle restatement of Net.forward as torch eager operators on the GPU (cuDNN convolutions, "
                               "cudnn.benchmark=True); device-resident inputs"},
            "modes_frames_per_sec": modes}
    emit(line)


# ----------------------------------------------------------------------------- GPU arm
FLOP_PER_PX_FLOW = 521_856              # the t-independent Flow U-Net's share of FLOP_PER_PX

# BASELINE.json configs[1..4]; "1080p" (configs[2], the configuration the metric is quoted on) is the default line.
CONFIGS = {
    "1080p": dict(h=1088, w=1920, batch=4, ts=None, clip=False,
                  what="1080p (1920x1088) 2x interpolation of a synthetic clip, one batch of {B} consecutive frame pair(s) per step, t=0.5"),
    "720p_b8": dict(h=736, w=1280, batch=8, ts=None, clip=False,
                    what="720p (1280x720 padded to 1280x736) batch of {B} frame pairs per step, t=0.5"),
    "1080p_t7": dict(h=1088, w=1920, batch=1, ts=[k / 8 for k in range(1, 8)], clip=False,
                     what="1080p (1920x1088) 8x slow motion: 7 timesteps t=k/8 of one frame pair per step, Flow U-Net computed once per pair"),
    "4k": dict(h=2176, w=3840, batch=1, ts=None, clip=False,
               what="4K (3840x2160 padded to 3840x2176) one frame pair per step, t=0.5"),
    "clip": dict(h=1088, w=1920, batch=4, ts=None, clip=True,
                 what="1080p (1920x1088) 2x interpolation of the whole 240-frame synthetic clip per step (239 frame pairs split into "
                      "contiguous shards over the ranks, batches of {B} pairs), t=0.5"),
}


def synth_frames(n, h, w, dev, seed):
    """n consecutive frames of a synthetic clip on the device: smooth content (bicubic-upsampled noise) shifted by (2, 3)
    pixels per frame plus 5 % white noise, U[0,1)."""
    import torch
    g = torch.Generator(device=dev).manual_seed(seed)
    lo = torch.rand(1, 3, h // 8 + 8 + n // 2, w // 8 + 8 + n // 2, generator=g, device=dev)
    big = torch.nn.functional.interpolate(lo, scale_factor=8, mode="bicubic", align_corners=False).clamp_(0, 1)
    noise = torch.rand(1, 3, h, w, generator=g, device=dev) * 0.05
    return [(big[:, :, 8 + 2 * i: 8 + 2 * i + h, 8 + 3 * i: 8 + 3 * i + w] * 0.95 + noise).contiguous() for i in range(n)]


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from rrin_b200 import ClipInterpolator, Net, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback in the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    cfg = CONFIGS[args.config]
    H, W, ts, clip_mode = cfg["h"], cfg["w"], cfg["ts"], cfg["clip"]
    B = args.batch if args.batch else cfg["batch"]
    metric = METRIC if args.config in ("1080p", "clip") else f"{args.config}_interpolated_frames_per_sec"

    # weights: random init of the reference architecture -- torch.manual_seed(0); Net() draws the same RNG stream as the
    # reference's model.Net() (tests/test_oracle.py pins the sha256) -- loaded via state_dict like convert.py:100-104
    torch.manual_seed(0)
    sd = {k: v.detach().clone().float() for k, v in Net().state_dict().items()}
    net = Net()
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    net.precision = args.precision

    K, Wm = args.steps, max(args.warmup, 3)
    lo_pair, hi_pair = sharding.pair_range(CLIP_FRAMES, rank, world)
    pairs_per_rank = hi_pair - lo_pair

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------- the step
    if clip_mode:
        # strong scaling: this rank's contiguous shard of the 240-frame clip, device resident (its pairs + the boundary frame)
        frames = synth_frames(pairs_per_rank + 1, H, W, dev, 1000)       # same clip on every rank; the rank takes its own range
        stack = torch.cat(frames)
        batches = [(stack[i:min(i + B, pairs_per_rank)], stack[i + 1:min(i + B, pairs_per_rank) + 1]) for i in range(0, pairs_per_rank, B)]
        frames_per_step_rank = pairs_per_rank
        frames_per_step_job = CLIP_FRAMES - 1

        def step(i):
            y = None
            for a, b in batches:
                y = net(a, b, t=0.5)
            return y
        flop_per_step_job = frames_per_step_job * FLOP_PER_PX * H * W
    else:
        n_frames = (min(pairs_per_rank, 12) if ts is None and H * W <= 1088 * 1920 else 4) + 1     # frames kept resident; steps cycle over them
        frames = synth_frames(n_frames, H, W, dev, 1000 + rank)
        if ts is None:
            def pair(i):
                js = [(i * B + k) % (n_frames - 1) for k in range(B)]
                if B == 1:
                    return frames[js[0]], frames[js[0] + 1]
                return torch.cat([frames[j] for j in js]), torch.cat([frames[j + 1] for j in js])
            batches = [pair(i) for i in range(max(1, min(K + Wm, (n_frames - 1) // B + 1)))]   # assembled outside the timed region
            frames_per_step_rank = B

            def step(i):
                a, b = batches[i % len(batches)]
                return net(a, b, t=0.5)
            flop_per_step_job = world * B * FLOP_PER_PX * H * W
        else:
            batches = [(frames[j], frames[j + 1]) for j in range(n_frames - 1)]
            frames_per_step_rank = len(ts)

            def step(i):
                a, b = batches[i % len(batches)]
                return net.forward_multi(a, b, ts)
            flop_per_step_job = world * H * W * (FLOP_PER_PX_FLOW + len(ts) * (FLOP_PER_PX - FLOP_PER_PX_FLOW))   # algorithmic: Flow once per pair
        frames_per_step_job = world * frames_per_step_rank

    # ---------------- device-resident throughput (`value`)
    clocks = ClockSampler(local)                 # started (and running) before the warm-up, see wait_ready
    clocks.wait_ready()
    for i in range(Wm):
        step(i)
    barrier()
    t_wall0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        y = step(i)
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms_max = max_over_ranks(e0.elapsed_time(e1))
    clk = clocks.window(t_wall0, t_wall1)
    value = K * frames_per_step_job / (ms_max * 1e-3)

    # ---------------- end to end with host buffers (`e2e`): the streaming clip pipeline (rrin_b200.ClipInterpolator) over a
    # pinned host clip -- every frame crosses PCIe once host->device, every interpolated frame once device->host, copies
    # overlapped with compute; the timed region covers all copies and ends with a host sync.
    sf = 1 if ts is None else len(ts)
    pb = B if ts is None else 2                                  # pairs per pipeline stage
    ke = max(3, min(K, 80 * 1088 * 1920 // (H * W * pb * sf)))          # about 80 1080p-size output frames of pinned host memory
    if clip_mode:
        n_clip_pairs = pairs_per_rank
    else:
        n_clip_pairs = ke * pb
    host_clip = torch.empty(n_clip_pairs + 1, 3, H, W).pin_memory()
    for i in range(host_clip.shape[0]):
        host_clip[i].copy_(frames[i % len(frames)][0])
    out_host = torch.empty(n_clip_pairs * sf, 3, H, W).pin_memory()
    pipe = ClipInterpolator(net, H, W, batch=pb, sf=sf)
    nw = min(n_clip_pairs, 4 * pb)                               # four batches: each staging slot is seen twice, so its CUDA graph exists
    pipe.run(host_clip[:nw + 1], out_host[:nw * sf])             # warm-up (engines, graphs, events)
    # The GPU has idled while the pinned clip was filled, and the timed pipeline run is shorter than the K-step loop above: bring
    # the chip back under the sustained load (power cap, clocks) of that loop first -- untimed whole-clip runs for about as long
    # as the K steps took -- so that `e2e` is not a cold-start burst number.
    for _ in range(1 if clip_mode else max(1, -(-K // ke) - 1)):
        pipe.run(host_clip, out_host)
    barrier()
    t_wall2 = time.time()
    e0.record()
    pipe.run(host_clip, out_host)
    e1.record()
    barrier()
    clk_e2e = clocks.window(t_wall2, time.time())
    clocks.close()
    t_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_frames_job = (CLIP_FRAMES - 1) if clip_mode else world * n_clip_pairs * sf
    e2e_steps = 1 if clip_mode else ke
    e2e_value = e2e_frames_job / (t_e2e * 1e-3)
    h2d_step, d2h_step = pipe.h2d_bytes / e2e_steps, pipe.d2h_bytes / e2e_steps
    e2e = {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d_step, "d2h_bytes_per_step": d2h_step, "steps": e2e_steps,
           "api": "rrin_b200.ClipInterpolator.run(pinned host clip) -> pinned host frames: each source frame uploaded once, "
                  "H2D / forward / D2H of successive batches on three streams, one host sync at the end",
           "sm_mhz": clk_e2e.get("sm_mhz")}
    del host_clip, out_host, pipe
    if args.config == "1080p":
        # same pipeline fed with the bytes an image decoder produces (uint8 HWC 1920x1080; Pad + ToTensor and to_pil + crop of
        # dataloader.py:93-118 / utils.py:51-58 on the device): 4x fewer PCIe bytes.  Reported beside `e2e`, not instead of it.
        clip8 = torch.empty(ke * B + 1, 1080, W, 3, dtype=torch.uint8).pin_memory()
        for i in range(clip8.shape[0]):
            clip8[i].copy_((frames[i % len(frames)][0, :, 8:, :] * 255).to(torch.uint8).permute(1, 2, 0))
        out8 = torch.empty(ke * B, 1080, W, 3, dtype=torch.uint8).pin_memory()
        pipe8 = ClipInterpolator(net, 1080, W, batch=B, sf=1, uint8=True)
        pipe8.run(clip8[:4 * B + 1], out8[:4 * B])
        barrier()
        e0.record()
        pipe8.run(clip8, out8)
        e1.record()
        barrier()
        t3 = max_over_ranks(e0.elapsed_time(e1))
        e2e["uint8_frames"] = {"value": world * ke * B / (t3 * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": pipe8.h2d_bytes / ke,
                               "d2h_bytes_per_step": pipe8.d2h_bytes / ke,
                               "api": "ClipInterpolator(uint8=True): 1920x1080x3 uint8 HWC frames in and out, pad/ToTensor/to_pil/crop on the device"}
        del clip8, out8, pipe8
        # the reference-shaped call (convert.py:130-133: upload both frames, forward, download, sync -- per step) for comparison
        hp = [(a.cpu().pin_memory(), b.cpu().pin_memory()) for a, b in batches[:2]]
        oh = torch.empty(B, 3, H, W).pin_memory()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(5):
            a, b = hp[i % len(hp)]
            oh.copy_(net(a.cuda(non_blocking=True), b.cuda(non_blocking=True), t=0.5), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        e2e["per_call_sync_frames_per_sec"] = 5 * B / (time.perf_counter() - t0) * world
        e2e["per_call_sync_api"] = "Net.forward(img1.cuda(), img2.cuda(), t) then .cpu() and a sync per step (convert.py:130-133)"
        del hp, oh

    # ---------------- single-pair calls (what convert.py:95-97 issues: batch_size=1), device resident
    batch1 = None
    if args.config == "1080p" and B != 1:
        a1, b1 = frames[0], frames[1]
        for _ in range(3):
            net(a1, b1, t=0.5)
        barrier()
        e0.record()
        for i in range(20):
            net(frames[i % (len(frames) - 1)], frames[i % (len(frames) - 1) + 1], t=0.5)
        e1.record()
        barrier()
        batch1 = {"value": world * 20 / (max_over_ranks(e0.elapsed_time(e1)) * 1e-3), "unit": "frames/s",
                  "what": "Net.forward on one frame pair per call (batch 1), device-resident inputs"}

    line = None
    if rank == 0:
        key = next(k for k in net._engines if k[2] == (B if ts is None else len(ts)) and k[3] == H and k[4] == W)
        eng = net._engines[key]
        w = net._weights(dev)
        pa, pbt = batches[0]
        tt = 0.5 if ts is None else list(ts)
        # per-launch device times (CUDA events on the launching stream): per-launch median over 5 profiled forwards, each
        # issued right behind two unprofiled ones so the clocks are those of the sustained step (a mean would let one stall --
        # a clock dip under the power cap -- land on whichever launch it hit).  Events between launches disable the
        # programmatic-dependent-launch overlap, so the sum of these is an upper bound of the kernel time inside a step.
        table = eng.launch_table()
        samples = []
        one = (lambda: net(pa, pbt, t=0.5)) if ts is None else (lambda: net.forward_multi(pa, pbt, ts))
        for r in range(5):
            one(); one()
            samples.append(eng.profile(w, pa[:eng.n_pairs], pbt[:eng.n_pairs], tt))
        acc = [statistics.median(s[i] for s in samples) for i in range(len(table))]
        classes = {}
        for (name, layer, fl, by), t in zip(table, acc):
            c = classes.setdefault(name, {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "launches": 0})
            c["ms"] += t; c["flops"] += fl; c["bytes"] += by; c["launches"] += 1
        sum_ms = sum(acc)
        dom_name, dom = max(classes.items(), key=lambda kv: kv[1]["ms"])
        conv_ms = sum(c["ms"] for n, c in classes.items() if n.startswith("conv"))
        conv_fl = sum(c["flops"] for n, c in classes.items() if n.startswith("conv"))
        # HBM-bound glue: pack_pair plus the four `last` convs whose epilogues carry the fused t-scale / warp / blend / clamp
        is_glue = lambda n: (not n.startswith("conv")) or n.endswith("+glue")
        glue_ms = sum(c["ms"] for n, c in classes.items() if is_glue(n))
        glue_by = sum(c["bytes"] for n, c in classes.items() if is_glue(n))
        peak_tf = peaks["tf_sustained"]             # kernels timed inside a long step -> sustained peak
        ach = dom["flops"] / (dom["ms"] * 1e-3) / 1e12
        # layer-wise roofline of the whole forward: every launch at max(tensor time, HBM time) of its algorithmic work
        roof_ms = sum(max(fl / (peak_tf * 1e12), by / (peaks["hbm_gbs"] * 1e9)) * 1e3 for (_, _, fl, by) in table)
        traffic, traffic_of = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")      # dram bytes of one launch of the class (ncu --set full capture)
        if os.path.exists(tp) and args.config == "1080p":
            with open(tp) as f:
                tj = json.load(f)
            ent = tj.get(dom_name)
            if isinstance(ent, dict):
                traffic = ent["traffic"]
                traffic_of = {k: ent[k] for k in ("launch", "algorithmic", "ratio") if k in ent}
                traffic_of["capture"] = tj.get("_note", "")
            else:
                traffic = ent
        per_launch = [{"i": i, "kernel": n, "layer": ly, "ms": round(t, 4)} for i, ((n, ly, _, _), t) in enumerate(zip(table, acc))]
        warp_i = next((i for i, (n, ly, _, _) in enumerate(table) if ly == "refine_flow.last" and n.endswith("+glue")), None)
        blend_i = next((i for i, (n, ly, _, _) in enumerate(table) if ly == "Mask.last" and n.endswith("+glue")), None)
        # Structural zeros inside the tensors these launches exchange (bytes per full-resolution pixel and sample, written or read):
        # the packed 16-channel bf16 head inputs carry 6 / 10 / 16 / 9 real channels (Flow / refine_flow / Mask / final), xt8 holds
        # 6 real floats of 8, out4 3 of 4.  They are real HBM traffic of this design, but not bytes the reference's tensors have:
        # `*_without_padding` rates leave them out.
        n_smp = eng.n
        pad_px = {"pack_pair": 20.0, "Flow.last": 12.0, "refine_flow.last": 8.0, "Mask.last": 8.0 + 4.0 + 14.0, "final.last": 4.0}
        pad_of = lambda n, ly: (pad_px.get("pack_pair", 0.0) * eng.n_pairs if n == "pack_pair" else pad_px.get(ly, 0.0) * n_smp) * H * W if is_glue(n) else 0.0
        glue_pad = sum(pad_of(n, ly) for (n, ly, _, _) in table)
        glue_launches = {}
        for nm, i in (("refine_flow.last + residue add + both warps + Mask head pack", warp_i), ("Mask.last + sigmoid + blend + final head pack", blend_i)):
            if i is not None:
                useful = table[i][3] - pad_of(table[i][0], table[i][1])
                glue_launches[nm] = {"ms": round(acc[i], 4), "gbs": round(table[i][3] / (acc[i] * 1e-3) / 1e9, 1),
                                     "frac_of_hbm_peak": round(table[i][3] / (acc[i] * 1e-3) / 1e9 / peaks["hbm_gbs"], 3),
                                     "gbs_without_padding": round(useful / (acc[i] * 1e-3) / 1e9, 1),
                                     "frac_without_padding": round(useful / (acc[i] * 1e-3) / 1e9 / peaks["hbm_gbs"], 3)}
        roofline = {"bound": "tensor", "kernel": dom_name, "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": ach / peak_tf, "traffic": traffic, "traffic_of": traffic_of,
                    "peak_source": peaks["source"] + ", bf16 sustained (kernel timed inside a long step)",
                    "launches_per_step": dom["launches"], "avg_launch_ms": dom["ms"] / dom["launches"],
                    "share_of_step": dom["ms"] / sum_ms,
                    "all_convs": {"achieved": conv_fl / (conv_ms * 1e-3) / 1e12, "frac": conv_fl / (conv_ms * 1e-3) / 1e12 / peak_tf,
                                  "share_of_step": conv_ms / sum_ms},
                    "whole_step": {"achieved": value * flop_per_step_job / frames_per_step_job / world / 1e12, "unit": "TFLOP/s per GPU",
                                   "frac_of_burst_peak": value * flop_per_step_job / frames_per_step_job / world / 1e12 / peaks["tf_burst"],
                                   "frac_of_sustained_peak": value * flop_per_step_job / frames_per_step_job / world / 1e12 / peaks["tf_sustained"],
                                   "layerwise_roofline_ms": roof_ms, "frac_of_layerwise_roofline": roof_ms / sum_ms,
                                   "note": "layer-wise roofline = sum over launches of max(FLOP / sustained bf16 peak, algorithmic bytes / HBM copy peak)"},
                    "warp_blend_glue": {"bound": "hbm", "kernels": sorted(n for n in classes if is_glue(n)),
                                        "achieved": glue_by / (glue_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                                        "unit": "GB/s", "frac": glue_by / (glue_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                        "frac_without_padding": (glue_by - glue_pad) / (glue_ms * 1e-3) / 1e9 / peaks["hbm_gbs"],
                                        "share_of_step": glue_ms / sum_ms, "launches": glue_launches},
                    "classes": {n: {"ms": round(c["ms"], 4), "launches": c["launches"],
                                    "tflops": round(c["flops"] / (c["ms"] * 1e-3) / 1e12, 1) if c["flops"] else 0,
                                    "gbs": round(c["bytes"] / (c["ms"] * 1e-3) / 1e9, 1)} for n, c in classes.items()}}
        if args.per_launch:
            roofline["per_launch"] = per_launch
        # the warp under the flows a trained model produces: Flow.last scaled x200 (|flow| of several pixels, out-of-bounds
        # taps) on the same smooth frames -- same launch, same bytes, scattered gathers
        if args.config in ("1080p", "4k") and warp_i is not None:
            sd2 = {k: v.clone() for k, v in sd.items()}
            sd2["Flow.last.weight"] *= 200.0
            sd2["Flow.last.bias"] *= 200.0
            net2 = Net()
            net2.load_state_dict(sd2, strict=True)
            net2 = net2.cuda().eval()
            net2.precision = args.precision
            for _ in range(2):
                net2(pa, pbt, t=0.5)
            eng2 = net2._engines[next(iter(net2._engines))]
            w2 = net2._weights(dev)
            s2 = []
            for r in range(3):
                net2(pa, pbt, t=0.5)
                s2.append(eng2.profile(w2, pa, pbt, 0.5)[warp_i])
            fl = eng2.tap(0)
            t2 = statistics.median(s2)
            roofline["warp_blend_glue"]["warp_under_stress_flow"] = {
                "weights": "Flow.last x200", "flow_abs_max_px": float(fl.abs().max().item()), "flow_abs_mean_px": float(fl.abs().mean().item()),
                "ms": round(t2, 4), "gbs": round(table[warp_i][3] / (t2 * 1e-3) / 1e9, 1),
                "frac_of_hbm_peak": round(table[warp_i][3] / (t2 * 1e-3) / 1e9 / peaks["hbm_gbs"], 3),
                "same_launch_at_random_init_ms": round(acc[warp_i], 4)}
            del net2, eng2, w2
        gs = eng.graph_stats()
        step_ms = ms_max / K
        n_fwd = len(batches) if clip_mode else 1
        line = {"metric": metric, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm,
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong" if clip_mode else "weak", "vs_baseline": None,
                "dtype": args.precision, "data": "synthetic",
                "config": {"workload": cfg["what"].format(B=B) + ", random-init weights (torch.manual_seed(0))", "name": args.config,
                           "batch": B, "arithmetic": f"{args.precision} operands on tcgen05 tensor cores, fp32 accumulation; flows / mask logits / blend in fp32",
                           "sharding": f"{world} rank(s), contiguous shards of the {CLIP_FRAMES}-frame clip, no collective",
                           "l2": "per-step working set (>= 0.7 GB of activations per frame pair) exceeds the 126 MB L2; no explicit flush",
                           "tflop_per_frame": flop_per_step_job / frames_per_step_job / 1e12,
                           "tensor_frac_of_burst_peak": value / world * flop_per_step_job / frames_per_step_job / 1e12 / peaks["tf_burst"],
                           "tensor_frac_of_sustained_peak": value / world * flop_per_step_job / frames_per_step_job / 1e12 / peaks["tf_sustained"]},
                "clocks": clk, "e2e": e2e, "gpu_launches": eng.num_launches * K * n_fwd,
                "launch_gap": {"step_ms": step_ms, "sum_kernel_ms": sum_ms * n_fwd if not clip_mode else None,
                               "gap_frac": (step_ms - sum_ms) / step_ms if not clip_mode else None,
                               "cuda_graph": {"replayed": gs[0], "direct": gs[1], "graphs": gs[2]},
                               "note": "sum_kernel_ms: per-launch CUDA-event times (events between launches switch the programmatic-dependent-launch "
                                       "overlap off, so this is an upper bound); `value` launches kernel by kernel (Net.forward allocates its result), "
                                       "the e2e pipeline replays one CUDA graph per forward (cuda_graph counts)"},
                "roofline": roofline}
        if batch1 is not None:
            line["batch1"] = batch1
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        if world == 1 and not args.no_cpu:
            # CPU baseline: oracle port on a bounded strip of the same workload, this box's host cores, N=1 only
            # (the only place the GPU arm touches oracle/; after the process group is gone, so no GPU waits for it)
            from oracle import rrin_oracle as O
            threads = os.cpu_count() or 1
            sdc, a, b, strip_h = cpu_reference_throughput(15.0, threads, H, W)
            t0 = time.perf_counter(); O.forward(sdc, a, b, 0.5); dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": (strip_h / H) / dt, "unit": "frames/s", "cores": threads, "kind": "port",
                                    "sample": f"one {strip_h}x{W} strip of the {H}x{W} pair (= {strip_h / H:.4f} frame), fp32, t=0.5, {dt:.1f} s"}
        emit(line)


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly one JSON line.  Libraries print there too (NCCL's version banner at communicator init goes
    through C stdio), so file descriptor 1 is pointed at stderr for the whole run and the line is written to the saved
    descriptor at the end."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=0, help="frame pairs per step (per GPU); 0 = the configuration's own")
    ap.add_argument("--config", default="1080p", choices=sorted(CONFIGS), help="named BASELINE.json configuration (default: the headline one)")
    ap.add_argument("--precision", default="bf16", help="operand format of the tensor-core path (Net.precision)")
    ap.add_argument("--per-launch", action="store_true", help="add the per-launch time table to the JSON line")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--impl", default="rrin_b200", choices=["rrin_b200", "reference", "library"])
    args = ap.parse_args()
    if args.impl == "library":
        return run_library(args)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
