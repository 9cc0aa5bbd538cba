"""Device-resident throughput of the other named configurations of BASELINE.json (configs[1], [3], [4]); bench.py's line is
configs[2].  Random-init weights, synthetic frames, CUDA-event timing after warm-up; one JSON object on stdout.
Usage: python tools/named_configs.py [iters]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rrin_b200 import Net

FLOP_PX = 1736064.0                 # conv FLOP per padded pixel per forward
FLOP_PX_FLOW = 521856.0             # the t-independent Flow U-Net's share
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
torch.manual_seed(0)
net = Net().cuda().eval()
g = torch.Generator(device="cuda").manual_seed(1)


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


out = {}
with torch.no_grad():
    # configs[1]: 720p padded to 1280x736, batch of 8 frame pairs, t = 0.5
    a, b = (torch.rand(8, 3, 736, 1280, generator=g, device="cuda") for _ in range(2))
    ms = timed(lambda: net(a, b, t=0.5), iters)
    out["720p_batch8"] = {"ms_per_batch": ms, "frames_per_sec": 8e3 / ms, "conv_tflops": 8 * 736 * 1280 * FLOP_PX / ms / 1e9}
    del a, b
    # configs[3]: 1080p, 7 intermediate timesteps t = k/8 per pair, Flow U-Net computed once per pair
    a, b = (torch.rand(1, 3, 1088, 1920, generator=g, device="cuda") for _ in range(2))
    ts = [k / 8 for k in range(1, 8)]
    ms = timed(lambda: net.forward_multi(a, b, ts), iters)
    px = 1088 * 1920
    out["1080p_7_timesteps"] = {"ms_per_pair": ms, "frames_per_sec": 7e3 / ms,
                                "conv_tflops_algorithmic": px * (FLOP_PX_FLOW + 7 * (FLOP_PX - FLOP_PX_FLOW)) / ms / 1e9,
                                "conv_tflops_reference_equivalent": 7 * px * FLOP_PX / ms / 1e9}
    del a, b
    # configs[4]: 4K padded to 3840x2176, one frame pair, t = 0.5
    a, b = (torch.rand(1, 3, 2176, 3840, generator=g, device="cuda") for _ in range(2))
    ms = timed(lambda: net(a, b, t=0.5), iters)
    out["4k_pair"] = {"ms_per_frame": ms, "frames_per_sec": 1e3 / ms, "conv_tflops": 2176 * 3840 * FLOP_PX / ms / 1e9}
print(json.dumps(out, indent=1))
