"""Small program for compute-sanitizer (memcheck / synccheck / initcheck): one pass over every launch class of the
library on frames just large enough to switch on the folded upsample convs and their border rings (H, W > 64).

Usage (under gpurun, see tools/r02_sanitize.sh):
    compute-sanitizer --tool memcheck python tools/sanitize_step.py [H W]

Covers: Net.forward batch 2 in both orientations (plain and transposed level >= 2 launches), forward_multi (Flow once,
three timesteps), the fp16 precision mode, the public warp(), forward_into through a CUDA graph, and the uint8 streaming
pipeline (frame_from_u8 / frame_to_u8).  Prints `ok` and a checksum per leg; any sanitizer finding is in its own report."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rrin_b200 import ClipInterpolator, Net, warp

h = int(sys.argv[1]) if len(sys.argv) > 1 else 128
w = int(sys.argv[2]) if len(sys.argv) > 2 else 208
legs = set(sys.argv[3].split(",")) if len(sys.argv) > 3 else {"fwd", "fwd_t", "multi", "fp16", "warp", "graph", "u8"}
torch.manual_seed(0)
net = Net().cuda().eval()
g = torch.Generator().manual_seed(3)


def frames(n, hh, ww):
    return torch.rand(n, 3, hh, ww, generator=g).cuda(), torch.rand(n, 3, hh, ww, generator=g).cuda()


def done(tag, y):
    torch.cuda.synchronize()
    print("ok", tag, tuple(y.shape), f"{float(y.double().mean()):.6f}", flush=True)


with torch.no_grad():
    if "fwd" in legs:
        a, b = frames(2, h, w)
        done("forward", net(a, b, t=0.5))
    if "fwd_t" in legs:
        a, b = frames(1, w, h)                                   # the other orientation
        done("forward (W x H)", net(a, b, t=0.25))
    if "multi" in legs:
        a, b = frames(1, h, w)
        done("forward_multi", net.forward_multi(a, b, [0.25, 0.5, 0.75]))
    if "graph" in legs:
        a, b = frames(2, h, w)
        out = torch.empty(2, 3, h, w, device="cuda")
        for _ in range(3):                                       # the second call per pointer set instantiates the graph
            net.forward_into(a, b, 0.5, out)
        done("forward_into (graph)", out)
    if "fp16" in legs:
        net.precision = "fp16"
        a, b = frames(1, h, w)
        done("forward fp16", net(a, b, t=0.5))
        net.precision = "bf16"
    if "warp" in legs:
        img = torch.rand(2, 3, 40, 56, generator=g).cuda()
        flow = (torch.rand(2, 2, 40, 56, generator=g) * 12 - 6).cuda()
        done("warp", warp(img, flow))
    if "u8" in legs:
        h0, w0 = h - 8, w                                        # rows padded up to h on the device like dataloader.py:93-108 (the width never is)
        clip = torch.randint(0, 256, (4, h0, w0, 3), dtype=torch.uint8, generator=g).pin_memory()
        ci = ClipInterpolator(net, h0, w0, batch=2, sf=1, uint8=True)
        done("ClipInterpolator uint8", ci.run(clip))
print("sanitize_step finished")
