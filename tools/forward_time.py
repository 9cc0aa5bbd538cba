"""Forward time only (CUDA events, after warm-up): BATCH=b python tools/forward_time.py [iters] -- for quick A/B of env switches."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rrin_b200 import Net

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
nb = int(os.environ.get("BATCH", "4"))
torch.manual_seed(0)
net = Net().cuda().eval()
g = torch.Generator(device="cuda").manual_seed(2)
a, b = (torch.rand(nb, 3, 1088, 1920, generator=g, device="cuda") for _ in range(2))
for _ in range(5):
    net(a, b, t=0.5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    net(a, b, t=0.5)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print(f"forward {nb}x1088x1920: {ms:.3f} ms  ({nb * 1e3 / ms:.1f} frames/s)")
