"""Minimal program for ncu: W warm-up forwards + 1 forward of the 1080p step (no timing, no CPU work).
Usage: [BATCH=b] python tools/profile_step.py [H W [warmups]]     (BATCH defaults to bench.py's 4 frame pairs per step)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rrin_b200 import Net

h = int(sys.argv[1]) if len(sys.argv) > 1 else 1088
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 2
torch.manual_seed(0)
net = Net().cuda().eval()          # random-init weights of the reference architecture
g = torch.Generator().manual_seed(2)
lo = torch.rand(1, 3, h // 8 + 2, w // 8 + 2, generator=g)          # smooth synthetic content, second frame shifted
big = torch.nn.functional.interpolate(lo, size=(h + 16, w + 16), mode="bicubic", align_corners=False).clamp(0, 1)
a, b = big[:, :, 8:8 + h, 8:8 + w].contiguous(), big[:, :, 5:5 + h, 3:3 + w].contiguous()
nb = int(os.environ.get("BATCH", "4"))
a, b = a.cuda().expand(nb, -1, -1, -1).contiguous(), b.cuda().expand(nb, -1, -1, -1).contiguous()
for _ in range(warm + 1):
    y = net(a, b, t=0.5)
torch.cuda.synchronize()
print("ok", float(y.mean()))
