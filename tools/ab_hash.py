"""Output fingerprints of a few forwards, for bit-equality checks between two builds of the library:
    RRIN_LIB=<build A> python tools/ab_hash.py > a.txt;  RRIN_LIB=<build B> python tools/ab_hash.py > b.txt;  diff a.txt b.txt"""
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rrin_b200 import Net

torch.manual_seed(0)
net = Net().cuda().eval()
g = torch.Generator().manual_seed(4)
with torch.no_grad():
    for prec in ("bf16", "fp16"):
        net.precision = prec
        for (n, h, w) in [(1, 64, 96), (2, 128, 208), (1, 208, 128), (1, 368, 368), (2, 1088, 1920), (1, 2176, 3840)]:
            a, b = torch.rand(n, 3, h, w, generator=g).cuda(), torch.rand(n, 3, h, w, generator=g).cuda()
            y = net(a, b, t=0.5)
            print(prec, n, h, w, hashlib.sha256(y.cpu().numpy().tobytes()).hexdigest()[:16], f"{float(y.double().mean()):.6f}")
        a, b = torch.rand(1, 3, 1088, 1920, generator=g).cuda(), torch.rand(1, 3, 1088, 1920, generator=g).cuda()
        y = net.forward_multi(a, b, [0.25, 0.5, 0.75])
        print(prec, "multi", hashlib.sha256(y.cpu().numpy().tobytes()).hexdigest()[:16])
