"""Per-launch-class device times of one 1080p Net.forward (CUDA events after every launch, averaged
over a few forwards) -- the quick look used while tuning kernels.  Usage: python tools/step_table.py [H W [reps]]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rrin_b200 import Net

h = int(sys.argv[1]) if len(sys.argv) > 1 else 1088
w = int(sys.argv[2]) if len(sys.argv) > 2 else 1920
reps = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 5
nb = int(os.environ.get("BATCH", "1"))
torch.manual_seed(0)
net = Net().cuda().eval()          # random-init weights of the reference architecture
g = torch.Generator().manual_seed(2)
lo = torch.rand(1, 3, h // 8 + 2, w // 8 + 2, generator=g)          # smooth synthetic content, second frame shifted
big = torch.nn.functional.interpolate(lo, size=(h + 16, w + 16), mode="bicubic", align_corners=False).clamp(0, 1)
a, b = big[:, :, 8:8 + h, 8:8 + w].contiguous(), big[:, :, 5:5 + h, 3:3 + w].contiguous()
a, b = a.cuda().expand(nb, -1, -1, -1).contiguous(), b.cuda().expand(nb, -1, -1, -1).contiguous()
for _ in range(3):
    net(a, b, t=0.5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    net(a, b, t=0.5)
e1.record()
torch.cuda.synchronize()
print(f"forward {nb}x{h}x{w}: {e0.elapsed_time(e1) / 20:.3f} ms  ({nb * 20e3 / e0.elapsed_time(e1):.1f} frames/s)")
eng = net._engines[next(iter(net._engines))]
wts = net._weights(a.device)
table = eng.launch_table()
acc = [0.0] * len(table)
for _ in range(reps):
    for i, v in enumerate(eng.profile(wts, a, b, 0.5)):
        acc[i] += v / reps
cls = {}
for (name, layer, fl, by), t in zip(table, acc):
    c = cls.setdefault(name, [0, 0.0, 0.0, 0.0]); c[0] += 1; c[1] += t; c[2] += fl; c[3] += by
tot = sum(acc)
print(f"sum of launches {tot:.3f} ms")
print(f"{'class':58s} {'n':>3s} {'ms':>7s} {'%':>5s} {'TF/s':>7s} {'GB/s':>7s}")
for n, (k, t, fl, by) in sorted(cls.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:58s} {k:3d} {t:7.3f} {100 * t / tot:5.1f} {fl / t / 1e9:7.1f} {by / t / 1e6:7.1f}")
if "-v" in sys.argv:
    for (name, layer, fl, by), t in zip(table, acc):
        print(f"{t * 1e3:8.1f} us  {fl / max(t, 1e-9) / 1e9:7.1f} TF/s  {name:55s} {layer}")
