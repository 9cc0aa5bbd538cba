"""End-to-end rate of convert_folder (PNG folder in -> PNG folder out) at 1080p: usage  python tools/convert_bench.py [frames]
Synthetic smooth frames with a little noise (PNG-compressible like video frames); random-init weights; prints output frames/s
(interpolated + copied originals, as the reference counts its progress bar) for PIL at its default zlib level and at level 1,
and for the lean writer (rrin_b200.fastpng) at levels 1 and 0."""
import os
import shutil
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from PIL import Image

from rrin_b200 import Net, convert_folder

n = int(sys.argv[1]) if len(sys.argv) > 1 else 33
root = tempfile.mkdtemp()
src = os.path.join(root, "frames")
os.makedirs(src)
g = torch.Generator().manual_seed(0)
lo = torch.rand(1, 3, 70, 122, generator=g)
big = torch.nn.functional.interpolate(lo, size=(1080 + 64, 1920 + 64), mode="bicubic", align_corners=False).clamp(0, 1)
rng = np.random.default_rng(0)


def make(i):
    dy, dx = i % 32, (2 * i) % 32                                  # a small drift inside the 32-pixel margin
    a = (big[0, :, 32 + dy:32 + dy + 1080, 32 + dx:32 + dx + 1920].permute(1, 2, 0).numpy() * 255).astype(np.uint8)
    a = a + np.random.default_rng(i).integers(0, 3, a.shape, dtype=np.uint8)
    Image.fromarray(a, "RGB").save(os.path.join(src, f"{i + 1:06d}.png"))


with ThreadPoolExecutor(os.cpu_count()) as ex:
    list(ex.map(make, range(n)))
torch.manual_seed(0)
net = Net().cuda().eval()
convert_folder(src, os.path.join(root, "warm"), 1, net=net, chunk_pairs=4)      # engine + weights + first-call costs
for writer, lvl in (("pil", None), ("pil", 1), ("fast", 1), ("fast", 0)):
    dst = os.path.join(root, f"out_{writer}_{lvl}")
    torch.cuda.synchronize()
    t0 = time.time()
    written = convert_folder(src, dst, 1, net=net, batch=4, chunk_pairs=16, png_compress_level=lvl, png_writer=writer)
    dt = time.time() - t0
    mb = sum(os.path.getsize(p) for p in written[1::2]) / max(1, len(written[1::2])) / 1e6
    print(f"convert_folder 1080p, {n} frames -> {len(written)} files, writer {writer}, png level {lvl}: {dt:.2f} s = {len(written) / dt:.1f} output frames/s "
          f"({(n - 1) / dt:.1f} interpolated/s) on {os.cpu_count()} host threads, {mb:.2f} MB per interpolated frame", flush=True)
shutil.rmtree(root)
