"""GPU tests of the composed convert path: PNG folder in, 9-digit PNG sequence out, checkpoint
found by name prefix -- ``rrin_b200.convert_folder`` against per-pair ``Net.forward`` + the reference's host-side
``to_pil_image`` / crop (convert.py:94-144, utils.py:51-58, dataloader.py:93-118); and, where the reference sources exist,
the unedited ``/root/reference/convert.py`` itself on the drop-in."""
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from oracle import rrin_oracle as O
    from rrin_b200 import Net, convert_folder, io as rio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_frames(folder, n, h0, w0, channels=3, seed=0):
    from PIL import Image
    os.makedirs(folder, exist_ok=True)
    frames = []
    for i in range(n):
        a, _ = O.seeded_frames(1, (h0 + 15) // 16 * 16, (w0 + 15) // 16 * 16, seed=seed + i, smooth=True)
        rgb = (a[0, :, :h0, :w0] * 255).byte().permute(1, 2, 0).numpy()
        if channels == 4:
            rgb = np.concatenate([rgb, np.full((h0, w0, 1), 200, np.uint8)], 2)
        Image.fromarray(rgb, "RGBA" if channels == 4 else "RGB").save(os.path.join(folder, f"{i + 1:09d}.png"))
        frames.append(rgb[..., :3])
    return frames


def _expected(net, f0, f1, t, h0, w0):
    """What the reference writes for one (pair, t): Pad(edge) + ToTensor (dataloader.py:93-118), forward, mul(255).byte() and
    the crop of the pad rows (utils.py:51-58)."""
    import torch.nn.functional as F
    top, bottom = rio.pad_amounts(h0, w0)
    prep = lambda f: F.pad(torch.from_numpy(f).permute(2, 0, 1).float().div(255).unsqueeze(0), (0, 0, top, bottom), mode="replicate")
    y = net(prep(f0).cuda(), prep(f1).cuda(), t=t).cpu()[0]
    return y.mul(255).byte()[:, top + bottom:, :].permute(1, 2, 0).numpy()


@pytest.mark.parametrize("channels", [3, 4], ids=["rgb", "rgba"])
def test_convert_folder_matches_reference_semantics(tmp_path, channels):
    from PIL import Image
    h0, w0, sf = 72, 96, 2                            # 72 rows -> padded to 80 (8 rows on top)
    src, dst, models = tmp_path / "frames", tmp_path / "out", tmp_path / "models"
    frames = _write_frames(str(src), 4, h0, w0, channels)
    sd = O.seeded_state_dict(stress_flow=50.0)
    os.makedirs(models)
    torch.save({"model": sd, "optim": {}, "epoch": 3}, models / "Demo0003.pth")            # train.py:158-161
    torch.save({"model": O.seeded_state_dict(), "optim": {}, "epoch": 1}, models / "Other0001.pth")
    written = convert_folder(str(src), str(dst), sf, model_name="demo", models_dir=str(models), batch=2)
    names = [f"{i:09d}.png" for i in range(1, 3 * (sf + 1) + 2)]
    assert [os.path.basename(p) for p in written] == names and sorted(os.listdir(dst)) == names
    net = Net()
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    for p in range(3):
        with open(src / f"{p + 1:09d}.png", "rb") as a, open(dst / f"{p * (sf + 1) + 1:09d}.png", "rb") as b:
            assert a.read() == b.read(), "originals are file copies (convert.py:121-123,136-139)"
        for i in range(1, sf + 1):
            got = np.asarray(Image.open(dst / f"{p * (sf + 1) + 1 + i:09d}.png"))
            want = _expected(net, frames[p], frames[p + 1], i / (sf + 1), h0, w0)
            assert got.shape == (h0, w0, 3) and np.array_equal(got, want), (p, i)
    with pytest.raises(RuntimeError, match="already in use"):
        convert_folder(str(src), str(dst), sf, net=net)


def test_convert_folder_resume_like_reference(tmp_path):
    h0, w0, sf = 48, 64, 1
    src, dst, dst2 = tmp_path / "frames", tmp_path / "out", tmp_path / "out2"
    _write_frames(str(src), 8, h0, w0)
    net = Net()
    net.load_state_dict(O.seeded_state_dict(stress_flow=50.0), strict=True)
    net = net.cuda().eval()
    convert_folder(str(src), str(dst), sf, net=net, batch=3, chunk_pairs=4)
    full = {n: open(dst / n, "rb").read() for n in sorted(os.listdir(dst))}
    assert len(full) == 15
    shutil.copytree(dst, dst2)
    for n in sorted(os.listdir(dst2))[9:]:            # an interrupted run left 9 files
        os.remove(dst2 / n)
    written = convert_folder(str(src), str(dst2), sf, net=net, batch=3, resume=True)
    # convert.py:50: resume_index = (9 - 1) // 2 = 4 -> pair 3, first number 7: files 8.. are (re)written
    assert os.path.basename(written[0]) == "000000008.png"
    assert {n: open(dst2 / n, "rb").read() for n in sorted(os.listdir(dst2))} == full


def test_convert_folder_sharded_over_ranks_equals_single_process(tmp_path):
    """Multi-GPU conversion: every rank converts its contiguous pair shard on its own (here: one after the other on one GPU);
    the union of what the ranks write is the single-process output, file for file."""
    h0, w0, sf = 48, 64, 2
    src, ref_dst, dst = tmp_path / "frames", tmp_path / "ref", tmp_path / "sharded"
    _write_frames(str(src), 8, h0, w0)
    net = Net()
    net.load_state_dict(O.seeded_state_dict(stress_flow=50.0), strict=True)
    net = net.cuda().eval()
    convert_folder(str(src), str(ref_dst), sf, net=net, batch=2)
    want = {n: open(ref_dst / n, "rb").read() for n in sorted(os.listdir(ref_dst))}
    written = []
    for rank in range(3):
        written += convert_folder(str(src), str(dst), sf, net=net, batch=2, rank=rank, world=3)
    got = {n: open(dst / n, "rb").read() for n in sorted(os.listdir(dst))}
    assert got == want
    assert len(written) == len(want) + 2          # the two shard-boundary originals are written by both neighbours (same bytes)
    assert convert_folder(str(src), str(tmp_path / "empty"), sf, net=net, rank=9, world=10) == []


_REAL_CONVERT = r"""
import os, sys, argparse
ROOT, REF, WORK = sys.argv[1:4]
sys.path[:0] = [os.path.join(ROOT, "dropin"), ROOT, REF]
import torch
import convert                                   # /root/reference/convert.py, unedited; `from model import Net` -> drop-in
import rrin_b200
assert convert.Net is rrin_b200.Net
os.chdir(WORK)
args = argparse.Namespace(input_video=None, output_video=None, image_folder="frames", resume=False, sf=2, fps="30",
                          no_cuda=False, model_name="demo", rm=False, mode="convert")
try:
    convert.convert(args)
except SystemExit as e:                          # ffmpeg is absent: _create_video exits after all frames are written
    assert e.code not in (0, None)
print("ok")
"""


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference sources do not travel to the GPU box")
def test_unedited_reference_convert_on_gpu(tmp_path):
    """python . convert --image_folder, i.e. /root/reference/convert.py unedited, with PYTHONPATH=dropin:. -- runs wherever
    both a GPU and the reference sources exist (the CPU twin with a stand-in forward is tests/test_host_logic.py)."""
    from PIL import Image
    h0, w0, sf = 72, 96, 2
    frames = _write_frames(str(tmp_path / "frames"), 3, h0, w0)
    sd = O.seeded_state_dict(stress_flow=50.0)
    os.makedirs(tmp_path / "models")
    torch.save({"model": sd, "optim": {}, "epoch": 3}, tmp_path / "models" / "Demo0003.pth")
    out = subprocess.run([sys.executable, "-c", _REAL_CONVERT, ROOT, "/root/reference", str(tmp_path)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-4000:]
    dest = tmp_path / "temp\\output"
    order = os.listdir(tmp_path / "frames")
    net = Net()
    net.load_state_dict(sd, strict=True)
    net = net.cuda().eval()
    by_name = {f"{i + 1:09d}.png": f for i, f in enumerate(frames)}
    fr = [by_name[n] for n in order]
    for p in range(2):
        for i in range(1, sf + 1):
            got = np.asarray(Image.open(dest / f"{p * (sf + 1) + 1 + i:09d}.png"))
            assert np.array_equal(got, _expected(net, fr[p], fr[p + 1], i / (sf + 1), h0, w0)), (p, i)
