"""CPU tests of the host-side logic: C-ABI symbols, the parameter tree, sharding (incl. a
world_size-2 gloo run), the drop-in module and the no-fallback rule."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    from rrin_b200 import _lib
    l = _lib.lib()
    syms = _lib.header_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(l, s)]
    assert not missing, missing
    assert set(_lib._SIGNATURES) == set(syms), set(_lib._SIGNATURES) ^ set(syms)
    assert l.rrin_num_convs() == 81


def test_conv_table_matches_reference_parameter_tree():
    """rrin_conv_info enumerates exactly the 81 convs of the module tree, with the true shapes."""
    from rrin_b200 import Net, engine
    net = Net()
    sd = net.state_dict()
    assert len(sd) == 162
    assert sum(v.numel() for v in sd.values()) == 19_194_445
    keys = list(sd)
    order = list(dict.fromkeys(k.split(".")[0] for k in keys))
    assert order == ["Mask", "Flow", "refine_flow", "final"]             # registration order, model.py:27-30
    table = engine.conv_table()
    assert len(table) == 81 and len({t[0] for t in table}) == 81
    for key, cin, cout, level, src, act in table:
        assert tuple(sd[key + ".weight"].shape) == (cout, cin, 3, 3), key
        assert tuple(sd[key + ".bias"].shape) == (cout,), key
        assert act == (0 if key.endswith("up.1") or key.endswith("last") else 1), key   # unet.py:47,60,63 only


def test_bench_flop_count_follows_from_the_conv_table():
    """bench.py's algorithmic FLOP per padded pixel (the roofline numerator) is the sum over the 81 convs of
    2 * 9 * Cin * Cout at the conv's resolution (level L convolves H * W / 4^L pixels) with the TRUE channel counts -- zero
    padding of packed tensors and structural zeros of the space-to-depth weights are not credited."""
    import importlib.util
    from fractions import Fraction
    from rrin_b200 import engine
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    per_px = sum(Fraction(2 * 9 * cin * cout, 4 ** level) for _key, cin, cout, level, _src, _act in engine.conv_table())
    assert per_px == bench.FLOP_PER_PX == 1_736_064
    assert abs(float(per_px) * bench.H * bench.W / 1e12 - 3.62656825344) < 1e-9          # TFLOP per 1080p frame, the bench line's config.tflop_per_frame


def test_net_accepts_optional_level_and_rejects_cpu_inputs():
    from rrin_b200 import Net
    Net(3)
    net = Net()
    x = torch.rand(1, 3, 32, 32)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net(x, x, 0.5)
    from rrin_b200 import warp
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        warp(x, torch.zeros(1, 2, 32, 32))


def test_engine_create_rejects_bad_shapes_without_gpu():
    import ctypes as C
    from rrin_b200._lib import lib
    l = lib()
    h = C.c_void_p()
    assert l.rrin_engine_create(1, 1, 72, 80, C.byref(h)) != 0         # 72 % 16 != 0 (reference: torch.cat error)
    assert b"multiples of 16" in l.rrin_last_error()
    assert l.rrin_engine_create(2, 3, 64, 64, C.byref(h)) != 0         # n_pairs must be 1 or n_samples
    assert l.rrin_engine_create(1, 7, 1088, 1920, C.byref(h)) == 0
    assert l.rrin_engine_num_launches(h) > 81
    assert l.rrin_engine_workspace_bytes(h) < 8 << 30                  # 7 timesteps of 1080p stay far below 180 GB
    l.rrin_engine_destroy(h)


def test_c99_consumer_links_and_runs(tmp_path):
    """include/rrin_b200.h is plain C (no C++ / torch types): a C99 program compiled with -pedantic links the library and
    drives the host-side entry points (conv table, engine planning, error codes) -- the maintainer-side binding of
    INTEGRATION.md section 1 in its smallest form."""
    import shutil
    from rrin_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("gcc not on PATH")
    _lib.lib()                                                          # builds the library if it is missing
    libdir = os.path.join(ROOT, "rrin_b200")
    exe = str(tmp_path / "consumer")
    cc = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                         os.path.join(ROOT, "tests", "cabi", "consumer.c"), "-o", exe, "-L", libdir, "-l:librrin_b200.so",
                         "-Wl,-rpath," + libdir], capture_output=True, text=True, timeout=120)
    assert cc.returncode == 0, cc.stderr
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stdout + run.stderr
    assert "convs 81 launches 90" in run.stdout, run.stdout


def test_conv_config_table_is_consistent():
    """Tile configurations behind rrin_conv3x3 (host-side queries only): ids, K-stage geometry, packed sizes."""
    import ctypes as C
    from rrin_b200._lib import lib
    l = lib()
    v = [C.c_int() for _ in range(4)]
    valid = {}
    for cfg in range(-1, 64):
        if l.rrin_conv_config_info(cfg, *map(C.byref, v)) == 0:
            valid[cfg] = tuple(x.value for x in v)                       # kcs, kb, nt, msub
    assert set(valid) == set(range(0, 9)) | set(range(10, 48)), sorted(valid)
    for cfg, (kcs, kb, nt, msub) in valid.items():
        assert kcs in (32, 64, 128) and kb in (16, 32, 64) and kb <= kcs and nt in (16, 64, 128) and 1 <= msub <= 4, (cfg, kcs, kb, nt, msub)
        if cfg >= 10:
            assert msub * nt <= 512, "a tile's accumulators fit TMEM"
    assert valid[21][:3] == (32, 32, 64) and valid[22][:3] == (64, 64, 64)
    # packed 9-tap weights: [n-tile][stage][9][KB x NT] bf16, independent of the CTA-pair split
    assert l.rrin_conv_packed_weight_bytes(16, 256, 4, 0) == 2 * 4 * 9 * 64 * 128 * 2
    assert l.rrin_conv_packed_weight_bytes(19, 256, 4, 0) == l.rrin_conv_packed_weight_bytes(16, 256, 4, 0)
    assert l.rrin_conv_packed_weight_bytes(21, 64, 1, 0) == 9 * 32 * 64 * 2
    assert l.rrin_conv_packed_weight_bytes(22, 64, 2, 0) == l.rrin_conv_packed_weight_bytes(15, 64, 2, 0)
    assert l.rrin_conv_packed_bias_count(21, 64) == 64


def test_dropin_module_is_importable_as_model():
    code = ("import sys; sys.path[:0] = [%r, %r]; from model import Net, warp; import rrin_b200; "
            "assert Net is rrin_b200.Net and warp is rrin_b200.warp; n = Net(); print(len(n.state_dict()))" % (os.path.join(ROOT, "dropin"), ROOT))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == "162"


def test_product_path_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "rrin_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dirpath, f)) as fh:
                    txt = fh.read()
                assert not re.search(r"^\s*(from|import)\s+oracle", txt, re.M), f
    for f in os.listdir(os.path.join(ROOT, "dropin")):
        if f.endswith(".py"):
            with open(os.path.join(ROOT, "dropin", f)) as fh:
                assert "oracle" not in fh.read()


# ----------------------------------------------------------------------------- sharding
def test_pair_ranges_partition_the_clip():
    from rrin_b200 import sharding as S
    for n_frames in (0, 1, 2, 3, 7, 240, 241):
        for world in (1, 2, 3, 4, 8):
            ranges = [S.pair_range(n_frames, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == max(n_frames - 1, 0)
            for (a, b), (c, d) in zip(ranges, ranges[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    assert [S.pair_range(240, r, 8) for r in range(8)][0] == (0, 30)      # 239 pairs: 7 shards of 30, one of 29
    assert S.frame_range(240, 7, 8) == (210, 240)
    with pytest.raises(ValueError):
        S.pair_range(10, 2, 2)


def test_plan_covers_every_output_frame_once():
    from rrin_b200 import sharding as S
    for world in (1, 2, 8):
        for sf in (1, 3, 7):
            p = S.plan(25, world, sf)
            idx = sorted(o for *_, o in p)
            want = sorted(i for i in range(24 * (sf + 1) + 1) if i % (sf + 1))
            assert idx == want
            assert all(abs(t - k / (sf + 1)) < 1e-15 for _, _, k, t, _ in p)


def _gloo_worker(rank, world, port, n_frames, q):
    import torch.distributed as dist
    from rrin_b200 import sharding as S
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = S.pair_range(n_frames, rank, world)
    # each rank "interpolates" its pairs: a stand-in result that depends only on the pair index
    mine = torch.zeros(n_frames - 1, dtype=torch.int64)
    mine[lo:hi] = torch.arange(lo, hi) * 2 + 1
    ms = torch.tensor([float(hi - lo)], dtype=torch.float64)            # bench.py's max-over-ranks reduction
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(mine, op=dist.ReduceOp.SUM)                         # test-only gather (no collective on the product path)
    if rank == 0:
        q.put((mine.tolist(), ms.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_shards_reassemble():
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_frames = 12
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, ms = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [2 * i + 1 for i in range(n_frames - 1)]
    assert ms == 6.0                                                    # 11 pairs over 2 ranks: 6 and 5


# ----------------------------------------------------------------------------- reference file / pad conventions
def test_pad_amounts_follow_the_reference_quirk():
    from rrin_b200 import io as rio
    assert rio.pad_amounts(1080, 1920) == (8, 0) and rio.padded_shape(1080, 1920) == (1088, 1920)
    assert rio.pad_amounts(720, 1280) == (0, 0)
    assert rio.pad_amounts(360, 640) == (8, 0) and rio.padded_shape(360, 640) == (368, 640)       # README.md:19 training size
    # a width that is not a multiple of 16 pads the BOTTOM (torchvision Pad order) and the model cannot run: 100x50 -> 100x76
    assert rio.pad_amounts(50, 100) == (14, 12)
    with pytest.raises(RuntimeError, match="Sizes of tensors must match"):
        rio.padded_shape(50, 100)
    assert rio.crop_rows(1088, 1080) == 8


def test_pad_matches_torchvision_pad_semantics():
    """dataloader.py:108: transforms.Pad((0, top_pad, 0, right_pad), 'edge') == F.pad(replicate) with (left, right, top, bottom)."""
    import torch.nn.functional as F
    from rrin_b200 import io as rio
    x = torch.arange(3 * 40 * 32, dtype=torch.float32).reshape(1, 3, 40, 32)
    top, bottom = rio.pad_amounts(40, 32)
    assert (top, bottom) == (8, 0)
    y = F.pad(x, (0, 0, top, bottom), mode="replicate")
    assert y.shape[2:] == rio.padded_shape(40, 32) and torch.equal(y[:, :, :top], x[:, :, :1].expand(-1, -1, top, -1))


def test_output_names_and_resume():
    from rrin_b200 import io as rio
    names = rio.output_names(3, 2)
    assert [n for n, _ in names] == [f"{i:09d}.png" for i in range(1, 8)]
    assert [s for _, s in names] == [None, (0, 1), (0, 2), None, (1, 1), (1, 2), None]
    # convert.py:46-53: the reference's 1-based resume_index is 1 unless more than 5 outputs exist
    assert rio.resume_index(0, 2) == 1 and rio.resume_index(4, 2) == 1 and rio.resume_index(5, 2) == 1
    assert rio.resume_index(7, 2) == 2 and rio.resume_index(10, 2) == 3 and rio.resume_index(6, 1) == 2
    assert rio.resume_first_pair(0, 2) == 0 and rio.resume_first_pair(7, 2) == 1          # ConvertSampler(dataset, resume_index - 1)
    # convert.py:118: first file number of the pair the conversion starts with
    assert rio.first_output_number(1, 2) == 1 and rio.first_output_number(2, 2) == 4 and rio.first_output_number(3, 1) == 5
    for sf in (1, 2, 7):
        for ridx in (1, 2, 5):
            assert rio.first_output_number(ridx, sf) == (ridx - 1) * (sf + 1) + 1


def test_checkpoint_lookup_and_load(tmp_path):
    from rrin_b200 import Net, io as rio
    (tmp_path / "other0001.pth").write_bytes(b"x")
    sd = Net().state_dict()
    for ep in (3, 12):
        torch.save({"model": sd, "optim": {}, "epoch": ep}, tmp_path / f"MyModel{ep:04d}.pth")
    got = rio.find_checkpoint(str(tmp_path), "mymodel")
    listing = [n for n in reversed(os.listdir(tmp_path)) if n.lower().startswith("mymodel")]
    assert os.path.basename(got) == listing[0]
    loaded = rio.load_checkpoint(got)
    Net().load_state_dict(loaded, strict=True)
    with pytest.raises(FileNotFoundError):
        rio.find_checkpoint(str(tmp_path), "absent")


def test_fastpng_round_trips_through_pil():
    """Every (level, row filter) of the lean PNG writer decodes with PIL to the bytes that went in; bad inputs are refused."""
    import io
    import numpy as np
    from PIL import Image
    from rrin_b200 import fastpng
    rng = np.random.default_rng(4)
    ramp = (np.add.outer(np.arange(37), np.arange(53))[:, :, None] * np.array([3, 5, 7])).astype(np.uint8)      # wraps past 255
    for img in (rng.integers(0, 256, (37, 53, 3), dtype=np.uint8), ramp, np.zeros((1, 1, 3), np.uint8),
                rng.integers(0, 256, (64, 48, 3), dtype=np.uint8)[::2, ::3]):                                   # a non-contiguous view
        for level in (0, 1, 6, 9):
            for flt in ("none", "sub", "up"):
                data = fastpng.encode_png(img, level, flt)
                back = Image.open(io.BytesIO(data))
                assert back.mode == "RGB" and back.size == (img.shape[1], img.shape[0])
                assert np.array_equal(np.asarray(back), img), (img.shape, level, flt)
    smooth = np.repeat(np.repeat(rng.integers(0, 256, (8, 8, 3), dtype=np.uint8), 16, 0), 16, 1)
    assert len(fastpng.encode_png(smooth, 1, "sub")) < len(fastpng.encode_png(smooth, 0)) // 10
    for bad in (np.zeros((4, 4), np.uint8), np.zeros((4, 4, 4), np.uint8), np.zeros((4, 4, 3), np.float32), np.zeros((0, 4, 3), np.uint8)):
        with pytest.raises(ValueError):
            fastpng.encode_png(bad)
    with pytest.raises(ValueError):
        fastpng.encode_png(smooth, 10)
    with pytest.raises(ValueError):
        fastpng.encode_png(smooth, 1, "paeth")


def test_convert_folder_host_logic_with_a_stand_in_pipeline(tmp_path, monkeypatch):
    """convert_folder's host side on CPU (decode-ahead pool, bounded writer backlog, numbering, copies, resume, sharding,
    error propagation) with ClipInterpolator replaced by a stand-in that averages neighbouring frames; the real pipeline is
    compared file by file in tests/test_gpu_convert.py."""
    import numpy as np
    from PIL import Image
    import rrin_b200.pipeline as pl
    from rrin_b200 import convert_folder

    class StandIn:
        def __init__(self, net, h, w, batch=2, sf=1, device=None, uint8=False, channels=3):
            assert uint8 and channels == 3
            self.sf = sf

        def run(self, clip, out_host=None):
            f = clip.numpy().astype(np.float32)
            outs = [((1 - k / (self.sf + 1)) * f[i] + k / (self.sf + 1) * f[i + 1]).astype(np.uint8)
                    for i in range(len(f) - 1) for k in range(1, self.sf + 1)]
            assert out_host is not None and tuple(out_host.shape) == (len(outs), *f.shape[1:3], 3)
            out_host.copy_(torch.from_numpy(np.stack(outs)))
            return out_host

    monkeypatch.setattr(pl, "ClipInterpolator", StandIn)
    net = torch.nn.Linear(1, 1)                                           # only .parameters() is touched when `net` is given
    src = tmp_path / "frames"
    src.mkdir()
    rng = np.random.default_rng(1)
    frames = [rng.integers(0, 256, (40, 32, 3), dtype=np.uint8) for _ in range(11)]
    for i, a in enumerate(frames):
        Image.fromarray(a, "RGB").save(src / f"f{i:03d}.png")
    dev = torch.device("cpu")

    def expect(i, k, sf):
        return ((1 - k / (sf + 1)) * frames[i].astype(np.float32) + k / (sf + 1) * frames[i + 1]).astype(np.uint8)

    for sf, chunk in ((1, 64), (2, 3), (3, 1)):                           # one chunk; several chunks with decode-ahead; chunk = pair
        dst = tmp_path / f"out_sf{sf}"
        written = convert_folder(str(src), str(dst), sf, net=net, device=dev, chunk_pairs=chunk, io_workers=3)
        assert [os.path.basename(p) for p in written] == [f"{i:09d}.png" for i in range(1, 10 * (sf + 1) + 2)]
        assert sorted(os.listdir(dst)) == [os.path.basename(p) for p in written]
        for i in range(11):
            name = f"{i * (sf + 1) + 1:09d}.png"
            assert (dst / name).read_bytes() == (src / f"f{i:03d}.png").read_bytes()           # originals are file copies
            if i < 10:
                for k in range(1, sf + 1):
                    got = np.asarray(Image.open(dst / f"{i * (sf + 1) + 1 + k:09d}.png"))
                    assert np.array_equal(got, expect(i, k, sf)), (sf, i, k)
    fast = tmp_path / "out_fast"
    convert_folder(str(src), str(fast), 2, net=net, device=dev, png_compress_level=1)                  # same pixels, other zlib level
    for n in os.listdir(fast):
        assert np.array_equal(np.asarray(Image.open(fast / n)), np.asarray(Image.open(tmp_path / "out_sf2" / n))), n
    lean = tmp_path / "out_lean"
    convert_folder(str(src), str(lean), 2, net=net, device=dev, png_writer="fast")                     # rrin_b200.fastpng
    for n in os.listdir(lean):
        assert np.array_equal(np.asarray(Image.open(lean / n)), np.asarray(Image.open(tmp_path / "out_sf2" / n))), n
    # two ranks write disjoint files whose union is the single-process output
    dst2 = tmp_path / "out_ranks"
    w0 = convert_folder(str(src), str(dst2), 2, net=net, device=dev, rank=0, world=2, chunk_pairs=2)
    w1 = convert_folder(str(src), str(dst2), 2, net=net, device=dev, rank=1, world=2, chunk_pairs=2)
    assert {os.path.basename(p) for p in set(w0) & set(w1)} == {"000000016.png"}   # the original between the shards: the same copy twice
    ref = tmp_path / "out_sf2"
    assert sorted(os.listdir(dst2)) == sorted(os.listdir(ref))
    for n in os.listdir(ref):
        assert (dst2 / n).read_bytes() == (ref / n).read_bytes(), n
    # a non-empty destination is refused like convert.py:57-59; resume continues where the count of files says
    with pytest.raises(RuntimeError, match="already in use"):
        convert_folder(str(src), str(ref), 2, net=net, device=dev)
    part = tmp_path / "out_part"
    part.mkdir()
    for n in sorted(os.listdir(ref))[:10]:                                # 10 files = 3 complete pairs + the 4th pair's first frame
        (part / n).write_bytes((ref / n).read_bytes())
    convert_folder(str(src), str(part), 2, net=net, device=dev, resume=True, chunk_pairs=4)
    assert sorted(os.listdir(part)) == sorted(os.listdir(ref))
    for n in os.listdir(ref):
        assert np.array_equal(np.asarray(Image.open(part / n)), np.asarray(Image.open(ref / n))), n
    # a frame of another size stops the conversion with the file's name (decoded ahead on a worker thread)
    Image.fromarray(rng.integers(0, 256, (24, 32, 3), dtype=np.uint8), "RGB").save(src / "f005.png")
    with pytest.raises(RuntimeError, match="f005.png"):
        convert_folder(str(src), str(tmp_path / "out_bad"), 1, net=net, device=dev, chunk_pairs=2)


# ----------------------------------------------------------------------------- the reference's own caller on the drop-in
_CONVERT_SCRIPT = r"""
import os, sys, argparse
ROOT, REF, WORK = sys.argv[1:4]
sys.path[:0] = [os.path.join(ROOT, "dropin"), ROOT, REF]      # `from model import Net` (convert.py:11) -> dropin/model.py
import numpy as np, torch
from PIL import Image
# no GPU in the build container: .cuda() becomes the identity (the same shim the oracle harness uses for model.py:11-12) and
# the drop-in's forward -- the only part that needs the device -- is replaced by a stand-in.  Everything else is the
# UNEDITED /root/reference/convert.py + dataloader.py + utils.py driving rrin_b200.Net through its public surface.
torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self
import rrin_b200
calls = []
def fake_forward(self, input0, input1, t=0.5):
    assert input0.shape == input1.shape and input0.dim() == 4 and input0.shape[:2] == (1, 3) and input0.dtype == torch.float32
    assert input0.shape[2] % 16 == 0 and input0.shape[3] % 16 == 0, input0.shape
    assert not torch.is_grad_enabled() and not self.training          # convert.py:111,117
    calls.append(float(t))
    return ((1 - t) * input0 + t * input1).clamp(0, 1)
rrin_b200.Net.forward = fake_forward
import convert                                                          # the reference's module, unedited
assert convert.Net is rrin_b200.Net, convert.Net
os.chdir(WORK)
os.makedirs("models"); os.makedirs("frames")
torch.manual_seed(0)
sd = rrin_b200.Net().state_dict()
torch.save({"model": sd, "optim": {}, "epoch": 7}, os.path.join("models", "Demo0007.pth"))      # train.py:158-161
rng = np.random.default_rng(0)
H0, W0 = 40, 32                                                         # 40 rows -> 8 rows of edge pad on top
for i in range(3):
    Image.fromarray(rng.integers(0, 256, (H0, W0, 3), dtype=np.uint8), "RGB").save(os.path.join("frames", f"{i + 1:09d}.png"))
args = argparse.Namespace(input_video=None, output_video=None, image_folder="frames", resume=False, sf=2, fps="30",
                          no_cuda=False, model_name="demo", rm=False, mode="convert")
try:
    convert.convert(args)                                               # convert.py:22-42
except SystemExit as e:                                                 # _create_video: ffmpeg is absent (convert.py:152-157)
    assert e.code not in (0, None)
dest = "temp\\output"
got = sorted(os.listdir(dest))
assert got == [f"{i:09d}.png" for i in range(1, 8)], got
assert calls == [1 / 3, 2 / 3, 1 / 3, 2 / 3], calls                     # convert.py:127-130
order = os.listdir("frames")                                            # the reference's frame order (dataloader.py:20)
fr = [np.asarray(Image.open(os.path.join("frames", n))) for n in order]
for p in range(2):
    assert np.array_equal(np.asarray(Image.open(os.path.join(dest, f"{p * 3 + 1:09d}.png"))), fr[p])          # copied originals
    a = torch.from_numpy(fr[p]).permute(2, 0, 1).float().div(255)
    b = torch.from_numpy(fr[p + 1]).permute(2, 0, 1).float().div(255)
    for i in (1, 2):
        t = i / 3
        want = ((1 - t) * a + t * b).clamp(0, 1).mul(255).byte().permute(1, 2, 0).numpy()     # pad rows are cropped again (utils.py:56-57)
        out = np.asarray(Image.open(os.path.join(dest, f"{p * 3 + 1 + i:09d}.png")))
        assert out.shape == (H0, W0, 3) and np.array_equal(out, want), (p, i)
print("ok")
"""


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference sources exist only in the build container")
def test_unedited_reference_convert_drives_the_dropin(tmp_path):
    """/root/reference/convert.py (with its dataloader.py / utils.py), unedited, against `from model import Net` resolved to
    the drop-in: constructor, strict checkpoint load found by name prefix, .cuda().eval(), the per-pair x per-timestep call
    signature, the .cpu() hand-off to the Writer thread and the 9-digit PNG sequence.  The device forward itself is a
    stand-in here (no GPU in this container); tests/test_gpu_convert.py runs the real one."""
    out = subprocess.run([sys.executable, "-c", _CONVERT_SCRIPT, ROOT, "/root/reference", str(tmp_path)],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.strip().endswith("ok")
