"""Memory-safety properties of the whole forward, checked through the C-ABI without a sanitizer (compute-sanitizer is
closed on this GPU pool):

* **workspace poison** -- the result does not depend on what the engine workspace held before the call: no kernel reads
  scratch memory that the same forward has not written (the launch plan relies on TMA's out-of-bounds zero fill, never on
  zero-initialised buffers);
* **guard bands** -- every buffer handed to ``rrin_engine_forward`` / ``rrin_engine_forward_graph`` (frames, result, packed
  weights, workspace) is a slice of a larger allocation; the bytes on either side are intact after the call, and filling them
  with NaN patterns instead of zeros does not change a bit of the result (an out-of-range read would pick them up).

Shapes cover frames too small to fold the upsample (64 x 96), the folded upsample convs with their border rings in both
orientations (128 x 208, 208 x 128: plain and transposed level >= 2 launches), batches and the multi-timestep engine."""
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from rrin_b200 import Net

GUARD = 1 << 20          # bytes on either side of a guarded buffer (a multiple of the 256-byte workspace alignment)
CASES = [(1, 64, 96, None), (2, 128, 208, None), (1, 208, 128, None), (1, 128, 208, [0.25, 0.5, 0.75])]
IDS = ["64x96", "2x128x208", "208x128", "128x208_multi_t"]
BIG = [(4, 1088, 1920, None)]         # bench.py's step: several tiles per CTA and band at levels 0-1, 3.8 GB of workspace
BIG_IDS = ["4x1088x1920"]


def _net(precision="bf16"):
    torch.manual_seed(0)
    net = Net().cuda().eval()
    net.precision = precision
    return net


def _frames(n, h, w, seed=5):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n, 3, h, w, generator=g).cuda(), torch.rand(n, 3, h, w, generator=g).cuda()


def _engine_and_coef(net, a, ts):
    n, _, h, w = a.shape
    dev = a.device
    if ts is None:
        eng = net._engine(dev, n, h, w)
        return eng, eng._coef(0.5), n
    eng = net._engine(dev, len(ts), h, w, n_pairs=1)
    return eng, eng._coef(list(ts)), len(ts)


class Guarded:
    """A byte buffer with GUARD bytes of `fill` on either side of the payload."""

    def __init__(self, nbytes, fill):
        self.fill, self.nbytes = fill, nbytes
        self.raw = torch.full((nbytes + 2 * GUARD,), fill, dtype=torch.uint8, device="cuda")
        self.payload = self.raw[GUARD:GUARD + nbytes]
        assert self.payload.data_ptr() % 256 == 0

    def like(self, t):                      # payload as a copy of tensor t
        self.payload.copy_(t.contiguous().view(torch.uint8).reshape(-1))
        return self.payload.view(t.dtype).reshape(t.shape)

    def intact(self):
        return bool((self.raw[:GUARD] == self.fill).all()) and bool((self.raw[GUARD + self.nbytes:] == self.fill).all())


@pytest.mark.parametrize("n,h,w,ts", CASES + BIG, ids=IDS + BIG_IDS)
@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_workspace_contents_never_reach_the_result(n, h, w, ts, precision):
    net = _net(precision)
    a, b = _frames(n, h, w)
    eng, coef, n_out = _engine_and_coef(net, a, ts)
    pk = net._weights(a.device)
    outs = []
    for fill in (0x00, 0xFF, 0x3C, 0x7F):                 # zeros; NaN in bf16 / fp16 / fp32; small finite values; NaN / huge
        eng.workspace.fill_(fill)
        out = torch.empty(n_out, 3, h, w, device="cuda")
        outs.append(eng.run(pk, a, b, coef, out=out).clone())
    torch.cuda.synchronize()
    assert torch.isfinite(outs[0]).all()
    for o in outs[1:]:
        assert torch.equal(o, outs[0]), "a kernel read workspace bytes that this forward had not written"


@pytest.mark.parametrize("n,h,w,ts", CASES + BIG, ids=IDS + BIG_IDS)
def test_guard_bands_around_every_buffer_stay_intact(n, h, w, ts):
    net = _net()
    a, b = _frames(n, h, w)
    eng, coef, n_out = _engine_and_coef(net, a, ts)
    pk = net._weights(a.device)
    ref = eng.run(pk, a, b, coef, out=torch.empty(n_out, 3, h, w, device="cuda")).clone()
    ws_saved, blob_saved = eng.workspace, pk.blob
    try:
        for fill in (0x00, 0xFF):
            g_a, g_b = Guarded(a.numel() * 4, fill), Guarded(b.numel() * 4, fill)
            g_out, g_ws = Guarded(n_out * 3 * h * w * 4, fill), Guarded(eng.workspace_bytes, fill)
            g_blob, g_coef = Guarded(blob_saved.numel(), fill), Guarded(coef.numel() * 4, fill)
            ga, gb, gcoef = g_a.like(a), g_b.like(b), g_coef.like(coef)
            pk.blob = g_blob.like(blob_saved)
            eng.workspace = g_ws.payload
            gout = g_out.payload.view(torch.float32).reshape(n_out, 3, h, w)
            for _ in range(3):                               # direct launches, graph capture, graph replay
                gout.fill_(float("nan"))
                eng.run(pk, ga, gb, gcoef, out=gout)
            torch.cuda.synchronize()
            assert torch.equal(gout, ref), f"guard fill {fill:#x} changed the result: something read outside its buffer"
            for name, g in (("in0", g_a), ("in1", g_b), ("out", g_out), ("workspace", g_ws), ("weights", g_blob), ("coef", g_coef)):
                assert g.intact(), f"bytes next to `{name}` were overwritten (guard fill {fill:#x})"
            assert torch.equal(ga, a) and torch.equal(gb, b) and torch.equal(pk.blob, blob_saved), "an input buffer was written"
    finally:
        eng.workspace, pk.blob = ws_saved, blob_saved


def test_public_warp_and_frame_io_respect_their_buffers():
    """rrin_warp, rrin_frame_from_u8 and rrin_frame_to_u8 on guarded buffers (odd sizes, flows that leave the image)."""
    from rrin_b200 import io as rio, warp
    from rrin_b200._lib import check, lib
    g = torch.Generator().manual_seed(9)
    img, flow = torch.rand(2, 3, 37, 53, generator=g).cuda(), (torch.rand(2, 2, 37, 53, generator=g) * 40 - 20).cuda()
    ref = warp(img, flow)
    for fill in (0x00, 0xFF):
        g_img, g_flow, g_out = Guarded(img.numel() * 4, fill), Guarded(flow.numel() * 4, fill), Guarded(img.numel() * 4, fill)
        gi, gf = g_img.like(img), g_flow.like(flow)
        go = g_out.payload.view(torch.float32).reshape(img.shape)
        check(lib().rrin_warp(gi.data_ptr(), gf.data_ptr(), 2, 3, 37, 53, go.data_ptr(), torch.cuda.current_stream().cuda_stream), "rrin_warp")
        torch.cuda.synchronize()
        assert torch.equal(go, ref) and g_img.intact() and g_flow.intact() and g_out.intact()
    h0, w0 = 41, 48                                               # pads to 48 x 48 (dataloader.py:93-108)
    top, bottom = rio.pad_amounts(h0, w0)
    hp, wp = rio.padded_shape(h0, w0)
    src = torch.randint(0, 256, (h0, w0, 3), dtype=torch.uint8, generator=g).cuda()
    st = torch.cuda.current_stream().cuda_stream
    results = []
    for fill in (0x00, 0xFF):
        g_src, g_f, g_u8 = Guarded(src.numel(), fill), Guarded(3 * hp * wp * 4, fill), Guarded(src.numel(), fill)
        gs = g_src.like(src)
        f = g_f.payload.view(torch.float32).reshape(3, hp, wp)
        check(lib().rrin_frame_from_u8(gs.data_ptr(), h0, w0, 3, top, bottom, f.data_ptr(), st), "rrin_frame_from_u8")
        u8 = g_u8.payload.reshape(h0, w0, 3)
        check(lib().rrin_frame_to_u8(f.data_ptr(), hp, wp, h0, w0, u8.data_ptr(), st), "rrin_frame_to_u8")
        torch.cuda.synchronize()
        assert g_src.intact() and g_f.intact() and g_u8.intact()
        assert torch.isfinite(f).all()
        results.append((f.clone(), u8.clone()))
    assert torch.equal(results[0][0], results[1][0]) and torch.equal(results[0][1], results[1][1])
    assert torch.equal(results[0][1], src)                       # pad -> ToTensor -> to_pil -> crop is the identity on bytes


def test_batch_of_36_pairs_crosses_2_31_elements_per_tensor():
    """Index arithmetic past 32 bits: 36 pairs of 1088 x 1920 make every level-0 tensor 2.4e9 elements (4.8 GB) and the
    workspace 34 GB of the 180 GB.  The K-sum order of a pixel depends on its row band and n-tile only (conv3x3_v2.cuh,
    `rot_of`), so each sample of the large batch must equal the single-pair call bit for bit -- checked on the first, a middle
    and the last sample (the last one lives entirely above the 2^31-element mark)."""
    net = _net()
    n, h, w = 36, 1088, 1920
    g = torch.Generator(device="cuda").manual_seed(11)
    a = torch.rand(n, 3, h, w, device="cuda", generator=g)
    b = (a + 0.05 * torch.rand(n, 3, h, w, device="cuda", generator=g)).clamp_(0, 1)
    with torch.no_grad():
        y = net(a, b, t=0.5)
        torch.cuda.synchronize()
        assert torch.isfinite(y).all() and float(y.min()) >= 0.0 and float(y.max()) <= 1.0
        for k in (0, 17, 35):
            yk = net(a[k:k + 1], b[k:k + 1], t=0.5)
            assert torch.equal(y[k:k + 1], yk), f"sample {k} of the 36-pair batch differs from its single-pair call"
    net._engines.clear()                     # give the 34 GB workspace back before the next test


def test_8k_frame_corner_matches_its_crop():
    """One 4320 x 7680 pair: level-0 tensors of 2.1 GB (byte offsets past 2^31), 270 x 480 level-0 tiles.  The far corner of the
    result is compared with a forward over the bottom-right 1024 x 1024 crop, away from the crop's inner edges (the receptive
    field of the four U-Nets in sequence stays below 256 px at these flow magnitudes); both runs round their 16-bit
    activations independently, hence the 1e-3 bar of the parity tests rather than equality."""
    net = _net()
    h, w, c = 4320, 7680, 1024
    g = torch.Generator(device="cuda").manual_seed(12)
    lo = torch.rand(1, 3, h // 16, w // 16, device="cuda", generator=g)
    a = torch.nn.functional.interpolate(lo, size=(h, w), mode="bilinear", align_corners=False)
    b = torch.roll(a, shifts=(1, 2), dims=(2, 3)).contiguous()
    with torch.no_grad():
        y = net(a, b, t=0.5)
        yc = net(a[:, :, h - c:, w - c:].contiguous(), b[:, :, h - c:, w - c:].contiguous(), t=0.5)
    torch.cuda.synchronize()
    assert torch.isfinite(y).all()
    m = 320                                                       # margin from the crop's inner (top / left) edges
    err = (y[:, :, h - c + m:, w - c + m:] - yc[:, :, m:, m:]).abs().max().item()
    print(f"8K corner vs crop: max-abs {err:.3e}")
    assert err <= 1e-3, err
    net._engines.clear()
