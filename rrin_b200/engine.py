"""Host side of the engine: packed weights, workspace and the forward call.

PyTorch is used only as plumbing here -- device allocations (weights blob, workspace, output)
and the current CUDA stream.  All arithmetic happens inside ``librrin_b200.so``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence, Union

import torch

from ._lib import check, lib


# operand formats of the tensor-core path -> code passed through the C-ABI
PRECISIONS = {"bf16": 0, "fp16": 1}

# RRIN_GRAPH=0: launch kernel by kernel instead of replaying CUDA graphs (A/B timing; results are bit-identical)
USE_GRAPH = os.environ.get("RRIN_GRAPH", "1") != "0"


def conv_table():
    """[(key, cin, cout, level, src_mode, act)] for the 81 convs in execution order."""
    l = lib()
    out = []
    buf = C.create_string_buffer(128)
    for i in range(l.rrin_num_convs()):
        v = [C.c_int() for _ in range(5)]
        check(l.rrin_conv_info(i, buf, 128, *[C.byref(x) for x in v]), "rrin_conv_info")
        out.append((buf.value.decode(), *[x.value for x in v]))
    return out


def _ptr(t: torch.Tensor) -> int:
    return t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class PackedWeights:
    """bf16 UMMA-ordered copy of all 81 conv weights + fp32 biases in one device blob (K7)."""

    def __init__(self, net: torch.nn.Module, device: torch.device, fingerprint=None, precision: str = "bf16"):
        l = lib()
        self.device, self.fingerprint, self.precision = device, fingerprint, precision
        if precision not in PRECISIONS:
            raise ValueError(f"unknown precision {precision!r}")
        sd = net.state_dict()
        with torch.cuda.device(device):
            self.blob = torch.zeros(l.rrin_packed_weights_bytes(), dtype=torch.uint8, device=device)
            keep = []
            for i, (key, cin, cout, *_rest) in enumerate(conv_table()):
                w = sd[key + ".weight"].detach().to(device=device, dtype=torch.float32).contiguous()
                b = sd[key + ".bias"].detach().to(device=device, dtype=torch.float32).contiguous()
                if tuple(w.shape) != (cout, cin, 3, 3) or tuple(b.shape) != (cout,):
                    raise RuntimeError(f"size mismatch for {key}: {tuple(w.shape)} vs {(cout, cin, 3, 3)}")
                keep += [w, b]
                check(l.rrin_pack_conv_ex(i, _ptr(w), _ptr(b), _ptr(self.blob), PRECISIONS[precision], _stream()), f"rrin_pack_conv({key})")
            torch.cuda.current_stream().synchronize()   # w/b temporaries may be freed after this


def time_coefficients(t: Union[float, torch.Tensor, Sequence[float]], n: int, device) -> torch.Tensor:
    """fp32 [n,6] = {-(1-t)t, t*t, (1-t)(1-t), t(1-t), 1-t, t} as model.py:38-39,54 evaluates them:
    Python-float t -> products in double, rounded once to fp32 (scalar * tensor semantics);
    tensor t -> the same expressions in fp32 tensor arithmetic."""
    if isinstance(t, torch.Tensor):
        tt = t.detach().to(dtype=torch.float32, device="cpu").reshape(-1)
        if tt.numel() == 1:
            tt = tt.expand(n)
        if tt.numel() != n:
            raise RuntimeError("tensor-valued t must be a scalar or broadcast per sample ([N,1,1,1]); "
                               f"got {tuple(t.shape)} for batch {n}")
        c = torch.stack([-(1 - tt) * tt, tt * tt, (1 - tt) * (1 - tt), tt * (1 - tt), 1 - tt, tt], 1)
        return c.contiguous().to(device)
    ts = [float(t)] * n if not isinstance(t, (list, tuple)) else [float(x) for x in t]
    if len(ts) != n:
        raise RuntimeError(f"expected {n} timesteps, got {len(ts)}")
    rows = [[-(1 - x) * x, x * x, (1 - x) * (1 - x), x * (1 - x), 1 - x, x] for x in ts]
    return torch.tensor(rows, dtype=torch.float64).to(torch.float32).to(device)


class Engine:
    """One problem shape (n_pairs, n_samples, H, W) on one device."""

    def __init__(self, device: torch.device, n: int, h: int, w: int, n_pairs: int | None = None, precision: str = "bf16"):
        l = lib()
        self.device, self.n, self.h, self.w, self.precision = device, n, h, w, precision
        if precision not in PRECISIONS:
            raise ValueError(f"unknown precision {precision!r}")
        self.n_pairs = n if n_pairs is None else n_pairs
        hnd = C.c_void_p()
        check(l.rrin_engine_create_ex(self.n_pairs, n, h, w, PRECISIONS[precision], C.byref(hnd)), "rrin_engine_create")
        self._h = hnd
        with torch.cuda.device(device):
            self.workspace = torch.empty(l.rrin_engine_workspace_bytes(hnd), dtype=torch.uint8, device=device)
            self._done = torch.cuda.Event()           # end of the last forward that used this engine's workspace
        self._last_stream = None
        self.num_launches = l.rrin_engine_num_launches(hnd)
        self.workspace_bytes = self.workspace.numel()
        self._coef_cache = {}

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().rrin_engine_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def graph_stats(self):
        """(forwards replayed from a CUDA graph, forwards launched kernel by kernel, graphs alive)."""
        v = [C.c_int() for _ in range(3)]
        check(lib().rrin_engine_graph_stats(self._h, *[C.byref(x) for x in v]), "rrin_engine_graph_stats")
        return tuple(x.value for x in v)

    @staticmethod
    def _frames(x: torch.Tensor) -> torch.Tensor:
        """fp32, contiguous, 16-byte aligned (the kernels read frames with 8-byte loads): a misaligned view -- e.g. a slice of
        a flat buffer at an odd element offset -- is copied rather than faulted on."""
        x = x.detach().to(torch.float32).contiguous()
        return x if x.data_ptr() % 16 == 0 else x.clone()

    def run(self, weights: "PackedWeights", in0: torch.Tensor, in1: torch.Tensor, coef: torch.Tensor,
            out: torch.Tensor | None = None) -> torch.Tensor:
        in0, in1 = self._frames(in0), self._frames(in1)
        # CUDA-graph replay needs stable pointers (a graph bakes them in): callers that own their output buffer
        # (forward_into / the streaming pipeline) cycle through a few buffers and replay; `forward`, which allocates a new
        # result tensor per call like the reference, launches kernel by kernel.
        graph = USE_GRAPH and out is not None
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.empty((self.n, 3, self.h, self.w), dtype=torch.float32, device=self.device)
            elif out.data_ptr() % 16:
                raise RuntimeError("`out` must be 16-byte aligned (the kernels write it with vector stores)")
            cur = torch.cuda.current_stream()
            # One workspace per engine: a forward issued on another stream than the previous one is ordered after it.
            if self._last_stream is not None and self._last_stream != cur.cuda_stream:
                cur.wait_event(self._done)
            fwd = lib().rrin_engine_forward_graph if graph else lib().rrin_engine_forward
            check(fwd(self._h, _ptr(weights.blob), _ptr(self.workspace), _ptr(in0), _ptr(in1),
                      _ptr(coef), _ptr(out), cur.cuda_stream), "rrin_engine_forward")
            self._done.record(cur)
            self._last_stream = cur.cuda_stream
        return out

    def _coef(self, t):
        if isinstance(t, torch.Tensor):
            return time_coefficients(t, self.n, self.device)
        key = tuple(t) if isinstance(t, (list, tuple)) else float(t)
        c = self._coef_cache.get(key)
        if c is None:
            if len(self._coef_cache) > 64:
                self._coef_cache.clear()
            c = self._coef_cache[key] = time_coefficients(t, self.n, self.device)
        return c

    def forward(self, weights, in0, in1, t) -> torch.Tensor:
        return self.run(weights, in0, in1, self._coef(t))

    def forward_multi(self, weights, in0, in1, ts: List[float]) -> torch.Tensor:
        if self.n_pairs != 1:
            raise RuntimeError("forward_multi needs an engine created with n_pairs=1")
        return self.run(weights, in0, in1, self._coef(list(ts)))

    def launch_table(self):
        """[(kernel_class, layer, flops, bytes)] for the launches of one forward, in stream order."""
        l, out = lib(), []
        nm, ly = C.create_string_buffer(128), C.create_string_buffer(128)
        for i in range(self.num_launches):
            fl, by = C.c_double(), C.c_double()
            check(l.rrin_engine_launch_info(self._h, i, nm, 128, ly, 128, C.byref(fl), C.byref(by)))
            out.append((nm.value.decode(), ly.value.decode(), fl.value, by.value))
        return out

    def profile(self, weights, in0, in1, t) -> List[float]:
        """Per-launch device milliseconds of one forward (CUDA events on the launching stream)."""
        coef = time_coefficients(t, self.n, self.device)
        in0, in1 = self._frames(in0), self._frames(in1)
        ms = (C.c_float * self.num_launches)()
        with torch.cuda.device(self.device):
            out = torch.empty((self.n, 3, self.h, self.w), dtype=torch.float32, device=self.device)
            check(lib().rrin_engine_forward_profiled(self._h, _ptr(weights.blob), _ptr(self.workspace), _ptr(in0), _ptr(in1),
                                                     _ptr(coef), _ptr(out), _stream(), ms), "rrin_engine_forward_profiled")
        return list(ms)

    def tap(self, which: int) -> torch.Tensor:
        n = self.n_pairs if which == 0 else self.n
        dst = torch.empty((n, self.h // 2, self.w // 2, 4, 4), dtype=torch.float32, device=self.device)
        check(lib().rrin_engine_tap(self._h, _ptr(self.workspace), which, _ptr(dst), _stream()), "rrin_engine_tap")
        # space-to-depth [n,H/2,W/2,phase,4] -> NCHW [n,4,H,W]
        return dst.reshape(n, self.h // 2, self.w // 2, 2, 2, 4).permute(0, 5, 1, 3, 2, 4).reshape(n, 4, self.h, self.w)
