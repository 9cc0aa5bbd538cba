"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference in this container.

TEST INFRASTRUCTURE.  Run once in the build container (``python oracle/make_golden.py``);
the GPU box has no ``/root/reference``, so tests there read only the committed fixtures.

It imports ``/root/reference/model.py`` as-is.  The single accommodation is a no-op
``torch.Tensor.cuda`` because ``warp`` hard-codes ``.cuda()`` (model.py:11-12) and this
container has no GPU; no reference file is modified or copied.

Weights are not stored (77 MB).  They are ``torch.manual_seed(0); Net()`` of the
reference; each fixture records the sha256 of the weight bytes so the tests can prove
that ``oracle.rrin_oracle.seeded_state_dict`` regenerates the identical tensors.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import rrin_oracle as O  # noqa: E402


def reference_net(seed=0):
    torch.Tensor.cuda = lambda self, *a, **k: self          # CPU shim
    sys.path.insert(0, REF)
    import model as ref_model                                 # /root/reference/model.py
    sys.path.pop(0)
    assert os.path.abspath(ref_model.__file__).startswith(REF)
    torch.manual_seed(seed)
    return ref_model.Net().eval()


CASES = [
    # name, N, H, W, t, stress_flow, stress_final, smooth
    ("rand64_t050", 1, 64, 64, 0.5, 1.0, 1.0, False),
    ("rand64_t0125", 1, 64, 64, 0.125, 1.0, 1.0, False),
    ("stress64_t050", 1, 64, 64, 0.5, 200.0, 20.0, False),
    ("stress_smooth_48x80_n2_t0875", 2, 48, 80, 0.875, 300.0, 20.0, True),
    ("rand_32x48_t030", 1, 32, 48, 0.3, 1.0, 1.0, False),
]


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    net = reference_net(0)
    base_sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    sha = O.weights_sha256(base_sd)
    print("reference weights sha256", sha)
    for name, n, h, w, t, sflow, sfinal, smooth in CASES:
        sd = {k: v.clone() for k, v in base_sd.items()}
        if sflow != 1.0:
            sd["Flow.last.weight"] *= sflow
            sd["Flow.last.bias"] *= sflow
        if sfinal != 1.0:
            sd["final.last.weight"] *= sfinal
        net.load_state_dict(sd, strict=True)
        a, b = O.seeded_frames(n, h, w, seed=1, smooth=smooth)
        with torch.no_grad():
            y = net(a, b, t=t)
            # intermediates of the reference, for finer-grained kernel tests
            x = torch.cat((a, b), 1)
            flow = net.Flow(x)
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"),
            in0=a.numpy(), in1=b.numpy(), out=y.numpy(), flow=flow.numpy(),
            t=np.float64(t), stress_flow=np.float64(sflow), stress_final=np.float64(sfinal),
            smooth=np.bool_(smooth), weights_sha256=np.array(sha),
            flow_last_bias=base_sd["Flow.last.bias"].numpy())
        print(f"{name}: out sum={float(y.double().sum()):.6f} max|flow|={float(flow.abs().max()):.3f} "
              f"clamped={(float(((y == 0) | (y == 1)).float().mean())):.4f}")

    # known answers at the CPU-reference config (368x368) -- summary only
    net.load_state_dict(base_sd, strict=True)
    a, b = O.seeded_frames(1, 368, 368, seed=1)
    kat = {}
    for t in (0.5, 0.125):
        with torch.no_grad():
            y = net(a, b, t=t)
        kat[f"sum_t{t}"] = float(y.double().sum())
        kat[f"px_t{t}"] = y[0, :, 183, 93].numpy()
        kat[f"corner_t{t}"] = y[0, :, -1, -1].numpy()
        # a strided sub-sample of the output keeps a dense-enough fingerprint small
        kat[f"sub_t{t}"] = y[0, :, ::8, ::8].numpy()
        print("368x368 t", t, "sum", kat[f"sum_t{t}"])
    np.savez_compressed(os.path.join(out_dir, "kat368.npz"), weights_sha256=np.array(sha), **kat)


if __name__ == "__main__":
    main()
