"""CPU: pin the oracle (torch-functional port) and the first-principles numpy restatement
against fixtures produced by the unmodified reference (oracle/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import rrin_numpy, rrin_oracle as O

CASES = ["rand64_t050", "rand64_t0125", "stress64_t050", "stress_smooth_48x80_n2_t0875", "rand_32x48_t030"]


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def sd_for(g):
    return O.seeded_state_dict(float(g["stress_flow"]), float(g["stress_final"]))


def test_seeded_weights_equal_reference_weights(golden_dir):
    g = load(golden_dir, "rand64_t050")
    sd = O.seeded_state_dict()
    assert len(sd) == 162 and sum(v.numel() for v in sd.values()) == 19_194_445
    assert O.weights_sha256(sd) == str(g["weights_sha256"])
    assert str(g["weights_sha256"]).startswith("83bf701ca6edf47d")
    np.testing.assert_array_equal(sd["Flow.last.bias"].numpy(), g["flow_last_bias"])
    assert list(sd)[0].startswith("Mask.") and list(sd)[-1] == "final.last.bias"


def test_seeded_frames_equal_fixture_inputs(golden_dir):
    for name in CASES:
        g = load(golden_dir, name)
        n, _, h, w = g["in0"].shape
        a, b = O.seeded_frames(n, h, w, seed=1, smooth=bool(g["smooth"]))
        np.testing.assert_array_equal(a.numpy(), g["in0"])
        np.testing.assert_array_equal(b.numpy(), g["in1"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_output(golden_dir, name):
    g = load(golden_dir, name)
    taps = {}
    y = O.forward(sd_for(g), torch.from_numpy(g["in0"]), torch.from_numpy(g["in1"]), float(g["t"]), taps)
    # same torch ops in the same order: bit-exact up to thread-count dependent conv summation
    assert np.abs(y.numpy() - g["out"]).max() <= 2e-6
    assert np.abs(taps["flow"].numpy() - g["flow"]).max() <= 1e-5 * max(1.0, float(g["stress_flow"]))


@pytest.mark.parametrize("name", ["rand64_t050", "stress64_t050", "rand_32x48_t030"])
def test_numpy_restatement_matches_reference_output(golden_dir, name):
    g = load(golden_dir, name)
    sd = {k: v.numpy() for k, v in sd_for(g).items()}
    y = rrin_numpy.forward(sd, g["in0"], g["in1"], float(g["t"]))
    err = np.abs(y - g["out"]).max()
    # fp32 summation-order noise only; amplified by |grad img| * flow error under stress weights
    assert err <= (5e-5 if float(g["stress_flow"]) == 1.0 else 2e-3), err


def test_known_answers_368(golden_dir):
    """config-1 size (BASELINE.json configs[0]); ~1 s per forward on 8 cores."""
    k = load(golden_dir, "kat368")
    sd = O.seeded_state_dict()
    a, b = O.seeded_frames(1, 368, 368)
    y = O.forward(sd, a, b, 0.5)
    assert abs(float(y.double().sum()) - float(k["sum_t0.5"])) < 0.05
    assert abs(float(y.double().sum()) - 203111.321613) < 0.05
    assert np.abs(y[0, :, ::8, ::8].numpy() - k["sub_t0.5"]).max() <= 2e-6
    np.testing.assert_allclose(y[0, :, 183, 93].numpy(), [0.486319, 0.643589, 0.413230], atol=2e-6)


def test_tensor_t_and_batch(golden_dir):
    g = load(golden_dir, "rand_32x48_t030")
    sd = sd_for(g)
    a, b = torch.from_numpy(g["in0"]), torch.from_numpy(g["in1"])
    yt = O.forward(sd, a, b, torch.full((1, 1, 1, 1), 0.3))
    assert np.abs(yt.numpy() - g["out"]).max() <= 1e-5
    y2 = O.forward(sd, torch.cat([a, b]), torch.cat([b, a]), 0.3)
    assert np.abs(y2[:1].numpy() - g["out"]).max() <= 1e-5


def test_zero_flow_is_not_identity():
    """align_corners=False semantics: zero flow samples at (x-0.5, y-0.5)."""
    img = torch.rand(1, 3, 16, 16, generator=torch.Generator().manual_seed(3))
    out = O.warp(img, torch.zeros(1, 2, 16, 16))
    assert (out - img).abs().max() > 0.1
    ref = rrin_numpy.warp(img.numpy(), np.zeros((1, 2, 16, 16), np.float32))
    assert np.abs(out.numpy() - ref).max() < 1e-6


def test_oracle_warp_equals_reference_warp(golden_dir):
    """``oracle.rrin_oracle.warp`` against the unmodified reference's ``warp`` (model.py:8-21; oracle/make_golden_warp.py):
    bit-identical, NaN positions included."""
    d = np.load(os.path.join(golden_dir, "warp_cases.npz"))
    names = sorted({k.split("/")[0] for k in d.files})
    assert len(names) == 4
    for name in names:
        img, flow, ref = (torch.from_numpy(d[f"{name}/{k}"]) for k in ("img", "flow", "out"))
        y = O.warp(img, flow)
        np.testing.assert_array_equal(y.numpy(), ref.numpy())
