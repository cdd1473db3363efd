"""tests/tools/reference_self_noise.py permutes the hidden units of every MLP layer to give the reference's arithmetic another
fp32 summation order.  The permutation must leave the networks unchanged in exact arithmetic: in float64 the permuted SpaceNet /
MotionNet agree with the originals to rounding (1e-12), while in float32 they differ (otherwise the study would measure nothing)."""
import importlib.util
import os

import torch

import cases as C
from oracle import stnerf_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tool():
    spec = importlib.util.spec_from_file_location("reference_self_noise", os.path.join(ROOT, "tests", "tools", "reference_self_noise.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)          # defines functions only; main() runs under __main__
    return mod


def test_hidden_unit_permutation_is_exact_in_float64_and_visible_in_float32():
    T = _tool()
    sd = O.synthetic_state_dict(2, True, seed=5)
    nets = O.split_state_dict(sd, 2)
    perm = T.permute_nets(nets, seed=11)
    g = torch.Generator().manual_seed(0)
    pos = torch.rand(512, 3, generator=g) * 4 - 2
    dirs = torch.nn.functional.normalize(torch.randn(512, 3, generator=g), dim=1)
    tm = torch.full((512, 1), 10.0)
    xyzt = torch.cat([pos, tm], 1)

    def f64(w):
        return {k: v.double() for k, v in w.items()}
    for a, b, use_time in ((nets["bkgd"], perm["bkgd"], False), (nets["space"][1], perm["space"][1], True),
                           (nets["space_fine"][0], perm["space_fine"][0], True)):
        assert any(not torch.equal(a[k], b[k]) for k in a)                       # something was permuted
        r0, s0 = O.spacenet_forward(f64(a), pos.double(), dirs.double(), tm.double() if use_time else None)
        r1, s1 = O.spacenet_forward(f64(b), pos.double(), dirs.double(), tm.double() if use_time else None)
        scale = s0.abs().max().clamp(min=1.0)
        assert (s0 - s1).abs().max() / scale < 1e-12 and (r0 - r1).abs().max() < 1e-10
        r0f, s0f = O.spacenet_forward(a, pos, dirs, tm if use_time else None)
        r1f, s1f = O.spacenet_forward(b, pos, dirs, tm if use_time else None)
        assert not torch.equal(s0f, s1f)                                         # fp32: another summation order, other bits
        assert (s0f - s1f).abs().max() / scale < 1e-4
    m0 = O.motionnet_forward(f64(nets["motion"][0]), xyzt.double())
    m1 = O.motionnet_forward(f64(perm["motion"][0]), xyzt.double())
    assert (m0 - m1).abs().max() < 1e-12
