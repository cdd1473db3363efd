"""Pin the CPU oracle against vectors produced by the unmodified reference (tests/golden/*.npz).

These run on CPU (`-m "not gpu"`).  Tolerances: the oracle uses the same ATen fp32 ops as the
reference, so most stages are bit-identical; matmul-bearing stages may differ by BLAS blocking.
"""
import numpy as np
import pytest
import torch

import cases as C
from oracle import stnerf_oracle as O

FN = C.load_golden("functions")
IN = C.function_inputs()


def close(a, b, atol, rtol=0.0):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    err = np.abs(a - b) - rtol * np.abs(b)
    assert err.max() <= atol, "max err %.3e" % np.abs(a - b).max()


def test_functions_npz_present():
    assert FN is not None, "tests/golden/functions.npz missing (run tests/golden/make_golden.py)"


def test_intersection():
    r = IN["isect.rays"]
    far, near = O.ray_box_intersect(r[:, :3], r[:, 3:], IN["isect.bmin"], IN["isect.bmax"])
    assert np.array_equal(torch.stack([far, near], 1).numpy(), FN["isect.t"])      # bit-exact


@pytest.mark.parametrize("layer", [0, 1])
def test_stratified_samples(layer):
    r = IN["isect.rays"]
    t, xyz, m = O.stratified_samples(r[:, :3], r[:, 3:], IN["isect.bmin"], IN["isect.bmax"], 48,
                                     IN["sample.jitter"][layer], is_bkgd=(layer == 0))
    assert np.array_equal(t.numpy(), FN["sample.t.%d" % layer])
    assert np.array_equal(xyz.numpy(), FN["sample.xyz.%d" % layer])
    assert np.array_equal(m.numpy().astype(np.uint8), FN["sample.mask.%d" % layer])


def test_composite():
    c, d, a, w = O.composite(IN["comp.t"], IN["comp.rgb"], IN["comp.sigma"])
    close(w, FN["comp.w"], 1e-7)
    close(c, FN["comp.color"], 1e-6)
    close(d, FN["comp.depth"], 1e-5)
    close(a, FN["comp.acc"], 1e-6)


def test_sample_pdf():
    z = O.sample_pdf(IN["pdf.t"], IN["pdf.w"][:, 1:-1], IN["pdf.u"])
    close(z, FN["pdf.z"], 1e-6)


def test_positional_encoding():
    assert np.array_equal(O.positional_encoding(IN["pe.x3"], 10).numpy(), FN["pe.x3_L10"])
    assert np.array_equal(O.positional_encoding(IN["pe.x3"], 4).numpy(), FN["pe.x3_L4"])
    assert np.array_equal(O.positional_encoding(IN["pe.x1"], 10).numpy(), FN["pe.x1_L10"])


def test_generate_rays():
    rays = O.generate_rays(IN["rays.K"], IN["rays.T"], 24, 40)
    close(rays, FN["rays.rays"], 2e-6)


def _nets(tag):
    if tag == "syn":
        return O.synthetic_state_dict(1, True, seed=5)
    p = C.find_checkpoint({"tkd": "taekwondo", "walk": "walking"}[tag])
    if p is None:
        pytest.skip("checkpoint copy not present (oracle/_ref/ckpt)")
    return torch.load(p, map_location="cpu")["model"]


@pytest.mark.parametrize("tag", ["syn", "tkd", "walk"])
def test_networks(tag):
    if "net.%s.perf.rgb" % tag not in FN:
        pytest.skip("golden for %s not generated" % tag)
    nets = O.split_state_dict(_nets(tag), 1)
    rgb, sig = O.spacenet_forward(nets["space"][0], IN["net.pos"], IN["net.dirs"], IN["net.time_int"])
    close(rgb, FN["net.%s.perf.rgb" % tag], 1e-4, 1e-5)
    close(sig, FN["net.%s.perf.sigma" % tag], 1e-3, 1e-5)
    rgb, sig = O.spacenet_forward(nets["bkgd_fine"], IN["net.pos"], IN["net.dirs"], None)
    close(rgb, FN["net.%s.bkgd.rgb" % tag], 1e-4, 1e-5)
    close(sig, FN["net.%s.bkgd.sigma" % tag], 1e-3, 1e-5)
    for kind, tcol in (("int", IN["net.time_int"]), ("frac", IN["net.time_frac"])):
        flow = O.motionnet_forward(nets["motion"][0], torch.cat([IN["net.pos"], tcol], 1))
        close(flow, FN["net.%s.motion_%s" % (tag, kind)], 1e-5, 1e-5)


@pytest.mark.parametrize("name", list(C.CASES))
def test_render_case(name):
    """End-to-end forward vs the reference's 5-tuple on identical rays / weights / uniforms."""
    case = C.CASES[name]
    gold = C.load_golden(name)
    assert gold is not None
    sd = C.state_dict_for(case)
    if sd is None:
        pytest.skip("checkpoint copy not present (oracle/_ref/ckpt)")
    nets = O.split_state_dict(sd, case["L"])
    jit, u = C.uniforms_for(case)
    out = O.render(nets, C.scene_for(case), C.rays_for(case), case["n1"], case["n2"], jit, u,
                   only_coarse=case.get("only_coarse", False),
                   density_threshold=case["thr"][0], bkgd_density_threshold=case["thr"][1],
                   shared_frame=case.get("seven", False))
    flat = C.flatten_outputs(out["fine_mixed"], out["coarse_mixed"], out["fine_layer"], out["coarse_layer"],
                             out["ray_mask"])
    assert set(flat) == set(gold)
    for k in sorted(gold):
        if k.startswith("ray_mask"):
            assert np.array_equal(flat[k], gold[k]), k
        elif k.endswith("rgb") or k.endswith("acc"):
            close(flat[k], gold[k], 2e-5)
        else:
            close(flat[k], gold[k], 2e-4, 1e-5)
