"""GPU edge cases of the hot path, checked against the CPU oracle on the same seeded inputs (the oracle itself is pinned
to the reference by tests/test_oracle_golden.py): maximum layer count, 64+192 samples, tiny sample counts, two rays,
every performer hidden, padded ray rows, bad arguments."""
import ctypes

import numpy as np
import pytest
import torch

import cases as C
from oracle import stnerf_oracle as O
from tests_support import build_case_model, make_cfg

pytestmark = pytest.mark.gpu


def _run_both(L, n1, n2, n_rays, seed, hidden=(), thr=(0.5, 0.1), near=0.0, precision="exact", space_time=True,
              extra_cols=0, frame_ids=None, keep=None):
    import modeling
    case = dict(weights="synthetic", seed=seed, L=L, space_time=space_time, n1=n1, n2=n2,
                frame_ids=frame_ids or ([0] + [10 + i for i in range(L)]), thr=thr, n_rays=n_rays, ray_seed=seed,
                hidden=list(hidden), near=near)
    sd = C.state_dict_for(case)
    rays = C.rays_for(case)
    jit, u = C.uniforms_for(case)
    if keep is not None:                            # render only the first `keep` rays of the generated set
        rays, jit, u, n_rays = rays[:keep].contiguous(), jit[:, :keep].contiguous(), u[:, :keep].contiguous(), keep
    nets = O.split_state_dict(sd, L)
    want = O.render(nets, C.scene_for(case), rays, n1, n2, jit, u, density_threshold=thr[0], bkgd_density_threshold=thr[1])
    want = C.flatten_outputs(want["fine_mixed"], want["coarse_mixed"], want["fine_layer"], want["coarse_layer"], want["ray_mask"])
    model = modeling.build_layered_model(make_cfg(L, n1, n2, space_time, precision))
    model.load_state_dict(sd)
    bkgd, frames = C.boxes_for(case)
    model.set_bkgd_bbox(bkgd); model.set_bboxes(frames); model.near = near
    for i in hidden:
        model.hide_layer(i)
    dev = torch.device("cuda", 0)
    r = rays.to(dev)
    if extra_cols:       # rays with more columns than 6+l are not a reference layout; exercise the C-ABI stride instead
        pad = torch.full((r.shape[0], r.shape[1] + extra_cols), 7.0, device=dev)
        pad[:, :r.shape[1]] = r
        r = pad[:, :r.shape[1]]                     # non-contiguous view with a wider row stride
        assert r.stride(0) == rays.shape[1] + extra_cols
        nat = model._ensure_native(dev)
        nat.set_scene(model._resolve_scene(rays[0, 6:], thr[0], thr[1]))
        from stnerf_b200 import _lib as Lb, split_planes
        out = torch.empty((2, L + 2, 5 * n_rays), device=dev)
        mask = torch.empty((L + 1, n_rays), dtype=torch.uint8, device=dev)
        jd, ud = jit.to(dev), u.to(dev)             # keep the device copies alive across the asynchronous call
        Lb.check(Lb.lib().stnerf_render(nat._h, Lb.ptr(r), n_rays, r.stride(0), n1, n2, 0, Lb.ptr(jd),
                                        Lb.ptr(ud), 1, Lb.ptr(out), Lb.ptr(mask), Lb.stream_ptr()), "render")
        torch.cuda.synchronize()
        fm, cm, fl, cl = split_planes(out, L + 1)
        got = C.flatten_outputs(fm, cm, fl, cl, [mask[i].bool() for i in range(L + 1)])
    else:
        model.inject_uniforms(jit.to(dev), u.to(dev))
        with torch.no_grad():
            got = C.flatten_outputs(*model(r, None, None, density_threshold=thr[0], bkgd_density_threshold=thr[1]))
    torch.cuda.synchronize()
    return got, want


def _assert_close(got, want, rgb_tol=1e-3):
    for k in sorted(want):
        if k.startswith("ray_mask"):
            assert np.array_equal(got[k], want[k]), k
        elif k.endswith("rgb") or k.endswith("acc"):
            err = np.abs(got[k].astype(np.float64) - want[k]).max()
            assert err <= rgb_tol, "%s %.3e" % (k, err)
        else:
            assert (np.abs(got[k] - want[k]) <= 2e-2 + 2e-3 * np.abs(want[k])).all(), k


def test_six_performers_64_192():
    """BASELINE config #5 shape: 6 performer layers + background, 64 coarse + 192 fine samples (1792-sample merge)."""
    got, want = _run_both(L=6, n1=64, n2=192, n_rays=70, seed=21, space_time=False)
    _assert_close(got, want)
    assert all(want["ray_mask.%d" % i].sum() > 0 for i in range(1, 7))


def test_tiny_sample_counts():
    got, want = _run_both(L=2, n1=8, n2=8, n_rays=96, seed=22)
    _assert_close(got, want)


def test_odd_sample_counts_fp32():
    got, want = _run_both(L=1, n1=37, n2=53, n_rays=64, seed=23, precision="fp32")
    _assert_close(got, want)


def test_two_rays():
    got, want = _run_both(L=2, n1=64, n2=128, n_rays=32, seed=24, keep=2)     # the smallest call the reference accepts
    assert got["fine_mixed.rgb"].shape == (2, 3)
    _assert_close(got, want)


def test_all_performers_hidden():
    got, want = _run_both(L=2, n1=64, n2=128, n_rays=96, seed=25, hidden=(1, 2))
    _assert_close(got, want)
    assert np.all(got["fine_layer.1.rgb"] == 0) and np.all(got["fine_layer.2.acc"] == 0)


def test_wide_ray_stride_through_c_abi():
    got, want = _run_both(L=2, n1=64, n2=128, n_rays=64, seed=26, extra_cols=3)
    _assert_close(got, want)


def test_c_abi_rejects_bad_arguments():
    from stnerf_b200 import NativeRenderer, _lib as Lb
    r = NativeRenderer(3, [False, True, True], "exact")
    lib = Lb.lib()
    rays = torch.zeros(16, 9, device="cuda")
    out = torch.zeros(2, 4, 80, device="cuda")
    args = lambda n1, n2, stride: (r._h, Lb.ptr(rays), 16, stride, n1, n2, 0, None, None, 0, Lb.ptr(out), None, Lb.stream_ptr())
    assert lib.stnerf_render(*args(64, 128, 9)) == -1            # no scene set yet
    r.set_scene(Lb.Scene())
    assert lib.stnerf_render(*args(64, 128, 9)) == -4            # weights never loaded -> STNERF_ENOWEIGHTS
    assert lib.stnerf_render(*args(2, 128, 9)) == -1             # n1 < 3
    assert lib.stnerf_render(*args(64, 600, 9)) == -1            # n1 + n2 > STNERF_MAX_S
    assert lib.stnerf_render(*args(64, 128, 8)) == -1            # ray row narrower than 6 + l
    torch.cuda.synchronize()
    r.close()


def test_coarse_fusion_is_not_observable(monkeypatch):
    """The coarse-pass fusion (per-layer compositing + resampling inside the SpaceNet kernel's spare warps, mlp_tc.cuh:
    FuseCoarse) and the stand-alone compositing kernel run the same arithmetic (csrc/resample.cuh): every output of forward() --
    coarse and fine images of every layer, masks -- is bit-identical with the fusion switched off (STNERF_NO_FUSE=1), with
    injected uniforms and with the in-kernel Philox stream, for 64+128 and 64+192 samples and with a hidden layer."""
    for name, n2, hidden in (("tkd_64_128", 128, []), ("walk_L4_64_128", 192, [2])):
        case = dict(C.CASES[name], n2=n2, hidden=hidden)
        if C.state_dict_for(case) is None:
            pytest.skip("checkpoint copy absent")
        rays = C.rays_for(case).cuda()
        jit, _ = C.uniforms_for(case)
        rs = np.random.RandomState(5)
        u = torch.from_numpy(rs.random_sample((case["L"] + 1, case["n_rays"], n2)).astype(np.float32)).clamp_(max=0.99999994)
        outs = {}
        for fused in (True, False):
            if fused:
                monkeypatch.delenv("STNERF_NO_FUSE", raising=False)
            else:
                monkeypatch.setenv("STNERF_NO_FUSE", "1")
            model = build_case_model(case, "exact")
            res = []
            for inject in (True, False):
                if inject:
                    model.inject_uniforms(jit.cuda(), u.cuda())
                model.seed = 41
                with torch.no_grad():
                    o = model(rays, None, None, density_threshold=case["thr"][0], bkgd_density_threshold=case["thr"][1])
                res.append(C.flatten_outputs(*o))
            outs[fused] = res
            del model
        for a, b in zip(outs[True], outs[False]):
            for k in a:
                assert np.array_equal(a[k], b[k]), (name, k, float(np.abs(a[k].astype(np.float64) - b[k]).max()))


def test_flow_reuse_is_not_observable(monkeypatch):
    """The fine pass takes the MotionNet flow of the n1 coarse depths from the coarse pass (same network, same points)
    and evaluates the MotionNet on the n2 new depths only; the origin map comes from the merge (csrc/resample.cuh).  Every output
    is bit-identical with the reuse switched off (STNERF_NO_REUSE=1): integer and fractional frame ids (MotionNet lerp), shift /
    scale edits, a `None` shift entry (which disables the fine-pass scale of that layer only -> no reuse there), 64+192 samples,
    fused and stand-alone merge."""
    variants = [("tkd_64_128", {}), ("tkd_edit_frac", {}), ("tkd_edit_frac", {"shift": [[0, 0, 0], None, [0, -2, 0]]}),
                ("walk_L4_64_128", {"n2": 192}), ("walk_90_30_hide", {})]
    for name, over in variants:
        case = dict(C.CASES[name], **over)
        if C.state_dict_for(case) is None:
            pytest.skip("checkpoint copy absent")
        rays = C.rays_for(case).cuda()
        outs = {}
        for reuse in (True, False):
            if reuse:
                monkeypatch.delenv("STNERF_NO_REUSE", raising=False)
            else:
                monkeypatch.setenv("STNERF_NO_REUSE", "1")
            model = build_case_model(case, "exact")
            model.seed = 17
            with torch.no_grad():
                o = model(rays, None, None, density_threshold=case["thr"][0], bkgd_density_threshold=case["thr"][1])
            outs[reuse] = C.flatten_outputs(*o)
            del model
        for k in outs[True]:
            assert np.array_equal(outs[True][k], outs[False][k]), (name, over, k, float(np.abs(outs[True][k].astype(np.float64) - outs[False][k]).max()))


def test_render_host_matches_device_render():
    """`stnerf_render_host` (the e2e path bench.py times: pinned host rays up and image planes / masks down, chunk by chunk on
    copy streams while other chunks render) returns exactly what `stnerf_render` leaves on the device for the same rays and seed --
    several chunks (chunk_rays = 1024), a ragged last chunk, coarse-only calls, staging pre-sized by `stnerf_reserve_host`."""
    name = "syn_L2_64_128"
    case = C.CASES[name]
    model = build_case_model(name, "exact", chunk_rays=1024)
    dev = torch.device("cuda", 0)
    nat = model._ensure_native(dev)
    nat.set_scene(model._resolve_scene(torch.tensor(case["frame_ids"], dtype=torch.float32), 0.3, 0.05))
    from stnerf_b200 import ops
    K, T = O.synthetic_camera(4, 16, 50, 75)
    rays = ops.generate_rays(K, T, 50, 75, frame_ids=case["frame_ids"])          # 3750 rays = 3 full chunks + 678
    nat.reserve_host(rays.shape[0], rays.shape[1])
    nat.render(rays[:8], 64, 128, seed=1)                                  # sizes the workspace (stnerf_reserve would too)
    ws0 = nat.workspace_bytes()
    rays_host = rays.cpu().pin_memory()
    for only_coarse in (False, True):
        out_d, mask_d = nat.render(rays, 64, 128, only_coarse=only_coarse, seed=77)
        out_h = torch.full((2, 4, 5 * rays.shape[0]), -7.0).pin_memory()
        mask_h = torch.full((3, rays.shape[0]), 9, dtype=torch.uint8).pin_memory()
        nat.render_host(rays_host, 64, 128, only_coarse=only_coarse, seed=77, out_host=out_h, mask_host=mask_h)
        assert torch.equal(mask_h, mask_d.cpu())
        assert torch.equal(out_h[0], out_d[0].cpu())                       # coarse images of every layer
        if only_coarse:
            assert bool((out_h[1] == -7.0).all())                          # fine planes untouched
        else:
            assert torch.equal(out_h[1], out_d[1].cpu())
    assert nat.workspace_bytes() == ws0
