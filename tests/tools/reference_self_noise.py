#!/usr/bin/env python
"""CPU study (oracle-based, test infrastructure): how far is the reference from ITSELF under a different fp32 summation order?

The hidden units of every MLP layer are permuted (rows of layer k together with the matching input columns of layer k+1): in
exact arithmetic the network is unchanged, in fp32 every dot product of the next layer is summed in a different order -- what any
other GEMM library, tiling or device does to the reference's arithmetic.  The permuted networks render the committed scale
fixtures (same rays, same uniforms) through the oracle, which is bit-identical to the unmodified reference on them, and the
result is compared with the reference's own fixture: rays over the 1e-3 gate, max / median error, PSNR.  That is the floor of
the gate for ANY implementation that is not bit-for-bit the reference's GEMM -- the number the GPU path's outlier counts
(tests/test_gpu_parity_scale.py) are to be read against.

    python tests/tools/reference_self_noise.py [case ...]        # default: all three scale cases; prints one JSON line
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "st-nerf_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import numpy as np, torch
import cases as C
from oracle import stnerf_oracle as O

torch.set_num_threads(os.cpu_count() or 1)


def permute_spacenet(w, g):
    w = {k: v.clone() for k, v in w.items()}

    def perm_out(name, consumers):
        n = w[name + ".weight"].shape[0]
        p = torch.randperm(n, generator=g)
        w[name + ".weight"] = w[name + ".weight"][p]
        w[name + ".bias"] = w[name + ".bias"][p]
        for cname in consumers:                       # the first n input columns of every consumer are this layer's outputs
            cw = w[cname + ".weight"]
            w[cname + ".weight"] = torch.cat([cw[:, :n][:, p], cw[:, n:]], 1)
    perm_out("stage1.0", ["stage1.2"]); perm_out("stage1.2", ["stage1.4"]); perm_out("stage1.4", ["stage1.6"])
    perm_out("stage1.6", ["stage2.0"])                # stage2.0 input = cat[x (256), PE(pos) (63)]
    perm_out("stage2.0", ["stage2.2"]); perm_out("stage2.2", ["stage2.4"])
    perm_out("stage2.4", ["density_net.0", "rgb_net.1"])   # rgb_net.1 input = cat[x (256), PE(dir) [, PE(t)]]
    perm_out("rgb_net.1", ["rgb_net.3"])
    return w


def permute_motionnet(w, g):
    w = {k: v.clone() for k, v in w.items()}
    for a, b in ((0, 2), (2, 4), (4, 6), (6, 8), (8, 10)):
        n = w["motion_net.%d.weight" % a].shape[0]
        p = torch.randperm(n, generator=g)
        w["motion_net.%d.weight" % a] = w["motion_net.%d.weight" % a][p]
        w["motion_net.%d.bias" % a] = w["motion_net.%d.bias" % a][p]
        w["motion_net.%d.weight" % b] = w["motion_net.%d.weight" % b][:, p]
    return w


def permute_nets(nets, seed):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in nets.items():
        if isinstance(v, list):
            out[k] = [(permute_motionnet(x, g) if k == "motion" else permute_spacenet(x, g)) if x is not None else None for x in v]
        elif isinstance(v, dict):
            out[k] = permute_spacenet(v, g)
        else:
            out[k] = v
    return out


def render(case, nets, rays, jit, u):
    sc = C.scene_for(case)
    outs = []
    with torch.no_grad():
        for c0 in range(0, rays.shape[0], 2048):
            w = O.render(nets, sc, rays[c0:c0 + 2048], case["n1"], case["n2"], jit[:, c0:c0 + 2048], u[:, c0:c0 + 2048],
                         density_threshold=case["thr"][0], bkgd_density_threshold=case["thr"][1])
            outs.append(w["fine_mixed"][0])
    return torch.cat(outs, 0).numpy()


def main():
    names = sys.argv[1:] or list(C.SCALE_CASES)
    res = {"what": "unmodified-reference arithmetic (oracle, bit-identical to it on these fixtures) with the hidden units of every "
                   "layer permuted = another fp32 summation order, vs the reference's own fixture; fine mixed rgb"}
    for name in names:
        case = dict(C.SCALE_CASES[name], name=name)
        gold = C.load_golden(name)
        sd = C.state_dict_for(case)
        if gold is None or sd is None:
            continue
        rays, jit, u = C.scale_inputs(case)
        nets = O.split_state_dict(sd, case["L"])
        ref = gold["fine_mixed.rgb"]
        rows = []
        base = render(case, nets, rays, jit, u)
        rows.append({"variant": "same order (sanity)", "max_abs_err": float(np.abs(base - ref).max())})
        for seed in (1, 2, 3):
            got = render(case, permute_nets(nets, seed), rays, jit, u)
            err = np.abs(got - ref).max(1)
            mse = float(((got.astype(np.float64) - ref) ** 2).mean())
            rows.append({"variant": "hidden units permuted, seed %d" % seed, "rays_over_1e-3": int((err > 1e-3).sum()),
                         "max_abs_err": float(err.max()), "median_err": float(np.median(err)),
                         "p999_err": float(np.sort(err)[int(0.999 * err.size)]),
                         "psnr_db": 99.0 if mse == 0 else float(10 * np.log10(1.0 / mse))})
            print(name, rows[-1], file=sys.stderr, flush=True)
        res[name] = {"rays": int(rays.shape[0]), "runs": rows}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
