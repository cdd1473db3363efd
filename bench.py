#!/usr/bin/env python
"""bench.py -- rays/sec of the layered ray-march hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--precision exact|mixed|fp32|fast]

Workload (BASELINE.json configs[1]): taekwondo 2-layer scene, 1080p, 16 views, 64 coarse + 128 fine samples.
A *step* renders one 1080p view (2 073 600 rays; view = step mod 16) through the whole hot path: bbox-clipped
sampling -> MotionNet/SpaceNet MLPs -> resampling -> fine MLPs -> per-layer + merged compositing of every
image plane.  Scene geometry and cameras are synthetic (the dataset is not shipped); weights are
the shipped taekwondo checkpoint when its copy is present under oracle/_ref/ckpt, else seeded random weights of
the same architecture (the cost of the path does not depend on weight values).

value  = device-resident throughput: rays already in HBM, CUDA events around the K steps, max over ranks.
e2e    = same metric through the host-buffer C-ABI call (stnerf_render_host): pinned host rays -> H2D ->
         render -> D2H of every image plane, inside the timed region.
N > 1  : each view's rows are interleaved over the ranks, one all-gather of the fine image planes per view
         (strong scaling: total work per step is fixed).
--impl reference : the CPU oracle port of the reference algorithm (torch fp32, all host threads) timed on a
         bounded sample of the same workload; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "st-nerf_b200"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

H, W, VIEWS, N1, N2, LAYERS = 1080, 1920, 16, 64, 128, 2
FRAME_IDS = [0.0, 10.0, 11.0]
FLOP_SPACE_BKGD, FLOP_SPACE_PERF, FLOP_MOTION = 924672.0, 930048.0, 153344.0   # SURVEY 8(d), 2*MAC per point
CKPT, SPACE_TIME, THR, NEAR = "taekwondo", True, (0.0, 0.0), 0.0                # demo/taekwondo_demo.py:44
WORKLOAD_NAME = "taekwondo 2-layer 1080p, 16 views, 64+128 samples (BASELINE configs[1])"

# The headline line is always configs[1] (default).  The other BASELINE configs can be timed for the record with --workload.
WORKLOADS = {
    "taekwondo2": None,
    # configs[2]: walking nets replicated round-robin to 4 performers, demo/walking_demo.py:43-50 thresholds
    "walking4": dict(H=1080, W=1920, VIEWS=16, N1=64, N2=128, LAYERS=4, CKPT="walking", SPACE_TIME=False, THR=(20.0, 0.8),
                     NEAR=4.0, FLOP_SPACE_PERF=924672.0,
                     WORKLOAD_NAME="walking 4-layer 1080p, 16 views, 64+128 samples (BASELINE configs[2])"),
    # configs[4]: 6 performers, 4K, 32 views, 64+192
    "walking6_4k": dict(H=2160, W=3840, VIEWS=32, N1=64, N2=192, LAYERS=6, CKPT="walking", SPACE_TIME=False,
                        THR=(20.0, 0.8), NEAR=4.0, FLOP_SPACE_PERF=924672.0,
                        WORKLOAD_NAME="walking 6-layer 4K, 32 views, 64+192 samples (BASELINE configs[4])"),
}


def select_workload(name):
    w = WORKLOADS[name]
    if w:
        globals().update(w)
        globals()["FRAME_IDS"] = [0.0] + [30.0 + i for i in range(w["LAYERS"])]


def load_weights():
    import torch
    from stnerf_b200 import checkpoint_io
    p = checkpoint_io.find_checkpoint(CKPT)
    if p is not None:
        sd = checkpoint_io.replicate_layers(torch.load(p, map_location="cpu")["model"], LAYERS)
        return sd, "%s checkpoint (oracle/_ref/ckpt)%s" % (CKPT, ", nets replicated round-robin to %d performers" % LAYERS if LAYERS > 2 else "")
    from stnerf_b200 import synthetic
    return synthetic.synthetic_state_dict(LAYERS, SPACE_TIME, seed=7), "seeded random weights (checkpoint copy absent)"


def scene_setup():
    from stnerf_b200 import synthetic              # synthetic scene description (inputs only)
    bkgd, frames = synthetic.synthetic_boxes(LAYERS)
    cams = [synthetic.synthetic_camera(v, VIEWS, H, W) for v in range(VIEWS)]
    return bkgd, frames, cams


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_rate(steps: int, warmup: int, sample_rays: int):
    """The reference algorithm on the host cores: oracle port (kind 'port'), bounded sample per step."""
    import torch
    from oracle import stnerf_oracle as O          # the CPU leg is the one place bench.py executes the oracle
    sd, data = load_weights()
    nets = O.split_state_dict(sd, LAYERS)
    bkgd, frames, cams = scene_setup()
    sc = O.resolve_scene(frames, bkgd, FRAME_IDS, None, None)
    sc.update(scale=None, shift=None, shown=[True] * (LAYERS + 1), near=NEAR, alpha=1.0, boarder=1e10)
    fid = torch.tensor(FRAME_IDS)[None]
    times = []
    gen = torch.Generator().manual_seed(1234)

    def probe(nthreads):
        """rays/s of a 256-ray sample with `nthreads` torch threads (picks the fastest host configuration)."""
        torch.set_num_threads(nthreads)
        K, T = cams[0]
        full = O.generate_rays(K, T, H, W)
        idx = torch.arange(256) + (H // 2) * W + (W - 256) // 2
        rays = torch.cat([full[idx], fid.expand(256, -1)], 1)
        jit, u = torch.rand((LAYERS + 1, 256, N1), generator=gen), torch.rand((LAYERS + 1, 256, N2), generator=gen)
        best = 0.0
        for _ in range(2):
            t0 = time.perf_counter()
            with torch.no_grad():
                O.render(nets, sc, rays, N1, N2, jit, u, density_threshold=THR[0], bkgd_density_threshold=THR[1])
            best = max(best, 256 / (time.perf_counter() - t0))
        return best

    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (ncpu, ncpu // 2, 64, 32, 16, 8) if 1 <= c <= ncpu}, reverse=True)
    rates = {c: probe(c) for c in cands}
    threads = max(rates, key=rates.get)
    torch.set_num_threads(threads)
    for s in range(warmup + steps):
        K, T = cams[s % VIEWS]
        full = O.generate_rays(K, T, H, W)
        # rows spread over the image so the layer hit fractions are representative
        per = max(1, sample_rays // 4)
        idx = torch.cat([torch.arange(per) + (H * (2 * q + 1) // 8) * W + (W - per) // 2 for q in range(4)])
        rays = torch.cat([full[idx], fid.expand(idx.numel(), -1)], 1)
        jit = torch.rand((LAYERS + 1, rays.shape[0], N1), generator=gen)
        u = torch.rand((LAYERS + 1, rays.shape[0], N2), generator=gen)
        t0 = time.perf_counter()
        with torch.no_grad():
            out = O.render(nets, sc, rays, N1, N2, jit, u, density_threshold=THR[0], bkgd_density_threshold=THR[1])
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append((rays.shape[0], dt))
    n = sum(a for a, _ in times); t = sum(b for _, b in times)
    cpu_reference_rate.last = (rays, jit, u, out)      # the parity leg of main() re-renders these rays on the GPU
    return n / t, t / len(times) * 1e3, rays.shape[0], data, threads


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--precision", default="exact", choices=["exact", "mixed", "fp32", "fast"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample-rays", type=int, default=2048)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="taekwondo2", choices=list(WORKLOADS))
    args = ap.parse_args()
    select_workload(args.workload)
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": WORKLOAD_NAME,
              "step": "one %dx%d view = %d rays (view = step mod %d)" % (W, H, H * W, VIEWS), "layers": LAYERS + 1, "n1": N1, "n2": N2,
              "parallelism": "rows interleaved over %d GPU(s) + 1 all-gather of fine image planes per view" % world,
              "l2": "no explicit flush: per-chunk working set (~1.2 GB of samples/raw/σ buffers) >> 126 MB L2"}

    if args.impl == "reference":
        if rank != 0:
            return
        rate, ms, nr, data, cores = cpu_reference_rate(args.steps, args.warmup, args.cpu_sample_rays)
        sample = "%d rays of the step's view (4 row bands), full 64+128 path, torch fp32 on %d of %d host threads (fastest of a probe)" % (nr, cores, os.cpu_count() or 1)
        print(json.dumps({"impl": "reference", "metric": "rays/sec", "value": rate, "unit": "rays/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                          "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic scene; " + data,
                          "config": config,
                          "cpu_baseline": {"value": rate, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
                          "e2e": {"value": rate, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import stnerf_b200 as S
    from stnerf_b200.dist import ShardedViewRenderer
    from stnerf_b200 import _lib as L
    from stnerf_b200.config import make_cfg
    import modeling

    sd, data = load_weights()
    bkgd, frames, cams = scene_setup()
    model = modeling.build_layered_model(make_cfg(LAYERS, N1, N2, SPACE_TIME, args.precision))
    model.load_state_dict(sd)
    model.set_bkgd_bbox(bkgd); model.set_bboxes(frames)
    nat = model._ensure_native(dev)
    model.near = NEAR
    nat.set_scene(model._resolve_scene(torch.tensor(FRAME_IDS), THR[0], THR[1]))   # the demo's thresholds
    svr = ShardedViewRenderer(nat, H, W, N1, N2, rank, world)
    rays_dev = [svr.rays_for(K, T, FRAME_IDS) for (K, T) in cams]               # inputs resident in HBM
    n_local = rays_dev[0].shape[0]
    rays_per_step = H * W

    def step(i):
        return svr.render(rays_dev[i % VIEWS], seed=i + 1)

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = ClockSampler(local_rank); clocks.start()
    launches0 = S.launch_count()
    nat.profile_begin()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms_total = e0.elapsed_time(e1)
    prof = nat.profile_end()
    launches = S.launch_count() - launches0
    clk = clocks.stop()
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        lt = torch.tensor([float(launches)], device=dev); dist.all_reduce(lt); launches = int(lt.item())
    ms_total = float(t.item())
    value = rays_per_step * args.steps / (ms_total * 1e-3)

    # ---- e2e: host buffers through the C-ABI (H2D rays + D2H every plane inside the timed region) ----------
    e2e = None
    # the same views as the first steps of the timed region (cost depends on how many rays hit the performers)
    e2e_views = [(args.warmup + i) % VIEWS for i in range(max(1, args.e2e_steps))]
    rays_host = [rays_dev[v].cpu().pin_memory() for v in e2e_views]
    out_host = torch.empty((2, LAYERS + 2, 5 * n_local), dtype=torch.float32).pin_memory()
    mask_host = torch.empty((LAYERS + 1, n_local), dtype=torch.uint8).pin_memory()
    nat.render_host(rays_host[0], N1, N2, seed=99, out_host=out_host, mask_host=mask_host)   # warm staging buffers
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(args.e2e_steps):
        nat.render_host(rays_host[i], N1, N2, seed=100 + i, out_host=out_host, mask_host=mask_host)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    te = torch.tensor([dt], device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = {"value": rays_per_step * args.e2e_steps / float(te.item()), "unit": "rays/s",
           "h2d_bytes_per_step": int(rays_host[0].numel() * 4 * world),
           "d2h_bytes_per_step": int((out_host.numel() * 4 + mask_host.numel()) * world),
           "api": "stnerf_render_host (C-ABI, pinned host buffers), %d steps" % args.e2e_steps}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (SpaceNet MLP) from live CUDA-event timings ----------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    sp = prof["spacenet"]
    # algorithmic FLOPs: 2*MAC per evaluated point (taekwondo performer nets have the wider rgb head);
    # the split between background and performer points comes from the per-launch point counts
    pts = sp["points"]
    bk_pts = float(n_local) * (N1 + N1 + N2) * args.steps
    flops = bk_pts * FLOP_SPACE_BKGD + max(0.0, pts - bk_pts) * FLOP_SPACE_PERF
    ach = flops / (sp["ms"] * 1e-3) / 1e12 if sp["ms"] > 0 else 0.0
    # executed MMA terms per algorithmic product: 3 (exact); mixed runs rgb_net.1 (256x128 of the 462 336 MAC/point) in one pass
    terms = {"exact": 3.0, "mixed": 3.0 - 2.0 * (256 * 128) / 462336.0}.get(args.precision, 1.0)
    traffic, traffic_note = None, None
    try:      # DRAM bytes per point of the same kernel from the committed `ncu --set full` capture, scaled to the mean launch
        cap = json.load(open(os.path.join(ROOT, "profiles", "r01_spacenet_traffic.json")))
        traffic = cap["dram_bytes_per_point"] * pts / max(1, sp["launches"])
        traffic_note = "dram__bytes_read+write per point (%.1f B, ncu --set full: %s) x mean points per launch" % (
            cap["dram_bytes_per_point"], cap["kernel"])
    except Exception:
        pass
    roof = {"kernel": "spacenet MLP (%s)" % args.precision, "bound": "tensor", "achieved": ach, "peak": peak_tf,
            "unit": "TFLOP/s", "frac": ach / peak_tf, "executed": ach * terms, "frac_executed": ach * terms / peak_tf,
            "peak_source": ("measured bf16_tflops_sustained (MEASURED_PEAKS.json)" if peaks else "fallback 1400 (B200_PROFILING.md)"),
            "traffic": traffic, "traffic_note": traffic_note, "launches": sp["launches"], "avg_launch_ms": sp["ms"] / max(1, sp["launches"]),
            "share_of_step": sp["ms"] / ms_total,
            "note": "achieved/frac = algorithmic FLOPs (2*MAC/point x points evaluated); exact mode executes 3 fp16 MMAs per product "
                    "(frac is capped at 1/3), executed/frac_executed = what the tensor pipe runs",
            "other_kernels_ms": {k: v["ms"] for k, v in prof.items() if k != "spacenet"}}
    # compositing kernel against the HBM roofline (algorithmic bytes: l*S*20 B in per ray-pass + outputs)
    cp = prof["composite"]
    hit_frac = max(0.0, pts - bk_pts) / max(1.0, bk_pts)
    bytes_comp = float(n_local) * args.steps * ((1 + hit_frac) * (N1 * 24 + (N1 + N2) * 20 + (N1 + N2) * 4) + (LAYERS + 2) * 40)
    roof["composite_hbm"] = {"achieved_GBps": bytes_comp / (cp["ms"] * 1e-3) / 1e9 if cp["ms"] > 0 else 0.0,
                             "peak_GBps": peaks.get("hbm_gbs", 6650.0)}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        rate, _, nr, _, cores = cpu_reference_rate(3, 1, args.cpu_sample_rays)
        cpu = {"value": rate, "unit": "rays/s", "cores": cores, "kind": "port",
               "sample": "3 steps x %d rays (4 row bands of the view), full 64+128 path, torch fp32, %d of %d host threads (fastest of a probe)" % (nr, cores, os.cpu_count() or 1)}

    parity = None
    if cpu is not None and getattr(cpu_reference_rate, "last", None) is not None:
        # same rays / weights / uniforms through the GPU path vs the CPU oracle (itself pinned to the reference)
        rays_c, jit_c, u_c, want = cpu_reference_rate.last
        nat.set_ray_ids(0, 0, 0)
        outp, _ = nat.render(rays_c.to(dev), N1, N2, jitter=jit_c.to(dev).contiguous(), u=u_c.to(dev).contiguous(), seed=1)
        fm, _, _, _ = S.split_planes(outp, LAYERS + 1)
        got = fm[0].float().cpu()
        ref = want["fine_mixed"][0]
        err = (got - ref).abs().max(dim=1)[0]
        mse = float(((got - ref) ** 2).mean())
        parity = {"rays": int(rays_c.shape[0]), "max_abs_rgb_err": float(err.max()),
                  "frac_pixels_over_1e-3": float((err > 1e-3).float().mean()),
                  "psnr_db": (99.0 if mse == 0 else float(10.0 * __import__("math").log10(1.0 / mse))),
                  "against": "CPU oracle port (pinned to the reference by tests/golden), identical rays/weights/uniforms"}

    dtype = {"exact": "f32 via fp16x3 split products (tcgen05), f32 accumulate", "fp32": "f32", "fast": "f16 products, f32 accumulate",
             "mixed": "f32 via fp16x3 split products on the density path, single f16 pass on the colour-only layer rgb_net.1 (tcgen05), f32 accumulate"}[args.precision]
    print(json.dumps({"metric": "rays/sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
                      "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                      "scaling": "strong", "vs_baseline": None, "dtype": dtype,
                      "data": "synthetic scene + cameras (SURVEY 8d); " + data, "config": config, "clocks": clk,
                      "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
                      "parity": parity, "precision": args.precision}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
