#!/usr/bin/env python
"""Golden vectors of the UNMODIFIED reference at scale (run in the build container, where /root/reference exists):

    python tests/golden/make_golden_scale.py [case ...]

For every case of `cases.SCALE_CASES` (16 384 rays of a 1080p taekwondo / 4-layer walking view, 4 096 rays of a 4K 6-layer
64+192 view) the reference's `LayeredRFRender.forward` runs on CPU through `oracle/run_reference.py` with injected uniforms
and its FINE images (mixed + per layer: rgb, depth, acc) and hit masks are stored as `tests/golden/<case>.npz`.
The inputs are regenerated from seeds (`cases.scale_inputs`), only reference OUTPUTS are stored.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases as C  # noqa: E402


def main(names):
    for name in names:
        case = C.SCALE_CASES[name]
        rays, jit, u = C.scale_inputs(case)
        t0 = time.time()
        res = C.run_reference_job(C.reference_job(case, rays, jit, u), workers=2, threads=os.cpu_count() or 1)
        flat = {k: v for k, v in res["flat"].items() if k.startswith(C.SCALE_KEYS_STORED)}
        np.savez_compressed(C.scale_golden_path(name), **flat)
        hits = [int(flat["ray_mask.%d" % i].sum()) for i in range(case["L"] + 1)]
        print("%-18s %6d rays  %.1fs (reference forward %.1fs)  hits=%s  %.2f MB" % (
            name, rays.shape[0], time.time() - t0, res["seconds"], hits, os.path.getsize(C.scale_golden_path(name)) / 1e6))


if __name__ == "__main__":
    main(sys.argv[1:] or list(C.SCALE_CASES))
