"""The harness that runs the UNMODIFIED reference out of process (oracle/run_reference.py; used by the gpu parity-at-scale test
and by bench.py's CPU leg on the GPU box) against a committed golden case: plain run split over two workers, the `variants`
mechanism (density scaling by exactly 1, its own fine depths fed back, a seeded weight perturbation), and the `multi` timing mode."""
import numpy as np
import pytest
import torch

import cases as C
from oracle import stash_reference

pytestmark = pytest.mark.skipif(stash_reference.reference_root() is None, reason="no reference checkout or archive")


def _job(**extra):
    name = "tkd_edit_frac"
    case = C.CASES[name]
    jit, u = C.uniforms_for(case)
    return name, C.reference_job(case, C.rays_for(case), jit, u, **extra)


def test_runner_reproduces_the_golden_case_over_two_workers():
    name, job = _job()
    res = C.run_reference_job(job, workers=2, threads=4)
    gold = C.load_golden(name)
    assert set(res["flat"]) == set(gold)
    for k in gold:
        assert np.array_equal(res["flat"][k], gold[k]), k
    assert res["rays"] == job["rays"].shape[0] and res["seconds"] > 0


def test_variants_and_multi():
    name, job = _job(record=True)
    base = C.run_reference_job(job, threads=4)
    z = torch.from_numpy(base["record"]["z"].copy())
    assert tuple(z.shape) == (3, job["rays"].shape[0], job["n2"]) and np.isfinite(base["record"]["denom"]).all()
    _, job_v = _job(variants=[dict(sigma_scale=1.0), dict(z_override=z), dict(perturb_seed=3, perturb_rel=1e-3), dict(perturb_seed=3, perturb_rel=1e-3)])
    var = C.run_reference_job(job_v, threads=4)["variants"]
    for k in base["flat"]:
        assert np.array_equal(var[0]["flat"][k], base["flat"][k]), k        # scaling every density by exactly 1
        assert np.array_equal(var[1]["flat"][k], base["flat"][k]), k        # its own fine depths fed back
        assert np.array_equal(var[2]["flat"][k], var[3]["flat"][k]), k      # the perturbation is seeded
    assert np.abs(var[2]["flat"]["fine_mixed.rgb"] - base["flat"]["fine_mixed.rgb"]).max() > 0      # ... and does perturb
    assert np.array_equal(var[2]["flat"]["coarse_mixed.rgb"], base["flat"]["coarse_mixed.rgb"])    # only the resampling input
    _, job_m = _job()
    job_m["multi"] = [dict(rays=job_m["rays"][:64], jitter=job_m["jitter"][:, :64], u=job_m["u"][:, :64]),
                      dict(rays=job_m["rays"][64:], jitter=job_m["jitter"][:, 64:], u=job_m["u"][:, 64:])]
    res = C.run_reference_job(job_m, workers=2, threads=4)
    assert res["rays_each"] == [64, job_m["rays"].shape[0] - 64] and all(s > 0 for s in res["seconds_each"])
