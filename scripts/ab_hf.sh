# A/B recipe of the HI-first hand-off: variant_base.so = the previous commit built with __graft_entry__.build_variant('base', []) in a worktree, copied next to the library
set -x
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_hf.log 2>&1; echo smoke rc=$?
tail -3 gpurun_out/smoke_hf.log
timeout 150 python -m pytest tests/test_gpu_stages.py tests/test_gpu_edge.py -m gpu -x -q > gpurun_out/pytest_hf_quick.log 2>&1; echo quick rc=$?
tail -3 gpurun_out/pytest_hf_quick.log
for i in 1 2; do
  STNERF_B200_LIB=$PWD/st-nerf_b200/stnerf_b200/variant_base.so timeout 120 python bench.py --steps 4 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_hf_base$i.json 2> gpurun_out/bench_hf_base$i.err
  timeout 120 python bench.py --steps 4 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_hf_new$i.json 2> gpurun_out/bench_hf_new$i.err
done
python - <<'PY'
import json
for n in ["base1","new1","base2","new2"]:
    try:
        d=json.loads(open("gpurun_out/bench_hf_%s.json"%n).read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("e2e",{}).get("value"), d.get("clocks"))
    except Exception as e: print(n, "ERR", e)
PY
