# A/B recipe (variants: python -c "import __graft_entry__ as g; g.build_variant('hf', ['SPACE_A_TMEM=0']); g.build_variant('atmem4', ['SPACE_RING=4'])")
# A/B of the tensor-memory activation path (SPACE_A_TMEM): bit comparison against the shared-memory build, then alternating benches
set -x
mkdir -p gpurun_out
V=$PWD/st-nerf_b200/stnerf_b200
STNERF_B200_LIB=$V/variant_hf.so timeout 100 python scripts/dump_networks.py gpurun_out/dump_hf.npz > gpurun_out/dump_hf.log 2>&1; echo dump_hf rc=$?
timeout 100 python scripts/dump_networks.py gpurun_out/dump_atmem8.npz > gpurun_out/dump_atmem8.log 2>&1; echo dump_atmem8 rc=$?
tail -3 gpurun_out/dump_atmem8.log
STNERF_B200_LIB=$V/variant_atmem4.so timeout 100 python scripts/dump_networks.py gpurun_out/dump_atmem4.npz > gpurun_out/dump_atmem4.log 2>&1; echo dump_atmem4 rc=$?
python scripts/dump_networks.py --compare gpurun_out/dump_hf.npz gpurun_out/dump_atmem8.npz > gpurun_out/cmp_atmem8.log 2>&1; OK8=$?
python scripts/dump_networks.py --compare gpurun_out/dump_hf.npz gpurun_out/dump_atmem4.npz > gpurun_out/cmp_atmem4.log 2>&1; OK4=$?
cat gpurun_out/cmp_atmem8.log; tail -1 gpurun_out/cmp_atmem4.log
rm -f gpurun_out/dump_*.npz
N=2; if [ $OK8 -ne 0 ]; then N=1; fi
for i in $(seq 1 $N); do
  for v in hf atmem4 main; do
    if [ $v = main ]; then L=$V/libstnerf_b200.so; else L=$V/variant_$v.so; fi
    STNERF_B200_LIB=$L timeout 120 python bench.py --steps 4 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_at_$v$i.json 2> gpurun_out/bench_at_$v$i.err
  done
done
python - <<'PY'
import json
for i in (1,2):
  for n in ["hf","atmem4","main"]:
    try:
        d=json.loads(open("gpurun_out/bench_at_%s%d.json"%(n,i)).read().strip().splitlines()[-1])
        print(n, i, round(d["value"]), round(d["ms_per_step"],1), round(d["roofline"]["frac"],4), round(d["roofline"]["avg_launch_ms"],3), d["clocks"]["sm_mhz"], d["clocks"]["power_w_median"], d.get("parity",{}).get("pixels_over_1e-3"))
    except Exception as e: print(n, i, "ERR", e)
PY
