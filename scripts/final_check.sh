# round-end check of the shipped build on one B200: full gpu test suite, default bench line, ncu capture of the MLP launches,
# launch list, sanitizer on the smoke render
set -x
mkdir -p gpurun_out
T=${1:-final}
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$T.log 2>&1; echo pytest rc=$?
tail -2 gpurun_out/pytest_$T.log
( time timeout 300 python bench.py > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err ) 2> gpurun_out/bench_$T.time; echo bench rc=$?
tail -c 600 gpurun_out/bench_$T.json | head -c 300; echo
timeout 240 ncu --set full --clock-control none --import-source on -k regex:mlp_tc_kernel -s 10 -c 6 -f -o gpurun_out/prof_spacenet_$T python scripts/profile_render.py --rays 65536 --calls 1 > gpurun_out/ncu_$T.log 2>&1; echo ncu rc=$?
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$T.csv python scripts/profile_render.py --rays 65536 --calls 2 > gpurun_out/launches_$T.log 2>&1; echo launches rc=$?
timeout 200 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_${T}_memcheck.log 2>&1; echo memcheck rc=$?
tail -3 gpurun_out/sanitizer_${T}_memcheck.log
