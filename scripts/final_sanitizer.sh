# compute-sanitizer on the shipped build: synccheck on smoke(), memcheck on the two-ray / hidden-layer / fusion / flow-reuse edge tests
mkdir -p gpurun_out
timeout 100 compute-sanitizer --tool synccheck python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_atmem_synccheck.log 2>&1; echo synccheck rc=$?
tail -2 gpurun_out/sanitizer_atmem_synccheck.log
timeout 140 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_edge.py -q -m gpu -k "two_rays or all_performers_hidden or coarse_fusion or flow_reuse" > gpurun_out/sanitizer_atmem_memcheck_edge.log 2>&1; echo memcheck rc=$?
tail -3 gpurun_out/sanitizer_atmem_memcheck_edge.log
