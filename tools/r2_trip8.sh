#!/bin/bash
# Round 2, GPU trip 8: simplifier v4 (per-round salt, dense E passes) + full suite + bench
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== 1. mesh tests"
timeout 900 python -m pytest tests/test_mesh_gpu.py -x -q 2>&1 | tail -6
echo "== 2. one 257^3 MeshTask body (ms)"
timeout 300 python tools/time_simplify.py 100 4 2>&1 | tail -2
IGN_SIMP_GMEM=1 timeout 300 python tools/time_simplify.py 100 2 2>&1 | tail -1
echo "== 3. full GPU suite"
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== 4. simplifier v4 full capture"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_simp_labels -c 1 \
  -o gpurun_out/r02_simp_labels_v4_full python tools/time_simplify.py 100 1 > /dev/null 2>&1
python tools/ncu_summary.py full gpurun_out/r02_simp_labels_v4_full.ncu-rep 2>&1 | tail -18
echo "== 5. racecheck / memcheck on the simplifier"
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --log-file gpurun_out/r02_racecheck_simplify_v4.log \
  python -m pytest tests/test_mesh_gpu.py -x -q -k "simplify_multilabel or shared_and_global" 2>&1 | tail -3
tail -6 gpurun_out/r02_racecheck_simplify_v4.log
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_memcheck_simplify_v4.log \
  python -m pytest tests/test_mesh_gpu.py -x -q -k "simplify" 2>&1 | tail -3
tail -4 gpurun_out/r02_memcheck_simplify_v4.log
echo "== 6. bench 2048 (headline)"
timeout 1500 python bench.py --steps 3 --warmup 3 --e2e-steps 1 > gpurun_out/bench2048.json 2> gpurun_out/bench2048.err; tail -c 600 gpurun_out/bench2048.json; tail -3 gpurun_out/bench2048.err
