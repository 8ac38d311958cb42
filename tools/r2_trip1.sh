#!/bin/bash
# Round 2, GPU trip 1: the per-label shared-memory simplifier (k_simp_labels) and the v2 CCL default.
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== 1. mesh tests"
timeout 900 python -m pytest tests/test_mesh_gpu.py -x -q 2>&1 | tail -15
echo "== 2. one 257^3 MeshTask body (ms)"
timeout 300 python tools/time_simplify.py 100 4 2>&1 | tail -3
IGN_SIMP_GMEM=1 timeout 300 python tools/time_simplify.py 100 2 2>&1 | tail -1
echo "== 3. full GPU suite"
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
echo "== 4. CCL tile kernel variants"
timeout 200 python tools/check_ccl_v2.py 1,3 2>&1 | tail -30
echo "== 5. launch list of one MeshTask body"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/r02_mesh_task_launches.csv python tools/time_simplify.py 100 1 > gpurun_out/mesh_task.log 2>&1
python tools/ncu_summary.py launches gpurun_out/r02_mesh_task_launches.csv | head -30
echo "== 6. ncu --set full of k_simp_labels"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_simp_labels -c 1 \
  -o gpurun_out/r02_simp_labels_full python tools/time_simplify.py 100 1 > /dev/null 2>&1
python tools/ncu_summary.py full gpurun_out/r02_simp_labels_full.ncu-rep 2>&1 | tail -40
echo "== 7. compute-sanitizer (memcheck, racecheck) on a small simplification"
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_sanitizer_memcheck_simplify.log \
  python -m pytest tests/test_mesh_gpu.py -q -k "simplify_multilabel" 2>&1 | tail -3
tail -5 gpurun_out/r02_sanitizer_memcheck_simplify.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --log-file gpurun_out/r02_sanitizer_racecheck_simplify.log \
  python -m pytest tests/test_mesh_gpu.py -q -k "simplify_multilabel and 100" 2>&1 | tail -3
tail -30 gpurun_out/r02_sanitizer_racecheck_simplify.log
