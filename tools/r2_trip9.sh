#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for cfg in "1024 1" "512 2" "256 4" "512 1" "256 2"; do
  set -- $cfg
  echo "== threads $1 ctas $2 (shared-memory class where it fits)"
  IGN_SIMP_THREADS=$1 IGN_SIMP_CTAS=$2 timeout 300 python tools/time_simplify.py 100 3 2>&1 | tail -1 | cut -c1-120
  echo "== threads $1 ctas $2 forced global-memory class"
  IGN_SIMP_GMEM=1 IGN_SIMP_THREADS=$1 IGN_SIMP_CTAS=$2 timeout 300 python tools/time_simplify.py 100 3 2>&1 | tail -1 | cut -c1-120
done
echo "== mesh tests with the default config"
timeout 600 python -m pytest tests/test_mesh_gpu.py tests/test_tasks_gpu.py -x -q 2>&1 | tail -4
echo "== ccl tests + timing (merge with 4 words / thread)"
timeout 600 python -m pytest tests/test_ccl_gpu.py -x -q 2>&1 | tail -3
timeout 300 python tools/microbench_ccl.py 2>&1 | tail -6
