#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== A. box test, incremental keys"
timeout 300 python -m pytest tests/test_mesh_gpu.py -x -q -k "simplify_box" 2>&1 | tail -3
echo "== B. box test, full keys every round"
IGN_SIMP_FULLKEYS=1 timeout 300 python -m pytest tests/test_mesh_gpu.py -x -q -k "simplify" 2>&1 | tail -3
IGN_SIMP_FULLKEYS=1 timeout 300 python tools/time_simplify.py 100 3 2>&1 | tail -1
echo "== C. cseg tests"
timeout 300 python -m pytest tests/test_cseg_gpu.py -x -q 2>&1 | tail -12
echo "== D. golden test"
timeout 300 python -m pytest tests/test_golden.py -x -q 2>&1 | tail -12
