#!/bin/bash
# Round 2, GPU trip 2: run/bitmask CCL (TMA mask kernel) first contact; simplifier flakiness under the sanitizer
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== 1. CCL tests"
timeout 900 python -m pytest tests/test_ccl_gpu.py -x -q 2>&1 | tail -25
echo "== 2. CCL timing 512^3 / 1024^3"
timeout 300 python tools/microbench_ccl.py 2>&1 | tail -12
echo "== 3. simplify tests repeated (flakiness?)"
for i in 1 2 3; do timeout 300 python -m pytest tests/test_mesh_gpu.py -q -k "simplify_multilabel" 2>&1 | tail -1; done
echo "== 4. simplify under memcheck, verbose"
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/memcheck2.log \
  python -m pytest tests/test_mesh_gpu.py -x -q -k "simplify_multilabel and 1000000000" 2>&1 | tail -40
tail -5 gpurun_out/memcheck2.log
echo "== 5. rest of the GPU suite"
timeout 1800 python -m pytest tests -m gpu -x -q --deselect tests/test_ccl_gpu.py 2>&1 | tail -8
