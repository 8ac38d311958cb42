#!/bin/bash
# Round 2, GPU trip 5: simplifier v3 (incremental keys, single-pass winners) + CCL race fix + leaner mask / expand kernels
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== 1. mesh + ccl tests"
timeout 900 python -m pytest tests/test_mesh_gpu.py tests/test_ccl_gpu.py -x -q 2>&1 | tail -8
echo "== 2. one 257^3 MeshTask body (ms)"
timeout 300 python tools/time_simplify.py 100 4 2>&1 | tail -2
IGN_SIMP_GMEM=1 timeout 300 python tools/time_simplify.py 100 2 2>&1 | tail -1
echo "== 3. CCL timings (x2: component counts must repeat)"
timeout 300 python tools/microbench_ccl.py 2>&1 | tail -13
timeout 300 python tools/microbench_ccl.py 2>&1 | head -5 | cut -c1-120
echo "== 4. full GPU suite"
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== 5. bench 1024 quick"
timeout 900 python bench.py --size 1024 --steps 2 --warmup 3 --e2e-steps 1 > gpurun_out/bench1024.json 2> gpurun_out/bench1024.err; tail -c 1500 gpurun_out/bench1024.json; tail -3 gpurun_out/bench1024.err
echo "== 6. simplifier v3 full capture"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_simp_labels -c 1 \
  -o gpurun_out/r02_simp_labels_v3_full python tools/time_simplify.py 100 1 > /dev/null 2>&1
python tools/ncu_summary.py full gpurun_out/r02_simp_labels_v3_full.ncu-rep 2>&1 | tail -18
