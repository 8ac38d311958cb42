#!/bin/bash
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== 1. simplifier v5"
timeout 300 python tools/time_simplify.py 100 4 2>&1 | tail -1 | cut -c1-130
timeout 600 python -m pytest tests/test_mesh_gpu.py -x -q 2>&1 | tail -3
echo "== 2. cseg task test"
timeout 300 python -m pytest tests/test_tasks_gpu.py -x -q -k compressed_segmentation 2>&1 | tail -30
echo "== 3. per-config bench lines"
timeout 300 python bench.py --config c2 --steps 5 --warmup 3 2>&1 | tail -2 | cut -c1-900
timeout 300 python bench.py --config c3 --steps 5 --warmup 3 2>&1 | tail -2 | cut -c1-1200
timeout 300 python bench.py --config c1 --steps 5 --warmup 3 2>&1 | tail -2 | cut -c1-900
echo "== 4. k_ccl_tiles full capture"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_ccl_tiles -c 1 -o gpurun_out/r02_k_ccl_tiles_v2_full \
    python tools/profile_ccl.py 1024 uint32 uint32 1 > /dev/null 2>&1
python tools/ncu_summary.py full gpurun_out/r02_k_ccl_tiles_v2_full.ncu-rep 2>&1 | tail -17
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_ccl_masks -c 1 -o gpurun_out/r02_k_ccl_masks_v2_full \
    python tools/profile_ccl.py 1024 uint32 uint32 1 > /dev/null 2>&1
python tools/ncu_summary.py full gpurun_out/r02_k_ccl_masks_v2_full.ncu-rep 2>&1 | tail -17
