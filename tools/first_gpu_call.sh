#!/bin/bash
# One GPU call that answers the open questions left at the end of round 1 (run from the repo
# root under gpurun; everything lands in gpurun_out/):
#   1. does the whole GPU suite pass with the opt-in kernels as defaults?
#   2. parity + timing of the CCL tile kernel variants (fast / v2 / half-height tiles)
#   3. launch list of one MeshTask body with the batched simplifier kernels
#   4. ncu --set full of the current v2 CCL tile kernel and of the batched collapse kernel
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== 1. GPU suite with IGN_CCL_V2=1 IGN_SIMP_BATCH=1"
IGN_CCL_V2=1 IGN_SIMP_BATCH=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== 2. CCL tile kernel variants"
timeout 120 python tools/check_ccl_v2.py 1,3 2>&1 | tail -30
echo "== 3. launch list of one MeshTask body (batched simplifier kernels)"
IGN_SIMP_BATCH=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/mesh_batch_launches.csv python tools/time_simplify.py 100 1 > gpurun_out/mesh_batch.log 2>&1
python tools/ncu_summary.py launches gpurun_out/mesh_batch_launches.csv | head -25
echo "== 4. full captures"
IGN_CCL_V2=1 timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_ccl_local_v2 -s 1 -c 1 \
  -o gpurun_out/ccl_local_v2b_full python tools/profile_ccl.py 512 > /dev/null 2>&1
python tools/ncu_summary.py full gpurun_out/ccl_local_v2b_full.ncu-rep
IGN_SIMP_BATCH=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_simp_collapse_b -s 10 -c 1 \
  -o gpurun_out/simp_collapse_b_full python tools/time_simplify.py 100 1 > /dev/null 2>&1
python tools/ncu_summary.py full gpurun_out/simp_collapse_b_full.ncu-rep
