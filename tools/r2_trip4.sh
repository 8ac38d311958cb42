#!/bin/bash
# Round 2, GPU trip 4: simplifier v2 (counter race fixed) + CCL kernels v2 (dense tile unions, 512-thread masks, ILP expand)
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== 1. mesh + ccl tests"
timeout 900 python -m pytest tests/test_mesh_gpu.py tests/test_ccl_gpu.py -x -q 2>&1 | tail -8
echo "== 2. one 257^3 MeshTask body (ms)"
timeout 300 python tools/time_simplify.py 100 4 2>&1 | tail -2
IGN_SIMP_GMEM=1 timeout 300 python tools/time_simplify.py 100 2 2>&1 | tail -1
echo "== 3. CCL timings"
timeout 300 python tools/microbench_ccl.py 2>&1 | tail -13
echo "== 4. CCL launch list 1024^3 u32->u32"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active \
  --clock-control none --csv --log-file gpurun_out/r02_ccl1024_launches_v2.csv python tools/profile_ccl.py 1024 uint32 uint32 1 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_ccl1024_launches_v2.csv", errors="replace")) if len(r) > 10]
h = rows[0]; ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
d = collections.OrderedDict()
for r in rows[1:]:
  d.setdefault((r[ii], r[ki].split("(")[0][-40:]), {})[r[mi]] = r[vi]
for (i, k), m in d.items():
  print(i, k, " ".join("%s=%s" % (a.split("__")[-1][:18], b) for a, b in m.items()))
PY
echo "== 5. simplifier v2 full capture"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_simp_labels -c 1 \
  -o gpurun_out/r02_simp_labels_v2_full python tools/time_simplify.py 100 1 > /dev/null 2>&1
python tools/ncu_summary.py full gpurun_out/r02_simp_labels_v2_full.ncu-rep 2>&1 | tail -18
echo "== 6. full GPU suite"
timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
echo "== 7. bench 1024 quick"
timeout 900 python bench.py --size 1024 --steps 2 --warmup 3 --e2e-steps 1 2>&1 | tail -3
