#!/bin/bash
# Round 2, GPU trip 3: simplifier v2 (alive lists, dense cost queue) + per-kernel view of the run/bitmask CCL
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== 1. mesh tests"
timeout 900 python -m pytest tests/test_mesh_gpu.py -x -q 2>&1 | tail -8
echo "== 2. one 257^3 MeshTask body (ms)"
timeout 300 python tools/time_simplify.py 100 4 2>&1 | tail -2
IGN_SIMP_GMEM=1 timeout 300 python tools/time_simplify.py 100 2 2>&1 | tail -1
echo "== 3. CCL launch list 1024^3 u32->u32"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct \
  --clock-control none --csv --log-file gpurun_out/r02_ccl1024_launches.csv python tools/profile_ccl.py 1024 uint32 uint32 1 > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/r02_ccl1024_launches.csv", errors="replace")) if len(r) > 10]
h = rows[0]; ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
d = collections.OrderedDict()
for r in rows[1:]:
  d.setdefault((r[ii], r[ki].split("(")[0][-60:]), {})[r[mi]] = r[vi]
for (i, k), m in d.items():
  print(i, k, " ".join("%s=%s" % (a.split("__")[-1][:22], b) for a, b in m.items()))
PY
echo "== 4. full captures"
for k in k_ccl_masks k_ccl_tiles k_ccl_expand4 k_ccl_merge; do
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:$k -c 1 -o gpurun_out/r02_${k}_full \
    python tools/profile_ccl.py 1024 uint32 uint32 1 > /dev/null 2>&1
  python tools/ncu_summary.py full gpurun_out/r02_${k}_full.ncu-rep 2>&1 | tail -18
done
echo "== 5. simplifier v2 full capture"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_simp_labels -c 1 \
  -o gpurun_out/r02_simp_labels_v2_full python tools/time_simplify.py 100 1 > /dev/null 2>&1
python tools/ncu_summary.py full gpurun_out/r02_simp_labels_v2_full.ncu-rep 2>&1 | tail -18
echo "== 6. simplify under memcheck / racecheck"
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/memcheck3.log \
  python -m pytest tests/test_mesh_gpu.py -x -q -k "simplify" 2>&1 | tail -15
tail -3 gpurun_out/memcheck3.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --log-file gpurun_out/racecheck3.log \
  python -m pytest tests/test_mesh_gpu.py -x -q -k "simplify_multilabel" 2>&1 | tail -5
tail -12 gpurun_out/racecheck3.log
