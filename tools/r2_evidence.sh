#!/bin/bash
# round-2 evidence run (one GPU): sanitizers on the simplifier, ncu --set full of k_simp_labels, ncu launch
# list of the bench command at a size ncu can replay in minutes; summaries go to gpurun_out/ (copy to profiles/)
mkdir -p gpurun_out
echo "== racecheck (simplify tests)"
timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_mesh_gpu.py -x -q -k "simplify_multilabel or memory_classes" > gpurun_out/r02_racecheck_simplify_v8.log 2>&1; tail -4 gpurun_out/r02_racecheck_simplify_v8.log
echo "== memcheck (simplify tests)"
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_mesh_gpu.py -x -q -k "simplify_multilabel or memory_classes" > gpurun_out/r02_memcheck_simplify_v8.log 2>&1; tail -4 gpurun_out/r02_memcheck_simplify_v8.log
echo "== synccheck (simplify tests)"
timeout 900 compute-sanitizer --tool synccheck python -m pytest tests/test_mesh_gpu.py -x -q -k "simplify_multilabel" > gpurun_out/r02_synccheck_simplify_v8.log 2>&1; tail -4 gpurun_out/r02_synccheck_simplify_v8.log
echo "== ncu --set full k_simp_labels"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_simp_labels -c 1 -o gpurun_out/r02_simp_labels_v8_full python tools/time_simplify.py 100 1 > /dev/null 2>&1
python tools/ncu_summary.py full gpurun_out/r02_simp_labels_v8_full.ncu-rep > gpurun_out/r02_simp_labels_v8_full_summary.txt 2>&1; tail -17 gpurun_out/r02_simp_labels_v8_full_summary.txt
echo "== ncu launch list of bench.py --size 1024 --steps 1 --warmup 1"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_bench1024_launches.csv python bench.py --size 1024 --steps 1 --warmup 1 --no-parity-check > gpurun_out/bench1024_under_ncu.log 2>&1
python tools/ncu_summary.py launches gpurun_out/r02_bench1024_launches.csv > gpurun_out/r02_bench1024_kernel_summary.txt 2>&1; head -24 gpurun_out/r02_bench1024_kernel_summary.txt
