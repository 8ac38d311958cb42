#!/bin/bash
# full GPU check of the round: whole -m gpu suite, smoke(), the default bench line, the reference arm
mkdir -p gpurun_out
echo "== gpu suite"; python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== bench (2048^3)"; python bench.py --steps 3 --warmup 3 > gpurun_out/bench2048.json 2> gpurun_out/bench2048.err; tail -c 300 gpurun_out/bench2048.err; python - <<'P'
import json
d=json.loads(open("gpurun_out/bench2048.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["config"]["stage_ms_per_step"], d.get("parity_check"), "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "ccl frac", d["roofline"]["stage_ccl"].get("frac"))
P
echo "== reference arm"; python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 200 gpurun_out/bench_ref.err; cut -c1-900 gpurun_out/bench_ref.json
