#!/bin/bash
# full GPU check of the round: whole -m gpu suite, then the default bench line and the reference arm
mkdir -p gpurun_out
echo "== gpu suite"; python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== bench (2048^3)"; python bench.py --steps 3 --warmup 3 > gpurun_out/bench2048.json 2> gpurun_out/bench2048.err; tail -c 600 gpurun_out/bench2048.err; python - <<'P'
import json
d=json.loads(open("gpurun_out/bench2048.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["config"]["stage_ms_per_step"], d.get("parity_check"), d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["stage_ccl"]["frac"] if "frac" in d["roofline"]["stage_ccl"] else None)
P
