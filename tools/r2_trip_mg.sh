#!/bin/bash
# Round 2, 2-GPU trip: sharded CCL in C (device link + replicated solve), N-rank parity, weak / strong scaling lines
set -u
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "== 1. multi-GPU test"
timeout 600 python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -5
echo "== 2. weak scaling, 2 ranks, 1024^3 per GPU, with the N-rank check"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus 2 --size 1024 --steps 2 --warmup 3 --check --e2e-steps 1 > gpurun_out/bench_mg_weak.json 2> gpurun_out/bench_mg_weak.err
tail -c 900 gpurun_out/bench_mg_weak.json; tail -3 gpurun_out/bench_mg_weak.err
echo "== 3. strong scaling, 2 ranks, ONE 1024^3 volume"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 \
  bench.py --gpus 2 --size 1024 --steps 2 --warmup 3 --scaling strong --no-e2e --no-cpu > gpurun_out/bench_mg_strong.json 2> gpurun_out/bench_mg_strong.err
tail -c 600 gpurun_out/bench_mg_strong.json; tail -3 gpurun_out/bench_mg_strong.err
echo "== 4. config c3 on 2 GPUs"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
  bench.py --gpus 2 --config c3 --steps 5 --warmup 3 2>&1 | tail -1 | cut -c1-1000
