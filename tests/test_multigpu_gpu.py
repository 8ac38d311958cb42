"""Multi-GPU CCL parity through NCCL (needs >= 2 GPUs; on a 1-GPU box the
world_size-2 logic is covered by tests/test_multigpu_cpu.py over gloo)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_ccl_over_nccl_matches_whole_volume(ctx):
  from igneous_b200 import _shim
  n = _shim.device_count()
  if n < 2:
    pytest.skip("needs >= 2 GPUs (single-GPU box): covered by the gloo test")
  world = 2
  out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", "29571",
                        os.path.join(ROOT, "tools", "check_multigpu.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
  assert "MULTIGPU_CCL_PARITY OK" in out.stdout, out.stdout[-3000:]
