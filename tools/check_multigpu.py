"""torchrun --nproc-per-node N tools/check_multigpu.py : parity of the sharded CCL
(NCCL all-gather of boundary planes) against a whole-volume oracle CCL."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from igneous_b200 import _shim, pipeline, multigpu
from oracle import oracle as O
ctx = _shim.Context(local)
shape = (96, 80, 40)
group = multigpu.Group(ctx, rank, world, dist)
pipe = pipeline.VolumePipeline(ctx, shape, np.uint32, pitch=32, num_ids=6, offset=(0, 0, rank * shape[2]),
                               group=group, simplification_factor=0, mesh_shape=(32, 32, 32))
pipe.synth()
pipe.ccl()
got = ctx.to_host(pipe.d_cc, shape, np.uint32)
whole = O.synth_seg((shape[0], shape[1], shape[2] * world), pitch=32, num_ids=6)
want, n_want = O.connected_components(whole, return_N=True)
ok = (pipe.n_components == n_want) and np.array_equal(got, want[:, :, rank * shape[2]:(rank + 1) * shape[2]].astype(np.uint32))
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
  print("MULTIGPU_CCL_PARITY", "OK" if int(flag.item()) == 1 else "FAIL", "components", pipe.n_components, "ranks", world)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
