"""world_size-2 gloo test (CPU) of the multi-GPU CCL host logic: plane exchange
by all_gather, boundary linkage, global union-find and table composition must
reproduce a single whole-volume CCL of the stacked dataset."""
import os
import socket

import numpy as np
import pytest


def _free_port():
  s = socket.socket()
  s.bind(("127.0.0.1", 0))
  p = s.getsockname()[1]
  s.close()
  return p


def _worker(rank, world, port, ret):
  import torch
  import torch.distributed as dist
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from igneous_b200 import multigpu
    from oracle import oracle as O
    shape = (48, 40, 24)
    slabs = [O.synth_seg(shape, pitch=16, num_ids=5, offset=(0, 0, r * shape[2])) for r in range(world)]
    whole = np.concatenate(slabs, axis=2)
    # step 1 (stands in for ign_ccl6_volume_begin_dev): local CCL of my slab
    local, n_local = O.connected_components(slabs[rank], return_N=True)
    npl = shape[0] * shape[1]
    rec = np.zeros(4 + 2 * npl * 2, dtype=np.int64)  # [n_local | first v | last v | first l | last l]
    rec[0] = n_local
    rec[4:4 + npl] = slabs[rank][:, :, 0].ravel(order="F")
    rec[4 + npl:4 + 2 * npl] = slabs[rank][:, :, -1].ravel(order="F")
    rec[4 + 2 * npl:4 + 3 * npl] = local[:, :, 0].ravel(order="F")
    rec[4 + 3 * npl:4 + 4 * npl] = local[:, :, -1].ravel(order="F")
    # step 2: the single collective
    gathered = [torch.zeros(len(rec), dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(rec))
    g = [t.numpy() for t in gathered]
    n_locals = [int(x[0]) for x in g]

    def link(r, off_lo, off_hi):
      a, b = g[r], g[r + 1]
      return multigpu.link_planes_numpy(a[4 + npl:4 + 2 * npl], a[4 + 3 * npl:4 + 4 * npl], off_lo,
                                        b[4:4 + npl], b[4 + 2 * npl:4 + 3 * npl], off_hi)
    # step 3: identical solve on every rank, then relabel my slab
    offs, lut, n_global = multigpu.solve_global(n_locals, link, multigpu.solve_pairs)
    final = np.where(local == 0, 0, lut[(local + offs[rank]).astype(np.int64)]).astype(np.uint64)
    want, n_want = O.connected_components(whole, return_N=True)
    ok = (n_global == n_want) and np.array_equal(final, want[:, :, rank * shape[2]:(rank + 1) * shape[2]])
    ret[rank] = bool(ok)
  finally:
    dist.destroy_process_group()


def test_two_rank_ccl_merge_matches_whole_volume():
  import torch.multiprocessing as mp
  world = 2
  port = _free_port()
  with mp.Manager() as mgr:
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_plane_record_layout_and_numpy_link():
  from igneous_b200 import multigpu
  assert multigpu.plane_record_bytes(100) == 256 + 1600 + 800
  va = np.array([0, 5, 5, 7], dtype=np.uint64)
  vb = np.array([0, 5, 6, 7], dtype=np.uint64)
  pairs = multigpu.link_planes_numpy(va, [0, 1, 1, 2], 10, vb, [0, 3, 4, 3], 20)
  assert pairs.tolist() == [[11, 23], [12, 23]]
  offs, lut, n = multigpu.solve_global([2, 4], lambda r, a, b: pairs - 10 + a if False else np.array([[1, 5], [2, 5]]),
                                       multigpu.solve_pairs)
  assert n == 4 and lut[1] == lut[2] == lut[5] == 1
