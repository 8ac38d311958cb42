"""igneous/shards.py:10-55 -- the voxel shape of one image shard.

With the identity hash the low `preshift_bits + minishard_bits` bits of a chunk's
compressed Morton code stay inside one shard, so a shard is a box of
2^bx x 2^by x 2^bz chunks where the bits are dealt x, y, z, x, ... and a
dimension stops receiving bits once it spans the whole chunk grid."""
import math

import numpy as np

from ._compat import Vec


def image_shard_shape_from_spec(spec, dataset_size, chunk_size):
  chunk_size = [int(c) for c in chunk_size][:3]
  dataset_size = [int(d) for d in dataset_size][:3]
  shape_bits = int(spec["preshift_bits"]) + int(spec["minishard_bits"])
  if shape_bits >= 64:
    raise ValueError("preshift_bits (%d) + minishard_bits (%d) must be < 64. Sum: %d"
                     % (int(spec["preshift_bits"]), int(spec["minishard_bits"]), shape_bits))
  grid = [int(math.ceil(d / c)) for d, c in zip(dataset_size, chunk_size)]
  bits = [0, 0, 0]
  dealt = 0
  while dealt < shape_bits:
    progressed = False
    for dim in range(3):
      if dealt < shape_bits and (1 << bits[dim]) < grid[dim]:
        bits[dim] += 1
        dealt += 1
        progressed = True
    if not progressed:
      break
  return Vec(*[c << b for c, b in zip(chunk_size, bits)], dtype=np.uint64)
