"""Grid enumeration shared by the task creators
(igneous/task_creation/common.py:11-104)."""
import copy
import os
import subprocess

import numpy as np

from .._compat import Bbox, Vec


def operator_contact():
  try:
    return str(subprocess.check_output("git config user.email", shell=True, stderr=subprocess.DEVNULL).rstrip())
  except Exception:
    return os.environ.get("USER", "")


def get_bounds(vol, bounds, mip, bounds_mip=0, chunk_size=None):
  if bounds is None:
    return vol.meta.bounds(mip)
  bounds = vol.bbox_to_mip(Bbox.create(bounds), mip=bounds_mip, to_mip=mip)
  if chunk_size is not None:
    bounds = bounds.expand_to_chunk_size(chunk_size, vol.meta.voxel_offset(mip))
  return Bbox.clamp(bounds, vol.meta.bounds(mip))


def num_tasks(bounds, shape):
  return int(np.prod(np.ceil(np.asarray(bounds.size3(), dtype=np.float64) / np.asarray(shape))))


class FinelyDividedTaskIterator:
  """Regular grid of non-overlapping tasks, x fastest (common.py:60-104)."""

  def __init__(self, bounds, shape):
    self.bounds = bounds
    self.shape = Vec(*shape)
    self.start = 0
    self.end = num_tasks(bounds, shape)

  def __len__(self):
    return self.end - self.start

  def __getitem__(self, slc):
    itr = copy.deepcopy(self)
    itr.start = max(self.start + slc.start, self.start)
    itr.end = min(self.start + slc.stop, self.end)
    return itr

  def to_coord(self, index):
    gx, gy, _ = np.ceil(np.asarray(self.bounds.size3(), dtype=np.float64) / np.asarray(self.shape)).astype(int)
    z, rem = divmod(index, gx * gy)
    y, x = divmod(rem, gx)
    return Vec(x, y, z)

  def __iter__(self):
    for i in range(self.start, self.end):
      offset = self.to_coord(i) * self.shape + self.bounds.minpt
      yield self.task(self.shape.clone(), Vec(*offset))
    self.on_finish()

  def task(self, shape, offset):
    raise NotImplementedError()

  def on_finish(self):
    pass
