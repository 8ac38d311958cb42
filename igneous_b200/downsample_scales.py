"""Host-side scale arithmetic of the downsample path (pure Python, no kernels).

Mirrors the functions of igneous/downsample_scales.py that sit on the hot path:
  compute_factors        :135-172   how many mips one task produces
  axis_to_factor         :174-182
  compute_scales         :184-212
  create_downsample_scales :214-244 adds the new scales to the info file
"""
import copy
import math

import numpy as np

from ._compat import CloudVolume, Vec, min2


def axis_to_factor(axis):
  table = {"x": (1, 2, 2), "y": (2, 1, 2), "z": (2, 2, 1)}
  if axis not in table:
    raise ValueError("Axis not supported: " + str(axis))
  return table[axis]


def compute_factors(ds_shape, factor, chunk_size, volume_size):
  """[factor] * N where N is the number of downsamples the least tolerant
  pooled axis allows for this task shape (float32 arithmetic and the +1e-4
  guard as in the reference; a partial last level is allowed only when the
  whole volume is already smaller than a chunk)."""
  pooled = [i for i, f in enumerate(factor) if f != 1]
  if not pooled:
    return []
  grid = np.array([ds_shape[i] for i in pooled], dtype=np.float32) / \
      np.array([chunk_size[i] for i in pooled], dtype=np.float32)
  fdiv = np.array([factor[i] for i in pooled], dtype=np.float32)
  eps = 0.0001
  n_float = float(np.min(np.log(grid) / np.log(fdiv) + np.float32(eps)))
  if n_float < eps:
    return []
  dsvol = np.array(volume_size, dtype=np.float64) / (np.array(factor, dtype=np.float64) ** int(math.ceil(n_float)))
  small = all(dsvol[i] < chunk_size[i] for i in pooled)
  n = int(n_float)
  if small and (n_float - n) > 0.05:
    n += 1
  return [tuple(factor)] * n


def _precision(x):
  s = repr(float(x))
  return len(s.split(".")[1].rstrip("0")) if "." in s else 0


def compute_scales(vol, mip, shape, axis, factor, chunk_size=None):
  shape = min2(vol.meta.volume_size(mip), shape)
  underlying = (mip + 1) if (mip + 1) in vol.available_mips else mip
  cs = np.asarray(chunk_size, dtype=np.float32) if chunk_size else \
      np.asarray(vol.meta.chunk_size(underlying), dtype=np.float32)
  if factor is None:
    factor = axis_to_factor(axis)
  factors = compute_factors(shape, factor, cs, vol.meta.volume_size(mip))
  base = [float(r) for r in vol.meta.resolution(mip)]
  prec = max(_precision(r) for r in base)
  scales, cur = [], base
  for f in factors:
    cur = [c * ff for c, ff in zip(cur, f)]
    scales.append([int(c) if prec == 0 else round(c, prec) for c in cur])
  return scales


def create_downsample_scales(layer_path, mip, ds_shape, axis="z", preserve_chunk_size=False,
                             chunk_size=None, encoding=None, factor=None, max_mips=None):
  vol = CloudVolume(layer_path, mip)
  resolutions = compute_scales(vol, mip, ds_shape, axis, factor, chunk_size)
  if max_mips is not None:
    resolutions = resolutions[:max_mips]
  if not resolutions:
    print("WARNING: No scales generated.")
  for res in resolutions:
    vol.meta.add_resolution(res, encoding=encoding, chunk_size=chunk_size)
  if chunk_size is None:
    src = mip if (preserve_chunk_size or not resolutions) else mip + 1
    new_cs = vol.scales[src]["chunk_sizes"]
  else:
    new_cs = [list(chunk_size)]
  for i in range(mip + 1, mip + len(resolutions) + 1):
    vol.scales[i]["chunk_sizes"] = new_cs
  vol.commit_info()
  return vol


def add_scales(layer_path, mip, num_mips, preserve_chunk_size=True, chunk_size=None, encoding=None, factor=None):
  """igneous/downsample_scales.py:246-278: append exactly `num_mips` scales above `mip`
  (no memory-driven truncation; used by the sharded downsample creator)."""
  vol = CloudVolume(layer_path, mip=mip)
  if factor is None:
    factor = (2, 2, 1)
  for _ in range(num_mips):
    res = [r * f for r, f in zip(vol.meta.resolution(mip), factor)]
    vol.meta.add_resolution(res, encoding=encoding, chunk_size=chunk_size)
    if chunk_size is None:
      new_cs = vol.scales[mip if preserve_chunk_size else mip + 1]["chunk_sizes"]
    else:
      new_cs = [list(chunk_size)]
    if encoding is None:
      encoding = vol.scales[mip]["encoding"]
    vol.scales[mip + 1]["chunk_sizes"] = copy.deepcopy(new_cs)
    mip += 1
    vol.mip = mip
  return vol
