"""DownsampleTask / TransferTask with the pooling done on the GPU.

Same names, arguments and side effects as igneous/tasks/image/image.py:
  downsample_method_to_fn :37-55, downsample_and_upload :57-100,
  TransferTask :434-516, DownsampleTask :518-549.
  ImageShardDownsampleTask :672-843.
Only the library behind `fn(image, factors[0], num_mips=...)` (:91) changes:
igneous_b200.tinybrain instead of the CPU tinybrain wheel.
"""
import math
from collections import defaultdict
from functools import partial

import numpy as np

from .. import downsample_scales, fastremap, sharding, shards, tinybrain
from .._compat import CloudVolume, CloudFiles, Bbox, Vec, min2, queueable
from ..types import DownsampleMethods


def downsample_method_to_fn(method, sparse, vol):
  if method == DownsampleMethods.AUTO:
    method = {"image": DownsampleMethods.AVERAGE_POOLING,
              "segmentation": DownsampleMethods.MODE_POOLING}.get(vol.layer_type, DownsampleMethods.STRIDING)
  if method == DownsampleMethods.AVERAGE_POOLING:
    return partial(tinybrain.downsample_with_averaging, sparse=sparse)
  if method == DownsampleMethods.MODE_POOLING:
    return partial(tinybrain.downsample_segmentation, sparse=sparse)
  if method == DownsampleMethods.MIN_POOLING:
    return tinybrain.downsample_with_min_pooling
  if method == DownsampleMethods.MAX_POOLING:
    return tinybrain.downsample_with_max_pooling
  return tinybrain.downsample_with_striding


def downsample_and_upload(image, bounds, vol, ds_shape, mip=0, axis="z", skip_first=False,
                          sparse=False, factor=None, max_mips=None, method=DownsampleMethods.AUTO):
  ds_shape = min2(vol.meta.volume_size(mip), Vec(*ds_shape[:3]))
  underlying = (mip + 1) if (mip + 1) in vol.available_mips else mip
  chunk = np.asarray(vol.meta.chunk_size(underlying), dtype=np.float32)
  if factor is None:
    factor = downsample_scales.axis_to_factor(axis)
  factors = downsample_scales.compute_factors(ds_shape, factor, chunk, vol.meta.volume_size(mip))
  if max_mips is not None:
    factors = factors[:max_mips]
  vol.mip = mip
  if not skip_first:
    vol[bounds] = image
  if not factors:
    return
  fn = downsample_method_to_fn(method, sparse, vol)
  mips = fn(image, factors[0], num_mips=len(factors))  # <- the kernel call (image.py:91)
  box = bounds.clone()
  for f3, mipped in zip(factors, mips):
    vol.mip += 1
    box //= f3
    box.maxpt = box.minpt + Vec(*mipped.shape[:3])
    vol[box] = mipped


@queueable
def TransferTask(src_path, dest_path, mip, shape, offset, translate=(0, 0, 0), fill_missing=False,
                 skip_first=False, skip_downsamples=False, delete_black_uploads=False,
                 background_color=0, sparse=False, axis="z", agglomerate=False, timestamp=None,
                 compress="gzip", factor=None, max_mips=None, stop_layer=None,
                 downsample_method=DownsampleMethods.AUTO, use_https_for_source=False):
  shape, offset, translate = Vec(*shape), Vec(*offset), Vec(*translate)
  src = CloudVolume(src_path, fill_missing=bool(fill_missing), mip=mip, bounded=False)
  dest = CloudVolume(dest_path, fill_missing=bool(fill_missing), mip=mip,
                     delete_black_uploads=bool(delete_black_uploads),
                     background_color=background_color, compress=compress)
  dst_box = Bbox.clamp(Bbox(offset, shape + offset), dest.meta.bounds(mip))
  image = src.download(dst_box - translate)
  if skip_downsamples:
    dest[dst_box] = image
    return
  downsample_and_upload(image, dst_box, dest, shape, mip=mip, skip_first=bool(skip_first),
                        sparse=bool(sparse), axis=axis, factor=factor, max_mips=max_mips,
                        method=downsample_method)


@queueable
def DownsampleTask(layer_path, mip, shape, offset, fill_missing=False, axis="z", sparse=False,
                   delete_black_uploads=False, background_color=0, dest_path=None, compress="gzip",
                   factor=None, max_mips=None, method=DownsampleMethods.AUTO):
  """2x2x1 (by default) downsample pyramid of one cutout.  As in the reference
  the `method` argument is accepted but AUTO is what runs (image.py:524,548)."""
  return TransferTask(layer_path, dest_path or layer_path, mip, shape, offset, translate=(0, 0, 0),
                      fill_missing=fill_missing, skip_first=True, skip_downsamples=False,
                      delete_black_uploads=delete_black_uploads, background_color=background_color,
                      sparse=sparse, axis=axis, compress=compress, factor=factor, max_mips=max_mips,
                      downsample_method=DownsampleMethods.AUTO)


@queueable
def ImageShardDownsampleTask(src_path, shape, offset, mip=0, fill_missing=False, sparse=False,
                             agglomerate=False, timestamp=None, factor=(2, 2, 1),
                             method=DownsampleMethods.AUTO, num_mips=1, progress=False):
  """Downsample the region under one (stack of) output shard(s) and write whole shard files
  for mips mip+1 .. mip+num_mips (image.py:672-843).

  As in the reference the region is walked in z-layers one chunk thick at the coarsest
  mip; segmentation layers are renumbered to a small dtype before pooling and mapped back
  afterwards (renumber / pooling / remap run on the GPU).  Chunks are collected per output
  shard -- keyed by the shard number their chunk id hashes to, which for the identity hash
  is the reference's (shard_x, shard_y, shard_z) box -- and every shard is written once."""
  shape, offset = Vec(*shape), Vec(*offset)
  mip, num_mips = int(mip), int(num_mips)
  factor = tuple(int(f) for f in factor)
  src = CloudVolume(src_path, fill_missing=bool(fill_missing), mip=mip, bounded=False)
  chunk_size = src.meta.chunk_size(mip)
  bbox = Bbox.clamp(Bbox(offset, offset + shape), src.meta.bounds(mip))
  bbox = bbox.expand_to_chunk_size(chunk_size, offset=src.meta.voxel_offset(mip))

  def shard_shape_at(m):
    return shards.image_shard_shape_from_spec(src.scales[m]["sharding"], src.meta.volume_size(m),
                                              src.meta.chunk_size(m))

  first = shard_shape_at(mip + 1)
  upper = offset // Vec(*factor)
  upper_box = Bbox.clamp(Bbox(upper, upper + Vec(*[int(v) for v in first])), src.meta.bounds(mip + 1))
  if upper_box.subvoxel():
    return
  fn = downsample_method_to_fn(method, sparse, src)
  renumber = src.layer_type == "segmentation"
  cz = int(chunk_size[2]) * factor[2] ** num_mips
  nz = int(math.ceil(int(bbox.size3()[2]) / cz))
  f3 = np.asarray(factor, dtype=int)
  pending = [defaultdict(dict) for _ in range(num_mips)]  # per mip: shard grid position -> {chunk id: bytes}
  zbox = bbox.clone()
  zbox.maxpt[2] = zbox.minpt[2] + cz
  for _ in range(nz):
    if renumber:
      img, mapping = src.download(zbox, agglomerate=agglomerate, timestamp=timestamp, renumber=True)
      back = {int(new): int(old) for old, new in mapping.items()}
      back[src.background_color] = src.background_color
    else:
      img = src.download(zbox, agglomerate=agglomerate, timestamp=timestamp)
    mips = fn(img, factor, num_mips=num_mips)  # <- the kernel call (image.py:764)
    del img
    for i in range(num_mips):
      m = mip + i + 1
      cutout = mips[i]
      if renumber:
        cutout = fastremap.remap(cutout.astype(src.dtype), back, preserve_missing_labels=False)
      lo = np.asarray(zbox.minpt, dtype=int) // (f3 ** (i + 1))
      box = Bbox(lo, lo + np.asarray(cutout.shape[:3], dtype=int))
      bounds_m = src.meta.bounds(m)
      if box.minpt[2] >= bounds_m.maxpt[2]:
        continue
      sshape = np.asarray(shard_shape_at(m), dtype=int)
      origin = np.asarray(src.meta.voxel_offset(m), dtype=int)
      g0 = (np.asarray(box.minpt) - origin) // sshape
      g1 = -((-(np.asarray(box.maxpt) - origin)) // sshape)
      for gz in range(int(g0[2]), int(g1[2])):
        for gy in range(int(g0[1]), int(g1[1])):
          for gx in range(int(g0[0]), int(g1[0])):
            smin = origin + np.asarray([gx, gy, gz]) * sshape
            part = Bbox.intersection(Bbox(smin, smin + sshape), box)
            if part.subvoxel():
              continue
            sl = tuple(slice(int(a - o), int(b - o)) for a, b, o in zip(part.minpt, part.maxpt, box.minpt))
            pending[i][(gx, gy, gz)].update(_shard_chunks(src, cutout[sl], part, m))
    del mips
    zbox.minpt[2] += cz
    zbox.maxpt[2] += cz
  for i in range(num_mips):
    m = mip + i + 1
    sshape = np.asarray(shard_shape_at(m), dtype=int)
    origin = np.asarray(src.meta.voxel_offset(m), dtype=int)
    base = src.meta.join(src.cloudpath, src.meta.key(m))
    done = Bbox(np.asarray(bbox.minpt, dtype=int) // (f3 ** (i + 1)), -((-np.asarray(bbox.maxpt, dtype=int)) // (f3 ** (i + 1))))
    for grid_pos, chunk_dict in pending[i].items():
      smin = origin + np.asarray(grid_pos) * sshape
      shard_box = Bbox(smin, smin + sshape)
      inside = Bbox.clamp(shard_box, src.meta.bounds(m))
      if not (np.all(inside.minpt >= done.minpt) and np.all(inside.maxpt <= done.maxpt)):
        # a coarser shard that is taller than this task (the reference would let the last task win):
        # keep the chunks other tasks already stored in it
        spec = sharding.ShardingSpecification(src.scales[m]["sharding"])
        name = spec.shard_filename(spec.locate(next(iter(chunk_dict)))[0])
        old = CloudFiles(base).get(name)
        if old is not None:
          merged = {cid: spec.read_chunk(old, cid) for cid in spec.chunk_ids(old)}
          merged.update(chunk_dict)
          chunk_dict = merged
      filename, blob = src.image.make_shard(chunk_dict, shard_box, m, progress=False)
      CloudFiles(base).put(filename, blob, compress=None)
    pending[i] = None


def _shard_chunks(vol, cutout, box, mip):
  """make_shard_chunks on a cutout whose far edges may be short of a chunk boundary: pad with the
  background colour (image.py:797-799) and let the chunker clamp to the dataset bounds."""
  cs = np.asarray(vol.meta.chunk_size(mip), dtype=int)
  off = np.asarray(vol.meta.voxel_offset(mip), dtype=int)
  want = np.ceil((np.asarray(box.maxpt) - off) / cs).astype(int) * cs + off
  pad = [(0, int(max(w - h, 0))) for w, h in zip(want, box.maxpt)]
  if any(p[1] for p in pad):
    if cutout.ndim == 4:
      pad = pad + [(0, 0)]
    cutout = np.pad(cutout, pad, mode="constant", constant_values=vol.background_color)
    box = Bbox(box.minpt, np.asarray(box.minpt) + np.asarray(cutout.shape[:3]))
  return vol.image.make_shard_chunks(cutout, box, mip)
