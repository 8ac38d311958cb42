"""DownsampleTask / TransferTask with the pooling done on the GPU.

Same names, arguments and side effects as igneous/tasks/image/image.py:
  downsample_method_to_fn :37-55, downsample_and_upload :57-100,
  TransferTask :434-516, DownsampleTask :518-549.
Only the library behind `fn(image, factors[0], num_mips=...)` (:91) changes:
igneous_b200.tinybrain instead of the CPU tinybrain wheel.
"""
from functools import partial

import numpy as np

from .. import downsample_scales, tinybrain
from .._compat import CloudVolume, Bbox, Vec, min2, queueable
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
