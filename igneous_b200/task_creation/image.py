"""create_downsampling_tasks, create_image_shard_downsample_tasks and the three CCL
task creators (igneous/task_creation/image.py:170-345, 639-770, 1726-1889): same
signatures, same info / provenance side effects, tasks from igneous_b200.tasks."""
import copy
import math
from functools import partial, reduce
from time import strftime

import numpy as np

from .. import downsample_scales, fastremap, sharding, shards
from .._compat import CloudVolume, CloudFiles, InfoUnavailableError, Vec
from ..tasks import DownsampleTask, ImageShardDownsampleTask, CCLFacesTask, CCLEquivalancesTask, RelabelCCLTask
from ..types import DownsampleMethods
from .common import FinelyDividedTaskIterator, get_bounds, operator_contact

MEMORY_TARGET = int(3.5e9)


def num_mips_from_memory_target(memory_target, dtype, chunk_size, num_channels, factor):
  voxels = memory_target / np.dtype(dtype).itemsize / num_channels
  chunks = voxels // reduce(lambda a, b: a * b, chunk_size)
  total = reduce(lambda a, b: a * b, factor)
  chunks /= total / (total - 1)  # the pyramid on top of mip 0 costs 1/(f-1) more
  side = chunks ** (1.0 / math.log2(total))
  if side <= 0:
    return 1
  n = math.log2(side)
  if math.ceil(n) - n <= 0.01:
    n = int(math.ceil(n))
  return max(1, int(n))


def _log(vol, task_name, **fields):
  vol.provenance.processing.append({"method": dict(task=task_name, **fields), "by": operator_contact(),
                                    "date": strftime("%Y-%m-%d %H:%M %Z")})
  vol.commit_provenance()


def create_downsampling_tasks(layer_path, mip=0, fill_missing=False, axis="z", num_mips=None,
                              preserve_chunk_size=True, sparse=False, bounds=None, chunk_size=None,
                              encoding=None, delete_black_uploads=False, background_color=0,
                              dest_path=None, compress=None, factor=None, bounds_mip=0,
                              memory_target=MEMORY_TARGET, encoding_level=None, encoding_effort=None,
                              method=DownsampleMethods.AUTO):
  vol = CloudVolume(layer_path, mip=mip)

  def task_shape(at_mip):
    nonlocal num_mips
    shape = Vec(*chunk_size) if chunk_size else Vec(*vol.meta.chunk_size(at_mip)[:3])
    f = factor if factor is not None else downsample_scales.axis_to_factor(axis)
    viable = num_mips_from_memory_target(memory_target, vol.dtype, shape, vol.num_channels, f)
    if num_mips is None:
      num_mips = viable
    if viable < num_mips:
      print("WARNING: memory limit (%d bytes) too low for %d mips at a time; %d possible."
            % (memory_target, num_mips, viable))
    return Vec(*[int(s) * int(ff) ** viable for s, ff in zip(shape, f)])

  shape = task_shape(mip)
  vol = downsample_scales.create_downsample_scales(
    layer_path, mip, shape, preserve_chunk_size=preserve_chunk_size, chunk_size=chunk_size,
    encoding=encoding, factor=factor, max_mips=num_mips)
  if encoding is not None:
    for m in range(mip + 1, min(mip + num_mips, len(vol.available_mips))):
      vol.scales[m]["encoding"] = encoding
    vol.commit_info()
  if not preserve_chunk_size or chunk_size:
    shape = task_shape(mip + 1)
  vol.mip = mip
  roi = get_bounds(vol, bounds, mip, bounds_mip=bounds_mip, chunk_size=vol.meta.chunk_size(mip))

  class DownsampleTaskIterator(FinelyDividedTaskIterator):
    def task(self, shape, offset):
      return partial(DownsampleTask, layer_path=layer_path, mip=mip, shape=shape.clone(),
                     offset=offset.clone(), axis=axis, fill_missing=fill_missing, sparse=sparse,
                     delete_black_uploads=delete_black_uploads, background_color=background_color,
                     dest_path=dest_path, compress=compress, factor=factor, max_mips=num_mips, method=method)

    def on_finish(self):
      _log(vol, "DownsampleTask", mip=mip, num_mips=num_mips, shape=[int(s) for s in shape], axis=axis,
           sparse=sparse, bounds=str(roi), chunk_size=(list(chunk_size) if chunk_size else None),
           preserve_chunk_size=preserve_chunk_size, encoding=encoding, fill_missing=bool(fill_missing),
           delete_black_uploads=bool(delete_black_uploads), background_color=background_color,
           dest_path=dest_path, compress=compress, factor=(tuple(factor) if factor else None),
           downsample_method=int(method))

  return DownsampleTaskIterator(roi, shape)


def set_encoding(cv, mip, encoding, encoding_level, encoding_effort):
  """task_creation/common.py:215-236 (the lossy-codec quality keys are recorded but those
  codecs are outside this implementation)."""
  scale = cv.scales[mip]
  if encoding is not None:
    scale["encoding"] = encoding
    if encoding == "compressed_segmentation" and "compressed_segmentation_block_size" not in scale:
      scale["compressed_segmentation_block_size"] = (8, 8, 8)
  if encoding_level is None:
    return
  key = {"jpeg": "jpeg_quality", "jxl": "jxl_quality", "png": "png_level", "fpzip": "fpzip_precision"}.get(encoding)
  if key:
    scale[key] = int(encoding_level)
  if encoding == "jxl" and encoding_effort is not None:
    scale["jxl_effort"] = int(encoding_effort)


def create_image_shard_downsample_tasks(cloudpath, mip=0, fill_missing=False, sparse=False, chunk_size=None,
                                        encoding=None, memory_target=MEMORY_TARGET, agglomerate=False,
                                        timestamp=None, factor=(2, 2, 1), bounds=None, bounds_mip=0,
                                        encoding_level=None, encoding_effort=None,
                                        method=DownsampleMethods.AUTO, num_mips=None, truncate_scales=True):
  """Downsample an (un)sharded layer into SHARDED scales mip+1 .. mip+num_mips
  (task_creation/image.py:639-770).  One task covers the footprint of one shard of
  mip+1 scaled up by factor^num_mips."""
  if num_mips is None:
    num_mips = 3
  cv = CloudVolume(cloudpath)
  if truncate_scales:
    cv.info["scales"] = cv.info["scales"][:mip + 1]
    cv.commit_info()
  cv = downsample_scales.add_scales(cloudpath, mip, num_mips, preserve_chunk_size=True, chunk_size=chunk_size,
                                    encoding=encoding, factor=factor)
  for i in range(1, num_mips + 1):
    scale = cv.scales[mip + i]
    scale["sharding"] = sharding.create_sharded_image_info(
      dataset_size=scale["size"], chunk_size=scale["chunk_sizes"][0], encoding=scale["encoding"],
      dtype=cv.dtype, uncompressed_shard_bytesize=int(memory_target))
  cv.mip = mip
  for i in range(num_mips):
    set_encoding(cv, mip + i + 1, encoding, encoding_level, encoding_effort)
  if num_mips > 1:  # keep the top level lossless so that further levels can be built on it
    if encoding == "jxl":
      set_encoding(cv, mip + num_mips, encoding, 100, encoding_effort)
    elif encoding == "jpeg":
      set_encoding(cv, mip + num_mips, "png", 9, encoding_effort)
  cv.commit_info()
  base_shape = shards.image_shard_shape_from_spec(cv.info["scales"][mip + 1]["sharding"],
                                                  cv.meta.volume_size(mip + 1), cv.meta.chunk_size(mip + 1))
  shape = Vec(*[int(b) * int(f) ** num_mips for b, f in zip(base_shape, factor)])
  cv.mip = mip
  roi = get_bounds(cv, bounds, mip, bounds_mip=bounds_mip, chunk_size=cv.meta.chunk_size(mip + 1))

  class ImageShardDownsampleTaskIterator(FinelyDividedTaskIterator):
    def task(self, shape, offset):
      return partial(ImageShardDownsampleTask, cloudpath, shape=tuple(int(v) for v in shape),
                     offset=tuple(int(v) for v in offset), mip=int(mip), fill_missing=bool(fill_missing),
                     sparse=bool(sparse), agglomerate=bool(agglomerate), timestamp=timestamp,
                     factor=tuple(factor), method=method, num_mips=int(num_mips))

    def on_finish(self):
      cv.provenance.sources = [cloudpath]
      _log(cv, "ImageShardDownsampleTask", cloudpath=cloudpath, shape=[int(v) for v in shape],
           fill_missing=fill_missing, sparse=bool(sparse), bounds=[roi.minpt.tolist(), roi.maxpt.tolist()],
           mip=mip, agglomerate=agglomerate, timestamp=timestamp, method=int(method),
           encoding_level=encoding_level, encoding_effort=encoding_effort, num_mips=int(num_mips))

  return ImageShardDownsampleTaskIterator(roi, shape)


def _ccl_creator(task_fn, task_name, cloudpath, mip, shape, **opts):
  vol = CloudVolume(cloudpath, mip=mip)
  shape = Vec(*shape)

  class CCLTaskIterator(FinelyDividedTaskIterator):
    def task(self, shape, offset):
      return partial(task_fn, cloudpath=cloudpath, mip=mip, shape=shape.clone(), offset=offset.clone(), **opts)

    def on_finish(self):
      _log(vol, task_name, cloudpath=cloudpath, mip=mip, shape=[int(s) for s in shape], **opts)

  return CCLTaskIterator(vol.meta.bounds(mip).clone(), shape)


def create_ccl_face_tasks(cloudpath, mip, shape=(512, 512, 512), threshold_gte=None, threshold_lte=None,
                          fill_missing=False, dust_threshold=0):
  """pass 1"""
  return _ccl_creator(CCLFacesTask, "CCLFacesTask", cloudpath, mip, shape, threshold_gte=threshold_gte,
                      threshold_lte=threshold_lte, fill_missing=fill_missing, dust_threshold=dust_threshold)


def create_ccl_equivalence_tasks(cloudpath, mip, shape=(512, 512, 512), threshold_gte=None,
                                 threshold_lte=None, fill_missing=False, dust_threshold=0):
  """pass 2 (shape must match pass 1)"""
  return _ccl_creator(CCLEquivalancesTask, "CCLEquivalancesTask", cloudpath, mip, shape,
                      threshold_gte=threshold_gte, threshold_lte=threshold_lte, fill_missing=fill_missing,
                      dust_threshold=dust_threshold)


def create_ccl_relabel_tasks(src_path, dest_path, mip, shape=(512, 512, 512), chunk_size=None, encoding=None,
                             threshold_gte=None, threshold_lte=None, fill_missing=False, dust_threshold=0):
  """pass 4: the destination layer gets the smallest dtype that holds max_label"""
  src = CloudVolume(src_path, mip=mip)
  cf = CloudFiles(src_path)
  max_label = int(cf.get_json(cf.join(src.key, "ccl", "max_label.json"))[0])
  dtype = fastremap.fit_dtype(np.uint64, max_label).name
  try:
    dest = CloudVolume(dest_path, mip=mip)
  except InfoUnavailableError:
    info = copy.deepcopy(src.info)
    info["data_type"] = dtype
    info["type"] = "segmentation"
    info["scales"] = info["scales"][:mip + 1]
    scale = info["scales"][mip]
    if chunk_size:
      scale["chunk_sizes"] = [list(chunk_size)]
    if encoding:
      scale["encoding"] = encoding
    scale.pop("sharding", None)
    dest = CloudVolume(dest_path, info=info, mip=mip)
    dest.commit_info()
  shape = Vec(*shape)

  class RelabelCCLTaskIterator(FinelyDividedTaskIterator):
    def task(self, shape, offset):
      return partial(RelabelCCLTask, src_path=src_path, dest_path=dest_path, mip=mip, shape=shape.clone(),
                     offset=offset.clone(), threshold_gte=threshold_gte, threshold_lte=threshold_lte,
                     fill_missing=fill_missing, dust_threshold=dust_threshold)

    def on_finish(self):
      _log(dest, "RelabelCCLTask", src_path=src_path, dest_path=dest_path, mip=mip,
           shape=[int(s) for s in shape], threshold_gte=threshold_gte, threshold_lte=threshold_lte,
           fill_missing=bool(fill_missing), dust_threshold=dust_threshold)

  return RelabelCCLTaskIterator(src.meta.bounds(mip).clone(), shape)
