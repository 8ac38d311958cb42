"""Minimal stand-ins for cloud-volume / cloud-files / task-queue, used ONLY when
those packages cannot be imported (they are absent from the build image, see
SURVEY.md 8(b) H8).  They implement just enough of the Precomputed `file://`
layout for the task layer to run end to end in tests:

  info JSON            {data_type, num_channels, type, scales:[{key, size,
                        resolution, voxel_offset, chunk_sizes, encoding}], mesh}
  chunk files          {key}/{x0}-{x1}_{y0}-{y1}_{z0}-{z1}[.gz]   raw, Fortran order
  mesh / ccl files     plain files under the layer directory

This is not a storage engine and it is never on the compute path: when the
real packages import, `igneous_b200._compat` uses them instead.
"""
import copy
import math
import gzip
import json
import os

import numpy as np


class EmptyVolumeException(Exception):
  pass


class InfoUnavailableError(Exception):
  pass


class OutOfBoundsError(Exception):
  pass


# ------------------------------------------------------------------ geometry
class Vec(np.ndarray):
  def __new__(cls, *args, dtype=int):
    if len(args) == 1 and hasattr(args[0], "__len__"):
      args = tuple(args[0])
    return np.array(args, dtype=dtype).view(cls)

  @property
  def x(self):
    return self[0]

  @x.setter
  def x(self, v):
    self[0] = v

  @property
  def y(self):
    return self[1]

  @y.setter
  def y(self, v):
    self[1] = v

  @property
  def z(self):
    return self[2]

  @z.setter
  def z(self, v):
    self[2] = v

  def clone(self):
    return Vec(*self, dtype=self.dtype)

  def rectVolume(self):
    return int(np.prod(self))


def min2(a, b):
  return Vec(*np.minimum(a, b), dtype=np.asarray(a).dtype)


def max2(a, b):
  return Vec(*np.maximum(a, b), dtype=np.asarray(a).dtype)


class Bbox:
  def __init__(self, a, b, dtype=int):
    a, b = np.asarray(a)[:3], np.asarray(b)[:3]
    self.minpt = Vec(*np.minimum(a, b), dtype=dtype)
    self.maxpt = Vec(*np.maximum(a, b), dtype=dtype)

  @classmethod
  def create(cls, obj):
    if isinstance(obj, Bbox):
      return obj.clone()
    if isinstance(obj, (list, tuple)) and len(obj) == 3 and isinstance(obj[0], slice):
      return cls([s.start for s in obj], [s.stop for s in obj])
    raise TypeError("cannot make a Bbox from %r" % (obj,))

  @classmethod
  def from_filename(cls, name):
    parts = os.path.basename(name).split(".")[0].split("_")
    lo, hi = zip(*[[int(v) for v in p.split("-")] for p in parts[-3:]])
    return cls(lo, hi)

  @classmethod
  def clamp(cls, box, bounds):
    box = box.clone()
    box.minpt = Vec(*np.clip(box.minpt, bounds.minpt, bounds.maxpt), dtype=box.minpt.dtype)
    box.maxpt = Vec(*np.clip(box.maxpt, bounds.minpt, bounds.maxpt), dtype=box.maxpt.dtype)
    return box

  @classmethod
  def intersection(cls, a, b):
    lo = np.maximum(a.minpt, b.minpt)
    hi = np.minimum(a.maxpt, b.maxpt)
    if np.any(hi <= lo):
      return cls((0, 0, 0), (0, 0, 0))
    return cls(lo, hi)

  def clone(self):
    return Bbox(self.minpt, self.maxpt, dtype=self.minpt.dtype)

  def size3(self):
    return Vec(*(self.maxpt - self.minpt), dtype=self.minpt.dtype)

  size = size3

  def volume(self):
    return int(np.prod(self.size3()))

  def subvoxel(self):
    return bool(np.any(self.size3() <= 0))

  empty = subvoxel

  def center(self):
    return (self.minpt + self.maxpt) / 2.0

  def to_slices(self):
    return tuple(slice(int(a), int(b)) for a, b in zip(self.minpt, self.maxpt))

  def to_list(self):
    return [v.item() if hasattr(v, "item") else v for v in list(self.minpt) + list(self.maxpt)]

  def to_filename(self, precision=None):
    def fmt(v):
      if precision:
        return ("%." + str(int(precision)) + "f") % float(v)
      return str(int(v))
    return "_".join("%s-%s" % (fmt(a), fmt(b)) for a, b in zip(self.minpt, self.maxpt))

  def astype(self, dtype):
    return Bbox(self.minpt.astype(dtype), self.maxpt.astype(dtype), dtype=dtype)

  def expand_to_chunk_size(self, chunk_size, offset=(0, 0, 0)):
    cs, off = np.asarray(chunk_size)[:3], np.asarray(offset)[:3]
    lo = np.floor((self.minpt - off) / cs) * cs + off
    hi = np.ceil((self.maxpt - off) / cs) * cs + off
    return Bbox(lo.astype(int), hi.astype(int))

  def __floordiv__(self, f):
    f = np.asarray(f)[:3]
    return Bbox(self.minpt // f, -(-self.maxpt // f))

  def __ifloordiv__(self, f):
    f = np.asarray(f)[:3]
    self.minpt = Vec(*(self.minpt // f), dtype=self.minpt.dtype)
    self.maxpt = Vec(*(-(-self.maxpt // f)), dtype=self.maxpt.dtype)
    return self

  def __mul__(self, f):
    f = np.asarray(f)[:3]
    return Bbox(self.minpt * f, self.maxpt * f, dtype=np.result_type(self.minpt.dtype, f.dtype))

  def __sub__(self, v):
    return Bbox(self.minpt - np.asarray(v)[:3], self.maxpt - np.asarray(v)[:3])

  def __add__(self, v):
    return Bbox(self.minpt + np.asarray(v)[:3], self.maxpt + np.asarray(v)[:3])

  def __eq__(self, o):
    return isinstance(o, Bbox) and np.array_equal(self.minpt, o.minpt) and np.array_equal(self.maxpt, o.maxpt)

  def __repr__(self):
    return "Bbox(%s, %s)" % (list(self.minpt), list(self.maxpt))


# --------------------------------------------------------------------- files
def _strip(path):
  if path.startswith("file://"):
    path = path[len("file://"):]
  elif "://" in path:
    raise NotImplementedError("the storage stand-in only implements file:// (got %s)" % path)
  return path.rstrip("/")


class CloudFiles:
  """file:// subset of cloudfiles.CloudFiles."""
  _EXT = {"gzip": ".gz", "br": ".br", None: "", False: "", "": ""}

  def __init__(self, cloudpath, progress=False, **kwargs):
    self.cloudpath = cloudpath
    self.root = _strip(cloudpath)

  def join(self, *parts):
    return "/".join(str(p).strip("/") for p in parts if str(p) != "")

  def _abs(self, key):
    return os.path.join(self.root, key)

  def put(self, key, content, compress=None, **kwargs):
    path = self._abs(key) + self._EXT.get(compress, "")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    if isinstance(content, str):
      content = content.encode("utf8")
    if compress in ("gzip", "br"):  # the stand-in stores both as gzip streams
      content = gzip.compress(content, compresslevel=1)
    with open(path, "wb") as f:
      f.write(content)

  def puts(self, files, compress=None, **kwargs):
    for item in files:
      if isinstance(item, dict):
        self.put(item["path"], item["content"], compress=item.get("compress", compress))
      else:
        self.put(item[0], item[1], compress=compress)

  def put_json(self, key, obj, compress=None, **kwargs):
    self.put(key, json.dumps(obj), compress=compress)

  def put_jsons(self, files, compress=None, **kwargs):
    for key, obj in files:
      self.put_json(key, obj, compress=compress)

  def _find(self, key):
    for ext in ("", ".gz", ".br"):
      path = self._abs(key) + ext
      if os.path.isfile(path):
        return path, ext
    return None, None

  def exists(self, key):
    return self._find(key)[0] is not None

  def get(self, key, return_dict=False, **kwargs):
    if isinstance(key, (list, tuple)) or return_dict:
      keys = list(key) if isinstance(key, (list, tuple)) else [key]
      return {k: self.get(k) for k in keys}
    path, ext = self._find(key)
    if path is None:
      return None
    with open(path, "rb") as f:
      data = f.read()
    return gzip.decompress(data) if ext else data

  def get_json(self, key):
    if isinstance(key, (list, tuple)):
      return [self.get_json(k) for k in key]
    data = self.get(key)
    return None if data is None else json.loads(data.decode("utf8"))

  def list(self, prefix="", flat=False):
    base = self._abs(prefix)
    top = base if os.path.isdir(base) else os.path.dirname(base)
    out = []
    for dirpath, _, files in os.walk(top):
      for fn in files:
        full = os.path.join(dirpath, fn)
        rel = os.path.relpath(full, self.root)
        for ext in (".gz", ".br"):
          if rel.endswith(ext):
            rel = rel[:-len(ext)]
        if rel.startswith(prefix):
          out.append(rel)
    return sorted(set(out))

  def delete(self, keys):
    if isinstance(keys, str):
      keys = [keys]
    for k in list(keys):
      path, _ = self._find(k)
      if path:
        os.remove(path)


# -------------------------------------------------------------------- volume
class _Provenance:
  def __init__(self):
    self.processing = []
    self.sources = []
    self.owners = []
    self.description = ""


class _SpatialIndex:
  precision = 0


class _MeshMeta:
  spatial_index = _SpatialIndex()


class _Meta:
  """cv.meta: per-mip accessors (cloudvolume's PrecomputedMetadata subset)."""

  def __init__(self, cv):
    self._cv = cv

  def resolution(self, mip):
    return self._cv.resolution_at(mip)

  def chunk_size(self, mip):
    return self._cv.chunk_size_at(mip)

  def volume_size(self, mip):
    return self._cv.volume_size_at(mip)

  def voxel_offset(self, mip):
    return self._cv.voxel_offset_at(mip)

  def bounds(self, mip):
    return self._cv.bounds_at(mip)

  def add_resolution(self, *args, **kwargs):
    return self._cv.add_resolution(*args, **kwargs)

  def join(self, *parts):
    return self._cv.join(*parts)

  def key(self, mip):
    return self._cv.key_at(mip)

  @property
  def cloudpath(self):
    return self._cv.cloudpath

  @property
  def info(self):
    return self._cv.info


class _ImageSource:
  """cv.image: the two shard builders ImageShardDownsampleTask calls
  (igneous/tasks/image/image.py:664-669,818,833)."""

  def __init__(self, cv):
    self._cv = cv

  def make_shard_chunks(self, img, bbox, mip):
    """Cut `img` (occupying `bbox` at `mip`) into the scale's chunks -> {chunk id: encoded bytes}."""
    cv = self._cv
    img = np.asarray(img)
    if img.ndim == 3:
      img = img[..., np.newaxis]
    bbox = Bbox.create(bbox) if not isinstance(bbox, Bbox) else bbox
    spec = cv._sharding(mip)
    if spec is None:
      raise ValueError("mip %d of %s is not sharded" % (mip, cv.cloudpath))
    cs, off = cv.chunk_size_at(mip), cv.voxel_offset_at(mip)
    if np.any((np.asarray(bbox.minpt) - np.asarray(off)) % np.asarray(cs)):
      raise ValueError("shard cutout %r is not chunk aligned" % (bbox,))
    out = {}
    for c in cv._chunks(mip, Bbox.clamp(bbox, cv.bounds_at(mip))):
      src = tuple(slice(int(a - o), int(b - o)) for a, b, o in zip(c.minpt, c.maxpt, bbox.minpt))
      block = np.asfortranarray(img[src].astype(cv.dtype, copy=False))
      if tuple(block.shape[:3]) != tuple(int(v) for v in c.size3()):
        raise ValueError("image %r does not cover chunk %r of %r" % (img.shape, c, bbox))
      out[cv._chunk_id(mip, c)] = cv._encode_chunk(block, mip)
    return out

  def make_shard(self, img, bbox, mip, progress=False):
    """-> (file name, shard bytes).  `img` is an array or a {chunk id: bytes} dict."""
    cv = self._cv
    spec = cv._sharding(mip)
    chunks = img if isinstance(img, dict) else self.make_shard_chunks(img, bbox, mip)
    if not chunks:
      raise ValueError("no chunks inside %r" % (bbox,))
    shard_no = spec.locate(next(iter(chunks)))[0]
    return spec.shard_filename(shard_no), spec.synthesize_shard(chunks)


class CloudVolume:
  """file:// Precomputed subset of cloudvolume.CloudVolume."""

  def __init__(self, cloudpath, mip=0, fill_missing=False, bounded=True, info=None, compress="gzip",
               delete_black_uploads=False, background_color=0, parallel=1, progress=False, **kwargs):
    self.cloudpath = cloudpath
    self.path = _strip(cloudpath)
    self.fill_missing = bool(fill_missing)
    self.bounded = bounded
    self.compress = compress
    self.delete_black_uploads = delete_black_uploads
    self.background_color = background_color
    self.cf = CloudFiles(cloudpath)
    self.provenance = _Provenance()
    self.mesh = _MeshMeta()
    if info is not None:
      self.info = copy.deepcopy(info)
    else:
      self.info = self.cf.get_json("info")
      if self.info is None:
        raise InfoUnavailableError("no info file at " + cloudpath)
    prov = self.cf.get_json("provenance")
    if prov:
      self.provenance.processing = prov.get("processing", [])
    self.mip = mip

  @property
  def meta(self):
    return _Meta(self)

  @property
  def image(self):
    return _ImageSource(self)

  @classmethod
  def create_new_info(cls, num_channels, layer_type, data_type, encoding, resolution, voxel_offset,
                      volume_size, chunk_size=(64, 64, 64), mesh=None, **kwargs):
    res = [int(r) if float(r).is_integer() else float(r) for r in resolution]
    info = {"num_channels": int(num_channels), "type": layer_type, "data_type": str(np.dtype(data_type)),
            "scales": [{"encoding": encoding, "chunk_sizes": [list(map(int, chunk_size))],
                        "key": "_".join(str(r) for r in res), "resolution": res,
                        "voxel_offset": list(map(int, voxel_offset)), "size": list(map(int, volume_size))}]}
    if mesh:
      info["mesh"] = mesh
    return info

  @classmethod
  def from_numpy(cls, arr, vol_path, resolution=(4, 4, 40), voxel_offset=(0, 0, 0), chunk_size=(128, 128, 64),
                 layer_type=None, max_mip=0, encoding="raw", compress=None):
    arr = np.asarray(arr)
    if arr.ndim == 3:
      arr = arr[..., np.newaxis]
    if layer_type is None:
      layer_type = "segmentation" if arr.dtype in (np.uint16, np.uint32, np.uint64) else "image"
    info = cls.create_new_info(arr.shape[3], layer_type, arr.dtype, encoding, resolution, voxel_offset,
                               arr.shape[:3], chunk_size)
    vol = cls(vol_path, info=info, compress=compress)
    vol.commit_info()
    vol[vol.bounds] = arr
    return vol

  # ---- info accessors
  @property
  def scales(self):
    return self.info["scales"]

  @property
  def available_mips(self):
    return list(range(len(self.info["scales"])))

  @property
  def dtype(self):
    return np.dtype(self.info["data_type"])

  data_type = dtype

  @property
  def layer_type(self):
    return self.info["type"]

  @property
  def num_channels(self):
    return int(self.info["num_channels"])

  def resolution_at(self, mip):
    return Vec(*self.info["scales"][mip]["resolution"], dtype=np.float32 if any(
      not float(r).is_integer() for r in self.info["scales"][mip]["resolution"]) else int)

  def chunk_size_at(self, mip):
    return Vec(*self.info["scales"][mip]["chunk_sizes"][0])

  def volume_size_at(self, mip):
    return Vec(*self.info["scales"][mip]["size"])

  def voxel_offset_at(self, mip):
    return Vec(*self.info["scales"][mip]["voxel_offset"])

  def bounds_at(self, mip):
    off = self.voxel_offset_at(mip)
    return Bbox(off, off + self.volume_size_at(mip))

  # current-mip properties, as on cloudvolume.CloudVolume
  resolution = property(lambda self: self.resolution_at(self._mip))
  chunk_size = property(lambda self: self.chunk_size_at(self._mip))
  volume_size = property(lambda self: self.volume_size_at(self._mip))
  voxel_offset = property(lambda self: self.voxel_offset_at(self._mip))
  bounds = property(lambda self: self.bounds_at(self._mip))

  def key_at(self, mip):
    return self.info["scales"][mip]["key"]

  @property
  def key(self):
    return self.key_at(self._mip)

  def join(self, *parts):
    return "/".join(str(p).rstrip("/") for p in parts)

  def mip_bounds(self, mip):
    return self.bounds_at(mip)

  def mip_volume_size(self, mip):
    return self.volume_size_at(mip)

  def bbox_to_mip(self, bbox, mip, to_mip):
    if mip == to_mip:
      return bbox.clone()
    f = np.asarray(self.resolution_at(to_mip), dtype=np.float64) / np.asarray(self.resolution_at(mip), dtype=np.float64)
    lo = np.floor(np.asarray(bbox.minpt) / f).astype(int)
    hi = np.ceil(np.asarray(bbox.maxpt) / f).astype(int)
    return Bbox(lo, hi)

  def add_resolution(self, res, encoding=None, chunk_size=None, info=None):
    base = self.info["scales"][0]
    res = [int(r) if float(r).is_integer() else float(r) for r in res]
    factor = np.asarray(res, dtype=np.float64) / np.asarray(base["resolution"], dtype=np.float64)
    key = "_".join(str(r) for r in res)
    scale = {"encoding": encoding or base["encoding"],
             "chunk_sizes": [list(map(int, chunk_size))] if chunk_size is not None else copy.deepcopy(base["chunk_sizes"]),
             "key": key, "resolution": res,
             "voxel_offset": [int(v) for v in np.floor(np.asarray(base["voxel_offset"]) / factor)],
             "size": [int(v) for v in np.ceil(np.asarray(base["size"]) / factor)]}
    for i, s in enumerate(self.info["scales"]):
      if s["key"] == key:
        self.info["scales"][i] = scale
        return scale
    self.info["scales"].append(scale)
    self.info["scales"].sort(key=lambda s: float(np.prod(s["resolution"])))
    return scale

  # ---- mip state: cv.mip and CloudVolume properties that depend on it
  @property
  def mip(self):
    return self._mip

  @mip.setter
  def mip(self, m):
    self._mip = int(m)

  def commit_info(self):
    self.cf.put_json("info", self.info)

  def refresh_info(self):
    self.info = self.cf.get_json("info")
    return self.info

  def commit_provenance(self):
    self.cf.put_json("provenance", {"processing": self.provenance.processing, "sources": [], "owners": [],
                                    "description": ""})

  # ---- IO
  def _chunk_name(self, mip, box):
    return self.key_at(mip) + "/" + box.to_filename()

  def _chunks(self, mip, box):
    cs, off = self.chunk_size_at(mip), self.voxel_offset_at(mip)
    vb = self.bounds_at(mip)
    grid = box.expand_to_chunk_size(cs, off)
    for z in range(int(grid.minpt[2]), int(grid.maxpt[2]), int(cs[2])):
      for y in range(int(grid.minpt[1]), int(grid.maxpt[1]), int(cs[1])):
        for x in range(int(grid.minpt[0]), int(grid.maxpt[0]), int(cs[0])):
          c = Bbox.clamp(Bbox((x, y, z), (x + cs[0], y + cs[1], z + cs[2])), vb)
          if not c.subvoxel():
            yield c

  # sharded scales (neuroglancer_uint64_sharded_v1, igneous_b200.sharding)
  def _sharding(self, mip):
    spec = self.info["scales"][mip].get("sharding")
    if not spec:
      return None
    from . import sharding
    return sharding.ShardingSpecification(spec)

  def _chunk_id(self, mip, chunk_box):
    from . import sharding
    cs, off = self.chunk_size_at(mip), self.voxel_offset_at(mip)
    grid = [int(math.ceil(int(v) / int(c))) for v, c in zip(self.volume_size_at(mip), cs)]
    pt = [int((int(a) - int(o)) // int(c)) for a, o, c in zip(chunk_box.minpt, off, cs)]
    return int(sharding.compressed_morton_code(pt, grid))

  def _read_chunk(self, mip, chunk_box, shard_cache):
    spec = self._sharding(mip)
    if spec is None:
      return self.cf.get(self._chunk_name(mip, chunk_box))
    cid = self._chunk_id(mip, chunk_box)
    name = self.key_at(mip) + "/" + spec.shard_filename(spec.locate(cid)[0])
    if name not in shard_cache:
      shard_cache[name] = self.cf.get(name)
    blob = shard_cache[name]
    return None if blob is None else spec.read_chunk(blob, cid)

  # chunk codecs: `raw` is the bytes of the Fortran-order array; `compressed_segmentation` goes
  # through the device codec (igneous_b200.codecs); anything else the Precomputed format knows
  # (jpeg, compresso, crackle, ...) is outside this stand-in
  def _encoding(self, mip):
    return self.info["scales"][mip].get("encoding", "raw")

  def _cseg_block(self, mip):
    return tuple(int(v) for v in self.info["scales"][mip].get("compressed_segmentation_block_size", (8, 8, 8)))

  def _encode_chunk(self, block, mip):
    enc = self._encoding(mip)
    if enc == "raw":
      return block.tobytes(order="F")
    if enc == "compressed_segmentation":
      from . import codecs
      return codecs.cseg_encode(block, self._cseg_block(mip))
    raise NotImplementedError("storage stand-in: chunk encoding %r is not supported" % enc)

  def _decode_chunk(self, data, mip, shape):
    enc = self._encoding(mip)
    if enc == "raw":
      return np.frombuffer(data, dtype=self.dtype).reshape(shape, order="F")
    if enc == "compressed_segmentation":
      from . import codecs
      return codecs.cseg_decode(data, shape, self.dtype, self._cseg_block(mip))
    raise NotImplementedError("storage stand-in: chunk encoding %r is not supported" % enc)

  def _to_bbox(self, key):
    if isinstance(key, Bbox):
      return key.clone()
    if isinstance(key, slice):
      key = (key,)
    if isinstance(key, tuple):
      b = self.bounds
      key = tuple(key) + (slice(None),) * (3 - len(key[:3]))
      lo = [b.minpt[i] if key[i].start is None else key[i].start for i in range(3)]
      hi = [b.maxpt[i] if key[i].stop is None else key[i].stop for i in range(3)]
      return Bbox(lo, hi)
    raise TypeError(key)

  def download(self, bbox, mip=None, renumber=False, **kwargs):
    """Cutout as an F-order [x, y, z, c] array.  renumber=True -> (array of the smallest
    dtype holding 1..N, {old: new}) with the relabelling done on the GPU
    (cloudvolume's download(renumber=True), image.py:745-752)."""
    if renumber:
      from . import fastremap
      img = self.download(bbox, mip=mip, **kwargs)
      small, mapping = fastremap.renumber(img, preserve_zero=True, in_place=False)
      return small, mapping
    mip = self._mip if mip is None else mip
    bbox = self._to_bbox(bbox)
    if self.bounded and not (np.all(bbox.minpt >= self.bounds_at(mip).minpt) and np.all(bbox.maxpt <= self.bounds_at(mip).maxpt)):
      raise OutOfBoundsError("%r is outside %r" % (bbox, self.bounds_at(mip)))
    out = np.zeros(tuple(int(v) for v in bbox.size3()) + (self.num_channels,), dtype=self.dtype, order="F")
    shard_cache = {}
    for c in self._chunks(mip, bbox):
      data = self._read_chunk(mip, c, shard_cache)
      inter = Bbox.intersection(c, bbox)
      if inter.subvoxel():
        continue
      if data is None:
        if not self.fill_missing:
          raise EmptyVolumeException(self._chunk_name(mip, c))
        continue
      chunk = self._decode_chunk(data, mip, tuple(int(v) for v in c.size3()) + (self.num_channels,))
      src = tuple(slice(int(a - o), int(b - o)) for a, b, o in zip(inter.minpt, inter.maxpt, c.minpt))
      dst = tuple(slice(int(a - o), int(b - o)) for a, b, o in zip(inter.minpt, inter.maxpt, bbox.minpt))
      out[dst] = chunk[src]
    return out

  def __getitem__(self, key):
    return self.download(key)

  def __setitem__(self, key, img):
    mip = self._mip
    bbox = self._to_bbox(key)
    img = np.asarray(img)
    if img.ndim == 3:
      img = img[..., np.newaxis]
    if tuple(img.shape[:3]) != tuple(int(v) for v in bbox.size3()):
      raise ValueError("image %r does not fit %r" % (img.shape, bbox))
    img = img.astype(self.dtype, copy=False)
    if self._sharding(mip) is not None:
      raise NotImplementedError("writes to a sharded scale go through image.make_shard (whole shards only)")
    for c in self._chunks(mip, bbox):
      inter = Bbox.intersection(c, bbox)
      if inter.subvoxel():
        continue
      if not (inter == c):
        raise ValueError("writes must be chunk aligned: %r vs chunk %r" % (bbox, c))
      src = tuple(slice(int(a - o), int(b - o)) for a, b, o in zip(c.minpt, c.maxpt, bbox.minpt))
      block = np.asfortranarray(img[src])
      name = self._chunk_name(mip, c)
      if self.delete_black_uploads and not np.any(block != self.background_color):
        self.cf.delete(name)
        continue
      self.cf.put(name, self._encode_chunk(block, mip), compress=self.compress)


# --------------------------------------------------------------------- queue
def queueable(fn):
  return fn


class RegisteredTask:
  def __init__(self, *args, **kwargs):
    self._args, self._kwargs = args, kwargs

  def execute(self):
    raise NotImplementedError()


class LocalTaskQueue:
  """In-process immediate execution (taskqueue.LocalTaskQueue(parallel=1))."""

  def __init__(self, parallel=1, **kwargs):
    self.parallel = parallel
    self.executed = 0

  def _run(self, task):
    if hasattr(task, "execute"):
      task.execute()
    else:
      task()
    self.executed += 1

  def insert(self, tasks, **kwargs):
    if hasattr(tasks, "execute") or callable(tasks):
      tasks = [tasks]
    for t in tasks:
      self._run(t)

  insert_all = insert

  def execute(self, **kwargs):
    pass
