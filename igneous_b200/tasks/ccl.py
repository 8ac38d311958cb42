"""Out-of-core 6-connected CCL in four passes, kernels on the GPU.

Mirror of igneous/tasks/image/ccl.py (same task names, arguments, file layout
under {key}/ccl/{faces,equivalences,relabel}, union rule and label offsets):
  (1) CCLFacesTask        :126-194   3 back faces of each 1-voxel-overlap task
  (2) CCLEquivalancesTask :196-294   links against the 3 neighbouring faces
  (3) create_relabeling   :358-420   global union-find -> relabel tables
  (4) RelabelCCLTask      :296-356   apply and write without the overlap
Each task's threshold -> rails blackout -> dust -> CCL -> offset chain is ONE
fused GPU call (cc3d.ccl_task); face linkage and the final remap use the GPU
fastremap kernels.  The reference writes its face files as crackle streams
(`*.ckl`, ccl.py:183); no crackle codec exists here, so the faces are written as
`.npy` files next to where the reference puts its `.ckl` files -- never under the
reference's names -- and pass 2 refuses to mix the two formats: a job must run all
four passes with one implementation.
"""
import io
from collections import defaultdict

import numpy as np

from .. import cc3d, fastremap
from .._compat import CloudVolume, CloudFiles, Bbox, Vec, queueable


class DisjointSet:
  """dict union-find, smaller id wins (ccl.py:48-73)."""

  def __init__(self):
    self.data = {}

  def makeset(self, x):
    self.data[x] = x
    return x

  def find(self, x):
    if x not in self.data:
      return None
    i = self.data[x]
    while i != self.data[i]:
      self.data[i] = self.data[self.data[i]]
      i = self.data[i]
    return i

  def union(self, x, y):
    i, j = self.find(x), self.find(y)
    if i is None:
      i = self.makeset(x)
    if j is None:
      j = self.makeset(y)
    if i < j:
      self.data[j] = i
    else:
      self.data[i] = j


def compute_task_number(grid_size, gridpoint):
  return int(gridpoint[0] + grid_size[0] * (gridpoint[1] + grid_size[1] * gridpoint[2]))


def compute_label_offset(shape, grid_size, gridpoint):
  return compute_task_number(grid_size, gridpoint) * int(shape[0]) * int(shape[1]) * int(shape[2])


def threshold_image(image, threshold_lte, threshold_gte):
  """ccl.py:89-101 (host form, kept for API parity; tasks use the fused kernel)."""
  if threshold_gte is None and threshold_lte is None:
    return image
  if threshold_gte is None:
    return image <= threshold_lte
  if threshold_lte is None:
    return image >= threshold_gte
  return (image >= threshold_gte) & (image <= threshold_lte)


def blackout_non_face_rails(labels, shape):
  """ccl.py:103-124 (host form)."""
  for slc in (np.s_[shape[0], shape[1], :], np.s_[shape[0], :, shape[2]], np.s_[:, shape[1], shape[2]]):
    try:
      labels[slc] = 0
    except IndexError:
      pass
  return labels


def _encode_face(face):
  buf = io.BytesIO()
  np.save(buf, np.ascontiguousarray(face), allow_pickle=False)
  return buf.getvalue()


FACE_SUFFIX = ".npy"  # the reference's `.ckl` names are reserved for real crackle streams


class ForeignFaceFormat(ValueError):
  """A CCL face file that was not written by this implementation."""


def _decode_face(data, name=""):
  if data[:4] == b"crkl":
    raise ForeignFaceFormat(
      "CCL face %s is a crackle stream written by a stock igneous worker; igneous_b200 has no crackle "
      "codec: run all four CCL passes of a job with one implementation" % name)
  if data[:6] != b"\x93NUMPY":
    raise ForeignFaceFormat("CCL face %s is not an igneous_b200 face file (unknown magic %r)" % (name, data[:6]))
  return np.load(io.BytesIO(data), allow_pickle=False)


def _task_frame(cloudpath, mip, shape, offset, fill_missing):
  shape, offset = Vec(*shape), Vec(*offset)
  bounds = Bbox(offset, offset + shape + 1)  # 1 voxel overlap
  if bounds.subvoxel():
    return None
  cv = CloudVolume(cloudpath, mip=mip, fill_missing=fill_missing)
  bounds = Bbox.clamp(bounds, cv.meta.bounds(mip))
  grid_size = np.ceil(np.asarray(cv.meta.bounds(mip).size3(), dtype=np.float64) / np.asarray(shape)).astype(int)
  gridpoint = np.floor(np.asarray(bounds.center()) / np.asarray(shape)).astype(int)
  offset_label = compute_label_offset(shape + 1, grid_size, gridpoint)
  return cv, shape, bounds, grid_size, gridpoint, offset_label


def _task_ccl(cv, bounds, shape, threshold_gte, threshold_lte, dust_threshold, label_offset):
  image = cv[bounds][..., 0]
  return cc3d.ccl_task(image, shape, threshold_gte=threshold_gte, threshold_lte=threshold_lte,
                       dust_threshold=dust_threshold, label_offset=label_offset)


@queueable
def CCLFacesTask(cloudpath, mip, shape, offset, threshold_gte=None, threshold_lte=None,
                 fill_missing=False, dust_threshold=0):
  frame = _task_frame(cloudpath, mip, shape, offset, fill_missing)
  if frame is None:
    return
  cv, shape, bounds, grid_size, gp, label_offset = frame
  cc_labels, _ = _task_ccl(cv, bounds, shape, threshold_gte, threshold_lte, dust_threshold, label_offset)
  faces = {"xy": cc_labels[:, :, -1], "xz": cc_labels[:, -1, :], "yz": cc_labels[-1, :, :]}
  cf = CloudFiles(cloudpath)
  cf.puts(((cf.join(cv.key, "ccl", "faces", "%d-%d-%d-%s%s" % (gp[0], gp[1], gp[2], k, FACE_SUFFIX)), _encode_face(v))
           for k, v in faces.items()), compress="br")


@queueable
def CCLEquivalancesTask(cloudpath, mip, shape, offset, threshold_gte=None, threshold_lte=None,
                        fill_missing=False, dust_threshold=0):
  frame = _task_frame(cloudpath, mip, shape, offset, fill_missing)
  if frame is None:
    return
  cv, shape, bounds, grid_size, gp, label_offset = frame
  cc_labels, n = _task_ccl(cv, bounds, shape, threshold_gte, threshold_lte, dust_threshold, label_offset)
  eq = DisjointSet()
  for i in range(1, n + 1):
    eq.makeset(i + label_offset)
  cf = CloudFiles(cloudpath)
  sx, sy, sz = (int(v) for v in shape)
  neighbours = [  # (file of the neighbouring task's back face, my front plane over the same voxels)
    ("%d-%d-%d-xy" % (gp[0], gp[1], gp[2] - 1), lambda f: f[:sx, :sy], cc_labels[:sx, :sy, 0]),
    ("%d-%d-%d-xz" % (gp[0], gp[1] - 1, gp[2]), lambda f: f[:sx, :sz], cc_labels[:sx, 0, :sz]),
    ("%d-%d-%d-yz" % (gp[0] - 1, gp[1], gp[2]), lambda f: f[:sy, :sz], cc_labels[0, :sy, :sz]),
  ]
  for stem, crop, cur in neighbours:
    fname = stem + FACE_SUFFIX
    data = cf.get(cf.join(cv.key, "ccl", "faces", fname))
    if data is None:
      if cf.exists(cf.join(cv.key, "ccl", "faces", stem + ".ckl")):  # pass 1 ran on a stock igneous worker
        raise ForeignFaceFormat(
          "face %s.ckl was written by a stock igneous worker (crackle); igneous_b200 cannot read it: run all "
          "four CCL passes of a job with one implementation" % stem)
      continue
    prev = crop(_decode_face(data, fname))
    cur = cur[:prev.shape[0], :prev.shape[1]]
    prev = prev[:cur.shape[0], :cur.shape[1]]
    for task_label, adj_labels in fastremap.inverse_component_map(cur, prev).items():
      if task_label == 0:
        continue
      for adj in adj_labels:
        if adj != 0:
          eq.union(int(task_label), int(adj))
  cf.put_json(cf.join(cv.key, "ccl", "equivalences", "%d-%d-%d.json" % (gp[0], gp[1], gp[2])),
              {str(k): int(v) for k, v in eq.data.items()}, compress="br")


def create_relabeling(cloudpath, mip, shape):
  """(3) global union-find over every equivalence file -> {key}/ccl/relabel/{task}.json
  and {key}/ccl/max_label.json."""
  cv = CloudVolume(cloudpath, mip=mip)
  cf = CloudFiles(cloudpath)
  eq = DisjointSet()
  for path in cf.list(cf.join(cv.key, "ccl", "equivalences")):
    for a, b in (cf.get_json(path) or {}).items():
      eq.union(int(a), int(b))
  relabel, next_label = {}, 1
  for key in eq.data.keys():
    root = eq.find(key)
    if root not in relabel:
      relabel[key] = relabel[root] = next_label
      next_label += 1
    else:
      relabel[key] = relabel[root]
  cf.put_json(cf.join(cv.key, "ccl", "max_label.json"), [next_label - 1])
  task_voxels = int(np.prod(np.asarray(shape) + 1))
  buckets = defaultdict(dict)
  for before, after in relabel.items():
    buckets[int(before // task_voxels)][before] = after
  cf.put_jsons(((cf.join(cv.key, "ccl", "relabel", "%d.json" % t), table) for t, table in buckets.items()),
               compress="br")


@queueable
def RelabelCCLTask(src_path, dest_path, mip, shape, offset, threshold_gte=None, threshold_lte=None,
                   fill_missing=False, dust_threshold=0):
  frame = _task_frame(src_path, mip, shape, offset, fill_missing)
  if frame is None:
    return
  cv, shape, bounds, grid_size, gp, label_offset = frame
  task_num = compute_task_number(grid_size, gp)
  cf = CloudFiles(src_path)
  mapping = cf.get_json(cf.join(cv.key, "ccl", "relabel", "%d.json" % task_num)) or {}
  mapping = {int(k): int(v) for k, v in mapping.items()}
  mapping[0] = 0
  cc_labels, _ = _task_ccl(cv, bounds, shape, threshold_gte, threshold_lte, dust_threshold, label_offset)
  cc_labels = fastremap.remap(cc_labels, mapping, in_place=True)
  dest = CloudVolume(dest_path, mip=mip)
  out_box = Bbox.clamp(Bbox(Vec(*offset), Vec(*offset) + shape), dest.meta.bounds(mip))
  s = out_box.size3()
  dest[out_box] = cc_labels[:s[0], :s[1], :s[2]].astype(dest.dtype)[..., np.newaxis]


def clean_intermediate_files(src, mip):
  cv = CloudVolume(src, mip)
  cf = CloudFiles(src)
  cf.delete(cf.list(cf.join(cv.key, "ccl")))
