"""Opt-in batched-gather simplifier kernels (IGN_SIMP_BATCH=1): bit-exact against the serial
kernels and the oracle, and timing of one 257^3 MeshTask body.
usage: python tools/check_simp_batch.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from igneous_b200 import _shim, pipeline, zmesh
from oracle import oracle


def meshes(seg, res, factor, max_error):
  m = zmesh.Mesher(res)
  m.mesh(seg)
  return {int(i): m.get(i, reduction_factor=factor, max_error=max_error, voxel_centered=True) for i in m.ids()}


def parity():
  ok = True
  for shape, pitch, factor in (((96, 80, 64), 24, 10), ((128, 128, 64), 32, 100)):
    seg = np.asfortranarray(oracle.synth_seg(shape, pitch=pitch, num_ids=9).astype(np.uint32))
    os.environ.pop("IGN_SIMP_BATCH", None)
    a = meshes(seg, (16, 16, 40), factor, 40.0)
    os.environ["IGN_SIMP_BATCH"] = "1"
    b = meshes(seg, (16, 16, 40), factor, 40.0)
    os.environ.pop("IGN_SIMP_BATCH", None)
    tl, tv = oracle.marching_cubes(seg)
    want, _ = oracle.simplify_welded(oracle.WeldedMeshes(tl, tv), (16, 16, 40), factor, 40.0, True)
    same = a.keys() == b.keys() == want.keys()
    for k in a:
      same &= bool(np.array_equal(a[k].vertices, b[k].vertices) and np.array_equal(a[k].faces, b[k].faces))
      same &= bool(np.array_equal(b[k].vertices, want[k][0]) and np.array_equal(b[k].faces, want[k][1]))
    ok &= same
    print("parity", shape, "factor", factor, "OK" if same else "MISMATCH", flush=True)
  return ok


def timing():
  ctx = _shim.default_context()
  pipe = pipeline.VolumePipeline(ctx, (1024, 1024, 256), np.uint32, simplification_factor=100, mesh_streams=1)
  pipe.synth(); pipe.pool(); ctx.sync()
  for name, env in (("serial", None), ("batch", "1")):
    if env:
      os.environ["IGN_SIMP_BATCH"] = env
    ts = []
    for _ in range(3):
      t0 = time.perf_counter(); pipe.mesh(); ts.append((time.perf_counter() - t0) * 1e3)
    os.environ.pop("IGN_SIMP_BATCH", None)
    print(name, "mesh ms per task:", " ".join("%.1f" % t for t in ts), pipe.mesh_stats["triangles"])


if __name__ == "__main__":
  good = parity()
  if good:
    timing()
  sys.exit(0 if good else 1)
