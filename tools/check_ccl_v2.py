"""Opt-in CCL tile kernels: k_ccl_local_v2 (IGN_CCL_V2=1) and the half-height tile experiment
k_ccl_local_v3 / k_ccl_merge_tiles_s (IGN_CCL_V2=3): parity against the oracle and timing
against k_ccl_local_fast.  usage: python tools/check_ccl_v2.py [modes, default "1,3"]"""
import ctypes as c
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from igneous_b200 import _shim, cc3d
from oracle import oracle


MODES = (sys.argv[1] if len(sys.argv) > 1 else "1,3").split(",")


def parity(mode):
  rng = np.random.default_rng(0)
  ok = True
  for dtype in (np.uint8, np.uint16, np.uint32, np.uint64):
    vols = [oracle.synth_seg((512, 64, 40), pitch=16, num_ids=9).astype(dtype),
            rng.integers(0, 3, size=(256, 16, 24)).astype(dtype),      # dense noise: overflow path
            oracle.synth_seg((300, 40, 20), pitch=16, num_ids=5).astype(dtype)]  # partial tiles, sx % 4 == 0
    for v in vols:
      v = np.asfortranarray(v)
      want = oracle.connected_components(v)
      os.environ["IGN_CCL_V2"] = mode
      got = cc3d.connected_components(v, connectivity=6, out_dtype=np.uint64)
      del os.environ["IGN_CCL_V2"]
      same = bool(np.array_equal(got, want))
      ok &= same
      print("parity mode", mode, np.dtype(dtype).name, v.shape, "OK" if same else "MISMATCH", flush=True)
  return ok


def timing(ctx, shape=(512, 512, 512)):
  sx, sy, sz = shape
  n = sx * sy * sz
  d_in, d_out = ctx.alloc(n * 4), ctx.alloc(n * 8)
  _shim.check(ctx.lib.ign_synth_seg_dev(ctx.handle, _shim.ptr(d_in), c.c_int(_shim.IGN_U32), c.c_uint64(sx), c.c_uint64(sy),
                                        c.c_uint64(sz), c.c_int64(0), c.c_int64(0), c.c_int64(0), c.c_uint32(64),
                                        c.c_uint64(1 << 20), c.c_uint64(0), c.c_uint64(0)))
  N = c.c_uint64(0)
  args = (ctx.handle, _shim.ptr(d_in), c.c_int(_shim.IGN_U32), c.c_uint64(sx), c.c_uint64(sy), c.c_uint64(sz),
          _shim.ptr(d_out), c.c_int(_shim.IGN_U64), c.byref(N))
  out = {}
  for name, env in [("fast", None)] + [("v2" if m == "1" else "mode" + m, m) for m in MODES]:
    if env:
      os.environ["IGN_CCL_V2"] = env
    _shim.check(ctx.lib.ign_prof_enable(ctx.handle, c.c_int(0)))
    for _ in range(2):
      _shim.check(ctx.lib.ign_ccl6_dev(*args))
    _shim.check(ctx.lib.ign_prof_enable(ctx.handle, c.c_int(1)))
    reps = 5
    for _ in range(reps):
      _shim.check(ctx.lib.ign_ccl6_dev(*args))
    ctx.sync()
    ms, cnt = c.c_float(0), c.c_uint64(0)
    _shim.check(ctx.lib.ign_prof_read(ctx.handle, c.c_int(0), c.byref(ms), c.byref(cnt)))
    out[name] = {"local_ms": round(ms.value / max(cnt.value, 1), 4), "components": int(N.value)}
    ms2, cnt2 = c.c_float(0), c.c_uint64(0)
    _shim.check(ctx.lib.ign_prof_read(ctx.handle, c.c_int(1), c.byref(ms2), c.byref(cnt2)))
    out[name]["merge_ms"] = round(ms2.value / max(cnt2.value, 1), 4)
    os.environ.pop("IGN_CCL_V2", None)
  print(json.dumps({"shape": shape, **out}))


if __name__ == "__main__":
  good = all([parity(m) for m in MODES])
  if good:
    timing(_shim.default_context())
  sys.exit(0 if good else 1)
