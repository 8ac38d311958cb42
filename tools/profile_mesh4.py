import ctypes as c, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from igneous_b200 import _shim, pipeline
ctx = _shim.default_context()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
pipe = pipeline.VolumePipeline(ctx, (S, S, S), np.uint32, simplification_factor=0)
pipe.synth(); ctx.sync()
lib = ctx.lib
src = pipe.d_mips[-1]; msx, msy, msz = pipe.mip_shapes[-1]
for rep in range(3):
  t = time.perf_counter(); pipe.pool(); ctx.sync(); tp = time.perf_counter() - t
  t = time.perf_counter(); pipe.ccl(); ctx.sync(); tc = time.perf_counter() - t
  out = []
  for (x0, y0, z0, bx, by, bz) in pipe.mesh_tasks():
    t0 = time.perf_counter()
    _shim.check(lib.ign_copy_box_dev(ctx.handle, _shim.ptr(src), c.c_int(pipe.code), c.c_uint64(msx), c.c_uint64(msy), c.c_uint64(msz), c.c_uint64(x0), c.c_uint64(y0), c.c_uint64(z0), c.c_uint64(bx), c.c_uint64(by), c.c_uint64(bz), _shim.ptr(pipe.d_task)))
    ctx.sync(); t1 = time.perf_counter()
    h = c.c_void_p()
    _shim.check(lib.ign_mesh_begin_dev(ctx.handle, _shim.ptr(pipe.d_task), c.c_int(pipe.code), c.c_uint64(bx), c.c_uint64(by), c.c_uint64(bz), c.byref(h)))
    ctx.sync(); t2 = time.perf_counter()
    lib.ign_mesh_free(h); t3 = time.perf_counter()
    out.append((round((t1-t0)*1e3,2), round((t2-t1)*1e3,2), round((t3-t2)*1e3,2)))
  print("pool %.1f ccl %.1f" % (tp*1e3, tc*1e3), out[:5], "sum begin", sum(o[1] for o in out))
