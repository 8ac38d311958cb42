import ctypes as c, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from igneous_b200 import _shim, pipeline
ctx = _shim.default_context()
pipe = pipeline.VolumePipeline(ctx, (2048, 2048, 256), np.uint32, simplification_factor=0)
pipe.synth(); pipe.pool(); ctx.sync()
for i in range(2):
  t=time.perf_counter(); ctx.timer_start(5); pipe.mesh(); ctx.timer_stop(5)
  print("A: 4 tasks 257x257x256: mesh ms", ctx.timer_ms(5), (time.perf_counter()-t)*1e3, pipe.mesh_stats)
big = ctx.alloc(100 << 30)   # memory pressure like the 2048^3 run
for i in range(2):
  ctx.timer_start(5); pipe.mesh(); ctx.timer_stop(5)
  print("B: +100GB allocated: mesh ms", ctx.timer_ms(5))
big.free()
# grow the arena like the volume CCL does
pipe.ccl(); ctx.sync()
for i in range(2):
  ctx.timer_start(5); pipe.mesh(); ctx.timer_stop(5)
  print("C: after volume CCL (big arena): mesh ms", ctx.timer_ms(5))
