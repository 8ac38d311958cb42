"""One 256^3 mesh task at mip 2 of the synthetic bench volume (for ncu / timing)."""
import ctypes as c, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from igneous_b200 import _shim, pipeline
ctx = _shim.default_context()
simplify = int(sys.argv[1]) if len(sys.argv) > 1 else 0
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pipe = pipeline.VolumePipeline(ctx, (1024, 1024, 256), np.uint32, simplification_factor=simplify)
pipe.synth(); pipe.pool(); ctx.sync()
for i in range(reps):
  ctx.timer_start(5); pipe.mesh(); ctx.timer_stop(5)
  print("mesh ms", ctx.timer_ms(5), pipe.mesh_stats)
