"""Wall time of one 256^3 MeshTask body (marching cubes + weld + simplification) at mip 2
of the synthetic bench volume, one mesh stream.  usage: time_simplify.py [factor] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from igneous_b200 import _shim, pipeline

ctx = _shim.default_context()
factor = int(sys.argv[1]) if len(sys.argv) > 1 else 100
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
pipe = pipeline.VolumePipeline(ctx, (1024, 1024, 256), np.uint32, simplification_factor=factor, mesh_streams=1)
pipe.synth(); pipe.pool(); ctx.sync()
times = []
for i in range(reps):
  t0 = time.perf_counter(); pipe.mesh(); times.append((time.perf_counter() - t0) * 1e3)
print("mesh ms per task:", " ".join("%.1f" % t for t in times), pipe.mesh_stats)
