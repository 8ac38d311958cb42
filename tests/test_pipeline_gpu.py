"""Device-resident pipeline (what bench.py times) against the oracle at a small size,
including a multi-slab CCL and concurrent mesh streams."""
import ctypes as c

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_volume_pipeline_matches_oracle(ctx, oracle):
  from igneous_b200 import pipeline, _shim
  shape = (128, 128, 96)
  pipe = pipeline.VolumePipeline(ctx, shape, np.uint32, num_mips=2, mesh_shape=(16, 16, 64), pitch=32,
                                 num_ids=9, simplification_factor=0, mesh_streams=3)
  try:
    pipe.synth()
    pipe.step(timers=True)
    ctx.sync()
    seg = oracle.synth_seg(shape, pitch=32, num_ids=9)
    assert np.array_equal(ctx.to_host(pipe.d_in, shape, np.uint32), seg)
    want_mips = oracle.downsample_segmentation(seg, (2, 2, 1), num_mips=2)
    for d, s, w in zip(pipe.d_mips, pipe.mip_shapes, want_mips):
      assert np.array_equal(ctx.to_host(d, s, np.uint32), w)
    cc, n = oracle.connected_components(seg, return_N=True)
    assert pipe.n_components == n
    assert np.array_equal(ctx.to_host(pipe.d_cc, shape, np.uint32), cc.astype(np.uint32))
    # mesh stage: triangle total equals the oracle's over the same task cutouts
    m2 = want_mips[1]
    total = 0
    for (x0, y0, z0, bx, by, bz) in pipe.mesh_tasks():
      tl, _ = oracle.marching_cubes(m2[x0:x0 + bx, y0:y0 + by, z0:z0 + bz])
      total += len(tl)
    assert pipe.mesh_stats["triangles"] == total and pipe.mesh_stats["tasks"] == 8
    ms = pipe.stage_ms()
    assert all(v >= 0 for v in ms.values())
    assert pipe.launch_count() > 0
  finally:
    pipe.free()


@pytest.mark.parametrize("shape,mesh_shape", [((128, 128, 96), (16, 16, 32)), ((96, 64, 70), (32, 32, 16))])
def test_streamed_step_equals_resident_step(ctx, oracle, shape, mesh_shape):
  """Layer-wise upload + overlapped meshing gives the same products as upload-then-step."""
  from igneous_b200 import pipeline
  seg = np.asfortranarray(oracle.synth_seg(shape, pitch=32, num_ids=9).astype(np.uint32))
  pipe = pipeline.VolumePipeline(ctx, shape, np.uint32, num_mips=2, mesh_shape=mesh_shape, pitch=32,
                                 num_ids=9, simplification_factor=0, mesh_streams=3)
  try:
    pipe.load_host(seg)
    pipe.step(timers=False)
    ctx.sync()
    want = {"mips": [ctx.to_host(d, s, np.uint32) for d, s in zip(pipe.d_mips, pipe.mip_shapes)],
            "cc": ctx.to_host(pipe.d_cc, shape, np.uint32), "stats": dict(pipe.mesh_stats),
            "n": pipe.n_components}
    for b in [pipe.d_in, pipe.d_cc] + pipe.d_mips:
      ctx.memset(b, 0xEE, b.nbytes)
    host = {"mips": [np.empty(s, dtype=np.uint32, order="F") for s in pipe.mip_shapes],
            "cc": np.empty(shape, dtype=np.uint32, order="F")}
    got_meshes = []

    def export(task, h, nv, nf, nl, wctx):
      got_meshes.append((task, nv, nf, nl))

    pipe.step_streamed(seg, host, export)
    assert pipe.n_components == want["n"]
    assert pipe.mesh_stats == want["stats"]
    assert sum(m[2] for m in got_meshes) == want["stats"]["triangles"]
    for g, w in zip(host["mips"], want["mips"]):
      assert np.array_equal(g, w)
    assert np.array_equal(host["cc"], want["cc"])
    assert np.array_equal(ctx.to_host(pipe.d_in, shape, np.uint32), seg)
  finally:
    pipe.free()
