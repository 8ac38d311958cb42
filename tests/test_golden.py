"""Committed golden vectors (tests/golden/make_golden.py): the oracle must keep
reproducing them (CPU), and the CUDA path must reproduce them too (GPU)."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hotpath_small.npz"))


def test_oracle_reproduces_golden(oracle):
  seg, img = G["seg"], G["img"]
  for k, m in enumerate(oracle.downsample_segmentation(seg, (2, 2, 1), num_mips=3)):
    assert np.array_equal(m, G["mode%d" % (k + 1)])
  for k, m in enumerate(oracle.downsample_with_averaging(img, (2, 2, 1), num_mips=5)):
    assert np.array_equal(m, G["avg%d" % (k + 1)])
  cc, n = oracle.connected_components(seg, return_N=True)
  assert n == int(G["n"]) and np.array_equal(cc.astype(np.uint32), G["cc"])
  assert np.array_equal(oracle.dust(seg, 40), G["dust40"])
  assert np.array_equal(oracle.renumber(seg)[0], G["renumber"])
  tl, tv = oracle.marching_cubes(seg[:25, :21, :13])
  assert len(tl) == int(G["n_triangles"])
  W = oracle.WeldedMeshes(tl, tv)
  v, f = W.get(int(G["mesh_label"]), (16, 16, 40), True)
  assert np.array_equal(v, G["mesh_vertices"]) and np.array_equal(f, G["mesh_faces"])
  simp, _ = oracle.simplify_welded(W, (16, 16, 40), 4, 1e9, True)
  sv, sf = simp[int(G["mesh_label"])]
  assert np.array_equal(sv, G["simp_vertices"]) and np.array_equal(sf, G["simp_faces"])


def test_oracle_reproduces_block_pooling_golden(oracle):
  B = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pooling_blocks.npz"))
  seg, img = G["seg"], G["img"]
  runs = {"mode222": oracle.downsample_segmentation(seg, (2, 2, 2), num_mips=2),
          "smode222": oracle.downsample_segmentation(seg, (2, 2, 2), num_mips=2, sparse=True),
          "mode122": oracle.downsample_segmentation(seg, (1, 2, 2), num_mips=2),
          "avg222": oracle.downsample_with_averaging(img, (2, 2, 2), num_mips=2),
          "savg221": oracle.downsample_with_averaging(np.where(img > 128, img, 0).astype(np.uint8), (2, 2, 1),
                                                      num_mips=2, sparse=True)}
  for name, arrs in runs.items():
    for k, a in enumerate(arrs):
      assert np.array_equal(a, B["%s_%d" % (name, k + 1)]), name


@pytest.mark.gpu
def test_cuda_reproduces_golden(ctx):
  from igneous_b200 import tinybrain, cc3d, fastremap, zmesh
  seg, img = G["seg"], G["img"]
  for k, m in enumerate(tinybrain.downsample_segmentation(seg, (2, 2, 1), num_mips=3)):
    assert np.array_equal(m, G["mode%d" % (k + 1)])
  for k, m in enumerate(tinybrain.downsample_with_averaging(img, (2, 2, 1), num_mips=5)):
    assert np.array_equal(m, G["avg%d" % (k + 1)])
  cc, n = cc3d.connected_components(seg, connectivity=6, out_dtype=np.uint32, return_N=True)
  assert n == int(G["n"]) and np.array_equal(cc, G["cc"])
  assert np.array_equal(cc3d.dust(seg, 40, connectivity=6), G["dust40"])
  assert np.array_equal(fastremap.renumber(seg)[0], G["renumber"])
  lab = int(G["mesh_label"])
  m = zmesh.Mesher((16, 16, 40))
  m.mesh(seg[:25, :21, :13])
  mesh = m.get(lab, reduction_factor=0, voxel_centered=True)
  assert np.array_equal(mesh.vertices, G["mesh_vertices"]) and np.array_equal(mesh.faces, G["mesh_faces"])
  m.mesh(seg[:25, :21, :13])
  s = m.get(lab, reduction_factor=4, max_error=1e9, voxel_centered=True)
  assert np.array_equal(s.vertices, G["simp_vertices"]) and np.array_equal(s.faces, G["simp_faces"])
