"""Seeded random sweeps over shapes / dtypes / label statistics: GPU vs oracle, bit exact.
Exercises partial tiles, rows that are not multiples of 32, degenerate extents and
dense label noise on every kernel of the path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

DTYPES = [np.uint8, np.uint16, np.uint32, np.uint64]


def _labels(rng, shape, dtype):
  kind = rng.integers(0, 4)
  if kind == 0:      # dense noise, few labels
    a = rng.integers(0, 3, size=shape)
  elif kind == 1:    # blobs
    small = rng.integers(0, 5, size=tuple((s + 4) // 5 for s in shape))
    a = np.repeat(np.repeat(np.repeat(small, 5, 0), 5, 1), 5, 2)[:shape[0], :shape[1], :shape[2]]
    a = np.where(rng.random(shape) < 0.15, 0, a)
  elif kind == 2:    # long runs along x
    a = np.repeat(rng.integers(0, 4, size=(1, shape[1], shape[2])), shape[0], axis=0)
    a = np.where(rng.random(shape) < 0.02, rng.integers(0, 4, size=shape), a)
  else:              # every voxel distinct
    a = rng.permutation(int(np.prod(shape))).reshape(shape) % 250 + 1
  a = a.astype(np.uint64)
  if np.dtype(dtype).itemsize == 8:
    a = a * np.uint64((1 << 40) + 3)
  return np.asfortranarray(a.astype(dtype))


def _shape(rng):
  pick = rng.integers(0, 4)
  if pick == 0:
    return tuple(int(v) for v in rng.integers(1, 12, size=3))
  if pick == 1:
    return (int(rng.integers(250, 270)), int(rng.integers(1, 20)), int(rng.integers(1, 12)))
  if pick == 2:
    return (int(rng.integers(20, 70)), int(rng.integers(7, 20)), int(rng.integers(7, 20)))
  return (int(rng.integers(30, 40)), int(rng.integers(1, 4)), int(rng.integers(15, 40)))


def test_ccl_random_sweep(ctx, oracle):
  from igneous_b200 import cc3d
  rng = np.random.default_rng(2024)
  for case in range(40):
    shape, dtype = _shape(rng), DTYPES[case % 4]
    labels = _labels(rng, shape, dtype)
    got, n = cc3d.connected_components(labels, connectivity=6, out_dtype=np.uint32, return_N=True)
    want, n_want = oracle.connected_components(labels, return_N=True)
    assert n == n_want and np.array_equal(got, want.astype(np.uint32)), (case, shape, dtype)


def test_pool_random_sweep(ctx, oracle):
  from igneous_b200 import tinybrain
  rng = np.random.default_rng(2025)
  for case in range(30):
    shape, dtype = _shape(rng), DTYPES[case % 4]
    labels = _labels(rng, shape, dtype)
    mips = int(rng.integers(1, 5))
    got = tinybrain.downsample_segmentation(labels, (2, 2, 1), num_mips=mips, sparse=bool(case % 3 == 0))
    want = oracle.downsample_segmentation(labels, (2, 2, 1), num_mips=mips, sparse=bool(case % 3 == 0))
    for g, w in zip(got, want):
      assert np.array_equal(g, w), (case, shape, dtype)
    if np.dtype(dtype).itemsize <= 4:
      r = int(case % 3)
      got = tinybrain.downsample_with_averaging(labels, (2, 2, 1), num_mips=mips, rounding=r)
      want = oracle.downsample_with_averaging(labels, (2, 2, 1), num_mips=mips, rounding=r)
      for g, w in zip(got, want):
        assert np.array_equal(g, w), (case, shape, dtype, "avg")


def test_mesh_random_sweep(ctx, oracle):
  from igneous_b200 import zmesh
  rng = np.random.default_rng(2026)
  for case in range(16):
    shape = tuple(int(v) for v in rng.integers(2, 20, size=3))
    labels = _labels(rng, shape, DTYPES[case % 4])
    m = zmesh.Mesher((4, 4, 40))
    m.mesh(labels)
    tl, tv = oracle.marching_cubes(labels)
    W = oracle.WeldedMeshes(tl, tv)
    assert sorted(m.ids()) == W.ids(), (case, shape)
    factor = [0, 3, 100][case % 3]
    want = oracle.simplify_welded(W, (4, 4, 40), factor, 40.0, True)[0] if factor else None
    for lab in W.ids():
      got = m.get(lab, reduction_factor=factor, max_error=40.0, voxel_centered=True)
      wv, wf = want[lab] if factor else W.get(lab, (4, 4, 40), True)
      assert np.array_equal(got.vertices, wv) and np.array_equal(got.faces, wf), (case, shape, lab, factor)


def test_remap_random_sweep(ctx, oracle):
  from igneous_b200 import fastremap
  rng = np.random.default_rng(2027)
  for case in range(20):
    shape, dtype = _shape(rng), DTYPES[case % 4]
    labels = _labels(rng, shape, dtype)
    got, gmap = fastremap.renumber(labels)
    want, wmap = oracle.renumber(labels)
    assert gmap == wmap and np.array_equal(got, want), (case, shape, dtype)
    u, c = fastremap.unique(labels, return_counts=True)
    wu, wc = oracle.unique(labels, return_counts=True)
    assert np.array_equal(u, wu) and np.array_equal(c.astype(np.int64), wc)
