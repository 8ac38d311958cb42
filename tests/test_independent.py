"""Independent cross-checks of the oracle's recalled rules (VERDICT r1 "pin the unpinned rules
as far as this image allows"): the upstream wheels are absent, so each rule is checked against
something that does not share code with oracle/igneous_oracle.c -- a numpy transcription of
the PUBLISHED COUNTLESS-2D algorithm the reference links (README.md:233,290), brute-force
statistics, and a geometric inside/outside test of the marching-cubes surfaces."""
import numpy as np
import pytest


def countless2d_published(data):
  """zero-corrected COUNTLESS 2D as published in the article linked from the reference's
  README.md:233 (numpy formulation: a, b, c, d are the four phase-shifted sub-images; matches
  PICK(a,b) | PICK(a,c) | PICK(b,c), else d; labels are offset by one so that 0 is a label)."""
  d64 = data.astype(np.uint64) + np.uint64(1)
  a, b, c, d = d64[0::2, 0::2], d64[1::2, 0::2], d64[0::2, 1::2], d64[1::2, 1::2]  # x, y: a=(0,0) b=(1,0) c=(0,1) d=(1,1)
  ab = a * (a == b)
  ac = a * (a == c)
  bc = b * (b == c)
  r = ab | ac | bc
  r = r + (r == 0) * d
  return (r - np.uint64(1)).astype(data.dtype)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint32, np.uint64])
def test_mode_pool_equals_published_countless2d(oracle, dtype):
  rng = np.random.default_rng(7)
  for hi in (2, 3, 5, 200):  # few labels: many ties, many zeros
    img = np.asfortranarray(rng.integers(0, hi, size=(64, 48, 5)).astype(dtype))
    got = oracle.downsample_segmentation(img, (2, 2, 1), num_mips=3)
    cur = img
    for m in range(3):
      want = np.stack([countless2d_published(cur[:, :, z]) for z in range(cur.shape[2])], axis=2)
      assert np.array_equal(got[m], want), (hi, m)
      cur = want


def test_mode_pool_mip1_is_a_statistical_mode(oracle):
  """README.md:233: 'mip 1 segmentation labels are exact mode computations': the picked value is
  always one of the most frequent values of its 2x2 block (brute force)."""
  rng = np.random.default_rng(8)
  img = np.asfortranarray(rng.integers(0, 4, size=(40, 40, 3)).astype(np.uint32))
  out = oracle.downsample_segmentation(img, (2, 2, 1), num_mips=1)[0]
  for z in range(3):
    for y in range(20):
      for x in range(20):
        blk = img[2 * x:2 * x + 2, 2 * y:2 * y + 2, z].ravel()
        vals, cnt = np.unique(blk, return_counts=True)
        assert out[x, y, z] in vals[cnt == cnt.max()]


def test_sparse_mode_pool_ignores_background(oracle):
  """README.md:290 (stippled COUNTLESS): zero is background, never a label: the result is 0 only
  for an all-zero block and otherwise one of the most frequent NON-ZERO values."""
  rng = np.random.default_rng(9)
  img = np.asfortranarray((rng.integers(0, 4, size=(32, 32, 2)) * (rng.random((32, 32, 2)) < 0.5)).astype(np.uint32))
  out = oracle.downsample_segmentation(img, (2, 2, 1), num_mips=1, sparse=True)[0]
  for z in range(2):
    for y in range(16):
      for x in range(16):
        blk = img[2 * x:2 * x + 2, 2 * y:2 * y + 2, z].ravel()
        nzv = blk[blk != 0]
        if len(nzv) == 0:
          assert out[x, y, z] == 0
        else:
          vals, cnt = np.unique(nzv, return_counts=True)
          assert out[x, y, z] in vals[cnt == cnt.max()]


def test_average_pool_is_the_block_mean_to_within_rounding(oracle):
  """Whatever the (unpinned) rounding rule, every level is within one unit of the exact mean of
  the mip-0 block it covers, and exact when the mean is an integer."""
  rng = np.random.default_rng(10)
  img = np.asfortranarray(rng.integers(0, 255, size=(64, 64, 2)).astype(np.uint8))
  outs = oracle.downsample_with_averaging(img, (2, 2, 1), num_mips=4)
  for m, o in enumerate(outs):
    k = 2 ** (m + 1)
    mean = img.astype(np.float64).reshape(64 // k, k, 64 // k, k, 2, order="F").mean(axis=(1, 3)) \
        if False else np.stack([img[:, :, z].astype(np.float64).reshape(64 // k, k, 64 // k, k).mean(axis=(1, 3)) for z in range(2)], axis=2)
    assert np.all(np.abs(o.astype(np.float64) - mean) < 1.0 + 1e-9)
    exact = mean == np.floor(mean)
    assert np.array_equal(o[exact].astype(np.float64), mean[exact])


def _ray_parity(tri, pts):
  """number of crossings (mod 2) of the +x ray from each point with the triangles"""
  p0, p1, p2 = tri[:, 0], tri[:, 1], tri[:, 2]
  inside = np.zeros(len(pts), dtype=bool)
  for i, p in enumerate(pts):
    # 2-D point-in-triangle in the (y, z) projection, then the x of the hit
    d1 = (p1[:, 1] - p0[:, 1]) * (p[2] - p0[:, 2]) - (p1[:, 2] - p0[:, 2]) * (p[1] - p0[:, 1])
    d2 = (p2[:, 1] - p1[:, 1]) * (p[2] - p1[:, 2]) - (p2[:, 2] - p1[:, 2]) * (p[1] - p1[:, 1])
    d3 = (p0[:, 1] - p2[:, 1]) * (p[2] - p2[:, 2]) - (p0[:, 2] - p2[:, 2]) * (p[1] - p2[:, 1])
    hit = ((d1 > 0) & (d2 > 0) & (d3 > 0)) | ((d1 < 0) & (d2 < 0) & (d3 < 0))
    if not hit.any():
      continue
    n = np.cross(p1[hit] - p0[hit], p2[hit] - p0[hit])
    # plane: n . (q - p0) = 0 with q = (x, p.y, p.z)
    x = p0[hit, 0] - (n[:, 1] * (p[1] - p0[hit, 1]) + n[:, 2] * (p[2] - p0[hit, 2])) / n[:, 0]
    inside[i] = (np.count_nonzero(x > p[0]) % 2) == 1
  return inside


def test_marching_cubes_surface_separates_inside_from_outside(oracle):
  """Geometric check that shares nothing with the 256-case table: for every label of a random
  multi-label volume (zero border: closed surfaces), a point next to each voxel centre is inside
  the label's mesh (odd number of ray crossings) exactly when the voxel carries the label."""
  rng = np.random.default_rng(11)
  vol = np.zeros((9, 8, 7), dtype=np.uint32, order="F")
  vol[1:-1, 1:-1, 1:-1] = rng.integers(0, 3, size=(7, 6, 5))
  tl, tv = oracle.marching_cubes(vol)
  xs, ys, zs = np.meshgrid(np.arange(9), np.arange(8), np.arange(7), indexing="ij")
  # voxel centres are lattice points of the half-voxel grid: step off the lattice (the surface
  # keeps >= 0.28 voxels away from every voxel centre)
  pts = np.stack([xs.ravel() + 0.031, ys.ravel() + 0.123, zs.ravel() + 0.077], axis=1)
  for lab in (1, 2):
    tri = tv[tl == lab].astype(np.float64) * 0.5  # half-voxel units -> voxels
    inside = _ray_parity(tri, pts).reshape(vol.shape)
    assert np.array_equal(inside, vol == lab), lab
