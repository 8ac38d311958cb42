"""CPU tests pinning the oracle (oracle/) against the reference's own known
answers and against independent implementations available in
this image (scipy.ndimage.label).  No GPU needed."""
import numpy as np
import pytest


# ---------------------------------------------------------------- mode pooling
MODE_KATS = [  # [a b; c d] -> out  (SURVEY.md 8(c) derived from the COUNTLESS rule)
  ((1, 1, 2, 3), 1), ((1, 2, 1, 3), 1), ((1, 2, 2, 3), 2), ((1, 2, 3, 3), 3),
  ((1, 2, 3, 4), 4), ((1, 1, 2, 2), 1), ((1, 2, 2, 1), 2), ((0, 0, 5, 5), 0),
]


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.uint32, np.uint64])
def test_mode_pool_kats(oracle, dtype):
  for (a, b, c, d), want in MODE_KATS:
    img = np.zeros((2, 2, 1), dtype=dtype, order="F")
    img[0, 0, 0], img[1, 0, 0], img[0, 1, 0], img[1, 1, 0] = a, b, c, d
    out, = oracle.downsample_segmentation(img, (2, 2, 1))
    assert out.shape == (1, 1, 1) and out[0, 0, 0] == want


def test_mode_pool_is_true_mode_when_unique(oracle):
  rng = np.random.default_rng(0)
  img = rng.integers(1, 4, size=(64, 64, 4)).astype(np.uint32)
  out, = oracle.downsample_segmentation(img, (2, 2, 1))
  for x in range(32):
    for y in range(32):
      blk = img[2 * x:2 * x + 2, 2 * y:2 * y + 2, 0].ravel()
      vals, cnt = np.unique(blk, return_counts=True)
      if (cnt == cnt.max()).sum() == 1 and cnt.max() >= 2:
        assert out[x, y, 0] == vals[np.argmax(cnt)]


def test_mode_pool_checker_equals_striding(oracle):
  # pooling invariant: block-constant volume -> mode pool == stride-2 subsample
  data = np.zeros((128, 128, 4), dtype=np.uint8, order="F")
  i = 1
  for x in range(4):
    for y in range(4):
      data[32 * x:32 * (x + 1), 32 * y:32 * (y + 1), :] = i
      i += 1
  mips = oracle.downsample_segmentation(data, (2, 2, 1), num_mips=4)
  cur = data
  for m in mips:
    cur = cur[::2, ::2, :]
    assert np.array_equal(m, cur)


def test_mode_pool_odd_and_4d(oracle):
  rng = np.random.default_rng(1)
  img = rng.integers(0, 3, size=(5, 7, 3, 2)).astype(np.uint16)
  out, = oracle.downsample_segmentation(img, (2, 2, 1, 1))
  assert out.shape == (3, 4, 3, 2)
  assert out[2, 0, 0, 0] == img[4, 0, 0, 0]  # lone column -> a
  assert out[0, 3, 1, 1] == img[0, 6, 1, 1]  # lone row -> a


# ------------------------------------------------------------------- averaging
def test_avg_pool_rounding_probes(oracle):
  # first-contact probes of SURVEY.md 8(c): floor is the recalled upstream rule
  def one(a, b, c, d, rounding):
    img = np.array([[[a], [c]], [[b], [d]]], dtype=np.uint8)
    return int(oracle.downsample_with_averaging(img, (2, 2, 1), rounding=rounding)[0][0, 0, 0])
  assert one(0, 1, 1, 1, 0) == 0 and one(0, 1, 1, 1, 1) == 1
  assert one(0, 0, 1, 1, 0) == 0 and one(0, 0, 1, 1, 1) == 1 and one(0, 0, 1, 1, 2) == 0
  assert one(1, 1, 2, 2, 1) == 2 and one(1, 1, 2, 2, 2) == 2


def test_avg_pool_exact_sums_in_groups_of_four(oracle):
  rng = np.random.default_rng(2)
  img = rng.integers(0, 255, size=(64, 64, 2), dtype=np.uint8)
  mips = oracle.downsample_with_averaging(img, (2, 2, 1), num_mips=5)
  wide = img.astype(np.uint64)
  for k in range(1, 5):
    f = 2 ** k
    s = wide.reshape(64 // f, f, 64 // f, f, 2).sum(axis=(1, 3))
    assert np.array_equal(mips[k - 1], (s >> (2 * k)).astype(np.uint8))
  # level 5 restarts from the truncated level-4 values
  m4 = mips[3].astype(np.uint64)
  s = m4.reshape(2, 2, 2, 2, 2).sum(axis=(1, 3))
  assert np.array_equal(mips[4], (s >> 2).astype(np.uint8))


def test_avg_pool_constant_and_mirror(oracle):
  img = np.full((7, 5, 2), 200, dtype=np.uint8)
  for m in oracle.downsample_with_averaging(img, (2, 2, 1), num_mips=3):
    assert (m == 200).all()


# ------------------------------------------------------------------------- CCL
def _scipy_ccl(labels):
  from scipy import ndimage
  out = np.zeros(labels.shape, dtype=np.uint64)
  nxt = 0
  for l in np.unique(labels):
    if l == 0:
      continue
    cc, k = ndimage.label(labels == l)  # default structure = 6-connectivity
    out[cc > 0] = cc[cc > 0].astype(np.uint64) + np.uint64(nxt)
    nxt += k
  return out, nxt


@pytest.mark.parametrize("dtype", [np.uint8, np.uint32, np.uint64])
def test_ccl_matches_scipy_after_renumber(oracle, dtype):
  rng = np.random.default_rng(3)
  labels = rng.integers(0, 4, size=(33, 29, 17)).astype(dtype)
  cc, n = oracle.connected_components(labels, return_N=True)
  ref, n_ref = _scipy_ccl(labels)
  assert n == n_ref
  a, _ = oracle.renumber(cc)
  b, _ = oracle.renumber(ref)
  assert np.array_equal(a, b)
  # the oracle numbers components by first voxel in Fortran order:
  # renumber (first appearance) must be the identity.
  assert np.array_equal(a, cc)


def test_ccl_checker_kat(oracle):
  # test/test_ccl_tasks.py:20-30 checker volume (scaled down 4x): 128 blocks
  data = np.zeros((128, 128, 32), dtype=np.uint8)
  i = 1
  for x in range(8):
    for y in range(8):
      for z in range(2):
        data[16 * x:16 * (x + 1), 16 * y:16 * (y + 1), 16 * z:16 * (z + 1)] = i
        i += 1
  cc, n = oracle.connected_components(data, return_N=True)
  assert n == 128
  assert np.array_equal(np.unique(cc), np.arange(1, 129))
  _, counts = np.unique(cc, return_counts=True)
  assert (counts == 16 ** 3).all()
  # thresholded to one component (test_ccl_tasks.py:200-203): image <= 255
  cc, n = oracle.connected_components(data <= 255, return_N=True)
  assert n == 1
  # dust removes components with fewer than threshold voxels (:113,192-198)
  assert not oracle.dust(data, 16 ** 3 + 1).any()
  assert np.array_equal(oracle.dust(data, 16 ** 3), data)


def test_ccl_synthetic_voronoi(oracle):
  seg = oracle.synth_seg((64, 48, 40), pitch=16, num_ids=6)
  cc, n = oracle.connected_components(seg, return_N=True)
  ref, n_ref = _scipy_ccl(seg)
  assert n == n_ref and n > 6
  assert np.array_equal(oracle.renumber(cc)[0], oracle.renumber(ref)[0])


# ------------------------------------------------------------------- fastremap
def test_renumber_remap_unique(oracle):
  arr = np.array([[5, 0, 9], [9, 5, 7]], dtype=np.uint64, order="F")
  out, mapping = oracle.renumber(arr)
  assert mapping == {5: 1, 9: 2, 0: 0, 7: 3}  # F order: 5,9,0,5,9,7
  assert out.dtype == np.uint8
  assert np.array_equal(out, [[1, 0, 2], [2, 1, 3]])
  back = oracle.remap(out, {v: k for k, v in mapping.items()})
  assert np.array_equal(back, arr.astype(np.uint8))
  with pytest.raises(KeyError):
    oracle.remap(arr, {5: 1})
  u, c = oracle.unique(arr, return_counts=True)
  assert list(u) == [0, 5, 7, 9] and list(c) == [1, 2, 1, 2]
  assert oracle.inverse_component_map([1, 1, 2, 0], [4, 5, 4, 0]) == {0: [0], 1: [4, 5], 2: [4]}


# --------------------------------------------------------------- marching cubes
def _edge_manifold(faces):
  e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]])
  fwd = {}
  for a, b in e:
    fwd[(a, b)] = fwd.get((a, b), 0) + 1
  return all(fwd.get((b, a), 0) == c for (a, b), c in fwd.items())


def _signed_volume(v, f):
  a, b, c = v[f[:, 0]].astype(np.float64), v[f[:, 1]].astype(np.float64), v[f[:, 2]].astype(np.float64)
  return np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0


def test_mc_box_kat(oracle):
  # SURVEY.md 8(c) box KAT from test/test_tasks.py:413-415: n=62 voxel cube
  n = 62
  data = np.zeros((64, 64, 64), dtype=np.uint32, order="F")
  data[1:-1, 1:-1, 1:-1] = 1
  tl, tv = oracle.marching_cubes(data)
  assert (tl == 1).all()
  F = 12 * (n - 1) ** 2 + 24 * (n - 1) + 8
  assert len(tl) == F == 46124
  v, f = oracle.mesh_for_label(tl, tv, 1, resolution=(1, 1, 1), voxel_centered=False)
  assert len(v) == F // 2 + 2 == 23064
  assert _edge_manifold(f)
  vol = _signed_volume(v, f)
  want = (n - 1) ** 3 + 3 * (n - 1) ** 2 + 1.5 * (n - 1) + 1.0 / 6.0
  assert abs(abs(vol) - want) < 1e-6 * want
  assert vol > 0, "triangles must wind counter-clockwise seen from outside"
  assert np.allclose(v * 2, np.round(v * 2))  # half-voxel lattice


def test_mc_watertight_random_multilabel(oracle):
  # every label's surface must be closed and consistently oriented when the
  # volume is zero padded: validates the 256-case table in all configurations
  rng = np.random.default_rng(5)
  data = np.zeros((18, 17, 16), dtype=np.uint8, order="F")
  data[1:-1, 1:-1, 1:-1] = rng.integers(0, 4, size=(16, 15, 14))
  tl, tv = oracle.marching_cubes(data)
  for l in (1, 2, 3):
    v, f = oracle.mesh_for_label(tl, tv, l, voxel_centered=False)
    assert len(f) > 0 and _edge_manifold(f)
    assert _signed_volume(v, f) > 0


def test_mc_all_256_cases_closed(oracle):
  # each cube configuration on its own, embedded in a zero volume
  for case in range(1, 255):
    data = np.zeros((4, 4, 4), dtype=np.uint8, order="F")
    for k, (cx, cy, cz) in enumerate([(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0),
                                      (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]):
      if (case >> k) & 1:
        data[1 + cx, 1 + cy, 1 + cz] = 1
    tl, tv = oracle.marching_cubes(data)
    v, f = oracle.mesh_for_label(tl, tv, 1, voxel_centered=False)
    assert _edge_manifold(f), case
    assert _signed_volume(v, f) > 0, case


def test_canonicalise_mesh_is_order_independent(oracle):
  data = np.zeros((8, 8, 8), dtype=np.uint8, order="F")
  data[2:6, 2:5, 3:6] = 1
  tl, tv = oracle.marching_cubes(data)
  v, f = oracle.mesh_for_label(tl, tv, 1)
  rng = np.random.default_rng(0)
  perm = rng.permutation(len(v))
  inv = np.argsort(perm)
  f2 = inv[f][rng.permutation(len(f))]
  f2 = np.roll(f2, 1, axis=1)
  v1, c1 = oracle.canonicalise_mesh(v, f)
  v2, c2 = oracle.canonicalise_mesh(v[perm], f2)
  assert np.array_equal(v1, v2) and np.array_equal(c1, c2)


def test_block_pooling_rules(oracle):
  """2x2x2 (and other 1/2 factors) mode / average: the restated rules on hand-made blocks."""
  blk = lambda vals, dt=np.uint32: np.array(vals, dtype=dt).reshape(2, 2, 2, order="F")
  mode = lambda a, **kw: int(oracle.downsample_segmentation(a, (2, 2, 2), **kw)[0][0, 0, 0])
  assert mode(blk([1, 2, 3, 4, 5, 6, 7, 8])) == 1          # all distinct -> earliest
  assert mode(blk([1, 2, 2, 1, 3, 3, 3, 4])) == 3          # plain majority
  assert mode(blk([5, 6, 6, 5, 7, 7, 8, 9])) == 5          # 2-2-2 tie -> earliest sample
  assert mode(blk([0, 0, 0, 0, 0, 9, 9, 4])) == 0          # dense counts zeros
  assert mode(blk([0, 0, 0, 0, 0, 4, 9, 9]), sparse=True) == 9
  assert mode(blk([0] * 8), sparse=True) == 0
  avg = lambda a, r=0: int(oracle.downsample_with_averaging(a, (2, 2, 2), rounding=r)[0][0, 0, 0])
  a = blk([1, 2, 3, 4, 5, 6, 7, 8], np.uint8)               # sum 36 -> 4.5
  assert (avg(a, 0), avg(a, 1), avg(a, 2)) == (4, 5, 4)
  assert avg(blk([255] * 8, np.uint8)) == 255
  # odd extents: the lone slice counts twice (divisor stays 8)
  odd = np.asfortranarray(np.arange(27, dtype=np.uint16).reshape(3, 3, 3, order="F"))
  out = oracle.downsample_with_averaging(odd, (2, 2, 2))[0]
  assert out.shape == (2, 2, 2) and out[1, 1, 1] == 26 and out[0, 0, 0] == (0 + 1 + 3 + 4 + 9 + 10 + 12 + 13) // 8
  # (2,2,1) through the block rule equals the 2x2x1 kernel, dense and sparse
  rng = np.random.default_rng(5)
  vol = np.asfortranarray(rng.integers(0, 4, size=(21, 10, 3)).astype(np.uint16))
  for sp in (0, 1):
    a = oracle.downsample_segmentation(vol, (2, 2, 1), num_mips=2, sparse=bool(sp))
    b = oracle._block_pool(vol, (2, 2, 1), 2, "mode", sp)
    assert all(np.array_equal(p, q) for p, q in zip(a, b))
  # invariants: constant volume stays constant; checker of 4^3 blocks subsamples exactly
  const = np.full((9, 8, 7), 77, dtype=np.uint8, order="F")
  for m in oracle.downsample_with_averaging(const, (2, 2, 2), num_mips=3) + \
      oracle.downsample_segmentation(const, (2, 2, 2), num_mips=3):
    assert (m == 77).all()
  x, y, z = np.meshgrid(np.arange(16), np.arange(16), np.arange(16), indexing="ij")
  checker = np.asfortranarray((1 + x // 4 + 4 * (y // 4) + 16 * (z // 4)).astype(np.uint32))
  cur = checker
  for m in oracle.downsample_segmentation(checker, (2, 2, 2), num_mips=2):
    cur = cur[::2, ::2, ::2]
    assert np.array_equal(m, cur)
  # sparse average: mean of the non-zero samples
  sp = lambda a, r=0: int(oracle.downsample_with_averaging(a, (2, 2, 2), sparse=True, rounding=r)[0][0, 0, 0])
  b = blk([0, 0, 0, 0, 0, 10, 11, 14], np.uint8)             # 35 / 3 = 11.67
  assert (sp(b, 0), sp(b, 1), sp(b, 2)) == (11, 12, 12)
  assert sp(blk([0] * 8, np.uint8)) == 0
  assert sp(blk([0, 0, 0, 0, 0, 0, 7, 8], np.uint8), 2) == 8  # 7.5 -> even
  q = np.array([[0, 5], [6, 0]], dtype=np.uint16).reshape(2, 2, 1, order="F")
  assert int(oracle.downsample_with_averaging(q, (2, 2, 1), sparse=True)[0][0, 0, 0]) == 5


# ------------------------------------------------------------------ simplifier
def _edge_use(faces):
  e = np.concatenate([faces[:, [0, 1]], faces[:, [1, 2]], faces[:, [2, 0]]]).astype(np.int64)
  e.sort(axis=1)
  _, counts = np.unique(e, axis=0, return_counts=True)
  return counts


def _signed_volume(v, f):
  a, b, c = (v[f[:, k]].astype(np.float64) for k in range(3))
  return float(np.einsum("ij,ij->i", a, np.cross(b, c)).sum() / 6.0)


def test_simplify_box_stays_closed_genus0(oracle):
  """The reference's own mesh test volume (test/test_tasks.py:413-415): simplification keeps
  the surface closed, manifold, genus 0 and outward oriented, moves no vertex further than the
  error bound allows and reaches the face target or a fixed point."""
  labels = np.zeros((64, 64, 64), dtype=np.uint32, order="F")
  labels[1:-1, 1:-1, 1:-1] = 7
  tl, tv = oracle.marching_cubes(labels)
  W = oracle.WeldedMeshes(tl, tv)
  v0, f0 = W.get(7)
  out, rounds = oracle.simplify_welded(W, resolution=(1, 1, 1), reduction_factor=100, max_error=40.0)
  v, f = out[7]
  assert len(f0) == 46124 and len(v0) == 23064
  assert 0 < rounds <= 400
  assert len(f) <= len(f0) // 20                      # heavily decimated (flat faces collapse freely)
  assert (_edge_use(f) == 2).all()                    # closed 2-manifold edges
  assert len(v) - len(f) * 3 // 2 + len(f) == 2       # Euler characteristic of a sphere
  assert f.max() == len(v) - 1 and len(np.unique(f)) == len(v)
  vol0, vol = _signed_volume(v0, f0), _signed_volume(v, f)
  assert vol0 > 0 and vol > 0 and abs(vol - vol0) / vol0 < 0.02
  lo, hi = v.min(axis=0), v.max(axis=0)
  assert np.all(lo >= v0.min(axis=0) - 1e-3) and np.all(hi <= v0.max(axis=0) + 1e-3)


def test_simplify_is_deterministic_and_label_order_free(oracle):
  """Same input -> same output; a label's result does not depend on which other labels
  share the task (keys are label-local)."""
  rng = np.random.default_rng(3)
  seg = oracle.synth_seg((48, 40, 36), pitch=16, num_ids=6).astype(np.uint32)
  tl, tv = oracle.marching_cubes(seg)
  W = oracle.WeldedMeshes(tl, tv)
  a, ra = oracle.simplify_welded(W, resolution=(16, 16, 40), reduction_factor=10, max_error=40.0)
  b, rb = oracle.simplify_welded(W, resolution=(16, 16, 40), reduction_factor=10, max_error=40.0)
  assert ra == rb and a.keys() == b.keys()
  for k in a:
    assert np.array_equal(a[k][0], b[k][0]) and np.array_equal(a[k][1], b[k][1])
  # one label alone: relabel everything else to background
  lab = int(W.labels[len(W.labels) // 2])
  only = np.where(seg == lab, seg, 0).astype(np.uint32)
  tl1, tv1 = oracle.marching_cubes(only)
  W1 = oracle.WeldedMeshes(tl1, tv1)
  c, _ = oracle.simplify_welded(W1, resolution=(16, 16, 40), reduction_factor=10, max_error=40.0)
  assert np.array_equal(c[lab][0], a[lab][0]) and np.array_equal(c[lab][1], a[lab][1])
  del rng


def test_simplify_locks_open_boundaries_and_respects_zero_error(oracle):
  """A label cut by the task border is an open surface: its boundary vertices (edges used once)
  must survive untouched so that neighbouring tasks stitch; max_error=0 only removes
  zero-cost (coplanar) geometry, so the enclosed volume is exactly preserved."""
  labels = np.zeros((24, 24, 24), dtype=np.uint32, order="F")
  labels[4:20, 4:20, :] = 3                       # a bar that leaves the block through both z faces
  tl, tv = oracle.marching_cubes(labels)
  W = oracle.WeldedMeshes(tl, tv)
  v0, f0 = W.get(3)
  e = np.concatenate([f0[:, [0, 1]], f0[:, [1, 2]], f0[:, [2, 0]]]).astype(np.int64)
  e.sort(axis=1)
  ue, cnt = np.unique(e, axis=0, return_counts=True)
  boundary = np.unique(ue[cnt == 1])
  assert len(boundary) > 0
  out, _ = oracle.simplify_welded(W, reduction_factor=100, max_error=40.0)
  v, f = out[3]
  assert len(f) < len(f0) // 4
  have = {tuple(p) for p in np.round(v * 2).astype(np.int64)}
  assert all(tuple(p) in have for p in np.round(v0[boundary] * 2).astype(np.int64))
  assert (_edge_use(f) <= 2).all() and (_edge_use(f) == 1).sum() == (cnt == 1).sum()
  # zero error budget on a closed box: only coplanar collapses, volume exactly kept
  box = np.zeros((20, 20, 20), dtype=np.uint32, order="F")
  box[2:-2, 2:-2, 2:-2] = 9
  tlb, tvb = oracle.marching_cubes(box)
  Wb = oracle.WeldedMeshes(tlb, tvb)
  vb0, fb0 = Wb.get(9)
  outb, _ = oracle.simplify_welded(Wb, reduction_factor=100, max_error=0.0)
  vb, fb = outb[9]
  assert len(fb) < len(fb0) and (_edge_use(fb) == 2).all()
  assert abs(_signed_volume(vb, fb) - _signed_volume(vb0, fb0)) < 1e-6 * abs(_signed_volume(vb0, fb0))


# ------------------------------------------------- compressed_segmentation codec
def test_cseg_format_kats(oracle):
  """Hand-checked streams of the compressed_segmentation layout: offsets are
  u32 words relative to the channel start, header = [table offset | bits << 24, indices offset]."""
  one = np.full((8, 8, 8), 7, dtype=np.uint32, order="F")
  # single value: 0 bits, no index words, the table follows the header directly
  assert oracle.cseg_encode(one).tolist() == [1, 2 | (0 << 24), 2, 7]
  two = one.copy()
  two[0, 0, 0] = 9
  w = oracle.cseg_encode(two)
  # 1 bit per voxel: 512 bits = 16 index words at offset 2, sorted table [7, 9] at offset 18
  assert len(w) == 1 + 2 + 16 + 2
  assert w[1] == (18 | (1 << 24)) and w[2] == 2
  assert w[3] == 1 and not w[4:19].any() and w[19:].tolist() == [7, 9]
  # uint64 labels: the table holds (low, high) word pairs
  big = np.full((8, 8, 8), (5 << 32) | 3, dtype=np.uint64, order="F")
  assert oracle.cseg_encode(big).tolist() == [1, 2, 2, 3, 5]
  # two blocks with the same label set share one table; voxel (1,0,0) of block 1 is bit 1
  pair = np.full((16, 8, 8), 4, dtype=np.uint32, order="F")
  pair[0, 0, 0] = 6
  pair[9, 0, 0] = 6
  w = oracle.cseg_encode(pair)
  t0, t1 = w[1] & 0xFFFFFF, w[3] & 0xFFFFFF
  assert t0 == t1 and (w[1] >> 24) == 1 and (w[3] >> 24) == 1
  assert w[1 + w[2]] == 1 and w[1 + w[4]] == 2
  assert len(w) == 1 + 4 + 16 + 2 + 16


@pytest.mark.parametrize("dtype", [np.uint32, np.uint64])
def test_cseg_roundtrip(oracle, dtype):
  rng = np.random.default_rng(12)
  vols = [oracle.synth_seg((64, 64, 64), pitch=16, num_ids=300).astype(dtype),
          rng.integers(0, 1 << 20, size=(17, 9, 5)).astype(dtype),            # ragged edges, 16/32-bit indices
          rng.integers(0, 3, size=(8, 8, 8, 3)).astype(dtype),                 # three channels, 2-bit indices
          np.zeros((5, 5, 5), dtype=dtype)]
  if dtype == np.uint64:
    vols[0] = vols[0] + np.uint64(1 << 40)
  for v in vols:
    for bs in ((8, 8, 8), (4, 4, 2)):
      w = oracle.cseg_encode(v, bs)
      back = oracle.cseg_decode(w, v.shape, dtype, bs)
      assert np.array_equal(back.reshape(v.shape), v)
  # the real segmentation compresses: 64^3 u32 = 262144 words raw
  assert len(oracle.cseg_encode(vols[0])) < 64 ** 3 // 3
  with pytest.raises(ValueError):
    oracle.cseg_decode(np.array([1, 0xFF000000, 2], dtype=np.uint32), (8, 8, 8), dtype)
