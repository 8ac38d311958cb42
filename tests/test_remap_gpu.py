"""GPU parity: fastremap-equivalent kernels vs the numpy oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _labels(rng, shape, dtype, nlab=40, order="F"):
  small = rng.integers(0, nlab, size=tuple((s + 3) // 4 for s in shape))
  big = small
  for ax in range(len(shape)):
    big = np.repeat(big, 4, axis=ax)
  big = big[tuple(slice(0, s) for s in shape)]
  vals = rng.integers(1, np.iinfo(dtype).max, size=nlab, dtype=np.uint64)
  vals[0] = 0
  out = vals[big].astype(dtype)
  return np.asfortranarray(out) if order == "F" else np.ascontiguousarray(out)


@pytest.mark.parametrize("dtype", [np.uint8, np.uint16, np.uint32, np.uint64])
@pytest.mark.parametrize("order", ["F", "C"])
def test_renumber_matches_oracle(ctx, oracle, dtype, order):
  from igneous_b200 import fastremap
  rng = np.random.default_rng(1)
  arr = _labels(rng, (37, 29, 18), dtype, order=order)
  got, gmap = fastremap.renumber(arr)
  ref_in = arr if order == "F" else np.asfortranarray(arr.T)  # oracle scans Fortran order
  want, wmap = oracle.renumber(ref_in)
  if order == "C":
    want = np.ascontiguousarray(want.T)
  assert gmap == wmap
  assert got.dtype == want.dtype and np.array_equal(got, want)


def test_renumber_many_labels_and_4d(ctx, oracle):
  from igneous_b200 import fastremap
  rng = np.random.default_rng(2)
  arr = np.asfortranarray(rng.integers(0, 1 << 40, size=(33, 17, 9, 1), dtype=np.uint64))
  got, gmap = fastremap.renumber(arr, in_place=True)
  want, wmap = oracle.renumber(arr)
  assert gmap == wmap and np.array_equal(got, want) and got.shape == arr.shape
  assert got.dtype == np.uint16  # 5049 labels < 65536


def test_remap_and_keyerror(ctx, oracle):
  from igneous_b200 import fastremap
  rng = np.random.default_rng(3)
  arr = _labels(rng, (40, 20, 10), np.uint64)
  table = {int(u): int(i * 3 + 1) for i, u in enumerate(np.unique(arr))}
  got = fastremap.remap(arr, table)
  assert np.array_equal(got, oracle.remap(arr, table))
  work = arr.copy(order="F")
  assert fastremap.remap(work, table, in_place=True) is work and np.array_equal(work, got)
  missing = dict(list(table.items())[:-1])
  with pytest.raises(KeyError):
    fastremap.remap(arr, missing)
  kept = fastremap.remap(arr, missing, preserve_missing_labels=True)
  assert np.array_equal(kept, oracle.remap(arr, missing, preserve_missing_labels=True))


@pytest.mark.parametrize("dtype", [np.uint8, np.uint32, np.uint64])
def test_unique_mask(ctx, oracle, dtype):
  from igneous_b200 import fastremap
  rng = np.random.default_rng(4)
  arr = _labels(rng, (31, 30, 12), dtype, nlab=25)
  u, c = fastremap.unique(arr, return_counts=True)
  wu, wc = oracle.unique(arr, return_counts=True)
  assert np.array_equal(u, wu) and np.array_equal(c.astype(np.int64), wc)
  assert np.array_equal(fastremap.unique(arr), wu)
  some = [int(x) for x in wu[::3]]
  assert np.array_equal(fastremap.mask(arr, some), oracle.mask(arr, some))
  assert np.array_equal(fastremap.mask_except(arr, some), oracle.mask_except(arr, some))
  assert np.array_equal(fastremap.mask_except(arr, []), np.zeros_like(arr))


def test_inverse_component_map(ctx, oracle):
  from igneous_b200 import fastremap
  rng = np.random.default_rng(5)
  cur = rng.integers(0, 9, size=(64, 48)).astype(np.uint64) * np.uint64(1 << 33)
  prev = rng.integers(0, 7, size=(64, 48)).astype(np.uint64)
  assert fastremap.inverse_component_map(cur, prev) == oracle.inverse_component_map(cur, prev)
  assert fastremap.inverse_component_map([1, 1, 2, 0], [4, 5, 4, 0]) == {0: [0], 1: [4, 5], 2: [4]}


def test_fit_dtype():
  from igneous_b200 import fastremap
  assert fastremap.fit_dtype(np.uint64, 255) == np.uint8
  assert fastremap.fit_dtype(np.uint64, 256) == np.uint16
  assert fastremap.fit_dtype(np.uint64, 1 << 32) == np.uint64
