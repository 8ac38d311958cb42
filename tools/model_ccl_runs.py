"""Word-level numpy/python model of the run/bitmask CCL of csrc/ccl.cu (masks S, Z, Ey, Ez,
run ids = exclusive scan of popc(S), one union per stretch of E in which neither row starts
a run, roots ranked by run id).  Checks the mask arithmetic against the oracle on CPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def popc(x):
  return bin(int(x)).count("1")


def mask_le(p):
  return 0xFFFFFFFF >> (31 - p)


def ccl_runs(vol):
  sx, sy, sz = vol.shape
  wpr = (sx + 31) // 32
  W = wpr * sy * sz
  S = np.zeros(W, np.uint64); Z = np.zeros(W, np.uint64); Ey = np.zeros(W, np.uint64); Ez = np.zeros(W, np.uint64)
  for z in range(sz):
    for y in range(sy):
      row = vol[:, y, z]
      left = np.concatenate([[0], row[:-1]])
      up = vol[:, y - 1, z] if y > 0 else np.zeros_like(row)
      back = vol[:, y, z - 1] if z > 0 else np.zeros_like(row)
      nz = row != 0
      for xw in range(wpr):
        wi = (z * sy + y) * wpr + xw
        for b in range(32):
          x = xw * 32 + b
          if x >= sx or not nz[x]:
            continue
          Z[wi] |= 1 << b
          if row[x] != left[x]: S[wi] |= 1 << b
          if row[x] == up[x]: Ey[wi] |= 1 << b
          if row[x] == back[x]: Ez[wi] |= 1 << b
  rbase = np.zeros(W + 1, np.int64)
  for w in range(W):
    rbase[w + 1] = rbase[w] + popc(S[w])
  R = int(rbase[W])
  parent = list(range(R))

  def find(i):
    while parent[i] != i:
      parent[i] = parent[parent[i]]
      i = parent[i]
    return i

  def union(a, b):
    a, b = find(a), find(b)
    if a == b: return
    if a < b: a, b = b, a
    parent[a] = b

  def word_unions(E, prev31, Sw, Sn, base, nbase):
    cand = E & (Sw | Sn | (~((E << 1) | prev31) & 0xFFFFFFFF))
    for p in range(32):
      if (cand >> p) & 1:
        le = mask_le(p)
        a = base + popc(Sw & le) - 1
        b = nbase + popc(Sn & le) - 1
        assert 0 <= a < R and 0 <= b < R
        union(a, b)

  for z in range(sz):
    for y in range(sy):
      for xw in range(wpr):
        g = (z * sy + y) * wpr + xw
        if y > 0 and Ey[g]:
          gn = g - wpr
          word_unions(int(Ey[g]), (int(Ey[g - 1]) >> 31) if xw > 0 else 0, int(S[g]), int(S[gn]), int(rbase[g]), int(rbase[gn]))
        if z > 0 and Ez[g]:
          gn = g - sy * wpr
          word_unions(int(Ez[g]), (int(Ez[g - 1]) >> 31) if xw > 0 else 0, int(S[g]), int(S[gn]), int(rbase[g]), int(rbase[gn]))
  roots = [find(r) for r in range(R)]
  rank = {}
  for r in range(R):
    if roots[r] == r:
      rank[r] = len(rank)
  label = [rank[roots[r]] + 1 for r in range(R)]
  out = np.zeros(vol.shape, np.uint64, order="F")
  for z in range(sz):
    for y in range(sy):
      for x in range(sx):
        wi = (z * sy + y) * wpr + (x >> 5)
        b = x & 31
        if (int(Z[wi]) >> b) & 1:
          out[x, y, z] = label[int(rbase[wi]) + popc(int(S[wi]) & mask_le(b)) - 1]
  return out, len(rank)


if __name__ == "__main__":
  from oracle import oracle as O
  rng = np.random.default_rng(1)
  vols = [O.synth_seg((70, 9, 7), pitch=8, num_ids=4), rng.integers(0, 3, size=(67, 6, 5)).astype(np.uint32),
          rng.integers(0, 2, size=(33, 5, 4)).astype(np.uint32), O.synth_seg((100, 12, 6), pitch=16, num_ids=3)]
  for v in vols:
    v = np.asfortranarray(v)
    got, n = ccl_runs(v)
    want, wn = O.connected_components(v, return_N=True)
    assert n == wn and np.array_equal(got, want), v.shape
    print("ok", v.shape, n)
