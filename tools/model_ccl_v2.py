"""Lane-level numpy model of local_tile_v2 (igneous_b200/csrc/ccl.cu): every statement of
passes 1-3 is mirrored on 32-wide lane vectors (ballot / shfl emulated), run on random
tiles and checked against the oracle's CCL of the tile.  This validates the mask and
index arithmetic of the kernel; it cannot validate CUDA-specific behaviour.
usage: python tools/model_ccl_v2.py [n_tiles]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle  # noqa: E402

TILE_X, TILE_Y, TILE_Z, QUAD, TASKS_PER_WARP = 256, 8, 8, 128, 128
BG = 0xFFFFFFFF
LANES = np.arange(32)


def ballot(pred):
  return int(sum(1 << int(l) for l in LANES[np.asarray(pred, bool)]))


def clz(x):
  return 32 - int(x).bit_length()


def popc(x):
  return bin(int(x)).count("1")


def shfl(v, src):
  return np.asarray(v)[np.asarray(src) % 32]


def shfl_up1(v):
  out = np.asarray(v).copy()
  out[1:] = np.asarray(v)[:-1]
  return out


def find(L, i):
  while L[i] != i:
    L[i] = L[L[i]]
    i = L[i]
  return i


def union(L, a, b):
  a, b = find(L, a), find(L, b)
  if a == b:
    return
  if a < b:
    a, b = b, a
  L[a] = b


def run_tile(tile, tasks_cap=TASKS_PER_WARP, starts_cap=128):
  """tile: [256, 8, 8] labels (x, y, z).  Returns parent (local root index or BG) per voxel
  and whether the overflow path was taken."""
  lin = lambda x, y, z: (z * TILE_Y + y) * TILE_X + x
  L = np.zeros(TILE_X * TILE_Y * TILE_Z, dtype=np.int64)
  queues, start_lists, overflow_any, soverflow_any = [], [], False, False
  for warp in range(16):
    lz, ly0 = warp >> 1, (warp & 1) * 4
    q, sq, overflow, soverflow = [], [], False, False
    prev = [None, None]
    for rr in range(4):
      ly = ly0 + rr
      r = lz * TILE_Y + ly
      carry, prev_last, cprev = 0, 0, 0
      for qd in range(2):
        xs = QUAD * qd + 4 * LANES
        a = np.stack([tile[xs + j, ly, lz] for j in range(4)])          # a[j][lane]
        if rr > 0:
          y = prev[qd]
        elif ly > 0:
          y = np.stack([tile[xs + j, ly - 1, lz] for j in range(4)])
        else:
          y = np.zeros_like(a)
        z = np.stack([tile[xs + j, ly, lz - 1] for j in range(4)]) if lz > 0 else np.zeros_like(a)
        left = shfl_up1(a[3])
        left[0] = prev_last
        nz = a != 0
        same = np.zeros_like(nz)
        same[0] = nz[0] & (a[0] == left)
        for j in range(1, 4):
          same[j] = nz[j] & (a[j] == a[j - 1])
        st = nz & ~same
        base = r * TILE_X + QUAD * qd + 4 * LANES
        has = st.any(axis=0)
        last_local = base + np.where(st[3], 3, np.where(st[2], 2, np.where(st[1], 1, 0)))
        m_has = ballot(has)
        incoming = np.zeros(32, dtype=np.int64)
        for l in range(32):
          below = m_has & ((1 << l) - 1)
          incoming[l] = last_local[31 - clz(below)] if below else carry
        c = np.zeros((4, 32), dtype=np.int64)
        c[0] = np.where(st[0], base, incoming)
        for j in range(1, 4):
          c[j] = np.where(st[j], base + j, c[j - 1])
        for j in range(4):
          L[base + j] = np.where(nz[j], c[j], BG)
        if m_has:
          carry = int(last_local[31 - clz(m_has)])
        prev_last = int(a[3][31])
        cy, cz = nz & (a == y), nz & (a == z)
        pk = cy[3].astype(np.int64) | (cz[3].astype(np.int64) << 1)
        pl = shfl_up1(pk)
        pl[0] = cprev
        cprev = int(pk[31])
        ty = np.zeros(32, dtype=np.int64)
        tz = np.zeros(32, dtype=np.int64)
        for j in range(4):
          py = (pl & 1) != 0 if j == 0 else cy[j - 1]
          pz = (pl & 2) != 0 if j == 0 else cz[j - 1]
          ty |= (cy[j] & ~(same[j] & py)).astype(np.int64) << j
          tz |= (cz[j] & ~(same[j] & pz)).astype(np.int64) << j
        stn = sum(st[j].astype(np.int64) << j for j in range(4))
        cnt = np.array([popc(ty[l]) + popc(tz[l]) for l in range(32)]) | (np.array([popc(stn[l]) for l in range(32)]) << 16)
        m_t = ballot(cnt != 0)
        if m_t:
          off = np.zeros(32, dtype=np.int64)
          total = 0
          m = m_t
          while m:
            src = (m & -m).bit_length() - 1
            k = int(cnt[src])
            off[LANES > src] += k
            total += k
            m &= m - 1
          tt, ts = total & 0xFFFF, total >> 16
          if len(q) + tt > tasks_cap:
            overflow = True
          if len(sq) + ts > starts_cap:
            soverflow = True
          if not overflow:
            slots = [None] * tt
            for l in range(32):
              w = int(off[l] & 0xFFFF)
              for nib, delta in ((ty, TILE_X), (tz, TILE_X * TILE_Y)):
                for j in range(4):
                  if nib[l] & (1 << j):
                    assert slots[w] is None
                    slots[w] = (int(c[j][l]), int(base[l] + j - delta))
                    w += 1
            assert all(x is not None for x in slots)
            q.extend(slots)
          if not soverflow:
            slots = [None] * ts
            for l in range(32):
              w = int(off[l] >> 16)
              for j in range(4):
                if stn[l] & (1 << j):
                  assert slots[w] is None
                  slots[w] = int(base[l] + j)
                  w += 1
            assert all(x is not None for x in slots)
            sq.extend(slots)
        prev[qd] = a
    queues.append(q)
    start_lists.append(sq)
    overflow_any |= overflow
    soverflow_any |= soverflow
  if overflow_any:  # classic pass: every y / z adjacency (a superset of the needed unions)
    for zz in range(TILE_Z):
      for yy in range(TILE_Y):
        for xx in range(TILE_X):
          v = tile[xx, yy, zz]
          if v == 0:
            continue
          if yy > 0 and tile[xx, yy - 1, zz] == v:
            union(L, lin(xx, yy, zz), lin(xx, yy - 1, zz))
          if zz > 0 and tile[xx, yy, zz - 1] == v:
            union(L, lin(xx, yy, zz), lin(xx, yy, zz - 1))
  else:
    for q in queues:
      for a_, b_ in q:
        assert L[a_] != BG and L[b_] != BG
        union(L, a_, b_)
  out = np.full(L.shape, BG, dtype=np.int64)
  if not soverflow_any:  # pass 2b: run starts -> roots, then ONE hop per voxel
    listed = [x for sq in start_lists for x in sq]
    assert len(listed) == len(set(listed))
    for s0 in listed:
      rt = s0
      while L[rt] != rt:
        rt = L[rt]
      L[s0] = rt
    for i in range(len(L)):
      if L[i] != BG:
        out[i] = L[L[i]]
        assert L[out[i]] == out[i], "single hop did not reach a root"
  else:
    for i in range(len(L)):
      if L[i] != BG:
        out[i] = find(L, i)
  return out, overflow_any


def check(tile, **kw):
  parent, ovf = run_tile(tile, **kw)
  cc = oracle.connected_components(np.asfortranarray(tile.astype(np.uint32)))
  flat_cc = cc.ravel(order="F").astype(np.int64)
  fg = flat_cc != 0
  assert ((parent == BG) == ~fg).all()
  # same partition, and every root is the minimum index of its component
  first = {}
  for i in np.nonzero(fg)[0]:
    first.setdefault(flat_cc[i], i)
  want = np.array([first[l] for l in flat_cc[fg]])
  assert (parent[fg] == want).all()
  return ovf


def main():
  n = int(sys.argv[1]) if len(sys.argv) > 1 else 6
  rng = np.random.default_rng(0)
  seen_overflow = 0
  for k in range(n):
    kind = k % 4
    if kind == 0:      # blobs: long runs, few tasks
      tile = oracle.synth_seg((TILE_X, TILE_Y, TILE_Z), pitch=16, num_ids=5, seed=k).astype(np.int64)
    elif kind == 1:    # noise with 2 labels: many tasks -> overflow path
      tile = rng.integers(0, 3, size=(TILE_X, TILE_Y, TILE_Z))
    elif kind == 2:    # runs crossing lane / quad boundaries at every phase
      tile = np.zeros((TILE_X, TILE_Y, TILE_Z), dtype=np.int64)
      for yy in range(TILE_Y):
        for zz in range(TILE_Z):
          pos = 0
          while pos < TILE_X:
            ln = int(rng.integers(1, 40))
            tile[pos:pos + ln, yy, zz] = int(rng.integers(0, 3))
            pos += ln
    else:              # sparse tasks, queue never overflows
      tile = np.zeros((TILE_X, TILE_Y, TILE_Z), dtype=np.int64)
      tile[3:200, :, :] = 1
      tile[100:130, 2:5, 3:6] = 2
      tile[255, :, :] = 3
      tile[0, ::2, :] = 4
    seen_overflow += int(check(tile))
  print("model_ccl_v2: %d tiles ok (%d through the overflow path)" % (n, seen_overflow))


if __name__ == "__main__":
  main()
