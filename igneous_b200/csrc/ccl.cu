// ccl.cu -- 6-connected multi-label connected components (K3), dust (K4)
//
// Replaces cc3d.connected_components(labels, connectivity=6, out_dtype=uint64)
// and cc3d.dust as called from igneous/tasks/image/ccl.py:169-175,231-240,335-344,
// and fuses the surrounding passes of CCLFacesTask (threshold_image :89-101,
// blackout_non_face_rails :103-124, `+= label_offset`, `[labels==0] = 0`
// :174-175) into the same kernels.
//
// Algorithm (two-level run-based union-find, union by minimum voxel index):
//   L  local  : a CTA owns a 256x8x8 voxel tile.  Warps walk tile rows as 8
//               sub-words of 32 voxels (one voxel per lane: every global access
//               is a coalesced 128 B line for any row pitch -- igneous's own
//               task shape is 513^3).  Ballots over "differs from my left
//               neighbour" find the x-runs; y / z adjacencies between runs are
//               queued per warp and resolved 32 at a time on a union-find held
//               in SHARED memory; every voxel then stores the global index of
//               its tile-local root (dense 4 B/voxel write) and local roots are
//               logged as root candidates.  k_ccl_local_fast (raw labels) /
//               k_ccl_local (threshold_image + rails fused in the reader).
//   G  merge  : one CTA per tile gathers the (local root, local root) pairs that
//               meet across the tile's low x / y / z faces in a shared-memory
//               hash set and unites the unique pairs with an atomicMin
//               union-find in global memory (path halving).  k_ccl_merge_tiles.
//   R  roots  : candidates that are still their own parent are the component
//               minima; sorted, their rank+1 is cc3d's id (first voxel in
//               raster order); it is written back into the root's entry, flagged.
//   F  label  : every voxel chases parent links to a flagged root (8 independent
//               chases per lane), optionally through a lookup table (dust,
//               multi-slab and multi-GPU relabelling), and writes u16/u32/u64.
// HBM traffic: in (L) + 4 (L) + face rows (G) + 4 + out (F) ~ in + 8 + out
// bytes/voxel (measured 7.65 B/voxel for L with u32 input); algorithmic bytes
// (cc3d contract) = in + out.
#include <cub/device/device_radix_sort.cuh>

#include <stdlib.h>

#include <type_traits>
#include <vector>

#include "common.cuh"

namespace ign {

constexpr uint32_t CCL_BG = 0xFFFFFFFFu;
constexpr uint32_t CCL_FLAG = 0x80000000u;
constexpr unsigned FULL = 0xFFFFFFFFu;

// ----------------------------------------------------------------- reader
// How a voxel value is obtained: raw label, or threshold_image() -> {0,1};
// rails of the +1 overlap shell are blacked out on the fly.
template <typename T, bool THR>
struct Reader {
  using V = std::conditional_t<THR, uint32_t, T>;
  using value_type = T;
  static constexpr bool thresholded = THR;
  const T* in;
  double gte, lte;
  int use_gte, use_lte;
  uint32_t rx, ry, rz;  // rail coordinates (0xFFFFFFFF = none)
  __device__ __forceinline__ V at(uint32_t idx, uint32_t x, uint32_t y, uint32_t z) const {
    const T raw = in[idx];
    V v;
    if constexpr (THR) {
      bool ok = true;
      if constexpr (std::is_same<T, float>::value) {
        if (use_gte) ok = ok && (raw >= (float)gte);
        if (use_lte) ok = ok && (raw <= (float)lte);
      } else {
        if (use_gte) ok = ok && ((double)raw >= gte);
        if (use_lte) ok = ok && ((double)raw <= lte);
      }
      v = ok ? 1u : 0u;
    } else {
      v = raw;
    }
    const int on = (int)(x == rx) + (int)(y == ry) + (int)(z == rz);
    if (on >= 2) v = 0;
    return v;
  }
};

template <typename V>
__device__ __forceinline__ V shfl_up1(V v) {
  if constexpr (sizeof(V) <= 4) return (V)__shfl_up_sync(FULL, (uint32_t)v, 1);
  else return (V)__shfl_up_sync(FULL, (unsigned long long)v, 1);
}

// ------------------------------------------------------------- union-find
__device__ __forceinline__ uint32_t uf_find(volatile uint32_t* P, uint32_t i) {
  uint32_t cur = i, p = P[cur];
  while (p != cur) {
    const uint32_t gp = P[p];
    if (gp != p) P[cur] = gp;  // path halving; cur is not a root here
    cur = p;
    p = gp;
  }
  return cur;
}

__device__ __forceinline__ void uf_union(uint32_t* P, uint32_t a, uint32_t b) {
  while (true) {
    a = uf_find(P, a);
    b = uf_find(P, b);
    if (a == b) return;
    if (a < b) {
      const uint32_t t = a;
      a = b;
      b = t;
    }
    const uint32_t old = atomicMin(&P[a], b);  // hook the larger root under the smaller
    if (old == a) return;
    a = old;  // lost a race: a had a parent already; unite that with b
  }
}

// ------------------------------------------------------------------ tiling
// A CTA owns a TILE_X x TILE_Y x TILE_Z voxel tile; a warp walks whole tile
// rows as SUBW sub-words of 32 voxels (one voxel per lane per sub-word, so all
// global accesses are 128 B coalesced whatever the row pitch -- igneous's own
// task shape is 513^3 -- and SUBW independent loads are in flight per lane).
constexpr int TILE_X = 256, TILE_Y = 8, TILE_Z = 8;
constexpr int SUBW = TILE_X / 32;             // sub-words per tile row
constexpr int TILE_ROWS = TILE_Y * TILE_Z;    // 64
constexpr int TILE_VOX = TILE_X * TILE_ROWS;  // 16384 -> 64 KB of u32 parents
constexpr int CCL_THREADS = 512;
constexpr int CCL_WARPS = CCL_THREADS / 32;
constexpr int ROWS_PER_WARP = TILE_ROWS / CCL_WARPS;  // 4
constexpr int TASKS_PER_WARP = 128;
static_assert(TILE_X == 256 && TILE_Y == 8, "k_ccl_local_fast decodes local indices with shifts");           // queued (a<<16|b) union tasks, 14-bit local indices

struct TilePos {
  uint32_t X0, Y0, Z0, warp, lane;
};

__device__ __forceinline__ TilePos tile_pos(uint32_t ntx, uint32_t nty) {
  TilePos t;
  const uint32_t b = blockIdx.x;
  t.X0 = (b % ntx) * TILE_X;
  t.Y0 = ((b / ntx) % nty) * TILE_Y;
  t.Z0 = (b / (ntx * nty)) * TILE_Z;
  t.warp = threadIdx.x >> 5;
  t.lane = threadIdx.x & 31;
  return t;
}

__device__ __forceinline__ uint32_t seg_start_lane(uint32_t sm, uint32_t lane) {
  return 31 - __clz(sm & (0xFFFFFFFFu >> (31 - lane)));
}

// shared-memory union-find (same algorithm as the global one)
__device__ __forceinline__ uint32_t sm_find(volatile uint32_t* L, uint32_t i) {
  uint32_t cur = i, p = L[cur];
  while (p != cur) {
    const uint32_t gp = L[p];
    if (gp != p) L[cur] = gp;
    cur = p;
    p = gp;
  }
  return cur;
}
__device__ __forceinline__ void sm_union(uint32_t* L, uint32_t a, uint32_t b) {
  while (true) {
    a = sm_find(L, a);
    b = sm_find(L, b);
    if (a == b) return;
    if (a < b) {
      const uint32_t t = a;
      a = b;
      b = t;
    }
    const uint32_t old = atomicMin(&L[a], b);
    if (old == a) return;
    a = old;
  }
}

// -------------------------------------------------------- L: tile-local CCL
// Resolves every tile completely in shared memory and writes, for each voxel,
// the GLOBAL linear index of its tile-local root (background -> BG).  Local
// roots are logged as root candidates.
template <typename R>
__global__ void __launch_bounds__(CCL_THREADS)
    k_ccl_local(R rd, uint32_t sx, uint32_t sy, uint32_t sz, uint32_t ntx, uint32_t nty,
                uint32_t* __restrict__ parent, uint32_t* __restrict__ cand, uint32_t cand_cap,
                uint32_t* counters) {
  extern __shared__ uint32_t L[];
  uint32_t* tasks = L + TILE_VOX;
  using V = typename R::V;
  const TilePos t = tile_pos(ntx, nty);

  // step 1: segment starts
#pragma unroll 1
  for (int rr = 0; rr < ROWS_PER_WARP; rr++) {
    const uint32_t r = t.warp * ROWS_PER_WARP + rr;
    const uint32_t y = t.Y0 + (r % TILE_Y), z = t.Z0 + (r / TILE_Y);
    const bool rowok = (y < sy) && (z < sz);
    const uint32_t rowbase = (z * sy + y) * sx;
    V v[SUBW];
#pragma unroll
    for (int k = 0; k < SUBW; k++) {
      const uint32_t x = t.X0 + 32 * k + t.lane;
      v[k] = (rowok && x < sx) ? rd.at(rowbase + x, x, y, z) : (V)0;
    }
#pragma unroll
    for (int k = 0; k < SUBW; k++) {
      const V vl = shfl_up1(v[k]);
      const bool start = (v[k] != 0) && (t.lane == 0 || vl != v[k]);
      const uint32_t sm = __ballot_sync(FULL, start);
      const uint32_t li = r * TILE_X + 32 * k + t.lane;
      L[li] = (v[k] != 0) ? (r * TILE_X + 32 * k + seg_start_lane(sm, t.lane)) : CCL_BG;
    }
  }
  __syncthreads();

  // step 2: unions inside the tile.  Union tasks (a,b) are queued per warp in
  // shared memory and executed 32 at a time: the dependent-load chains of one
  // sub-word's unions would otherwise run back to back on a single lane.
  uint32_t* q = tasks + t.warp * TASKS_PER_WARP;
  uint32_t nq = 0;
  auto flush = [&]() {
    __syncwarp();
    for (uint32_t i = t.lane; i < nq; i += 32) {
      const uint32_t ab = q[i];
      sm_union(L, ab >> 16, ab & 0xFFFFu);
    }
    __syncwarp();
    nq = 0;
  };
#pragma unroll 1
  for (int rr = 0; rr < ROWS_PER_WARP; rr++) {
    const uint32_t r = t.warp * ROWS_PER_WARP + rr;
    const uint32_t ly = r % TILE_Y, lz = r / TILE_Y;
    const uint32_t y = t.Y0 + ly, z = t.Z0 + lz;
    if (!((y < sy) && (z < sz))) continue;
    const uint32_t rowbase = (z * sy + y) * sx;
    const uint32_t sxy = sx * sy;
    V v[SUBW], vy[SUBW], vz[SUBW];
#pragma unroll
    for (int k = 0; k < SUBW; k++) {
      const uint32_t x = t.X0 + 32 * k + t.lane;
      const bool inb = x < sx;
      v[k] = inb ? rd.at(rowbase + x, x, y, z) : (V)0;
      vy[k] = (inb && ly > 0) ? rd.at(rowbase + x - sx, x, y - 1, z) : (V)0;
      vz[k] = (inb && lz > 0) ? rd.at(rowbase + x - sxy, x, y, z - 1) : (V)0;
    }
    V prev_last = 0;
#pragma unroll
    for (int k = 0; k < SUBW; k++) {
      const V vl = shfl_up1(v[k]);
      const bool same_left = (t.lane > 0) && (v[k] == vl);
      const bool start = (v[k] != 0) && !same_left;
      const uint32_t sm = __ballot_sync(FULL, start);
      const bool cy = (v[k] != 0) && (v[k] == vy[k]);
      const bool cz = (v[k] != 0) && (v[k] == vz[k]);
      const bool cy_l = __shfl_up_sync(FULL, (int)cy, 1) != 0;
      const bool cz_l = __shfl_up_sync(FULL, (int)cz, 1) != 0;
      const bool tx = (t.lane == 0) && (k > 0) && (v[k] != 0) && (v[k] == prev_last);
      const bool ty = cy && !(same_left && cy_l);
      const bool tz = cz && !(same_left && cz_l);
      const uint32_t li = r * TILE_X + 32 * k + t.lane;
      const uint32_t node = r * TILE_X + 32 * k + seg_start_lane(sm | 1u, t.lane);
      const uint32_t my = __ballot_sync(FULL, ty), mz = __ballot_sync(FULL, tz);
      const uint32_t mx = __ballot_sync(FULL, tx);
      const uint32_t below = (1u << t.lane) - 1u;
      if (nq + __popc(my) + __popc(mz) + __popc(mx) > TASKS_PER_WARP) flush();
      if (tx) q[nq] = (li << 16) | (li - 1);
      uint32_t o = nq + __popc(mx);
      if (ty) q[o + __popc(my & below)] = (node << 16) | (li - TILE_X);
      o += __popc(my);
      if (tz) q[o + __popc(mz & below)] = (node << 16) | (li - TILE_X * TILE_Y);
      nq = o + __popc(mz);
      if constexpr (sizeof(V) <= 4) prev_last = (V)__shfl_sync(FULL, (uint32_t)v[k], 31);
      else prev_last = (V)__shfl_sync(FULL, (unsigned long long)v[k], 31);
    }
  }
  flush();
  __syncthreads();

  // step 3: flatten, translate to global indices, log local roots
#pragma unroll 1
  for (int rr = 0; rr < ROWS_PER_WARP; rr++) {
    const uint32_t r = t.warp * ROWS_PER_WARP + rr;
    const uint32_t y = t.Y0 + (r % TILE_Y), z = t.Z0 + (r / TILE_Y);
    if (!((y < sy) && (z < sz))) continue;
    const uint32_t rowbase = (z * sy + y) * sx;
#pragma unroll
    for (int k = 0; k < SUBW; k++) {
      const uint32_t x = t.X0 + 32 * k + t.lane;
      const uint32_t li = r * TILE_X + 32 * k + t.lane;
      uint32_t p = L[li];
      bool is_root = false;
      uint32_t g = CCL_BG;
      if (p != CCL_BG) {
        is_root = (p == li);
        uint32_t cur = p;
        while (true) {  // read-only chase: no writer after the barrier
          const uint32_t q = L[cur];
          if (q == cur) break;
          cur = q;
        }
        const uint32_t rr2 = cur / TILE_X, lx = cur % TILE_X;
        g = ((t.Z0 + rr2 / TILE_Y) * sy + (t.Y0 + rr2 % TILE_Y)) * sx + t.X0 + lx;
      }
      if (x < sx) parent[rowbase + x] = g;
      const uint32_t cm = __ballot_sync(FULL, is_root);
      if (cm) {
        const int leader = __ffs(cm) - 1;
        uint32_t base = 0;
        if ((int)t.lane == leader) base = atomicAdd(&counters[0], (uint32_t)__popc(cm));
        base = __shfl_sync(FULL, base, leader);
        if (is_root) {
          const uint32_t pos = base + __popc(cm & ((1u << t.lane) - 1u));
          if (pos < cand_cap) cand[pos] = g;
          else counters[1] = 1;
        }
      }
    }
  }
}

// ------------------------------------------------ L (fast path): plain labels
// Same algorithm and same results as k_ccl_local, specialised for the common
// case (raw labels, no threshold, no rails): loads are pointer + immediate
// offset, interior tiles carry no bounds predicates, per-lane flags are derived
// from warp-uniform ballot masks, and x-runs continue across the 32-voxel
// sub-words of a tile row (no x-boundary unions inside a tile).  The generic
// kernel above was issue bound: 27 % IMAD + 20 % ISETP of 221 warp
// instructions per sub-word (profiles/r01_ccl_local_full_512_raw.csv).
template <typename T>
__device__ __forceinline__ T shfl_idx(T v, int src) {
  if constexpr (sizeof(T) <= 4) return (T)__shfl_sync(FULL, (uint32_t)v, src);
  else return (T)__shfl_sync(FULL, (unsigned long long)v, src);
}

template <typename T, bool FULLTILE>
__device__ __forceinline__ void local_tile_fast(const T* __restrict__ in, uint32_t sx, uint32_t sy,
                                                uint32_t sz, const TilePos& t, uint32_t* L,
                                                uint32_t* tasks, uint32_t* __restrict__ parent,
                                                uint32_t* __restrict__ cand, uint32_t cand_cap,
                                                uint32_t* counters) {
  const uint32_t sxy = sx * sy;
  const uint32_t lemask = 0xFFFFFFFFu >> (31 - t.lane);  // lanes <= mine
  const uint32_t ltmask = lemask >> 1;                   // lanes <  mine
  const uint32_t tile_g0 = (t.Z0 * sy + t.Y0) * sx + t.X0;

  // ---- phase 1: segment starts (runs continue across sub-words of the row)
#pragma unroll 1
  for (int rr = 0; rr < ROWS_PER_WARP; rr++) {
    const uint32_t r = t.warp * ROWS_PER_WARP + rr;
    const uint32_t ly = r % TILE_Y, lz = r / TILE_Y;
    const bool rowok = FULLTILE || ((t.Y0 + ly < sy) && (t.Z0 + lz < sz));
    const T* p = in + (tile_g0 + lz * sxy + ly * sx + t.lane);
    T v[SUBW];
#pragma unroll
    for (int k = 0; k < SUBW; k++)
      v[k] = (FULLTILE || (rowok && (t.X0 + 32 * k + t.lane < sx))) ? p[32 * k] : (T)0;
    uint32_t carry = 0;
    T prev_last = 0;
#pragma unroll
    for (int k = 0; k < SUBW; k++) {
      const T vl = shfl_up1(v[k]);
      const T v0 = shfl_idx(v[k], 0);
      const bool cont = (k > 0) && (v0 != 0) && (v0 == prev_last);
      const bool nz = v[k] != 0;
      const bool same = (t.lane > 0) ? (v[k] == vl) : cont;
      const uint32_t m_start = __ballot_sync(FULL, nz && !same);
      const uint32_t below = m_start & lemask;
      const uint32_t base = r * TILE_X + 32 * k;
      const uint32_t start = below ? (base + 31 - __clz(below)) : carry;
      L[base + t.lane] = nz ? start : CCL_BG;
      if (m_start) carry = base + 31 - __clz(m_start);
      prev_last = shfl_idx(v[k], 31);
    }
  }
  __syncthreads();

  // ---- phase 2: y / z unions, queued per warp and executed 32 wide
  uint32_t* q = tasks + t.warp * TASKS_PER_WARP;
  uint32_t nq = 0;
  auto flush = [&]() {
    __syncwarp();
    for (uint32_t i = t.lane; i < nq; i += 32) {
      const uint32_t ab = q[i];
      sm_union(L, ab >> 16, ab & 0xFFFFu);
    }
    __syncwarp();
    nq = 0;
  };
#pragma unroll 1
  for (int rr = 0; rr < ROWS_PER_WARP; rr++) {
    const uint32_t r = t.warp * ROWS_PER_WARP + rr;
    const uint32_t ly = r % TILE_Y, lz = r / TILE_Y;
    if (!FULLTILE && !((t.Y0 + ly < sy) && (t.Z0 + lz < sz))) continue;
    if (ly == 0 && lz == 0) continue;  // no in-tile neighbour below: nothing to unite
    const T* p = in + (tile_g0 + lz * sxy + ly * sx + t.lane);
    T v[SUBW], vy[SUBW], vz[SUBW];
#pragma unroll
    for (int k = 0; k < SUBW; k++) {
      const bool inb = FULLTILE || (t.X0 + 32 * k + t.lane < sx);
      v[k] = inb ? p[32 * k] : (T)0;
      vy[k] = (inb && ly > 0) ? *(p + 32 * k - sx) : (T)0;
      vz[k] = (inb && lz > 0) ? *(p + 32 * k - sxy) : (T)0;
    }
    T prev_last = 0;
    uint32_t cy_prev = 0, cz_prev = 0;  // bit 31 of the previous sub-word's masks
#pragma unroll
    for (int k = 0; k < SUBW; k++) {
      const T vl = shfl_up1(v[k]);
      const T v0 = shfl_idx(v[k], 0);
      const bool cont = (k > 0) && (v0 != 0) && (v0 == prev_last);
      const bool nz = v[k] != 0;
      const bool same = (t.lane > 0) ? (v[k] == vl) : cont;
      const uint32_t m_same = __ballot_sync(FULL, nz && same);
      const uint32_t m_cy = __ballot_sync(FULL, nz && (v[k] == vy[k]));
      const uint32_t m_cz = __ballot_sync(FULL, nz && (v[k] == vz[k]));
      // a union is needed where the connection starts or the x-run breaks
      const uint32_t t_y = m_cy & ~(m_same & ((m_cy << 1) | cy_prev));
      const uint32_t t_z = m_cz & ~(m_same & ((m_cz << 1) | cz_prev));
      const uint32_t both = t_y | t_z;
      if (both) {
        const uint32_t ny = __popc(t_y), nz_ = __popc(t_z);
        if (nq + ny + nz_ > TASKS_PER_WARP) flush();
        if ((both >> t.lane) & 1u) {
          const uint32_t li = r * TILE_X + 32 * k + t.lane;
          // any member of my run's set represents it: the entry written in phase 1
          // (or an ancestor another warp's path halving put there meanwhile)
          const uint32_t node = ((volatile uint32_t*)L)[li];
          if ((t_y >> t.lane) & 1u) q[nq + __popc(t_y & ltmask)] = (node << 16) | (li - TILE_X);
          if ((t_z >> t.lane) & 1u) q[nq + ny + __popc(t_z & ltmask)] = (node << 16) | (li - TILE_X * TILE_Y);
        }
        nq += ny + nz_;
      }
      prev_last = shfl_idx(v[k], 31);
      cy_prev = m_cy >> 31;
      cz_prev = m_cz >> 31;
    }
  }
  flush();
  __syncthreads();

  // ---- phase 3: flatten, translate to global indices, log local roots
#pragma unroll 1
  for (int rr = 0; rr < ROWS_PER_WARP; rr++) {
    const uint32_t r = t.warp * ROWS_PER_WARP + rr;
    const uint32_t ly = r % TILE_Y, lz = r / TILE_Y;
    if (!FULLTILE && !((t.Y0 + ly < sy) && (t.Z0 + lz < sz))) continue;
    uint32_t* out = parent + (tile_g0 + lz * sxy + ly * sx + t.lane);
#pragma unroll
    for (int k = 0; k < SUBW; k++) {
      const uint32_t li = r * TILE_X + 32 * k + t.lane;
      const uint32_t p0 = L[li];
      const bool bgv = (p0 == CCL_BG);
      // two unrolled hops cover almost every voxel after path halving (voxel -> run
      // start -> root); the loop only runs for the rare deeper chains
      uint32_t cur = bgv ? li : p0;
      uint32_t nxt = L[cur];
      if (__any_sync(FULL, !bgv && nxt != cur)) {
        while (!bgv && nxt != cur) {
          cur = nxt;
          nxt = L[cur];
        }
      }
      const uint32_t rr2 = cur >> 8;  // TILE_X == 256
      const uint32_t g = bgv ? CCL_BG : (tile_g0 + (rr2 >> 3) * sxy + (rr2 & 7) * sx + (cur & 255));
      if (FULLTILE || (t.X0 + 32 * k + t.lane < sx)) out[32 * k] = g;
      if (!bgv && p0 == li) {  // tile-local root (a handful per tile): log it as a root candidate
        const uint32_t pos = atomicAdd(&counters[0], 1u);
        if (pos < cand_cap) cand[pos] = g;
        else counters[1] = 1;
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(CCL_THREADS)
    k_ccl_local_fast(const T* __restrict__ in, uint32_t sx, uint32_t sy, uint32_t sz, uint32_t ntx,
                     uint32_t nty, uint32_t* __restrict__ parent, uint32_t* __restrict__ cand,
                     uint32_t cand_cap, uint32_t* counters) {
  extern __shared__ uint32_t L[];
  uint32_t* tasks = L + TILE_VOX;
  const TilePos t = tile_pos(ntx, nty);
  const bool full = (t.X0 + TILE_X <= sx) && (t.Y0 + TILE_Y <= sy) && (t.Z0 + TILE_Z <= sz);
  if (full) local_tile_fast<T, true>(in, sx, sy, sz, t, L, tasks, parent, cand, cand_cap, counters);
  else local_tile_fast<T, false>(in, sx, sy, sz, t, L, tasks, parent, cand, cand_cap, counters);
}

// --------------------------------------- L (v2, experimental): 4 voxels per lane
// Same results as k_ccl_local_fast.  k_ccl_local_fast is issue bound (137 warp
// instructions per 32-voxel sub-word, profiles/r01_ccl_local_fast_full_512_raw.csv);
// this variant cuts the instruction count per voxel:
//   * a lane owns FOUR consecutive voxels (one 128-bit load for u32): compares stay
//     per voxel, but shuffles, ballots, address arithmetic and shared-memory traffic
//     are paid once per 128 voxels instead of once per 32;
//   * ONE pass over the labels: the y neighbour row is the row the warp read just
//     before (registers) for 3 of its 4 rows, and the y / z union tasks are queued
//     while the run starts are written (a task only needs indices, not the parents);
//     they are executed after the barrier as before.
//   * the run starts are listed, pointed at their roots after the unions, and every
//     voxel then reaches its root with a single shared-memory hop (no chase loop).
// Requires full tiles and sx % 4 == 0 (vector loads); everything else takes the
// existing paths.  OPT-IN (IGN_CCL_V2=1): validated bit-exact on a B200 for u8 / u16 /
// u32 / u64 including partial tiles and the overflow paths (tools/check_ccl_v2.py,
// tests/test_ccl_gpu.py::test_ccl_v2_kernel_matches_oracle;
// tools/model_ccl_v2.py is a lane-level numpy model of the mask arithmetic) and 11 %
// faster than k_ccl_local_fast at 512^3, but the full GPU suite and the bench have not
// been run with it yet, so the default stays k_ccl_local_fast.
template <typename T>
struct Vec4Load;
template <>
struct Vec4Load<uint8_t> {
  static __device__ __forceinline__ void ld(const uint8_t* p, uint8_t (&a)[4]) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(p);
    a[0] = (uint8_t)(w & 0xFFu); a[1] = (uint8_t)((w >> 8) & 0xFFu);
    a[2] = (uint8_t)((w >> 16) & 0xFFu); a[3] = (uint8_t)(w >> 24);
  }
};
template <>
struct Vec4Load<uint16_t> {
  static __device__ __forceinline__ void ld(const uint16_t* p, uint16_t (&a)[4]) {
    const uint2 w = *reinterpret_cast<const uint2*>(p);
    a[0] = (uint16_t)(w.x & 0xFFFFu); a[1] = (uint16_t)(w.x >> 16);
    a[2] = (uint16_t)(w.y & 0xFFFFu); a[3] = (uint16_t)(w.y >> 16);
  }
};
template <>
struct Vec4Load<uint32_t> {
  static __device__ __forceinline__ void ld(const uint32_t* p, uint32_t (&a)[4]) {
    const uint4 w = *reinterpret_cast<const uint4*>(p);
    a[0] = w.x; a[1] = w.y; a[2] = w.z; a[3] = w.w;
  }
};
template <>
struct Vec4Load<uint64_t> {
  static __device__ __forceinline__ void ld(const uint64_t* p, uint64_t (&a)[4]) {
    const ulonglong2 w0 = *reinterpret_cast<const ulonglong2*>(p);
    const ulonglong2 w1 = *reinterpret_cast<const ulonglong2*>(p + 2);
    a[0] = w0.x; a[1] = w0.y; a[2] = w1.x; a[3] = w1.y;
  }
};

constexpr int QUAD = 128;                 // voxels a warp covers per step (4 per lane)
constexpr int QUADS = TILE_X / QUAD;      // 2
static_assert(ROWS_PER_WARP == 4 && TILE_Y == 8 && QUADS == 2, "local_tile_v2 row ownership");

constexpr int STARTS_PER_WARP = 128;      // run starts a warp can list for the compression pass

template <typename T>
__device__ __forceinline__ void local_tile_v2(const T* __restrict__ in, uint32_t sx, uint32_t sy,
                                              const TilePos& t, uint32_t* L, uint32_t* tasks, uint32_t* starts,
                                              uint32_t* __restrict__ parent, uint32_t* __restrict__ cand,
                                              uint32_t cand_cap, uint32_t* counters) {
  const uint32_t sxy = sx * sy;
  const uint32_t ltmask = (1u << t.lane) - 1u;  // lanes < mine
  const uint32_t tile_g0 = (t.Z0 * sy + t.Y0) * sx + t.X0;
  // a warp owns 4 consecutive y rows of one z slice of the tile
  const uint32_t lz = t.warp >> 1, ly0 = (t.warp & 1u) * ROWS_PER_WARP;
  uint32_t* q = tasks + t.warp * TASKS_PER_WARP;
  uint32_t* sq = starts + t.warp * STARTS_PER_WARP;
  uint32_t nq = 0, ns = 0;
  bool overflow = false, soverflow = false;

  // ---- pass 1: run starts + queued y / z union tasks + list of run starts
  T prev[QUADS][4];  // the row this warp handled before (y neighbour of the next one)
#pragma unroll
  for (int rr = 0; rr < ROWS_PER_WARP; rr++) {
    const uint32_t ly = ly0 + rr, r = lz * TILE_Y + ly;
    const T* row = in + (tile_g0 + lz * sxy + ly * sx + 4 * t.lane);
    uint32_t carry = 0;   // local index of the last run start seen in this row
    T prev_last = 0;      // last voxel of the previous quad (0 never equals a foreground label)
    uint32_t cprev = 0;   // bit 0 / 1: y / z connection of the previous quad's last voxel
#pragma unroll
    for (int qd = 0; qd < QUADS; qd++) {
      T a[4], y[4], z[4];
      Vec4Load<T>::ld(row + QUAD * qd, a);
      if (rr > 0) {
#pragma unroll
        for (int j = 0; j < 4; j++) y[j] = prev[qd][j];
      } else if (ly > 0) {
        Vec4Load<T>::ld(row + QUAD * qd - sx, y);
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++) y[j] = 0;
      }
      if (lz > 0) {
        Vec4Load<T>::ld(row + QUAD * qd - sxy, z);
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++) z[j] = 0;
      }
      T left = shfl_up1(a[3]);
      if (t.lane == 0) left = prev_last;
      bool nzv[4], same[4], st[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        nzv[j] = a[j] != 0;
        same[j] = nzv[j] && (a[j] == (j == 0 ? left : a[j - 1]));
        st[j] = nzv[j] && !same[j];
      }
      const uint32_t base = r * TILE_X + QUAD * qd + 4 * t.lane;  // local index of a[0]
      const bool has = st[0] || st[1] || st[2] || st[3];
      const uint32_t last_local = base + (st[3] ? 3u : (st[2] ? 2u : (st[1] ? 1u : 0u)));
      const uint32_t m_has = __ballot_sync(FULL, has);
      const uint32_t below = m_has & ltmask;
      uint32_t incoming = __shfl_sync(FULL, last_local, below ? (31 - __clz(below)) : 0);
      if (!below) incoming = carry;
      uint32_t c[4];
      c[0] = st[0] ? base : incoming;
#pragma unroll
      for (int j = 1; j < 4; j++) c[j] = st[j] ? (base + j) : c[j - 1];
      uint4 o;
      o.x = nzv[0] ? c[0] : CCL_BG;
      o.y = nzv[1] ? c[1] : CCL_BG;
      o.z = nzv[2] ? c[2] : CCL_BG;
      o.w = nzv[3] ? c[3] : CCL_BG;
      *reinterpret_cast<uint4*>(L + base) = o;
      if (m_has) carry = __shfl_sync(FULL, last_local, 31 - __clz(m_has));
      prev_last = shfl_idx(a[3], 31);

      // y / z connections: a union is needed where the connection starts or the x-run breaks
      bool cy[4], cz[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        cy[j] = nzv[j] && (a[j] == y[j]);
        cz[j] = nzv[j] && (a[j] == z[j]);
      }
      const uint32_t pk = (cy[3] ? 1u : 0u) | (cz[3] ? 2u : 0u);
      uint32_t pl = __shfl_up_sync(FULL, pk, 1);
      if (t.lane == 0) pl = cprev;
      cprev = __shfl_sync(FULL, pk, 31);
      uint32_t ty = 0, tz = 0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const bool py = (j == 0) ? ((pl & 1u) != 0) : cy[j - 1];
        const bool pz = (j == 0) ? ((pl & 2u) != 0) : cz[j - 1];
        if (cy[j] && !(same[j] && py)) ty |= 1u << j;
        if (cz[j] && !(same[j] && pz)) tz |= 1u << j;
      }
      // slots for this quad's tasks (low half-word) and run starts (high half-word): only the
      // few lanes that hold any take part in the reservation
      const uint32_t stn = (st[0] ? 1u : 0u) | (st[1] ? 2u : 0u) | (st[2] ? 4u : 0u) | (st[3] ? 8u : 0u);
      const uint32_t cnt = (uint32_t)(__popc(ty) + __popc(tz)) | ((uint32_t)__popc(stn) << 16);
      const uint32_t m_t = __ballot_sync(FULL, cnt != 0);
      if (m_t) {
        uint32_t off = 0, total = 0;
        for (uint32_t m = m_t; m; m &= m - 1) {
          const int src = __ffs(m) - 1;
          const uint32_t k = __shfl_sync(FULL, cnt, src);
          if ((int)t.lane > src) off += k;
          total += k;
        }
        const uint32_t tt = total & 0xFFFFu, ts = total >> 16;
        // a full queue before the barrier: the classic pass below redoes every union
        if (nq + tt > (uint32_t)TASKS_PER_WARP) overflow = true;
        // a full start list: pass 3 falls back to chasing every voxel
        if (ns + ts > (uint32_t)STARTS_PER_WARP) soverflow = true;
        if (!overflow) {
          uint32_t w = nq + (off & 0xFFFFu);
#pragma unroll
          for (int j = 0; j < 4; j++)
            if (ty & (1u << j)) q[w++] = (c[j] << 16) | (base + j - (uint32_t)TILE_X);
#pragma unroll
          for (int j = 0; j < 4; j++)
            if (tz & (1u << j)) q[w++] = (c[j] << 16) | (base + j - (uint32_t)(TILE_X * TILE_Y));
          nq += tt;
        }
        if (!soverflow) {
          uint32_t w = ns + (off >> 16);
#pragma unroll
          for (int j = 0; j < 4; j++)
            if (stn & (1u << j)) sq[w++] = base + j;
          ns += ts;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; j++) prev[qd][j] = a[j];
    }
  }
  const int any_overflow = __syncthreads_or(overflow ? 1 : 0);

  // ---- pass 2: execute the queued unions 32 wide on the shared-memory union-find
  auto flush = [&]() {
    __syncwarp();
    for (uint32_t i = t.lane; i < nq; i += 32) {
      const uint32_t ab = q[i];
      sm_union(L, ab >> 16, ab & 0xFFFFu);
    }
    __syncwarp();
    nq = 0;
  };
  if (any_overflow) {
    // rare (more than TASKS_PER_WARP tasks in 4 rows): drop the queues and redo the y / z
    // unions the classic way, 32 voxels at a time with a flush whenever the queue fills
    nq = 0;
    const uint32_t lemask = 0xFFFFFFFFu >> (31 - t.lane);
#pragma unroll 1
    for (int rr = 0; rr < ROWS_PER_WARP; rr++) {
      const uint32_t ly = ly0 + rr, r = lz * TILE_Y + ly;
      if (ly == 0 && lz == 0) continue;
      const T* p = in + (tile_g0 + lz * sxy + ly * sx + t.lane);
      T prev_last = 0;
      uint32_t cy_prev = 0, cz_prev = 0;
#pragma unroll 1
      for (int k = 0; k < SUBW; k++) {
        const T v = p[32 * k];
        const T vy = (ly > 0) ? *(p + 32 * k - sx) : (T)0;
        const T vz = (lz > 0) ? *(p + 32 * k - sxy) : (T)0;
        const T vl = shfl_up1(v);
        const T v0 = shfl_idx(v, 0);
        const bool cont = (k > 0) && (v0 != 0) && (v0 == prev_last);
        const bool nz = v != 0;
        const bool sm = (t.lane > 0) ? (v == vl) : cont;
        const uint32_t m_same = __ballot_sync(FULL, nz && sm);
        const uint32_t m_cy = __ballot_sync(FULL, nz && (v == vy));
        const uint32_t m_cz = __ballot_sync(FULL, nz && (v == vz));
        const uint32_t t_y = m_cy & ~(m_same & ((m_cy << 1) | cy_prev));
        const uint32_t t_z = m_cz & ~(m_same & ((m_cz << 1) | cz_prev));
        if (t_y | t_z) {
          const uint32_t ny = __popc(t_y), nz_ = __popc(t_z);
          if (nq + ny + nz_ > (uint32_t)TASKS_PER_WARP) flush();
          if (((t_y | t_z) >> t.lane) & 1u) {
            const uint32_t li = r * TILE_X + 32 * k + t.lane;
            const uint32_t node = ((volatile uint32_t*)L)[li];
            if ((t_y >> t.lane) & 1u) q[nq + __popc(t_y & (lemask >> 1))] = (node << 16) | (li - TILE_X);
            if ((t_z >> t.lane) & 1u) q[nq + ny + __popc(t_z & (lemask >> 1))] = (node << 16) | (li - TILE_X * TILE_Y);
          }
          nq += ny + nz_;
        }
        prev_last = shfl_idx(v, 31);
        cy_prev = m_cy >> 31;
        cz_prev = m_cz >> 31;
      }
    }
  }
  flush();
  const int any_soverflow = __syncthreads_or(soverflow ? 1 : 0);

  // ---- pass 2b: point every run start at its root.  Every node of a parent chain is a run
  // start (only roots are hooked, and a root is the first voxel of its run), so afterwards
  // any voxel reaches its root in ONE hop: parent entry -> L[entry].
  if (!any_soverflow) {
    for (uint32_t i = t.lane; i < ns; i += 32) {
      const uint32_t s0 = sq[i];
      uint32_t rt = s0, pp = ((volatile uint32_t*)L)[rt];
      while (pp != rt) {
        rt = pp;
        pp = ((volatile uint32_t*)L)[rt];
      }
      ((volatile uint32_t*)L)[s0] = rt;
    }
    __syncthreads();
  }

  // ---- pass 3: translate to global indices, log local roots (4 voxels per lane)
#pragma unroll 1
  for (int rr = 0; rr < ROWS_PER_WARP; rr++) {
    const uint32_t ly = ly0 + rr, r = lz * TILE_Y + ly;
    uint32_t* out = parent + (tile_g0 + lz * sxy + ly * sx + 4 * t.lane);
#pragma unroll
    for (int qd = 0; qd < QUADS; qd++) {
      const uint32_t base = r * TILE_X + QUAD * qd + 4 * t.lane;
      const uint4 P = *reinterpret_cast<const uint4*>(L + base);
      const uint32_t p0[4] = {P.x, P.y, P.z, P.w};
      uint32_t cur[4], nxt[4];
      bool bgv[4], more = false;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        bgv[j] = (p0[j] == CCL_BG);
        cur[j] = bgv[j] ? (base + j) : p0[j];
      }
#pragma unroll
      for (int j = 0; j < 4; j++) nxt[j] = L[cur[j]];
      if (!any_soverflow) {  // compressed: the entry's parent is the root
#pragma unroll
        for (int j = 0; j < 4; j++) cur[j] = bgv[j] ? cur[j] : nxt[j];
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++) more |= (!bgv[j] && nxt[j] != cur[j]);
        if (__any_sync(FULL, more)) {
#pragma unroll
          for (int j = 0; j < 4; j++)
            while (!bgv[j] && nxt[j] != cur[j]) {
              cur[j] = nxt[j];
              nxt[j] = L[cur[j]];
            }
        }
      }
      uint32_t g[4];
#pragma unroll
      for (int j = 0; j < 4; j++)
        g[j] = bgv[j] ? CCL_BG : (tile_g0 + (cur[j] >> 11) * sxy + ((cur[j] >> 8) & 7u) * sx + (cur[j] & 255u));
      *reinterpret_cast<uint4*>(out + QUAD * qd) = make_uint4(g[0], g[1], g[2], g[3]);
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (!bgv[j] && p0[j] == base + j) {  // tile-local root: log it as a root candidate
          const uint32_t pos = atomicAdd(&counters[0], 1u);
          if (pos < cand_cap) cand[pos] = g[j];
          else counters[1] = 1;
        }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(CCL_THREADS)
    k_ccl_local_v2(const T* __restrict__ in, uint32_t sx, uint32_t sy, uint32_t sz, uint32_t ntx,
                   uint32_t nty, uint32_t* __restrict__ parent, uint32_t* __restrict__ cand,
                   uint32_t cand_cap, uint32_t* counters) {
  extern __shared__ __align__(16) uint32_t Lv2[];
  uint32_t* tasks = Lv2 + TILE_VOX;
  uint32_t* starts = tasks + CCL_WARPS * TASKS_PER_WARP;
  const TilePos t = tile_pos(ntx, nty);
  const bool full = (t.X0 + TILE_X <= sx) && (t.Y0 + TILE_Y <= sy) && (t.Z0 + TILE_Z <= sz);
  if (full) local_tile_v2<T>(in, sx, sy, t, Lv2, tasks, starts, parent, cand, cand_cap, counters);
  else local_tile_fast<T, false>(in, sx, sy, sz, t, Lv2, tasks, parent, cand, cand_cap, counters);
}

// ------------------------------ L / G2 on 256 x 8 x 4 tiles (experiment, IGN_CCL_V2=3)
// k_ccl_local_v2 is no longer issue bound: two 512-thread CTAs per SM spend a large share
// of their time in block barriers and in the latency-bound shared-memory union phase
// (DESIGN.md section 8).  The same device code runs unchanged on half-height tiles with
// 256 threads (a warp still owns 4 consecutive y rows of one z slice): 40 KB of shared
// memory and 16 K registers per CTA let 4-5 CTAs in different phases share an SM, at the
// price of twice as many z faces to merge.  NOT VALIDATED ON A GPU YET (written after the
// round's GPU budget was spent); nothing selects it unless IGN_CCL_V2=3 is set.
constexpr int TILE_Z_S = 4;
constexpr int TILE_VOX_S = TILE_X * TILE_Y * TILE_Z_S;
constexpr int CCL_THREADS_S = 256;
constexpr int CCL_WARPS_S = CCL_THREADS_S / 32;
static_assert(TILE_Y * TILE_Z_S / CCL_WARPS_S == ROWS_PER_WARP, "small tiles keep 4 rows per warp");

__device__ __forceinline__ TilePos tile_pos_s(uint32_t ntx, uint32_t nty) {
  TilePos t;
  const uint32_t b = blockIdx.x;
  t.X0 = (b % ntx) * TILE_X;
  t.Y0 = ((b / ntx) % nty) * TILE_Y;
  t.Z0 = (b / (ntx * nty)) * TILE_Z_S;
  t.warp = threadIdx.x >> 5;
  t.lane = threadIdx.x & 31;
  return t;
}

template <typename T>
__global__ void __launch_bounds__(CCL_THREADS_S)
    k_ccl_local_v3(const T* __restrict__ in, uint32_t sx, uint32_t sy, uint32_t sz, uint32_t ntx,
                   uint32_t nty, uint32_t* __restrict__ parent, uint32_t* __restrict__ cand,
                   uint32_t cand_cap, uint32_t* counters) {
  extern __shared__ __align__(16) uint32_t Lv3[];
  uint32_t* tasks = Lv3 + TILE_VOX_S;
  uint32_t* starts = tasks + CCL_WARPS_S * TASKS_PER_WARP;
  const TilePos t = tile_pos_s(ntx, nty);
  const bool full = (t.X0 + TILE_X <= sx) && (t.Y0 + TILE_Y <= sy) && (t.Z0 + TILE_Z_S <= sz);
  if (full) local_tile_v2<T>(in, sx, sy, t, Lv3, tasks, starts, parent, cand, cand_cap, counters);
  else local_tile_fast<T, false>(in, sx, sy, sz, t, Lv3, tasks, parent, cand, cand_cap, counters);
}

// ------------------------------------------------- G: merges across tile faces
// Flat mapping over the voxel pairs that straddle a tile face, one 32-voxel
// sub-word per warp (y and z faces) or 32 rows per warp (x faces), so that the
// dependent global-memory chains of the unions are hidden by warp parallelism.
template <typename R>
__global__ void __launch_bounds__(256)
    k_ccl_merge(R rd, uint32_t sx, uint32_t sy, uint32_t sz, uint32_t nfy, uint32_t nfz,
                uint32_t nfx, uint64_t items_y, uint64_t items_z, uint64_t items_x,
                uint32_t* parent) {
  using V = typename R::V;
  const uint64_t wid = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t w32 = (sx + 31) / 32;
  const uint32_t sxy = sx * sy;
  if (wid < items_y + items_z) {
    uint32_t x0, y, z, stride;
    if (wid < items_y) {  // (k32, z, fy)
      x0 = (uint32_t)(wid % w32) * 32;
      const uint64_t r = wid / w32;
      z = (uint32_t)(r % sz);
      y = ((uint32_t)(r / sz) + 1) * TILE_Y;
      stride = sx;
    } else {  // (k32, y, fz)
      const uint64_t w2 = wid - items_y;
      x0 = (uint32_t)(w2 % w32) * 32;
      const uint64_t r = w2 / w32;
      y = (uint32_t)(r % sy);
      z = ((uint32_t)(r / sy) + 1) * TILE_Z;
      stride = sxy;
    }
    const uint32_t x = x0 + lane;
    const bool inb = x < sx;
    const uint32_t idx = (z * sy + y) * sx + x;
    const V v = inb ? rd.at(idx, x, y, z) : (V)0;
    const V vn = inb ? ((stride == sx) ? rd.at(idx - sx, x, y - 1, z) : rd.at(idx - sxy, x, y, z - 1))
                     : (V)0;
    const V vl = shfl_up1(v);
    // a 32-voxel sub-word may straddle a tile x face: segments break there too,
    // which only costs a redundant union.
    const bool same_left = (lane > 0) && (v == vl);
    const bool c = (v != 0) && (v == vn);
    const bool c_l = __shfl_up_sync(FULL, (int)c, 1) != 0;
    if (c && !(same_left && c_l)) uf_union(parent, idx, idx - stride);
  } else if (wid < items_y + items_z + items_x) {  // (row block of 32, fx)
    const uint64_t w2 = wid - items_y - items_z;
    const uint64_t nrows = (uint64_t)sy * sz;
    const uint64_t rb = (nrows + 31) / 32;
    const uint32_t fx = (uint32_t)(w2 / rb) + 1;
    const uint64_t row = (w2 % rb) * 32 + lane;
    if (row < nrows) {
      const uint32_t x = fx * TILE_X;
      const uint32_t y = (uint32_t)(row % sy), z = (uint32_t)(row / sy);
      const uint32_t idx = (uint32_t)row * sx + x;
      const V a = rd.at(idx, x, y, z);
      if (a != 0 && a == rd.at(idx - 1, x - 1, y, z)) uf_union(parent, idx, idx - 1);
    }
  }
  (void)nfy; (void)nfz; (void)nfx;
}

// ----------------------------------------- G2: face merges, deduplicated per tile
// One CTA per tile handles the tile's low y / z / x faces.  Along a face the same
// pair of tile-local components meets in up to 64 sub-words; instead of running
// a global union-find for each meeting, the (local root A, local root B) pairs
// are first collected in a shared-memory hash set and only the unique pairs are
// united in global memory.
constexpr int MERGE_SLOTS = 1024;  // power of two, 8 KB of u64

__device__ __forceinline__ void merge_emit(unsigned long long* set, uint32_t* parent, uint32_t a,
                                           uint32_t b) {
  const unsigned long long key = ((unsigned long long)a << 32) | b;
  uint32_t h = (uint32_t)((key * 0x9E3779B97F4A7C15ull) >> 54) & (MERGE_SLOTS - 1);
  for (int probe = 0; probe < 16; probe++) {
    const unsigned long long cur = atomicCAS(&set[h], 0xFFFFFFFFFFFFFFFFull, key);
    if (cur == 0xFFFFFFFFFFFFFFFFull || cur == key) return;
    h = (h + 1) & (MERGE_SLOTS - 1);
  }
  uf_union(parent, a, b);  // table crowded: unite directly
}

template <typename R>
__global__ void __launch_bounds__(256)
    k_ccl_merge_tiles(R rd, uint32_t sx, uint32_t sy, uint32_t sz, uint32_t ntx, uint32_t nty,
                      uint32_t* parent) {
  using V = typename R::V;
  __shared__ unsigned long long set[MERGE_SLOTS];
  for (int i = threadIdx.x; i < MERGE_SLOTS; i += blockDim.x) set[i] = 0xFFFFFFFFFFFFFFFFull;
  __syncthreads();
  const TilePos t = tile_pos(ntx, nty);
  const uint32_t sxy = sx * sy;
  const uint32_t nwarps = blockDim.x >> 5;
  // faces as (row list, neighbour stride): y face = rows (ly=0, lz=0..7); z face = rows (lz=0, ly=0..7)
  for (uint32_t item = t.warp; item < 2 * TILE_Y * SUBW; item += nwarps) {
    const uint32_t face = item / (TILE_Y * SUBW);      // 0: y face, 1: z face
    const uint32_t rowi = (item / SUBW) % TILE_Y, k = item % SUBW;
    const uint32_t ly = face == 0 ? 0 : rowi, lz = face == 0 ? rowi : 0;
    const uint32_t y = t.Y0 + ly, z = t.Z0 + lz;
    if (y >= sy || z >= sz) continue;
    if (face == 0 ? (y == 0) : (z == 0)) continue;
    const uint32_t stride = face == 0 ? sx : sxy;
    const uint32_t x = t.X0 + 32 * k + t.lane;
    const bool inb = x < sx;
    const uint32_t idx = (z * sy + y) * sx + x;
    const V v = inb ? rd.at(idx, x, y, z) : (V)0;
    const V vn = inb ? (face == 0 ? rd.at(idx - sx, x, y - 1, z) : rd.at(idx - sxy, x, y, z - 1)) : (V)0;
    const V vl = shfl_up1(v);
    const bool same_left = (t.lane > 0) && (v == vl);
    const bool c = (v != 0) && (v == vn);
    const bool c_l = __shfl_up_sync(FULL, (int)c, 1) != 0;
    if (c && !(same_left && c_l)) merge_emit(set, parent, parent[idx], parent[idx - stride]);
  }
  if (t.X0 > 0) {  // x face: one voxel pair per tile row
    for (uint32_t r = threadIdx.x; r < TILE_ROWS; r += blockDim.x) {
      const uint32_t y = t.Y0 + (r % TILE_Y), z = t.Z0 + (r / TILE_Y);
      if (y >= sy || z >= sz) continue;
      const uint32_t idx = (z * sy + y) * sx + t.X0;
      const V a = rd.at(idx, t.X0, y, z);
      if (a != 0 && a == rd.at(idx - 1, t.X0 - 1, y, z)) merge_emit(set, parent, parent[idx], parent[idx - 1]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < MERGE_SLOTS; i += blockDim.x) {
    const unsigned long long key = set[i];
    if (key != 0xFFFFFFFFFFFFFFFFull) uf_union(parent, (uint32_t)(key >> 32), (uint32_t)(key & 0xFFFFFFFFu));
  }
}

// G2 for the 256 x 8 x 4 tiles of k_ccl_local_v3: y faces have TILE_Z_S rows, z faces TILE_Y
template <typename R>
__global__ void __launch_bounds__(256)
    k_ccl_merge_tiles_s(R rd, uint32_t sx, uint32_t sy, uint32_t sz, uint32_t ntx, uint32_t nty,
                        uint32_t* parent) {
  using V = typename R::V;
  __shared__ unsigned long long set[MERGE_SLOTS];
  for (int i = threadIdx.x; i < MERGE_SLOTS; i += blockDim.x) set[i] = 0xFFFFFFFFFFFFFFFFull;
  __syncthreads();
  const TilePos t = tile_pos_s(ntx, nty);
  const uint32_t sxy = sx * sy;
  const uint32_t nwarps = blockDim.x >> 5;
  constexpr uint32_t Y_ITEMS = TILE_Z_S * SUBW, Z_ITEMS = TILE_Y * SUBW;
  for (uint32_t item = t.warp; item < Y_ITEMS + Z_ITEMS; item += nwarps) {
    const uint32_t face = item < Y_ITEMS ? 0u : 1u;  // 0: y face (ly = 0), 1: z face (lz = 0)
    const uint32_t it = face == 0 ? item : item - Y_ITEMS;
    const uint32_t rowi = it / SUBW, k = it % SUBW;
    const uint32_t ly = face == 0 ? 0 : rowi, lz = face == 0 ? rowi : 0;
    const uint32_t y = t.Y0 + ly, z = t.Z0 + lz;
    if (y >= sy || z >= sz) continue;
    if (face == 0 ? (y == 0) : (z == 0)) continue;
    const uint32_t stride = face == 0 ? sx : sxy;
    const uint32_t x = t.X0 + 32 * k + t.lane;
    const bool inb = x < sx;
    const uint32_t idx = (z * sy + y) * sx + x;
    const V v = inb ? rd.at(idx, x, y, z) : (V)0;
    const V vn = inb ? (face == 0 ? rd.at(idx - sx, x, y - 1, z) : rd.at(idx - sxy, x, y, z - 1)) : (V)0;
    const V vl = shfl_up1(v);
    const bool same_left = (t.lane > 0) && (v == vl);
    const bool c = (v != 0) && (v == vn);
    const bool c_l = __shfl_up_sync(FULL, (int)c, 1) != 0;
    if (c && !(same_left && c_l)) merge_emit(set, parent, parent[idx], parent[idx - stride]);
  }
  if (t.X0 > 0) {  // x face: one voxel pair per tile row
    for (uint32_t r = threadIdx.x; r < (uint32_t)(TILE_Y * TILE_Z_S); r += blockDim.x) {
      const uint32_t y = t.Y0 + (r % TILE_Y), z = t.Z0 + (r / TILE_Y);
      if (y >= sy || z >= sz) continue;
      const uint32_t idx = (z * sy + y) * sx + t.X0;
      const V a = rd.at(idx, t.X0, y, z);
      if (a != 0 && a == rd.at(idx - 1, t.X0 - 1, y, z)) merge_emit(set, parent, parent[idx], parent[idx - 1]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < MERGE_SLOTS; i += blockDim.x) {
    const unsigned long long key = set[i];
    if (key != 0xFFFFFFFFFFFFFFFFull) uf_union(parent, (uint32_t)(key >> 32), (uint32_t)(key & 0xFFFFFFFFu));
  }
}

// -------------------------------------------------------------------- roots
__global__ void __launch_bounds__(256)
    k_ccl_roots(const uint32_t* __restrict__ parent, const uint32_t* __restrict__ cand,
                uint32_t ncand, uint32_t* __restrict__ roots, uint32_t* counters) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  bool is_root = false;
  uint32_t c = 0;
  if (i < ncand) {
    c = cand[i];
    is_root = (parent[c] == c);
  }
  const uint32_t m = __ballot_sync(FULL, is_root);
  if (m) {
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(&counters[2], (uint32_t)__popc(m));
    base = __shfl_sync(FULL, base, leader);
    if (is_root) roots[base + __popc(m & ((1u << lane) - 1u))] = c;
  }
}

__global__ void __launch_bounds__(256)
    k_ccl_rank(uint32_t* __restrict__ parent, const uint32_t* __restrict__ roots_sorted,
               uint32_t nroots) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nroots) parent[roots_sorted[i]] = CCL_FLAG | (i + 1);
}

// ------------------------------------------------------------------ F label
// flat mapping: a warp owns ROWCHUNK = 256 consecutive voxels of one row
struct ChunkPos {
  uint32_t lane, rowbase, x0;
  bool ok;
};
__device__ __forceinline__ ChunkPos chunk_pos(uint32_t sx, uint32_t cpr, uint64_t nchunks) {
  ChunkPos c;
  const uint64_t wid = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  c.ok = wid < nchunks;
  c.lane = threadIdx.x & 31;
  const uint32_t row = (uint32_t)(wid / cpr);
  c.x0 = (uint32_t)(wid % cpr) * TILE_X;
  c.rowbase = row * sx;
  return c;
}

// rank (1..N) of the component of a voxel whose parent entry is e; BG -> BG
__device__ __forceinline__ uint32_t chase(const uint32_t* __restrict__ parent, uint32_t e) {
  while (!(e & CCL_FLAG)) e = parent[e];
  return e;
}

template <typename OUT>
__global__ void __launch_bounds__(256)
    k_ccl_label(const uint32_t* __restrict__ parent, uint32_t sx, uint32_t cpr, uint64_t nchunks,
                uint64_t offset, const uint32_t* __restrict__ rank_map, OUT* __restrict__ out) {
  const ChunkPos c = chunk_pos(sx, cpr, nchunks);
  if (!c.ok) return;
  uint32_t e[SUBW];
#pragma unroll
  for (int k = 0; k < SUBW; k++) {
    const uint32_t x = c.x0 + 32 * k + c.lane;
    e[k] = (x < sx) ? parent[c.rowbase + x] : CCL_BG;
  }
#pragma unroll
  for (int k = 0; k < SUBW; k++) e[k] = chase(parent, e[k]);
#pragma unroll
  for (int k = 0; k < SUBW; k++) {
    const uint32_t x = c.x0 + 32 * k + c.lane;
    uint32_t label = (e[k] == CCL_BG) ? 0u : (e[k] & ~CCL_FLAG);
    if (rank_map != nullptr && label != 0) label = rank_map[label];  // dust: 0 = removed
    if (x < sx) out[c.rowbase + x] = (label == 0) ? (OUT)0 : (OUT)((uint64_t)label + offset);
  }
}

// component sizes: one atomicAdd per run of equal ids inside a sub-word
__global__ void __launch_bounds__(256)
    k_ccl_count(const uint32_t* __restrict__ parent, uint32_t sx, uint32_t cpr, uint64_t nchunks,
                uint32_t* __restrict__ counts) {
  const ChunkPos c = chunk_pos(sx, cpr, nchunks);
  if (!c.ok) return;
#pragma unroll 1
  for (int k = 0; k < SUBW; k++) {
    const uint32_t x = c.x0 + 32 * k + c.lane;
    const uint32_t e = chase(parent, (x < sx) ? parent[c.rowbase + x] : CCL_BG);
    const bool bg = (e == CCL_BG);
    const uint32_t label = bg ? 0u : (e & ~CCL_FLAG);
    const uint32_t ll = __shfl_up_sync(FULL, label, 1);
    const bool head = !bg && (c.lane == 0 || ll != label);
    const uint32_t hm = __ballot_sync(FULL, head || bg);
    if (head) {
      const uint32_t above = (c.lane == 31) ? 0u : (hm & ~((2u << c.lane) - 1u));
      const uint32_t end = above ? (uint32_t)(__ffs(above) - 1) : 32u;
      atomicAdd(&counts[label], end - c.lane);
    }
  }
}

__global__ void __launch_bounds__(256)
    k_dust_flags(const uint32_t* __restrict__ counts, uint32_t n, uint64_t threshold,
                 uint32_t* __restrict__ keep) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i <= n && i > 0) keep[i] = ((uint64_t)counts[i] >= threshold) ? 1u : 0u;
  if (i == 0) keep[0] = 0;
}

// single-block inclusive scan: component counts are tiny next to voxel counts
__global__ void __launch_bounds__(1024)
    k_scan_keep(const uint32_t* __restrict__ keep, uint32_t n_plus1, uint32_t* __restrict__ rank_map,
                uint32_t* __restrict__ total) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n_plus1; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t k = (i < n_plus1) ? keep[i] : 0;
    uint32_t v = k;
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t t = __shfl_up_sync(FULL, v, d);
      if ((threadIdx.x & 31) >= d) v += t;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t s = warp_sums[threadIdx.x];
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, s, d);
        if (threadIdx.x >= d) s += t;
      }
      warp_sums[threadIdx.x] = s;
    }
    __syncthreads();
    const uint32_t prev_warps = (threadIdx.x >> 5) ? warp_sums[(threadIdx.x >> 5) - 1] : 0;
    const uint32_t incl = carry + prev_warps + v;
    if (i < n_plus1) rank_map[i] = k ? incl : 0;
    __syncthreads();
    if (threadIdx.x == 1023) carry = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

// in-place dust on the caller's labels: zero voxels of removed components
template <typename T>
__global__ void __launch_bounds__(256)
    k_dust_apply(const uint32_t* __restrict__ parent, uint32_t sx, uint32_t cpr, uint64_t nchunks,
                 const uint32_t* __restrict__ keep, T* __restrict__ labels) {
  const ChunkPos c = chunk_pos(sx, cpr, nchunks);
  if (!c.ok) return;
#pragma unroll
  for (int k = 0; k < SUBW; k++) {
    const uint32_t x = c.x0 + 32 * k + c.lane;
    if (x >= sx) continue;
    const uint32_t e = chase(parent, parent[c.rowbase + x]);
    if (e != CCL_BG && keep[e & ~CCL_FLAG] == 0) labels[c.rowbase + x] = 0;
  }
}

// ------------------------------------------------------------- host driver
struct CclScratch {
  uint32_t* parent;
  uint32_t* cand;
  uint32_t* roots;
  uint32_t* roots_sorted;
  uint32_t* counters;  // [0] ncand [1] overflow [2] nroots [3] kept
  void* cub_tmp;
  size_t cub_bytes;
  uint32_t cap;
};

static size_t ccl_cub_bytes(uint32_t cap) {
  size_t b = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, b, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)cap);
  return b;
}

static size_t ccl_scratch_bytes(uint64_t n, uint32_t cap) {
  return align_up(n * 4, 256) + 3 * align_up((size_t)cap * 4, 256) + 256 + align_up(ccl_cub_bytes(cap), 256) + 4096;
}

static int ccl_take(ign_ctx* ctx, uint64_t n, uint32_t cap, CclScratch& s,
                    uint32_t* ext_parent = nullptr) {
  s.cap = cap;
  s.parent = ext_parent ? ext_parent : (uint32_t*)scratch_take(ctx, n * 4);
  s.cand = (uint32_t*)scratch_take(ctx, (size_t)cap * 4);
  s.roots = (uint32_t*)scratch_take(ctx, (size_t)cap * 4);
  s.roots_sorted = (uint32_t*)scratch_take(ctx, (size_t)cap * 4);
  s.counters = (uint32_t*)scratch_take(ctx, 256);
  s.cub_bytes = ccl_cub_bytes(cap);
  s.cub_tmp = scratch_take(ctx, s.cub_bytes);
  IGN_REQUIRE(s.parent && s.cand && s.roots && s.roots_sorted && s.counters && s.cub_tmp,
              IGN_ERR_NOMEM, "CCL scratch arena too small");
  return IGN_OK;
}

static uint32_t default_cap(uint64_t n) {
  uint64_t c = n / 8 + 4096;
  return (uint32_t)c;
}

// runs A1, A2, roots, sort, rank.  On return parent[] holds flagged roots and
// *n_roots the number of components.  *overflow set if the candidate buffer
// was too small (nothing else valid then).
template <typename R>
static int ccl_core(ign_ctx* ctx, const R& rd, uint32_t sx, uint32_t sy, uint32_t sz,
                    CclScratch& s, uint32_t* n_roots, bool* overflow) {
  // IGN_CCL_V2=3: the half-height tile experiment (k_ccl_local_v3 / k_ccl_merge_tiles_s)
  bool small = false;
  if constexpr (!R::thresholded) {
    const char* e = getenv("IGN_CCL_V2");
    small = e != nullptr && e[0] == '3' && (rd.rx & rd.ry & rd.rz) == 0xFFFFFFFFu &&
            getenv("IGN_CCL_GENERIC") == nullptr && getenv("IGN_CCL_FLATMERGE") == nullptr && (sx % 4 == 0) &&
            ((uintptr_t)rd.in % 16 == 0) && ((uintptr_t)s.parent % 16 == 0);
  }
  const uint32_t tile_z = small ? (uint32_t)TILE_Z_S : (uint32_t)TILE_Z;
  const uint32_t ntx = (sx + TILE_X - 1) / TILE_X, nty = (sy + TILE_Y - 1) / TILE_Y,
                 ntz = (sz + tile_z - 1) / tile_z;
  const uint64_t n = (uint64_t)sx * sy * sz;
  const uint64_t ntiles = (uint64_t)ntx * nty * ntz;
  IGN_REQUIRE(ntiles < 0x7FFFFFFFull, IGN_ERR_OVERFLOW, "too many CCL tiles");
  const unsigned grid = (unsigned)ntiles;
  constexpr size_t smem = (TILE_VOX + CCL_WARPS * TASKS_PER_WARP) * sizeof(uint32_t);
  *overflow = false;
  // per device and cheap: set on every call (a process may drive several devices)
  IGN_CUDA(cudaFuncSetAttribute(k_ccl_local<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if constexpr (!R::thresholded)
    IGN_CUDA(cudaFuncSetAttribute(k_ccl_local_fast<typename R::value_type>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  IGN_CUDA(cudaMemsetAsync(s.counters, 0, 256, ctx->stream));
  bool fast = false;
  if constexpr (!R::thresholded) {  // raw labels (no threshold)
    fast = (rd.rx & rd.ry & rd.rz) == 0xFFFFFFFFu && getenv("IGN_CCL_GENERIC") == nullptr;
    // experimental 4-voxels-per-lane kernel: opt-in until it has been validated on a GPU
    const bool v2 = fast && getenv("IGN_CCL_V2") != nullptr && (sx % 4 == 0) &&
                    ((uintptr_t)rd.in % 16 == 0) && ((uintptr_t)s.parent % 16 == 0);
    if (small) {
      constexpr size_t smem3 = (size_t)(TILE_VOX_S + CCL_WARPS_S * (TASKS_PER_WARP + STARTS_PER_WARP)) * sizeof(uint32_t);
      IGN_LAUNCH_PROF(ctx, IGN_PROF_CCL_LOCAL, (k_ccl_local_v3<typename R::value_type>), grid, CCL_THREADS_S, smem3,
                      rd.in, sx, sy, sz, ntx, nty, s.parent, s.cand, s.cap, s.counters);
    } else if (v2) {
      constexpr size_t smem2 = smem + (size_t)CCL_WARPS * STARTS_PER_WARP * sizeof(uint32_t);
      IGN_CUDA(cudaFuncSetAttribute(k_ccl_local_v2<typename R::value_type>,
                                    cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      IGN_LAUNCH_PROF(ctx, IGN_PROF_CCL_LOCAL, (k_ccl_local_v2<typename R::value_type>), grid, CCL_THREADS, smem2,
                      rd.in, sx, sy, sz, ntx, nty, s.parent, s.cand, s.cap, s.counters);
    } else if (fast) {
      IGN_LAUNCH_PROF(ctx, IGN_PROF_CCL_LOCAL, (k_ccl_local_fast<typename R::value_type>), grid, CCL_THREADS, smem,
                      rd.in, sx, sy, sz, ntx, nty, s.parent, s.cand, s.cap, s.counters);
    }
  }
  if (!fast)
    IGN_LAUNCH_PROF(ctx, IGN_PROF_CCL_LOCAL, (k_ccl_local<R>), grid, CCL_THREADS, smem, rd, sx, sy, sz, ntx, nty,
                    s.parent, s.cand, s.cap, s.counters);
  if (ntiles > 1) {
    if (small) {
      if constexpr (!R::thresholded)
        IGN_LAUNCH_PROF(ctx, IGN_PROF_CCL_MERGE, (k_ccl_merge_tiles_s<R>), grid, 256, 0, rd, sx, sy, sz, ntx, nty, s.parent);
    } else if (getenv("IGN_CCL_FLATMERGE") == nullptr) {
      IGN_LAUNCH_PROF(ctx, IGN_PROF_CCL_MERGE, (k_ccl_merge_tiles<R>), grid, 256, 0, rd, sx, sy, sz, ntx, nty, s.parent);
    } else {
      const uint32_t w32 = (sx + 31) / 32;
      const uint64_t items_y = (uint64_t)(nty - 1) * sz * w32;
      const uint64_t items_z = (uint64_t)(ntz - 1) * sy * w32;
      const uint64_t items_x = (uint64_t)(ntx - 1) * (((uint64_t)sy * sz + 31) / 32);
      const uint64_t items = items_y + items_z + items_x;
      if (items > 0)
        IGN_LAUNCH_PROF(ctx, IGN_PROF_CCL_MERGE, (k_ccl_merge<R>), blocks_for(items * 32, 256), 256, 0, rd, sx, sy, sz, nty - 1,
                   ntz - 1, ntx - 1, items_y, items_z, items_x, s.parent);
    }
  }
  uint32_t* h = (uint32_t*)ctx->pinned;
  IGN_CUDA(cudaMemcpyAsync(h, s.counters, 16, cudaMemcpyDeviceToHost, ctx->stream));
  IGN_CUDA(cudaStreamSynchronize(ctx->stream));
  if (h[1] != 0 || h[0] > s.cap) {
    *overflow = true;
    return IGN_OK;
  }
  const uint32_t ncand = h[0];
  uint32_t nroots = 0;
  if (ncand > 0) {
    IGN_LAUNCH(ctx, k_ccl_roots, blocks_for(ncand, 256), 256, 0, s.parent, s.cand, ncand, s.roots,
               s.counters);
    IGN_CUDA(cudaMemcpyAsync(h, s.counters, 16, cudaMemcpyDeviceToHost, ctx->stream));
    IGN_CUDA(cudaStreamSynchronize(ctx->stream));
    nroots = h[2];
  }
  if (nroots > 0) {
    int end_bit = 1;
    while (end_bit < 32 && (1ull << end_bit) < n) end_bit++;
    size_t tb = s.cub_bytes;
    IGN_CUDA(cub::DeviceRadixSort::SortKeys(s.cub_tmp, tb, s.roots, s.roots_sorted, (int)nroots, 0,
                                            end_bit, ctx->stream));
    ctx->launches += 2;  // cub radix sort: library kernels, counted conservatively
    IGN_LAUNCH(ctx, k_ccl_rank, blocks_for(nroots, 256), 256, 0, s.parent, s.roots_sorted, nroots);
  }
  *n_roots = nroots;
  return IGN_OK;
}

template <typename T>
static Reader<T, false> plain_reader(const void* in) {
  Reader<T, false> r;
  r.in = (const T*)in;
  r.gte = r.lte = 0;
  r.use_gte = r.use_lte = 0;
  r.rx = r.ry = r.rz = 0xFFFFFFFFu;
  return r;
}

static int check_ccl_dims(uint64_t sx, uint64_t sy, uint64_t sz) {
  IGN_REQUIRE(sx > 0 && sy > 0 && sz > 0, IGN_ERR_INVALID, "empty volume");
  IGN_REQUIRE(sx * sy * sz <= 0x7FFFFFF0ull, IGN_ERR_OVERFLOW,
              "CCL chunk of %llu voxels exceeds the 2^31 voxel limit of 32-bit provisional labels; "
              "split the volume into tasks (igneous uses 512^3)",
              (unsigned long long)(sx * sy * sz));
  return IGN_OK;
}

static int launch_label(ign_ctx* ctx, const CclScratch& s, uint32_t sx, uint32_t sy, uint32_t sz,
                        uint64_t offset, const uint32_t* rank_map, void* out, int out_dtype,
                        uint64_t max_label) {
  const uint32_t wpr = (sx + TILE_X - 1) / TILE_X;
  const uint64_t nwords = (uint64_t)wpr * sy * sz;
  const unsigned grid = blocks_for(nwords * 32, 256);
  switch (out_dtype) {
    case IGN_U16:
      IGN_REQUIRE(max_label + offset <= 0xFFFFull, IGN_ERR_OVERFLOW, "%llu labels do not fit uint16", (unsigned long long)max_label);
      IGN_LAUNCH_PROF(ctx, IGN_PROF_CCL_LABEL, (k_ccl_label<uint16_t>), grid, 256, 0, s.parent, sx, wpr, nwords, offset, rank_map, (uint16_t*)out);
      break;
    case IGN_U32:
      IGN_REQUIRE(max_label + offset <= 0xFFFFFFFFull, IGN_ERR_OVERFLOW, "labels do not fit uint32");
      IGN_LAUNCH_PROF(ctx, IGN_PROF_CCL_LABEL, (k_ccl_label<uint32_t>), grid, 256, 0, s.parent, sx, wpr, nwords, offset, rank_map, (uint32_t*)out);
      break;
    case IGN_U64:
      IGN_LAUNCH_PROF(ctx, IGN_PROF_CCL_LABEL, (k_ccl_label<uint64_t>), grid, 256, 0, s.parent, sx, wpr, nwords, offset, rank_map, (uint64_t*)out);
      break;
    default:
      set_error("CCL out_dtype must be u16/u32/u64 (got %d)", out_dtype);
      return IGN_ERR_UNSUPPORTED;
  }
  return IGN_OK;
}

// dust on the component structure held in s.parent: builds keep/rank maps.
// returns device pointers (inside the arena) and the number of kept components.
static int dust_maps(ign_ctx* ctx, const CclScratch& s, uint32_t sx, uint32_t sy, uint32_t sz,
                     uint32_t nroots, uint64_t threshold, uint32_t** keep_out,
                     uint32_t** rank_map_out, uint32_t* kept) {
  const uint32_t wpr = (sx + TILE_X - 1) / TILE_X;
  const uint64_t nwords = (uint64_t)wpr * sy * sz;
  const unsigned grid = blocks_for(nwords * 32, 256);
  const size_t bytes = ((size_t)nroots + 1) * 4;
  uint32_t* counts = (uint32_t*)scratch_take(ctx, bytes);
  uint32_t* keep = (uint32_t*)scratch_take(ctx, bytes);
  uint32_t* rank_map = (uint32_t*)scratch_take(ctx, bytes);
  IGN_REQUIRE(counts && keep && rank_map, IGN_ERR_NOMEM, "scratch arena too small for dust maps");
  IGN_CUDA(cudaMemsetAsync(counts, 0, bytes, ctx->stream));
  IGN_LAUNCH(ctx, k_ccl_count, grid, 256, 0, s.parent, sx, wpr, nwords, counts);
  IGN_LAUNCH(ctx, k_dust_flags, blocks_for(nroots + 1, 256), 256, 0, counts, nroots, threshold, keep);
  IGN_LAUNCH(ctx, k_scan_keep, 1, 1024, 0, keep, nroots + 1, rank_map, s.counters + 3);
  uint32_t* h = (uint32_t*)ctx->pinned;
  IGN_CUDA(cudaMemcpyAsync(h, s.counters, 16, cudaMemcpyDeviceToHost, ctx->stream));
  IGN_CUDA(cudaStreamSynchronize(ctx->stream));
  *kept = h[3];
  *keep_out = keep;
  *rank_map_out = rank_map;
  return IGN_OK;
}

// full pipeline for one reader type; out may be null when only dust-in-place is wanted
template <typename R, typename TL>
static int ccl_run(ign_ctx* ctx, const R& rd, uint64_t sx, uint64_t sy, uint64_t sz,
                   uint64_t dust_threshold, uint64_t offset, void* out, int out_dtype,
                   TL* dust_labels_inplace, uint64_t* n_components) {
  IGN_TRY(check_ccl_dims(sx, sy, sz));
  const uint64_t n = sx * sy * sz;
  const bool own_arena = (ctx->scratch_used == 0);
  uint32_t cap = default_cap(n);
  for (int attempt = 0; attempt < 2; attempt++) {
    const size_t keep_used = ctx->scratch_used;
    const size_t need = ccl_scratch_bytes(n, cap) + 3 * ((size_t)cap + 1) * 4 + 1024;
    if (own_arena) IGN_TRY(scratch_reserve(ctx, need));
    CclScratch s;
    IGN_TRY(ccl_take(ctx, n, cap, s));
    uint32_t nroots = 0;
    bool overflow = false;
    int rc = ccl_core(ctx, rd, (uint32_t)sx, (uint32_t)sy, (uint32_t)sz, s, &nroots, &overflow);
    if (rc != IGN_OK) {
      ctx->scratch_used = keep_used;
      return rc;
    }
    if (overflow) {
      ctx->scratch_used = keep_used;
      IGN_REQUIRE(attempt == 0, IGN_ERR_OVERFLOW,
                  "CCL candidate buffer overflow (more than %u isolated segments)", cap);
      cap = (uint32_t)n + 1024;  // worst case: every voxel its own component
      continue;
    }
    uint32_t* rank_map = nullptr;
    uint32_t* keep = nullptr;
    uint32_t kept = nroots;
    if (dust_threshold > 0 && nroots > 0) {
      rc = dust_maps(ctx, s, (uint32_t)sx, (uint32_t)sy, (uint32_t)sz, nroots, dust_threshold, &keep, &rank_map, &kept);
      if (rc != IGN_OK) {
        ctx->scratch_used = keep_used;
        return rc;
      }
    }
    if (dust_labels_inplace != nullptr && keep != nullptr) {
      const uint32_t wpr = ((uint32_t)sx + TILE_X - 1) / TILE_X;
      const uint64_t nwords = (uint64_t)wpr * sy * sz;
      IGN_LAUNCH(ctx, (k_dust_apply<TL>), blocks_for(nwords * 32, 256), 256, 0, s.parent, (uint32_t)sx, wpr, nwords, keep, dust_labels_inplace);
    }
    if (out != nullptr) {
      rc = launch_label(ctx, s, (uint32_t)sx, (uint32_t)sy, (uint32_t)sz, offset, rank_map, out, out_dtype, kept);
      if (rc != IGN_OK) {
        ctx->scratch_used = keep_used;
        return rc;
      }
    }
    if (n_components) *n_components = kept;
    ctx->scratch_used = keep_used;
    return IGN_OK;
  }
  return IGN_ERR_OVERFLOW;
}

template <typename T>
static int ccl_task_typed(ign_ctx* ctx, const void* in, uint64_t sx, uint64_t sy, uint64_t sz,
                          int use_gte, double gte, int use_lte, double lte, uint64_t rx, uint64_t ry,
                          uint64_t rz, uint64_t dust, uint64_t offset, uint64_t* out, uint64_t* n) {
  auto rail = [](uint64_t r, uint64_t s) { return (r < s) ? (uint32_t)r : 0xFFFFFFFFu; };
  if (use_gte || use_lte) {
    Reader<T, true> r;
    r.in = (const T*)in;
    r.gte = gte;
    r.lte = lte;
    r.use_gte = use_gte;
    r.use_lte = use_lte;
    r.rx = rail(rx, sx);
    r.ry = rail(ry, sy);
    r.rz = rail(rz, sz);
    return ccl_run(ctx, r, sx, sy, sz, dust, offset, out, IGN_U64, (uint8_t*)nullptr, n);
  }
  if constexpr (std::is_same<T, float>::value) {
    set_error("CCL on float input requires a threshold");
    return IGN_ERR_UNSUPPORTED;
  } else {
    Reader<T, false> r = plain_reader<T>(in);
    r.rx = rail(rx, sx);
    r.ry = rail(ry, sy);
    r.rz = rail(rz, sz);
    return ccl_run(ctx, r, sx, sy, sz, dust, offset, out, IGN_U64, (uint8_t*)nullptr, n);
  }
}


// ------------------------------------------------------ multi-slab building blocks
// (igneous/tasks/image/ccl.py passes 1-4 without the file exchange: slabs are
// disjoint in z, linked through their facing planes.)

// structure only (phases L, G, roots, rank) into a caller-owned parent array
template <typename R>
static int ccl_build(ign_ctx* ctx, const R& rd, uint64_t sx, uint64_t sy, uint64_t sz,
                     uint32_t* parent, uint64_t* n_local) {
  IGN_TRY(check_ccl_dims(sx, sy, sz));
  const uint64_t n = sx * sy * sz;
  const bool own_arena = (ctx->scratch_used == 0);
  uint32_t cap = default_cap(n);
  for (int attempt = 0; attempt < 2; attempt++) {
    const size_t keep_used = ctx->scratch_used;
    if (own_arena) IGN_TRY(scratch_reserve(ctx, ccl_scratch_bytes(0, cap) + 4096));
    CclScratch s;
    int rc = ccl_take(ctx, n, cap, s, parent);
    if (rc != IGN_OK) {
      ctx->scratch_used = keep_used;
      return rc;
    }
    uint32_t nroots = 0;
    bool overflow = false;
    rc = ccl_core(ctx, rd, (uint32_t)sx, (uint32_t)sy, (uint32_t)sz, s, &nroots, &overflow);
    ctx->scratch_used = keep_used;
    if (rc != IGN_OK) return rc;
    if (overflow) {
      IGN_REQUIRE(attempt == 0, IGN_ERR_OVERFLOW, "CCL candidate buffer overflow");
      cap = (uint32_t)n + 1024;
      continue;
    }
    *n_local = nroots;
    return IGN_OK;
  }
  return IGN_ERR_OVERFLOW;
}

// one z-plane of a slab: voxel values widened to u64 and local component ids
template <typename T>
__global__ void __launch_bounds__(256)
    k_ccl_plane(const T* __restrict__ in, const uint32_t* __restrict__ parent, uint64_t plane_base,
                uint64_t nplane, const uint32_t* __restrict__ lut, uint64_t* __restrict__ values,
                uint32_t* __restrict__ labels) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= nplane) return;
  values[i] = (uint64_t)in[plane_base + i];
  const uint32_t e = chase(parent, parent[plane_base + i]);
  uint32_t l = (e == CCL_BG) ? 0u : (e & ~CCL_FLAG);
  if (lut != nullptr && l != 0) l = lut[l];
  labels[i] = l;
}

// equivalence pairs between two facing planes (same x,y; adjacent z)
__global__ void __launch_bounds__(256)
    k_ccl_link(const uint64_t* __restrict__ va, const uint32_t* __restrict__ la, uint64_t offa,
               const uint64_t* __restrict__ vb, const uint32_t* __restrict__ lb, uint64_t offb,
               uint64_t nplane, uint64_t* __restrict__ pairs, uint32_t cap, uint32_t* counters) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint32_t lane = threadIdx.x & 31;
  bool emit = false;
  uint64_t a = 0, b = 0;
  if (i < nplane) {
    const uint64_t v = va[i];
    if (v != 0 && v == vb[i]) {
      a = offa + la[i];
      b = offb + lb[i];
      // runs of the same pair along x are emitted once
      emit = !(i > 0 && la[i - 1] == la[i] && lb[i - 1] == lb[i] && va[i - 1] == v && vb[i - 1] == v);
    }
  }
  const uint32_t m = __ballot_sync(FULL, emit);
  if (m) {
    const int leader = __ffs(m) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(&counters[0], (uint32_t)__popc(m));
    base = __shfl_sync(FULL, base, leader);
    if (emit) {
      const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
      if (pos < cap) {
        pairs[2 * (uint64_t)pos] = a;
        pairs[2 * (uint64_t)pos + 1] = b;
      }
    }
  }
}

}  // namespace ign

using namespace ign;

extern "C" {

int ign_ccl6_dev(ign_ctx* ctx, const void* in, int in_dtype, uint64_t sx, uint64_t sy, uint64_t sz,
                 void* out, int out_dtype, uint64_t* n_components) {
  IGN_TRY(activate(ctx));
  IGN_REQUIRE(in && out, IGN_ERR_INVALID, "null buffer");
  switch (in_dtype) {
    case IGN_U8: return ccl_run(ctx, plain_reader<uint8_t>(in), sx, sy, sz, 0, 0, out, out_dtype, (uint8_t*)nullptr, n_components);
    case IGN_U16: return ccl_run(ctx, plain_reader<uint16_t>(in), sx, sy, sz, 0, 0, out, out_dtype, (uint16_t*)nullptr, n_components);
    case IGN_U32: return ccl_run(ctx, plain_reader<uint32_t>(in), sx, sy, sz, 0, 0, out, out_dtype, (uint32_t*)nullptr, n_components);
    case IGN_U64: return ccl_run(ctx, plain_reader<uint64_t>(in), sx, sy, sz, 0, 0, out, out_dtype, (uint64_t*)nullptr, n_components);
  }
  set_error("CCL: unsupported input dtype %d", in_dtype);
  return IGN_ERR_UNSUPPORTED;
}

int ign_dust_dev(ign_ctx* ctx, void* labels, int dtype, uint64_t sx, uint64_t sy, uint64_t sz,
                 uint64_t threshold) {
  IGN_TRY(activate(ctx));
  IGN_REQUIRE(labels, IGN_ERR_INVALID, "null buffer");
  if (threshold == 0) return IGN_OK;
  switch (dtype) {
    case IGN_U8: return ccl_run(ctx, plain_reader<uint8_t>(labels), sx, sy, sz, threshold, 0, nullptr, IGN_U64, (uint8_t*)labels, nullptr);
    case IGN_U16: return ccl_run(ctx, plain_reader<uint16_t>(labels), sx, sy, sz, threshold, 0, nullptr, IGN_U64, (uint16_t*)labels, nullptr);
    case IGN_U32: return ccl_run(ctx, plain_reader<uint32_t>(labels), sx, sy, sz, threshold, 0, nullptr, IGN_U64, (uint32_t*)labels, nullptr);
    case IGN_U64: return ccl_run(ctx, plain_reader<uint64_t>(labels), sx, sy, sz, threshold, 0, nullptr, IGN_U64, (uint64_t*)labels, nullptr);
  }
  set_error("dust: unsupported dtype %d", dtype);
  return IGN_ERR_UNSUPPORTED;
}

int ign_ccl_task_dev(ign_ctx* ctx, const void* in, int in_dtype, uint64_t sx, uint64_t sy,
                     uint64_t sz, int use_gte, double gte, int use_lte, double lte,
                     uint64_t rail_x, uint64_t rail_y, uint64_t rail_z, uint64_t dust_threshold,
                     uint64_t label_offset, uint64_t* out, uint64_t* n_components) {
  IGN_TRY(activate(ctx));
  IGN_REQUIRE(in && out, IGN_ERR_INVALID, "null buffer");
  switch (in_dtype) {
    case IGN_U8: return ccl_task_typed<uint8_t>(ctx, in, sx, sy, sz, use_gte, gte, use_lte, lte, rail_x, rail_y, rail_z, dust_threshold, label_offset, out, n_components);
    case IGN_U16: return ccl_task_typed<uint16_t>(ctx, in, sx, sy, sz, use_gte, gte, use_lte, lte, rail_x, rail_y, rail_z, dust_threshold, label_offset, out, n_components);
    case IGN_U32: return ccl_task_typed<uint32_t>(ctx, in, sx, sy, sz, use_gte, gte, use_lte, lte, rail_x, rail_y, rail_z, dust_threshold, label_offset, out, n_components);
    case IGN_U64: return ccl_task_typed<uint64_t>(ctx, in, sx, sy, sz, use_gte, gte, use_lte, lte, rail_x, rail_y, rail_z, dust_threshold, label_offset, out, n_components);
    case IGN_F32: return ccl_task_typed<float>(ctx, in, sx, sy, sz, use_gte, gte, use_lte, lte, rail_x, rail_y, rail_z, dust_threshold, label_offset, out, n_components);
  }
  set_error("CCL task: unsupported input dtype %d", in_dtype);
  return IGN_ERR_UNSUPPORTED;
}

// ---- host-buffer wrappers
int ign_ccl6(ign_ctx* ctx, const void* in, int in_dtype, uint64_t sx, uint64_t sy, uint64_t sz,
             void* out, int out_dtype, uint64_t* n_components) {
  IGN_TRY(activate(ctx));
  IGN_REQUIRE(in && out, IGN_ERR_INVALID, "null buffer");
  IGN_TRY(check_ccl_dims(sx, sy, sz));
  const int es = dtype_size(in_dtype), os = dtype_size(out_dtype);
  IGN_REQUIRE(es > 0 && os > 0, IGN_ERR_UNSUPPORTED, "unsupported dtype");
  const uint64_t n = sx * sy * sz;
  scratch_reset(ctx);
  const uint32_t cap = (uint32_t)n + 1024;  // host path: size for the worst case once
  IGN_TRY(scratch_reserve(ctx, align_up(n * es, 256) + align_up(n * os, 256) + ccl_scratch_bytes(n, cap) + 3 * ((size_t)cap + 1) * 4 + 8192));
  void* d_in = scratch_take(ctx, n * es);
  void* d_out = scratch_take(ctx, n * os);
  IGN_CUDA(cudaMemcpyAsync(d_in, in, n * es, cudaMemcpyHostToDevice, ctx->stream));
  int rc = ign_ccl6_dev(ctx, d_in, in_dtype, sx, sy, sz, d_out, out_dtype, n_components);
  if (rc == IGN_OK) {
    cudaError_t e = cudaMemcpyAsync(out, d_out, n * os, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      set_error("CCL D2H: %s", cudaGetErrorString(e));
      rc = IGN_ERR_CUDA;
    }
  }
  scratch_reset(ctx);
  return rc;
}

int ign_dust(ign_ctx* ctx, void* labels, int dtype, uint64_t sx, uint64_t sy, uint64_t sz,
             uint64_t threshold) {
  IGN_TRY(activate(ctx));
  IGN_REQUIRE(labels, IGN_ERR_INVALID, "null buffer");
  if (threshold == 0) return IGN_OK;
  IGN_TRY(check_ccl_dims(sx, sy, sz));
  const int es = dtype_size(dtype);
  IGN_REQUIRE(es > 0 && dtype != IGN_F32, IGN_ERR_UNSUPPORTED, "unsupported dtype");
  const uint64_t n = sx * sy * sz;
  scratch_reset(ctx);
  const uint32_t cap = (uint32_t)n + 1024;
  IGN_TRY(scratch_reserve(ctx, align_up(n * es, 256) + ccl_scratch_bytes(n, cap) + 3 * ((size_t)cap + 1) * 4 + 8192));
  void* d = scratch_take(ctx, n * es);
  IGN_CUDA(cudaMemcpyAsync(d, labels, n * es, cudaMemcpyHostToDevice, ctx->stream));
  int rc = ign_dust_dev(ctx, d, dtype, sx, sy, sz, threshold);
  if (rc == IGN_OK) {
    cudaError_t e = cudaMemcpyAsync(labels, d, n * es, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      set_error("dust D2H: %s", cudaGetErrorString(e));
      rc = IGN_ERR_CUDA;
    }
  }
  scratch_reset(ctx);
  return rc;
}


// ------------------------------------------------------------ multi-slab API
int ign_ccl6_build_dev(ign_ctx* ctx, const void* in, int in_dtype, uint64_t sx, uint64_t sy,
                       uint64_t sz, uint32_t* work, uint64_t* n_local) {
  IGN_TRY(activate(ctx));
  IGN_REQUIRE(in && work && n_local, IGN_ERR_INVALID, "null argument");
  switch (in_dtype) {
    case IGN_U8: return ccl_build(ctx, plain_reader<uint8_t>(in), sx, sy, sz, work, n_local);
    case IGN_U16: return ccl_build(ctx, plain_reader<uint16_t>(in), sx, sy, sz, work, n_local);
    case IGN_U32: return ccl_build(ctx, plain_reader<uint32_t>(in), sx, sy, sz, work, n_local);
    case IGN_U64: return ccl_build(ctx, plain_reader<uint64_t>(in), sx, sy, sz, work, n_local);
  }
  set_error("CCL build: unsupported input dtype %d", in_dtype);
  return IGN_ERR_UNSUPPORTED;
}

static int ccl_plane(ign_ctx* ctx, const void* in, int in_dtype, const uint32_t* work, uint64_t sx,
                     uint64_t sy, uint64_t sz, uint64_t z, const uint32_t* lut, uint64_t* values,
                     uint32_t* labels);

int ign_ccl6_plane_dev(ign_ctx* ctx, const void* in, int in_dtype, const uint32_t* work, uint64_t sx,
                       uint64_t sy, uint64_t sz, uint64_t z, uint64_t* values, uint32_t* labels) {
  return ccl_plane(ctx, in, in_dtype, work, sx, sy, sz, z, nullptr, values, labels);
}

static int ccl_plane(ign_ctx* ctx, const void* in, int in_dtype, const uint32_t* work, uint64_t sx,
                     uint64_t sy, uint64_t sz, uint64_t z, const uint32_t* lut, uint64_t* values,
                     uint32_t* labels) {
  IGN_TRY(activate(ctx));
  IGN_REQUIRE(in && work && values && labels && z < sz, IGN_ERR_INVALID, "bad plane argument");
  const uint64_t np = sx * sy, base = z * np;
  const unsigned g = blocks_for(np, 256);
  switch (in_dtype) {
    case IGN_U8: IGN_LAUNCH(ctx, (k_ccl_plane<uint8_t>), g, 256, 0, (const uint8_t*)in, work, base, np, lut, values, labels); break;
    case IGN_U16: IGN_LAUNCH(ctx, (k_ccl_plane<uint16_t>), g, 256, 0, (const uint16_t*)in, work, base, np, lut, values, labels); break;
    case IGN_U32: IGN_LAUNCH(ctx, (k_ccl_plane<uint32_t>), g, 256, 0, (const uint32_t*)in, work, base, np, lut, values, labels); break;
    case IGN_U64: IGN_LAUNCH(ctx, (k_ccl_plane<uint64_t>), g, 256, 0, (const uint64_t*)in, work, base, np, lut, values, labels); break;
    default: set_error("CCL plane: unsupported input dtype %d", in_dtype); return IGN_ERR_UNSUPPORTED;
  }
  return IGN_OK;
}

// device address of the pair list written by the last ign_ccl6_link_dev call on
// this thread; valid until the next arena allocation at the same bump position
static thread_local uint64_t* g_last_link_pairs = nullptr;
static uint64_t* link_pairs_dev(ign_ctx*) { return g_last_link_pairs; }

int ign_ccl6_link_dev(ign_ctx* ctx, const uint64_t* values_a, const uint32_t* labels_a,
                      uint64_t offset_a, const uint64_t* values_b, const uint32_t* labels_b,
                      uint64_t offset_b, uint64_t n_plane, uint64_t* pairs_host, uint64_t capacity,
                      uint64_t* n_pairs) {
  IGN_TRY(activate(ctx));
  IGN_REQUIRE(values_a && labels_a && values_b && labels_b && n_pairs, IGN_ERR_INVALID, "null argument");
  *n_pairs = 0;
  if (n_plane == 0) return IGN_OK;
  const size_t keep = ctx->scratch_used;
  const bool own = (keep == 0);
  const uint32_t cap = (uint32_t)(n_plane < 0x7FFFFFFFull ? n_plane : 0x7FFFFFFFull);
  if (own) IGN_TRY(scratch_reserve(ctx, (size_t)cap * 16 + 8192));
  uint64_t* d_pairs = (uint64_t*)scratch_take(ctx, (size_t)cap * 16);
  uint32_t* counters = (uint32_t*)scratch_take(ctx, 256);
  g_last_link_pairs = d_pairs;
  if (!d_pairs || !counters) {
    ctx->scratch_used = keep;
    set_error("scratch arena too small (CCL link)");
    return IGN_ERR_NOMEM;
  }
  IGN_CUDA(cudaMemsetAsync(counters, 0, 256, ctx->stream));
  IGN_LAUNCH(ctx, k_ccl_link, blocks_for(n_plane, 256), 256, 0, values_a, labels_a, offset_a, values_b,
             labels_b, offset_b, n_plane, d_pairs, cap, counters);
  uint32_t* h = (uint32_t*)ctx->pinned;
  IGN_CUDA(cudaMemcpyAsync(h, counters, 4, cudaMemcpyDeviceToHost, ctx->stream));
  IGN_CUDA(cudaStreamSynchronize(ctx->stream));
  const uint64_t total = h[0];
  *n_pairs = total;
  if (pairs_host && total) {
    const uint64_t m = total < capacity ? total : capacity;
    IGN_CUDA(cudaMemcpy(pairs_host, d_pairs, m * 16, cudaMemcpyDeviceToHost));
  }
  ctx->scratch_used = keep;
  return IGN_OK;
}

// Host-side global union-find over provisional ids 1..total (the B200-native
// stand-in for create_relabeling, igneous/tasks/image/ccl.py:358-420):
// smaller id wins (ccl.py:70-73); final ids are the ranks of the component
// minima, i.e. identical to a whole-volume cc3d numbering.
int ign_ccl6_solve(const uint64_t* pairs, uint64_t n_pairs, uint64_t total, uint32_t* lut,
                   uint64_t* n_global) {
  IGN_REQUIRE(lut && n_global && (n_pairs == 0 || pairs), IGN_ERR_INVALID, "null argument");
  IGN_REQUIRE(total < 0xFFFFFFF0ull, IGN_ERR_OVERFLOW, "too many provisional components");
  std::vector<uint32_t> p(total + 1);
  for (uint64_t i = 0; i <= total; i++) p[i] = (uint32_t)i;
  auto find = [&](uint32_t i) {
    while (p[i] != i) {
      p[i] = p[p[i]];
      i = p[i];
    }
    return i;
  };
  for (uint64_t k = 0; k < n_pairs; k++) {
    const uint64_t a64 = pairs[2 * k], b64 = pairs[2 * k + 1];
    IGN_REQUIRE(a64 >= 1 && a64 <= total && b64 >= 1 && b64 <= total, IGN_ERR_INVALID,
                "equivalence pair (%llu,%llu) out of range", (unsigned long long)a64, (unsigned long long)b64);
    const uint32_t a = find((uint32_t)a64), b = find((uint32_t)b64);
    if (a < b) p[b] = a;
    else if (b < a) p[a] = b;
  }
  uint32_t next = 0;
  lut[0] = 0;
  for (uint64_t i = 1; i <= total; i++) {
    const uint32_t r = find((uint32_t)i);
    if (r == i) lut[i] = ++next;  // roots are minima: met before their members
    else lut[i] = lut[r];
  }
  *n_global = next;
  return IGN_OK;
}

int ign_ccl6_label_dev(ign_ctx* ctx, const uint32_t* work, uint64_t sx, uint64_t sy, uint64_t sz,
                       const uint32_t* lut_dev, uint64_t offset, void* out, int out_dtype,
                       uint64_t max_label) {
  IGN_TRY(activate(ctx));
  IGN_REQUIRE(work && out, IGN_ERR_INVALID, "null argument");
  IGN_TRY(check_ccl_dims(sx, sy, sz));
  CclScratch s;
  s.parent = const_cast<uint32_t*>(work);
  return launch_label(ctx, s, (uint32_t)sx, (uint32_t)sy, (uint32_t)sz, offset, lut_dev, out, out_dtype, max_label);
}

// ---------------------------------------------------------------- volume CCL
// z-slabs of <= 2^30 voxels are resolved independently, linked through their
// facing planes, solved on the host and labelled once through the composed
// lookup table.  begin/finish are split so that a multi-GPU run can exchange the
// outer planes of every rank's volume in between (ONE all-gather) and fold the
// global relabelling into the same single label pass.
struct ign_ccl_volume {
  ign_ctx* ctx;
  const void* in;
  int in_dtype;
  uint64_t sx, sy, sz, slab_sz, nslabs;
  uint32_t* work;
  std::vector<uint64_t> nloc, off;   // per slab component counts / offsets
  std::vector<uint32_t> local_lut;   // provisional slab id -> volume-local id (1..n_local)
  uint64_t n_local;
};

int ign_ccl6_volume_abort(ign_ccl_volume* v) {
  if (!v) return IGN_OK;
  scratch_reset(v->ctx);
  delete v;
  return IGN_OK;
}

int ign_ccl6_volume_begin_dev(ign_ctx* ctx, const void* in, int in_dtype, uint64_t sx, uint64_t sy,
                              uint64_t sz, uint64_t max_slab_voxels, uint64_t* first_values,
                              uint32_t* first_labels, uint64_t* last_values, uint32_t* last_labels,
                              ign_ccl_volume** out, uint64_t* n_local) {
  IGN_TRY(activate(ctx));
  IGN_REQUIRE(in && out && n_local, IGN_ERR_INVALID, "null argument");
  *out = nullptr;
  IGN_REQUIRE(sx > 0 && sy > 0 && sz > 0, IGN_ERR_INVALID, "empty volume");
  const int es = dtype_size(in_dtype);
  IGN_REQUIRE(es > 0 && in_dtype != IGN_F32, IGN_ERR_UNSUPPORTED, "unsupported dtype");
  const uint64_t np = sx * sy;
  if (max_slab_voxels == 0 || max_slab_voxels > 0x40000000ull) max_slab_voxels = 0x40000000ull;
  IGN_REQUIRE(np <= max_slab_voxels, IGN_ERR_OVERFLOW, "one z-plane exceeds the slab limit");
  IGN_REQUIRE(ctx->scratch_used == 0, IGN_ERR_INVALID, "volume CCL must own the scratch arena");
  ign_ccl_volume* v = new ign_ccl_volume();
  v->ctx = ctx;
  v->in = in;
  v->in_dtype = in_dtype;
  v->sx = sx; v->sy = sy; v->sz = sz;
  v->slab_sz = max_slab_voxels / np;
  v->nslabs = (sz + v->slab_sz - 1) / v->slab_sz;
  const uint64_t n = np * sz, slab_sz = v->slab_sz, nslabs = v->nslabs;
  const uint64_t slab_vox = np * (slab_sz < sz ? slab_sz : sz);
  const uint32_t cap = (uint32_t)slab_vox + 1024;
  const size_t need = align_up(n * 4, 256) + ccl_scratch_bytes(0, cap) + 2 * (align_up(np * 8, 256) + align_up(np * 4, 256)) +
                      2 * align_up(np * 16, 256) + 2 * align_up((n / 8 + 4096) * 4, 256) + (4 << 20);
  int rc = scratch_reserve(ctx, need);
  if (rc != IGN_OK) { delete v; return rc; }
  uint32_t* work = v->work = (uint32_t*)scratch_take(ctx, n * 4);
  uint64_t* va = (uint64_t*)scratch_take(ctx, np * 8);
  uint64_t* vb = (uint64_t*)scratch_take(ctx, np * 8);
  uint32_t* la = (uint32_t*)scratch_take(ctx, np * 4);
  uint32_t* lb = (uint32_t*)scratch_take(ctx, np * 4);
  if (!work || !va || !vb || !la || !lb) {
    set_error("scratch arena too small (volume CCL)");
    ign_ccl6_volume_abort(v);
    return IGN_ERR_NOMEM;
  }
  v->nloc.assign(nslabs, 0);
  v->off.assign(nslabs + 1, 0);
  for (uint64_t s = 0; s < nslabs && rc == IGN_OK; s++) {
    const uint64_t z0 = s * slab_sz, zs = (z0 + slab_sz <= sz) ? slab_sz : sz - z0;
    rc = ign_ccl6_build_dev(ctx, (const char*)in + z0 * np * es, in_dtype, sx, sy, zs, work + z0 * np, &v->nloc[s]);
    v->off[s + 1] = v->off[s] + v->nloc[s];
  }
  const uint64_t total = v->off[nslabs];
  std::vector<uint64_t> pairs;
  for (uint64_t s = 0; s + 1 < nslabs && rc == IGN_OK; s++) {
    const uint64_t z0 = s * slab_sz, z1 = (s + 1) * slab_sz;
    const uint64_t zs1 = (z1 + slab_sz <= sz) ? slab_sz : sz - z1;
    rc = ign_ccl6_plane_dev(ctx, (const char*)in + z0 * np * es, in_dtype, work + z0 * np, sx, sy, slab_sz, slab_sz - 1, va, la);
    if (rc == IGN_OK) rc = ign_ccl6_plane_dev(ctx, (const char*)in + z1 * np * es, in_dtype, work + z1 * np, sx, sy, zs1, 0, vb, lb);
    if (rc != IGN_OK) break;
    uint64_t cnt = 0;
    rc = ign_ccl6_link_dev(ctx, va, la, v->off[s], vb, lb, v->off[s + 1], np, nullptr, 0, &cnt);
    if (rc != IGN_OK) break;
    if (cnt) {
      const size_t at = pairs.size();
      pairs.resize(at + 2 * cnt);
      cudaError_t e = cudaMemcpy(pairs.data() + at, link_pairs_dev(ctx), cnt * 16, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) {
        set_error("volume CCL: pairs D2H: %s", cudaGetErrorString(e));
        rc = IGN_ERR_CUDA;
      }
    }
  }
  if (rc == IGN_OK) {
    v->local_lut.assign(total + 1, 0);
    if (nslabs > 1) {
      rc = ign_ccl6_solve(pairs.data(), pairs.size() / 2, total, v->local_lut.data(), &v->n_local);
    } else {
      for (uint64_t i = 0; i <= total; i++) v->local_lut[i] = (uint32_t)i;
      v->n_local = total;
    }
  }
  // outer planes with volume-local ids, for a caller that links several volumes
  if (rc == IGN_OK && first_values && first_labels && last_values && last_labels) {
    uint32_t* d_lut = (uint32_t*)scratch_take(ctx, (total + 1) * 4);
    if (!d_lut) {
      set_error("scratch arena too small for the relabel table (%llu components)", (unsigned long long)total);
      rc = IGN_ERR_NOMEM;
    } else {
      cudaError_t e = cudaMemcpyAsync(d_lut, v->local_lut.data(), (total + 1) * 4, cudaMemcpyHostToDevice, ctx->stream);
      if (e != cudaSuccess) { set_error("volume CCL: lut H2D: %s", cudaGetErrorString(e)); rc = IGN_ERR_CUDA; }
      const uint64_t zl = (nslabs - 1) * slab_sz, zsl = sz - zl;
      if (rc == IGN_OK) rc = ccl_plane(ctx, in, in_dtype, work, sx, sy, (slab_sz < sz ? slab_sz : sz), 0, d_lut + v->off[0], first_values, first_labels);
      if (rc == IGN_OK) rc = ccl_plane(ctx, (const char*)in + zl * np * es, in_dtype, work + zl * np, sx, sy, zsl, zsl - 1, d_lut + v->off[nslabs - 1], last_values, last_labels);
      if (rc == IGN_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) { set_error("volume CCL: sync failed"); rc = IGN_ERR_CUDA; }
    }
  }
  if (rc != IGN_OK) {
    ign_ccl6_volume_abort(v);
    return rc;
  }
  // the arena stays held (work[] lives in it) until finish/abort; later
  // allocations of this call sequence bump above the structure
  *n_local = v->n_local;
  *out = v;
  return IGN_OK;
}

// global_lut: NULL, or [n_local+1] volume-local id -> final id (from the caller's
// cross-volume solve).  Labels every slab once and releases the arena.
int ign_ccl6_volume_finish_dev(ign_ccl_volume* v, const uint32_t* global_lut, uint64_t max_label,
                               void* out, int out_dtype) {
  IGN_REQUIRE(v && out, IGN_ERR_INVALID, "null argument");
  ign_ctx* ctx = v->ctx;
  IGN_TRY(activate(ctx));
  const int os = dtype_size(out_dtype);
  const uint64_t np = v->sx * v->sy, total = v->off[v->nslabs];
  int rc = IGN_OK;
  uint32_t* d_lut = nullptr;
  std::vector<uint32_t> composed;
  const uint32_t* lut_host = nullptr;
  if (global_lut) {
    composed.resize(total + 1);
    for (uint64_t i = 0; i <= total; i++) composed[i] = global_lut[v->local_lut[i]];
    lut_host = composed.data();
  } else if (v->nslabs > 1) {
    lut_host = v->local_lut.data();
    max_label = v->n_local;
  } else {
    max_label = v->n_local;
  }
  if (os <= 0) { set_error("unsupported out dtype"); rc = IGN_ERR_UNSUPPORTED; }
  if (rc == IGN_OK && lut_host) {
    d_lut = (uint32_t*)scratch_take(ctx, (total + 1) * 4);
    if (!d_lut) {
      set_error("scratch arena too small for the relabel table (%llu components)", (unsigned long long)total);
      rc = IGN_ERR_NOMEM;
    } else {
      cudaError_t e = cudaMemcpyAsync(d_lut, lut_host, (total + 1) * 4, cudaMemcpyHostToDevice, ctx->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);  // host table is a temporary
      if (e != cudaSuccess) { set_error("volume CCL: lut H2D: %s", cudaGetErrorString(e)); rc = IGN_ERR_CUDA; }
    }
  }
  for (uint64_t s = 0; s < v->nslabs && rc == IGN_OK; s++) {
    const uint64_t z0 = s * v->slab_sz, zs = (z0 + v->slab_sz <= v->sz) ? v->slab_sz : v->sz - z0;
    rc = ign_ccl6_label_dev(ctx, v->work + z0 * np, v->sx, v->sy, zs, d_lut ? d_lut + v->off[s] : nullptr, 0,
                            (char*)out + z0 * np * os, out_dtype, max_label);
  }
  scratch_reset(ctx);
  delete v;
  return rc;
}

int ign_ccl6_volume_dev(ign_ctx* ctx, const void* in, int in_dtype, uint64_t sx, uint64_t sy,
                        uint64_t sz, void* out, int out_dtype, uint64_t max_slab_voxels,
                        uint64_t* n_components) {
  IGN_REQUIRE(in && out, IGN_ERR_INVALID, "null buffer");
  ign_ccl_volume* v = nullptr;
  uint64_t n = 0;
  IGN_TRY(ign_ccl6_volume_begin_dev(ctx, in, in_dtype, sx, sy, sz, max_slab_voxels, nullptr, nullptr, nullptr,
                                    nullptr, &v, &n));
  IGN_TRY(ign_ccl6_volume_finish_dev(v, nullptr, n, out, out_dtype));
  if (n_components) *n_components = n;
  return IGN_OK;
}

}  // extern "C"
