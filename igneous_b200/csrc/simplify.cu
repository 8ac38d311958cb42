// simplify.cu -- quadric edge-collapse mesh simplification (K10)
//
// Replaces the simplifier inside zmesh.Mesher.get(id, reduction_factor,
// max_error) (igneous/tasks/mesh/mesh.py:376-381) for ALL labels of a task at
// once.  zmesh's simplifier is a sequential heap-ordered collapse per label; a
// GPU needs a data-parallel formulation, so this is a round-based variant:
//
//   init    per-vertex Garland-Heckbert plane quadrics (unit normals, summed in
//           face order), boundary vertices locked (chunk borders must stitch),
//           per-vertex incident-face arrays (half-edge nodes, fixed capacity,
//           merged and compacted on collapse).
//   round   E  one thread per half-edge: the half-edge with u < v of every edge
//              of a label still above its face target computes the cheap
//              quadric cost (min over {u, v, midpoint} of p^T (Qu+Qv) p) and,
//              if cost <= max_error^2, posts a (float cost, hashed label-local
//              half-edge id) key to both endpoints (atomicMin);
//           K2 one thread per vertex: minimum key over its 1-ring;
//           C  an edge WINS iff its key is the minimum of both endpoints'
//              rings -> winners never touch each other's faces or vertices
//              and are processed concurrently: a winner collapses iff the
//              link condition holds and no incident face flips, otherwise it
//              is parked until one of its endpoints' rings changes.
//   compact scans renumber surviving vertices / faces per label.
//
// All arithmetic is double precision WITHOUT fused multiply-add (this file is
// compiled with -fmad=false) so that oracle/igneous_oracle.c::orc_simplify
// reproduces it bit for bit.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <stdlib.h>

#include "mesher.h"

namespace ign {

constexpr uint32_t S_NONE = 0xFFFFFFFFu;
constexpr uint64_t S_KEYMAX = 0xFFFFFFFFFFFFFFFFull;
constexpr int S_MAXV = 32;
constexpr int S_VCAP = 64;  // two rings of < S_MAXV alive faces always fit after a collapse
constexpr int VB = 11;  // vertex key coordinate bits (mesh.cu V_COORD_BITS)

struct Simp {
  uint64_t U, T;
  double* pos;      // 3U
  double* Q;        // 10U
  uint32_t* face;   // 3T global vertex ids
  uint32_t* flabel; // T dense labels
  uint8_t* falive;
  uint8_t* valive;
  uint8_t* vbound;
  // incident half-edge nodes (3f+c) of every vertex as a fixed-capacity array: ring
  // enumeration is a set of independent loads instead of a linked-list pointer chase
  uint32_t* vf;   // [U * S_VCAP]
  uint32_t* vn;   // [U] entries in use (dead faces are skipped, compacted when the vertex is kept)
  unsigned long long *key1, *key2;
  uint32_t* alive_faces;   // [K+2]
  const uint32_t* target;  // [K+2]
  uint8_t* label_active;   // [K+2]
  const uint32_t* tri_off; // [K+2] first face of each label: keys use label-local half-edge ids
  uint8_t* estate;  // [3T] see k_simp_edge_keys
  float* ecost;     // [3T] cached float cost of state 2
  uint8_t* vdirty;  // [U] ring changed by a collapse of the previous round
  // compacted work lists (rebuilt every few rounds; pure work skipping)
  uint32_t* elist;  // candidate half-edges
  uint32_t* vlist;  // alive vertices of active labels
  uint32_t ne, nv;
};

__device__ __forceinline__ uint32_t s_mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t s_unmix(uint32_t x) {
  x ^= x >> 16; x *= 0x43021123U; x ^= x >> 15 ^ x >> 30; x *= 0x1d69e2a5U; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ unsigned long long s_key(double cost, uint32_t h, uint32_t salt) {
  const float c = __double2float_rn(cost);
  return ((unsigned long long)__float_as_uint(c) << 32) | s_mix(h ^ salt);
}

__device__ __forceinline__ double s_qeval(const double* q, const double* p) {
  const double x = p[0], y = p[1], z = p[2];
  return q[0] * x * x + 2.0 * q[1] * x * y + 2.0 * q[2] * x * z + 2.0 * q[3] * x + q[4] * y * y +
         2.0 * q[5] * y * z + 2.0 * q[6] * y + q[7] * z * z + 2.0 * q[8] * z + q[9];
}

__device__ int s_twins(const Simp& s, uint32_t f, uint32_t u, uint32_t v, uint32_t* twin) {
  int cnt = 0;
  const uint32_t* lu = s.vf + (uint64_t)u * S_VCAP;
  const uint32_t cu = s.vn[u];
  for (uint32_t j = 0; j < cu; j++) {
    const uint32_t h = lu[j];
    const uint32_t g = h / 3;
    if (g == f || !s.falive[g]) continue;
    const uint32_t* fv = s.face + 3 * (uint64_t)g;
    if (fv[0] == v || fv[1] == v || fv[2] == v) {
      if (cnt == 0) *twin = h;
      cnt++;
    }
  }
  return cnt;
}

__device__ bool s_ring(const Simp& s, uint32_t w, uint32_t* faces, uint32_t* nbr, int* nf, int* nn) {
  *nf = 0;
  *nn = 0;
  const uint32_t* lw = s.vf + (uint64_t)w * S_VCAP;
  const uint32_t cw = s.vn[w];
  for (uint32_t j = 0; j < cw; j++) {
    const uint32_t h = lw[j];
    const uint32_t g = h / 3;
    if (!s.falive[g]) continue;
    if (*nf >= S_MAXV) return false;
    faces[(*nf)++] = g;
    const uint32_t* fv = s.face + 3 * (uint64_t)g;
    for (int k = 0; k < 3; k++) {
      const uint32_t x = fv[k];
      if (x == w) continue;
      bool seen = false;
      for (int j = 0; j < *nn; j++) seen |= (nbr[j] == x);
      if (!seen) {
        if (*nn >= S_MAXV) return false;
        nbr[(*nn)++] = x;
      }
    }
  }
  return true;
}

struct SEval {
  bool valid;
  double cost;
  uint32_t keep, remove;
  double p[3];
};

// cheap part: placement and quadric cost (no ring walks)
__device__ void s_cost(const Simp& s, uint32_t u, uint32_t v, double max_err2, SEval* e) {
  e->valid = false;
  if (s.vbound[u] && s.vbound[v]) return;
  double q[10];
  for (int i = 0; i < 10; i++) q[i] = s.Q[10 * (uint64_t)u + i] + s.Q[10 * (uint64_t)v + i];
  const double* pu = s.pos + 3 * (uint64_t)u;
  const double* pv = s.pos + 3 * (uint64_t)v;
  double best[3], cost;
  if (s.vbound[u]) {
    e->keep = u;
    e->remove = v;
    best[0] = pu[0]; best[1] = pu[1]; best[2] = pu[2];
    cost = s_qeval(q, best);
  } else if (s.vbound[v]) {
    e->keep = v;
    e->remove = u;
    best[0] = pv[0]; best[1] = pv[1]; best[2] = pv[2];
    cost = s_qeval(q, best);
  } else {
    e->keep = u < v ? u : v;
    e->remove = u < v ? v : u;
    const double* pk = s.pos + 3 * (uint64_t)e->keep;
    const double* pr = s.pos + 3 * (uint64_t)e->remove;
    double mid[3] = {(pk[0] + pr[0]) * 0.5, (pk[1] + pr[1]) * 0.5, (pk[2] + pr[2]) * 0.5};
    const double ck = s_qeval(q, pk), cr = s_qeval(q, pr), cm = s_qeval(q, mid);
    cost = ck;
    best[0] = pk[0]; best[1] = pk[1]; best[2] = pk[2];
    if (cr < cost) { cost = cr; best[0] = pr[0]; best[1] = pr[1]; best[2] = pr[2]; }
    if (cm < cost) { cost = cm; best[0] = mid[0]; best[1] = mid[1]; best[2] = mid[2]; }
  }
  if (cost < 0.0) cost = 0.0;
  if (!(cost <= max_err2)) return;
  e->valid = true;
  e->cost = cost;
  e->p[0] = best[0]; e->p[1] = best[1]; e->p[2] = best[2];
}

// full validation of a round winner: link condition + no face flips
__device__ void s_evaluate(const Simp& s, uint32_t u, uint32_t v, double max_err2, SEval* e) {
  s_cost(s, u, v, max_err2, e);
  if (!e->valid) return;
  e->valid = false;
  const double* best = e->p;
  uint32_t fu[S_MAXV], fv[S_MAXV], nu[S_MAXV], nv[S_MAXV];
  int nfu, nfv, nnu, nnv;
  if (!s_ring(s, u, fu, nu, &nfu, &nnu)) return;
  if (!s_ring(s, v, fv, nv, &nfv, &nnv)) return;
  int common = 0;
  for (int i = 0; i < nnu; i++)
    for (int j = 0; j < nnv; j++) common += (nu[i] == nv[j]);
  int shared = 0;
  for (int i = 0; i < nfu; i++)
    for (int j = 0; j < nfv; j++) shared += (fu[i] == fv[j]);
  if (shared != 2 || common != 2) return;
  for (int pass = 0; pass < 2; pass++) {
    const uint32_t* fl = pass ? fv : fu;
    const int n = pass ? nfv : nfu;
    const uint32_t w = pass ? v : u, other = pass ? u : v;
    for (int i = 0; i < n; i++) {
      const uint32_t* fx = s.face + 3 * (uint64_t)fl[i];
      if (fx[0] == other || fx[1] == other || fx[2] == other) continue;
      const double* P[3];
      const double* N[3];
      for (int k = 0; k < 3; k++) {
        P[k] = s.pos + 3 * (uint64_t)fx[k];
        N[k] = (fx[k] == w) ? best : P[k];
      }
      const double ax = P[1][0] - P[0][0], ay = P[1][1] - P[0][1], az = P[1][2] - P[0][2];
      const double bx = P[2][0] - P[0][0], by = P[2][1] - P[0][1], bz = P[2][2] - P[0][2];
      const double n0x = ay * bz - az * by, n0y = az * bx - ax * bz, n0z = ax * by - ay * bx;
      const double cx = N[1][0] - N[0][0], cy = N[1][1] - N[0][1], cz = N[1][2] - N[0][2];
      const double dx = N[2][0] - N[0][0], dy = N[2][1] - N[0][1], dz = N[2][2] - N[0][2];
      const double n1x = cy * dz - cz * dy, n1y = cz * dx - cx * dz, n1z = cx * dy - cy * dx;
      const double dot = n0x * n1x + n0y * n1y + n0z * n1z;
      if (!(dot > 0.0)) return;
    }
  }
  e->valid = true;
}

// ------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(256)
    k_simp_init_verts(const uint64_t* __restrict__ vkeys, uint64_t U, double rx, double ry, double rz,
                      double* __restrict__ pos, uint8_t* __restrict__ valive,
                      uint8_t* __restrict__ vbound, uint32_t* __restrict__ vn) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= U) return;
  const uint64_t k = vkeys[i];
  const double x = (double)(k & ((1u << VB) - 1));
  const double y = (double)((k >> VB) & ((1u << VB) - 1));
  const double z = (double)((k >> (2 * VB)) & ((1u << VB) - 1));
  pos[3 * i + 0] = x * 0.5 * rx;
  pos[3 * i + 1] = y * 0.5 * ry;
  pos[3 * i + 2] = z * 0.5 * rz;
  valive[i] = 1;
  vbound[i] = 0;
  vn[i] = 0;
}

// faces: local ids + per-label vertex base -> global ids; flabel by offsets search
__global__ void __launch_bounds__(256)
    k_simp_init_faces(const uint32_t* __restrict__ faces_local, const uint32_t* __restrict__ tri_off,
                      const uint32_t* __restrict__ vert_off, uint32_t K, uint64_t T,
                      uint32_t* __restrict__ face, uint32_t* __restrict__ flabel,
                      uint8_t* __restrict__ falive, uint32_t* __restrict__ node_vertex,
                      uint32_t* __restrict__ node_id) {
  const uint64_t f = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (f >= T) return;
  // largest l in [1,K] with tri_off[l] <= f
  uint32_t lo = 1, hi = K;
  while (lo < hi) {
    const uint32_t mid = (lo + hi + 1) >> 1;
    if (tri_off[mid] <= f) lo = mid;
    else hi = mid - 1;
  }
  flabel[f] = lo;
  falive[f] = 1;
  for (int k = 0; k < 3; k++) {
    const uint32_t g = faces_local[3 * f + k] + vert_off[lo];
    face[3 * f + k] = g;
    node_vertex[3 * f + k] = g;
    node_id[3 * f + k] = (uint32_t)(3 * f + k);
  }
}

// sorted (vertex, node) pairs -> per-vertex arrays in ascending node order
__global__ void __launch_bounds__(256)
    k_simp_link(const uint32_t* __restrict__ sv, const uint32_t* __restrict__ sh, uint64_t n,
                uint32_t* __restrict__ vf, uint32_t* __restrict__ vn, uint32_t* overflow) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t v = sv[i];
  if (i > 0 && sv[i - 1] == v) return;  // the first pair of a run writes the whole run
  uint32_t c = 0;
  for (uint64_t j = i; j < n && sv[j] == v; j++) {
    if (c < S_VCAP) vf[(uint64_t)v * S_VCAP + c] = sh[j];
    c++;
  }
  if (c > S_VCAP) {
    *overflow = 1;
    c = S_VCAP;
  }
  vn[v] = c;
}

__global__ void __launch_bounds__(128) k_simp_quadrics(Simp s) {
  const uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (v >= s.U) return;
  double q[10];
  for (int i = 0; i < 10; i++) q[i] = 0.0;
  for (uint32_t j = 0; j < s.vn[v]; j++) {
    const uint32_t h = s.vf[(uint64_t)v * S_VCAP + j];
    const uint32_t* fv = s.face + 3 * (uint64_t)(h / 3);
    const double* a = s.pos + 3 * (uint64_t)fv[0];
    const double* b = s.pos + 3 * (uint64_t)fv[1];
    const double* c = s.pos + 3 * (uint64_t)fv[2];
    const double ux = b[0] - a[0], uy = b[1] - a[1], uz = b[2] - a[2];
    const double vx = c[0] - a[0], vy = c[1] - a[1], vz = c[2] - a[2];
    double nx = uy * vz - uz * vy, ny = uz * vx - ux * vz, nz = ux * vy - uy * vx;
    const double len = sqrt(nx * nx + ny * ny + nz * nz);
    if (!(len > 0.0)) continue;
    nx = nx / len; ny = ny / len; nz = nz / len;
    const double d = -(nx * a[0] + ny * a[1] + nz * a[2]);
    q[0] += nx * nx; q[1] += nx * ny; q[2] += nx * nz; q[3] += nx * d;
    q[4] += ny * ny; q[5] += ny * nz; q[6] += ny * d;
    q[7] += nz * nz; q[8] += nz * d; q[9] += d * d;
  }
  for (int i = 0; i < 10; i++) s.Q[10 * v + i] = q[i];
}

__global__ void __launch_bounds__(256) k_simp_boundary(Simp s) {
  const uint64_t h = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (h >= 3 * s.T) return;
  const uint32_t f = (uint32_t)(h / 3), c = (uint32_t)(h % 3);
  const uint32_t u = s.face[3 * (uint64_t)f + c], v = s.face[3 * (uint64_t)f + (c + 1) % 3];
  uint32_t tw;
  if (s_twins(s, f, u, v, &tw) != 1) {
    s.vbound[u] = 1;
    s.vbound[v] = 1;
  }
}

// work-list construction (warp-aggregated append; order is irrelevant)
__device__ __forceinline__ void s_append(bool take, uint32_t value, uint32_t* list, uint32_t* counter) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t m = __ballot_sync(0xFFFFFFFFu, take);
  if (!m) return;
  const int leader = __ffs(m) - 1;
  uint32_t base = 0;
  if ((int)lane == leader) base = atomicAdd(counter, (uint32_t)__popc(m));
  base = __shfl_sync(0xFFFFFFFFu, base, leader);
  if (take) list[base + __popc(m & ((1u << lane) - 1u))] = value;
}

__global__ void __launch_bounds__(256) k_simp_build_elist(Simp s, uint32_t* counters) {
  const uint64_t h = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  bool take = false;
  if (h < 3 * s.T) {
    const uint32_t f = (uint32_t)(h / 3), c = (uint32_t)(h % 3);
    // every half-edge of an alive face of an active label: faces only die and labels
    // only deactivate, so the list stays a superset between rebuilds (u<v can flip
    // when a collapse renames a vertex, so that filter stays in the edge pass)
    (void)c;
    take = s.falive[f] && s.alive_faces[s.flabel[f]] > s.target[s.flabel[f]];
  }
  s_append(take, (uint32_t)h, s.elist, &counters[0]);
}

__global__ void __launch_bounds__(256) k_simp_build_vlist(Simp s, uint32_t* counters) {
  const uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  bool take = false;
  if (v < s.U && s.valive[v]) {
    const uint32_t l = s.flabel[s.vf[(uint64_t)v * S_VCAP] / 3];
    take = s.alive_faces[l] > s.target[l];
  }
  s_append(take, (uint32_t)v, s.vlist, &counters[1]);
}

__global__ void __launch_bounds__(256)
    k_simp_round_begin(Simp s, uint32_t K, uint32_t* flags /* [0] any label active, [1] progress, [2] collapses */) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < s.nv) s.key1[s.vlist[i]] = S_KEYMAX;
  if (i >= 1 && i <= K) {
    const bool a = s.alive_faces[i] > s.target[i];
    s.label_active[i] = a ? 1 : 0;
    if (a) flags[0] = 1;
  }
}

__global__ void __launch_bounds__(128) k_simp_edge_keys(Simp s, double max_err2, uint32_t salt) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= s.ne) return;
  const uint32_t h = s.elist[i];
  const uint32_t f = (uint32_t)(h / 3), c = (uint32_t)(h % 3);
  if (!s.falive[f] || !s.label_active[s.flabel[f]]) return;
  const uint32_t u = s.face[3 * (uint64_t)f + c], v = s.face[3 * (uint64_t)f + (c + 1) % 3];
  if (!(u < v)) return;  // one key per edge
  // estate: 0 unknown, 1 parked (won a round, failed validation), 2 cost cached in
  // ecost, 3 known to exceed max_error.  Any cached state is dropped when one of the
  // endpoints' rings changed in the previous round (pure memoisation).
  uint8_t st = s.estate[h];
  if (st != 0 && (s.vdirty[u] || s.vdirty[v])) st = 0;
  if (st == 1 || st == 3) return;
  float cf;
  if (st == 2) {
    cf = s.ecost[h];
  } else {
    SEval e;
    s_cost(s, u, v, max_err2, &e);
    if (!e.valid) {
      s.estate[h] = 3;
      return;
    }
    cf = __double2float_rn(e.cost);
    s.ecost[h] = cf;
    s.estate[h] = 2;
  }
  const unsigned long long key = ((unsigned long long)__float_as_uint(cf) << 32) |
                                 s_mix(((uint32_t)h - 3 * s.tri_off[s.flabel[f]]) ^ salt);
  atomicMin(&s.key1[u], key);
  atomicMin(&s.key1[v], key);
}

__global__ void __launch_bounds__(256) k_simp_key2(Simp s) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= s.nv) return;
  const uint32_t w = s.vlist[i];
  s.vdirty[w] = 0;  // consumed by this round's edge pass; the collapse pass sets it again
  if (!s.valive[w]) {
    s.key2[w] = S_KEYMAX;
    return;
  }
  unsigned long long m = s.key1[w];
  const uint32_t* lw = s.vf + (uint64_t)w * S_VCAP;
  const uint32_t cw = s.vn[w];
  for (uint32_t j = 0; j < cw; j++) {
    const uint32_t g = lw[j] / 3;
    if (!s.falive[g]) continue;
    for (int k = 0; k < 3; k++) {
      const unsigned long long kk = s.key1[s.face[3 * (uint64_t)g + k]];
      if (kk < m) m = kk;
    }
  }
  s.key2[w] = m;
}

// winners of the round -> dense list (one winner per ~50 vertices: processing
// them in place would leave one active lane per warp)
__global__ void __launch_bounds__(256)
    k_simp_select(Simp s, uint32_t salt, uint32_t* wlist, uint32_t* counters) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  bool win = false;
  uint32_t h = 0;
  if (i < s.nv) {
    const uint32_t a = s.vlist[i];
    const unsigned long long key = s.key1[a];
    if (s.valive[a] && key != S_KEYMAX) {
      // the key holds a label-local half-edge id; a's label is that of any of its faces
      const uint32_t hl = s_unmix((uint32_t)(key & 0xFFFFFFFFu)) ^ salt;
      h = hl + 3 * s.tri_off[s.flabel[s.vf[(uint64_t)a * S_VCAP] / 3]];
      const uint32_t f = h / 3, c = h % 3;
      const uint32_t u = s.face[3 * (uint64_t)f + c], v = s.face[3 * (uint64_t)f + (c + 1) % 3];
      win = (a == u) && s.key2[u] == key && s.key2[v] == key;
    }
  }
  s_append(win, h, wlist, &counters[3]);
}

__global__ void __launch_bounds__(128)
    k_simp_collapse(Simp s, double max_err2, const uint32_t* __restrict__ wlist, uint32_t* flags) {
  const uint32_t nw = flags[3];  // written by k_simp_select (stream ordered)
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nw;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t h = wlist[i];
    const uint32_t f = h / 3, c = h % 3;
    const uint32_t u = s.face[3 * (uint64_t)f + c], v = s.face[3 * (uint64_t)f + (c + 1) % 3];
    SEval e;
    s_evaluate(s, u, v, max_err2, &e);
    if (!e.valid) {  // park the edge until one of its endpoints' rings changes
      s.estate[h] = 1;
      flags[1] = 1;
      continue;
    }
    const uint32_t k = e.keep, rm = e.remove;
    s.pos[3 * (uint64_t)k + 0] = e.p[0];
    s.pos[3 * (uint64_t)k + 1] = e.p[1];
    s.pos[3 * (uint64_t)k + 2] = e.p[2];
    for (int q = 0; q < 10; q++)
      s.Q[10 * (uint64_t)k + q] = s.Q[10 * (uint64_t)k + q] + s.Q[10 * (uint64_t)rm + q];
    uint32_t* lk = s.vf + (uint64_t)k * S_VCAP;
    const uint32_t* lr = s.vf + (uint64_t)rm * S_VCAP;
    const uint32_t ck = s.vn[k], cr = s.vn[rm];
    for (uint32_t j = 0; j < cr; j++) {
      const uint32_t hh = lr[j];
      const uint32_t g = hh / 3;
      if (!s.falive[g]) continue;
      uint32_t* fv = s.face + 3 * (uint64_t)g;
      if (fv[0] == k || fv[1] == k || fv[2] == k) {
        s.falive[g] = 0;
        atomicSub(&s.alive_faces[s.flabel[g]], 1u);
      } else {
        fv[hh % 3] = k;
      }
    }
    // k's new ring = alive entries of k's array (compacted in place) ++ alive entries of
    // rm's.  Both rings had < S_MAXV alive faces (validated), so S_VCAP entries suffice.
    // Only this thread touches k and rm this round.
    uint32_t nk = 0;
    for (int pass = 0; pass < 2; pass++) {
      const uint32_t* src = pass ? lr : lk;
      const uint32_t cnt = pass ? cr : ck;
      for (uint32_t j = 0; j < cnt; j++) {
        const uint32_t hh = src[j];
        if (!s.falive[hh / 3]) continue;
        lk[nk++] = hh;
        const uint32_t* fv = s.face + 3 * (uint64_t)(hh / 3);
        s.vdirty[fv[0]] = 1;
        s.vdirty[fv[1]] = 1;
        s.vdirty[fv[2]] = 1;
      }
    }
    s.vn[k] = nk;
    s.vdirty[k] = 1;
    s.valive[rm] = 0;
    flags[1] = 1;
    atomicAdd(&flags[2], 1u);
  }
}

// ---------------------------------------------------------------- batched gathers
// Opt-in variants (IGN_SIMP_BATCH=1) of the three latency-bound ring walkers.  The serial
// versions above chase node -> alive flag -> vertex ids -> positions one ring entry at a
// time (k_simp_collapse: 19-21 % issue-active, profiles/r01_simp_round_metrics.csv); here
// the loads of S_B ring entries are issued together before anything depends on them.
// Only the order of LOADS changes: every comparison, every floating-point operation and
// every store happens in the same order on the same values, so results are bit-identical
// to the serial kernels (and to oracle/igneous_oracle.c::orc_simplify).
constexpr int S_B = 8;

// s_ring + the vertex ids of every collected face (3 per face, the flip test reuses them)
__device__ bool s_ring_b(const Simp& s, uint32_t w, uint32_t* faces, uint32_t* fverts, uint32_t* nbr,
                         int* nf, int* nn) {
  *nf = 0;
  *nn = 0;
  const uint32_t* lw = s.vf + (uint64_t)w * S_VCAP;
  const uint32_t cw = s.vn[w];
  for (uint32_t j0 = 0; j0 < cw; j0 += S_B) {
    uint32_t h[S_B], fv[S_B][3];
    uint8_t al[S_B];
#pragma unroll
    for (int i = 0; i < S_B; i++) h[i] = (j0 + i < cw) ? lw[j0 + i] : S_NONE;
#pragma unroll
    for (int i = 0; i < S_B; i++) al[i] = (h[i] != S_NONE) ? s.falive[h[i] / 3] : (uint8_t)0;
#pragma unroll
    for (int i = 0; i < S_B; i++) {
      const uint32_t* p = s.face + 3 * (uint64_t)((al[i] ? h[i] : 0u) / 3);  // dead entry: any valid face
      fv[i][0] = p[0];
      fv[i][1] = p[1];
      fv[i][2] = p[2];
    }
#pragma unroll
    for (int i = 0; i < S_B; i++) {
      if (!al[i]) continue;
      if (*nf >= S_MAXV) return false;
      faces[*nf] = h[i] / 3;
      fverts[3 * *nf + 0] = fv[i][0];
      fverts[3 * *nf + 1] = fv[i][1];
      fverts[3 * *nf + 2] = fv[i][2];
      (*nf)++;
      for (int k = 0; k < 3; k++) {
        const uint32_t x = fv[i][k];
        if (x == w) continue;
        bool seen = false;
        for (int j = 0; j < *nn; j++) seen |= (nbr[j] == x);
        if (!seen) {
          if (*nn >= S_MAXV) return false;
          nbr[(*nn)++] = x;
        }
      }
    }
  }
  return true;
}

__device__ void s_evaluate_b(const Simp& s, uint32_t u, uint32_t v, double max_err2, SEval* e) {
  s_cost(s, u, v, max_err2, e);
  if (!e->valid) return;
  e->valid = false;
  const double* best = e->p;
  uint32_t fu[S_MAXV], fv[S_MAXV], nu[S_MAXV], nv[S_MAXV], xu[3 * S_MAXV], xv[3 * S_MAXV];
  int nfu, nfv, nnu, nnv;
  if (!s_ring_b(s, u, fu, xu, nu, &nfu, &nnu)) return;
  if (!s_ring_b(s, v, fv, xv, nv, &nfv, &nnv)) return;
  int common = 0;
  for (int i = 0; i < nnu; i++)
    for (int j = 0; j < nnv; j++) common += (nu[i] == nv[j]);
  int shared = 0;
  for (int i = 0; i < nfu; i++)
    for (int j = 0; j < nfv; j++) shared += (fu[i] == fv[j]);
  if (shared != 2 || common != 2) return;
  for (int pass = 0; pass < 2; pass++) {
    const uint32_t* xl = pass ? xv : xu;
    const int n = pass ? nfv : nfu;
    const uint32_t w = pass ? v : u, other = pass ? u : v;
    for (int i0 = 0; i0 < n; i0 += 2) {  // positions of two faces (18 doubles) per batch
      double pos[2][3][3];
      uint32_t ids[2][3];
#pragma unroll
      for (int b = 0; b < 2; b++) {
        const int i = (i0 + b < n) ? (i0 + b) : i0;
#pragma unroll
        for (int k = 0; k < 3; k++) ids[b][k] = xl[3 * i + k];
      }
#pragma unroll
      for (int b = 0; b < 2; b++)
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const double* p = s.pos + 3 * (uint64_t)ids[b][k];
          pos[b][k][0] = p[0];
          pos[b][k][1] = p[1];
          pos[b][k][2] = p[2];
        }
#pragma unroll
      for (int b = 0; b < 2; b++) {
        if (i0 + b >= n) continue;
        const uint32_t* fx = ids[b];
        if (fx[0] == other || fx[1] == other || fx[2] == other) continue;
        const double* P[3];
        const double* N[3];
        for (int k = 0; k < 3; k++) {
          P[k] = pos[b][k];
          N[k] = (fx[k] == w) ? best : P[k];
        }
        const double ax = P[1][0] - P[0][0], ay = P[1][1] - P[0][1], az = P[1][2] - P[0][2];
        const double bx = P[2][0] - P[0][0], by = P[2][1] - P[0][1], bz = P[2][2] - P[0][2];
        const double n0x = ay * bz - az * by, n0y = az * bx - ax * bz, n0z = ax * by - ay * bx;
        const double cx = N[1][0] - N[0][0], cy = N[1][1] - N[0][1], cz = N[1][2] - N[0][2];
        const double dx = N[2][0] - N[0][0], dy = N[2][1] - N[0][1], dz = N[2][2] - N[0][2];
        const double n1x = cy * dz - cz * dy, n1y = cz * dx - cx * dz, n1z = cx * dy - cy * dx;
        const double dot = n0x * n1x + n0y * n1y + n0z * n1z;
        if (!(dot > 0.0)) return;
      }
    }
  }
  e->valid = true;
}

__global__ void __launch_bounds__(256) k_simp_key2_b(Simp s) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= s.nv) return;
  const uint32_t w = s.vlist[i];
  s.vdirty[w] = 0;
  if (!s.valive[w]) {
    s.key2[w] = S_KEYMAX;
    return;
  }
  unsigned long long m = s.key1[w];
  const uint32_t* lw = s.vf + (uint64_t)w * S_VCAP;
  const uint32_t cw = s.vn[w];
  for (uint32_t j0 = 0; j0 < cw; j0 += S_B) {
    uint32_t h[S_B], fv[S_B][3];
    uint8_t al[S_B];
#pragma unroll
    for (int b = 0; b < S_B; b++) h[b] = (j0 + b < cw) ? lw[j0 + b] : S_NONE;
#pragma unroll
    for (int b = 0; b < S_B; b++) al[b] = (h[b] != S_NONE) ? s.falive[h[b] / 3] : (uint8_t)0;
#pragma unroll
    for (int b = 0; b < S_B; b++) {
      const uint32_t* p = s.face + 3 * (uint64_t)((al[b] ? h[b] : 0u) / 3);
      fv[b][0] = al[b] ? p[0] : w;  // dead entry: w itself (its key1 is already in m)
      fv[b][1] = al[b] ? p[1] : w;
      fv[b][2] = al[b] ? p[2] : w;
    }
#pragma unroll
    for (int b = 0; b < S_B; b++)
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const unsigned long long kk = s.key1[fv[b][k]];
        if (kk < m) m = kk;
      }
  }
  s.key2[w] = m;
}

__global__ void __launch_bounds__(128)
    k_simp_collapse_b(Simp s, double max_err2, const uint32_t* __restrict__ wlist, uint32_t* flags) {
  const uint32_t nw = flags[3];
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < nw;
       i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t h = wlist[i];
    const uint32_t f = h / 3, c = h % 3;
    const uint32_t u = s.face[3 * (uint64_t)f + c], v = s.face[3 * (uint64_t)f + (c + 1) % 3];
    SEval e;
    s_evaluate_b(s, u, v, max_err2, &e);
    if (!e.valid) {
      s.estate[h] = 1;
      flags[1] = 1;
      continue;
    }
    const uint32_t k = e.keep, rm = e.remove;
    s.pos[3 * (uint64_t)k + 0] = e.p[0];
    s.pos[3 * (uint64_t)k + 1] = e.p[1];
    s.pos[3 * (uint64_t)k + 2] = e.p[2];
    {
      double qk[10], qr[10];
#pragma unroll
      for (int q = 0; q < 10; q++) {
        qk[q] = s.Q[10 * (uint64_t)k + q];
        qr[q] = s.Q[10 * (uint64_t)rm + q];
      }
#pragma unroll
      for (int q = 0; q < 10; q++) s.Q[10 * (uint64_t)k + q] = qk[q] + qr[q];
    }
    uint32_t* lk = s.vf + (uint64_t)k * S_VCAP;
    const uint32_t* lr = s.vf + (uint64_t)rm * S_VCAP;
    const uint32_t ck = s.vn[k], cr = s.vn[rm];
    // faces of rm: those that also hold k die, the others get k in rm's corner
    for (uint32_t j0 = 0; j0 < cr; j0 += S_B) {
      uint32_t hh[S_B], fx[S_B][3];
      uint8_t al[S_B];
#pragma unroll
      for (int b = 0; b < S_B; b++) hh[b] = (j0 + b < cr) ? lr[j0 + b] : S_NONE;
#pragma unroll
      for (int b = 0; b < S_B; b++) al[b] = (hh[b] != S_NONE) ? s.falive[hh[b] / 3] : (uint8_t)0;
#pragma unroll
      for (int b = 0; b < S_B; b++) {
        const uint32_t* p = s.face + 3 * (uint64_t)((al[b] ? hh[b] : 0u) / 3);
        fx[b][0] = p[0];
        fx[b][1] = p[1];
        fx[b][2] = p[2];
      }
#pragma unroll
      for (int b = 0; b < S_B; b++) {
        if (!al[b]) continue;
        const uint32_t g = hh[b] / 3;
        if (fx[b][0] == k || fx[b][1] == k || fx[b][2] == k) {
          s.falive[g] = 0;
          atomicSub(&s.alive_faces[s.flabel[g]], 1u);
        } else {
          s.face[3 * (uint64_t)g + hh[b] % 3] = k;
        }
      }
    }
    // k's new ring = alive entries of k's array (compacted in place) ++ alive entries of rm's
    uint32_t nk = 0;
    for (int pass = 0; pass < 2; pass++) {
      const uint32_t* src = pass ? lr : lk;
      const uint32_t cnt = pass ? cr : ck;
      for (uint32_t j0 = 0; j0 < cnt; j0 += S_B) {
        uint32_t hh[S_B], fx[S_B][3];
        uint8_t al[S_B];
#pragma unroll
        for (int b = 0; b < S_B; b++) hh[b] = (j0 + b < cnt) ? src[j0 + b] : S_NONE;
#pragma unroll
        for (int b = 0; b < S_B; b++) al[b] = (hh[b] != S_NONE) ? s.falive[hh[b] / 3] : (uint8_t)0;
#pragma unroll
        for (int b = 0; b < S_B; b++) {
          const uint32_t* p = s.face + 3 * (uint64_t)((al[b] ? hh[b] : 0u) / 3);
          fx[b][0] = p[0];
          fx[b][1] = p[1];
          fx[b][2] = p[2];
        }
#pragma unroll
        for (int b = 0; b < S_B; b++) {
          if (!al[b]) continue;
          lk[nk++] = hh[b];
          s.vdirty[fx[b][0]] = 1;
          s.vdirty[fx[b][1]] = 1;
          s.vdirty[fx[b][2]] = 1;
        }
      }
    }
    s.vn[k] = nk;
    s.vdirty[k] = 1;
    s.valive[rm] = 0;
    flags[1] = 1;
    atomicAdd(&flags[2], 1u);
  }
}

__global__ void __launch_bounds__(256)
    k_simp_flags_u32(const uint8_t* __restrict__ a, uint64_t n, uint32_t* __restrict__ out) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = a[i];
}

__global__ void __launch_bounds__(256)
    k_simp_new_offsets(const uint32_t* __restrict__ old_off, const uint32_t* __restrict__ scan,
                       uint32_t K2, uint64_t n, uint32_t total, uint32_t* __restrict__ new_off) {
  const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < K2) new_off[l] = (old_off[l] < n) ? scan[old_off[l]] : total;
}

__global__ void __launch_bounds__(256)
    k_simp_compact_verts(Simp s, const uint32_t* __restrict__ vscan, float* __restrict__ pos_f) {
  const uint64_t v = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (v >= s.U || !s.valive[v]) return;
  const uint32_t n = vscan[v];
  pos_f[3 * (uint64_t)n + 0] = __double2float_rn(s.pos[3 * v + 0]);
  pos_f[3 * (uint64_t)n + 1] = __double2float_rn(s.pos[3 * v + 1]);
  pos_f[3 * (uint64_t)n + 2] = __double2float_rn(s.pos[3 * v + 2]);
}

__global__ void __launch_bounds__(256)
    k_simp_compact_faces(Simp s, const uint32_t* __restrict__ vscan, const uint32_t* __restrict__ fscan,
                         const uint32_t* __restrict__ new_vert_off, uint32_t* __restrict__ faces_out) {
  const uint64_t f = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (f >= s.T || !s.falive[f]) return;
  const uint32_t n = fscan[f];
  const uint32_t base = new_vert_off[s.flabel[f]];
  for (int k = 0; k < 3; k++) faces_out[3 * (uint64_t)n + k] = vscan[s.face[3 * f + k]] - base;
}

__global__ void __launch_bounds__(256)
    k_simp_export(const float* __restrict__ pos_f, uint64_t first, uint64_t count, float sx, float sy,
                  float sz, float* __restrict__ out) {
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (i >= count) return;
  out[3 * i + 0] = __fadd_rn(pos_f[3 * (first + i) + 0], sx);
  out[3 * i + 1] = __fadd_rn(pos_f[3 * (first + i) + 1], sy);
  out[3 * i + 2] = __fadd_rn(pos_f[3 * (first + i) + 2], sz);
}

// exposed to mesh.cu for export of simplified positions
int simp_export_positions(ign_ctx* ctx, const float* pos_f, uint64_t first, uint64_t count,
                          const float shift[3], float* d_out) {
  if (count == 0) return IGN_OK;
  IGN_LAUNCH(ctx, k_simp_export, blocks_for(count, 256), 256, 0, pos_f, first, count, shift[0], shift[1],
             shift[2], d_out);
  return IGN_OK;
}

}  // namespace ign

using namespace ign;

extern "C" int ign_mesh_simplify(ign_mesher* m, const float resolution[3], int reduction_factor,
                                 float max_error) {
  IGN_REQUIRE(m && resolution, IGN_ERR_INVALID, "null argument");
  ign_ctx* ctx = m->ctx;
  IGN_TRY(activate(ctx));
  IGN_REQUIRE(!m->simplified, IGN_ERR_INVALID, "mesher is already simplified; call mesh() again");
  IGN_REQUIRE(reduction_factor >= 1, IGN_ERR_INVALID, "reduction_factor must be >= 1");
  m->res[0] = resolution[0];
  m->res[1] = resolution[1];
  m->res[2] = resolution[2];
  m->simp_factor = reduction_factor;
  m->simp_max_error = max_error;
  m->simp_rounds = 0;
  const uint64_t U = m->U, T = m->T, K = m->K;
  if (T == 0 || U == 0) {
    m->simplified = true;
    m->d_pos_f = nullptr;
    return IGN_OK;
  }
  IGN_REQUIRE(m->pooled, IGN_ERR_UNSUPPORTED, "simplification needs the pooled mesher buffers");
  IGN_REQUIRE(ctx->scratch_used == 0, IGN_ERR_INVALID, "ign_mesh_simplify must own the scratch arena");

  size_t sortb = 0, scanb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sortb, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)(3 * T));
  cub::DeviceScan::ExclusiveSum(nullptr, scanb, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                (int)(3 * T));
  const size_t tmpb = (sortb > scanb ? sortb : scanb) + 256;
  const size_t need = align_up(U * 24, 256) + align_up(U * 80, 256) + 4 * align_up(3 * T * 4, 256) +
                      align_up(U * S_VCAP * 4, 256) +
                      3 * align_up(T * 4, 256) + 2 * align_up(T, 256) + 3 * align_up(U, 256) +
                      4 * align_up(U * 4, 256) + 2 * align_up(U * 8, 256) + 6 * align_up((K + 2) * 4, 256) +
                      2 * align_up(3 * T * 4, 256) + align_up(3 * T, 256) + align_up(3 * T * 4, 256) +
                      align_up(U, 256) + align_up(3 * T * 4, 256) + 2 * align_up(U * 4, 256) + tmpb + (1 << 20);
  IGN_TRY(scratch_reserve(ctx, need));
  Simp s;
  s.U = U;
  s.T = T;
  s.pos = (double*)scratch_take(ctx, U * 24);
  s.Q = (double*)scratch_take(ctx, U * 80);
  s.face = (uint32_t*)scratch_take(ctx, 3 * T * 4);
  s.vf = (uint32_t*)scratch_take(ctx, U * S_VCAP * 4);
  uint32_t* node_v = (uint32_t*)scratch_take(ctx, 3 * T * 4);   // reused as face scan later
  uint32_t* node_h = (uint32_t*)scratch_take(ctx, 3 * T * 4);
  s.flabel = (uint32_t*)scratch_take(ctx, T * 4);
  s.falive = (uint8_t*)scratch_take(ctx, T);
  s.valive = (uint8_t*)scratch_take(ctx, U);
  s.vbound = (uint8_t*)scratch_take(ctx, U);
  s.vn = (uint32_t*)scratch_take(ctx, U * 4);
  uint32_t* vscan = (uint32_t*)scratch_take(ctx, U * 4);
  uint32_t* vflag = (uint32_t*)scratch_take(ctx, U * 4);
  s.key1 = (unsigned long long*)scratch_take(ctx, U * 8);
  s.key2 = (unsigned long long*)scratch_take(ctx, U * 8);
  s.alive_faces = (uint32_t*)scratch_take(ctx, (K + 2) * 4);
  uint32_t* d_target = (uint32_t*)scratch_take(ctx, (K + 2) * 4);
  s.label_active = (uint8_t*)scratch_take(ctx, K + 2);
  uint32_t* d_tri_off = (uint32_t*)scratch_take(ctx, (K + 2) * 4);
  uint32_t* d_vert_off = (uint32_t*)scratch_take(ctx, (K + 2) * 4);
  uint32_t* d_new_tri_off = (uint32_t*)scratch_take(ctx, (K + 2) * 4);
  uint32_t* d_new_vert_off = (uint32_t*)scratch_take(ctx, (K + 2) * 4);
  uint32_t* flags = (uint32_t*)scratch_take(ctx, 256);
  s.estate = (uint8_t*)scratch_take(ctx, 3 * T);
  s.ecost = (float*)scratch_take(ctx, 3 * T * 4);
  s.vdirty = (uint8_t*)scratch_take(ctx, U);
  s.elist = (uint32_t*)scratch_take(ctx, 3 * T * 4);
  s.vlist = (uint32_t*)scratch_take(ctx, U * 4);
  uint32_t* wlist = (uint32_t*)scratch_take(ctx, U * 4);
  s.ne = s.nv = 0;
  void* tmp = scratch_take(ctx, tmpb);
  uint32_t* sorted_v = (uint32_t*)scratch_take(ctx, 3 * T * 4);
  uint32_t* sorted_h = (uint32_t*)scratch_take(ctx, 3 * T * 4);
  if (!s.pos || !s.Q || !s.face || !s.vf || !node_v || !node_h || !s.flabel || !s.falive || !s.valive ||
      !s.vbound || !s.vn || !vscan || !vflag || !s.key1 || !s.key2 || !s.alive_faces ||
      !d_target || !s.label_active || !d_tri_off || !d_vert_off || !d_new_tri_off || !d_new_vert_off ||
      !flags || !tmp || !sorted_v || !sorted_h || !s.estate || !s.ecost || !s.vdirty || !s.elist ||
      !s.vlist || !wlist) {
    scratch_reset(ctx);
    set_error("scratch arena too small (simplify: %llu faces)", (unsigned long long)T);
    return IGN_ERR_NOMEM;
  }
  s.target = d_target;
  s.tri_off = d_tri_off;
  auto done = [&](int code) {
    scratch_reset(ctx);
    return code;
  };
#define S_CUDA(call)                                                                  \
  do {                                                                                \
    cudaError_t _e = (call);                                                          \
    if (_e != cudaSuccess) {                                                          \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return done(IGN_ERR_CUDA);                                                      \
    }                                                                                 \
  } while (0)
#define S_TRY(call)                       \
  do {                                    \
    const int _s = (call);                \
    if (_s != IGN_OK) return done(_s);    \
  } while (0)
#define S_LAUNCH(kernel, g, b, ...)                       \
  do {                                                    \
    kernel<<<(g), (b), 0, ctx->stream>>>(__VA_ARGS__);    \
    ctx->launches++;                                      \
    S_CUDA(cudaGetLastError());                           \
  } while (0)

  std::vector<uint32_t> target(K + 2, 0);
  for (uint64_t l = 1; l <= K; l++)
    target[l] = (m->tri_off[l + 1] - m->tri_off[l]) / (uint32_t)reduction_factor;
  S_TRY(small_h2d(ctx, d_target, target.data(), (K + 2) * 4));
  S_TRY(small_h2d(ctx, d_tri_off, m->tri_off.data(), (K + 2) * 4));
  S_TRY(small_h2d(ctx, d_vert_off, m->vert_off.data(), (K + 2) * 4));
  S_CUDA(cudaMemsetAsync(s.alive_faces, 0, (K + 2) * 4, ctx->stream));
  S_CUDA(cudaMemsetAsync(s.label_active, 0, K + 2, ctx->stream));
  S_CUDA(cudaMemsetAsync(s.estate, 0, 3 * T, ctx->stream));
  S_CUDA(cudaMemsetAsync(s.vdirty, 0, U, ctx->stream));
  S_LAUNCH(k_simp_init_verts, blocks_for(U, 256), 256, m->d_uniq_vkeys, U, (double)resolution[0],
           (double)resolution[1], (double)resolution[2], s.pos, s.valive, s.vbound, s.vn);
  S_LAUNCH(k_simp_init_faces, blocks_for(T, 256), 256, m->d_faces, d_tri_off, d_vert_off, (uint32_t)K, T,
           s.face, s.flabel, s.falive, node_v, node_h);
  {
    int bits = 1;
    while (bits < 32 && (1ull << bits) < U) bits++;
    size_t tb = tmpb;
    S_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, node_v, sorted_v, node_h, sorted_h, (int)(3 * T), 0, bits,
                                           ctx->stream));
    ctx->launches += 3;
  }
  S_CUDA(cudaMemsetAsync(flags, 0, 64, ctx->stream));
  S_LAUNCH(k_simp_link, blocks_for(3 * T, 256), 256, sorted_v, sorted_h, (uint64_t)(3 * T), s.vf, s.vn,
           flags + 12);
  {
    uint32_t* hf = (uint32_t*)ctx->pinned;
    S_TRY(small_d2h(ctx, hf, flags + 12, 4));
    S_TRY(small_sync(ctx));
    if (hf[0] != 0) {
      set_error("simplify: a vertex has more than %d incident faces", S_VCAP);
      return done(IGN_ERR_UNSUPPORTED);
    }
  }
  S_LAUNCH(k_simp_quadrics, blocks_for(U, 128), 128, s);
  S_LAUNCH(k_simp_boundary, blocks_for(3 * T, 256), 256, s);
  {
    std::vector<uint32_t> af(K + 2, 0);
    for (uint64_t l = 1; l <= K; l++) af[l] = m->tri_off[l + 1] - m->tri_off[l];
    S_TRY(small_h2d(ctx, s.alive_faces, af.data(), (K + 2) * 4));
    S_TRY(small_sync(ctx));
  }

  const double max_err2 = (double)max_error * (double)max_error;
  const int max_rounds = 400;
  const bool trace = getenv("IGN_SIMP_TRACE") != nullptr;  // per-round progress on stderr
  // batched-gather ring walkers: same results, opt-in until validated on a GPU
  const bool batch = getenv("IGN_SIMP_BATCH") != nullptr;
  // work-list rebuild period and collapse grid cap: measured on B200 (tools/time_simplify.py),
  // 4 / 8 / 16 rounds -> 187.7 / 177.4 / 175.3 ms per 257^3 task; the grid cap has no effect
  const int rebuild_every = 16;
  const unsigned collapse_cap = 1184u;
  uint32_t* hflags = (uint32_t*)ctx->pinned;
  int r = 0, slow = 0;
  uint64_t cum_collapses = 0;
  for (; r < max_rounds; r++) {
    const uint32_t salt = (uint32_t)r * 0x9E3779B9u;
    if (r % rebuild_every == 0) {  // rebuild the compact work lists
      S_CUDA(cudaMemsetAsync(flags + 8, 0, 8, ctx->stream));
      S_LAUNCH(k_simp_build_elist, blocks_for(3 * T, 256), 256, s, flags + 8);
      S_LAUNCH(k_simp_build_vlist, blocks_for(U, 256), 256, s, flags + 8);
      S_TRY(small_d2h(ctx, hflags + 8, flags + 8, 8));
      S_TRY(small_sync(ctx));
      s.ne = hflags[8];
      s.nv = hflags[9];
    }
    S_CUDA(cudaMemsetAsync(flags, 0, 16, ctx->stream));
    const uint64_t nb = (s.nv > K + 1 ? s.nv : K + 1);
    S_LAUNCH(k_simp_round_begin, blocks_for(nb, 256), 256, s, (uint32_t)K, flags);
    if (s.ne) S_LAUNCH(k_simp_edge_keys, blocks_for(s.ne, 128), 128, s, max_err2, salt);
    if (s.nv) {
      if (batch) S_LAUNCH(k_simp_key2_b, blocks_for(s.nv, 256), 256, s);
      else S_LAUNCH(k_simp_key2, blocks_for(s.nv, 256), 256, s);
      S_LAUNCH(k_simp_select, blocks_for(s.nv, 256), 256, s, salt, wlist, flags);
      // grid-stride over the device-side winner count: no host round trip in between
      const unsigned cg = blocks_for(s.nv / 16 + 1, 128);
      if (batch) S_LAUNCH(k_simp_collapse_b, cg < collapse_cap ? cg : collapse_cap, 128, s, max_err2, wlist, flags);
      else S_LAUNCH(k_simp_collapse, cg < collapse_cap ? cg : collapse_cap, 128, s, max_err2, wlist, flags);
    }
    S_TRY(small_d2h(ctx, hflags, flags, 16));
    S_TRY(small_sync(ctx));
    if (hflags[0] == 0) break;           // every label reached its target before this round
    if (hflags[1] == 0) { r++; break; }  // nothing collapsed or parked: fixed point
    // early stop (mirrored by the oracle): four consecutive rounds that each remove
    // fewer than 0.2% of the remaining faces
    cum_collapses += hflags[2];
    const uint64_t alive_total = T - 2 * cum_collapses;
    if ((uint64_t)hflags[2] * 1000 < alive_total) slow++; else slow = 0;
    if (trace)
      fprintf(stderr, "round %d collapses %u alive %llu ne %u nv %u\n", r, hflags[2],
              (unsigned long long)alive_total, s.ne, s.nv);
    if (slow >= 4) { r++; break; }
  }
  m->simp_rounds = r;

  // ---- compaction
  uint32_t* fscan = node_v;   // 3T u32 >= T
  uint32_t* fflag = node_h;
  size_t tb = tmpb;
  S_LAUNCH(k_simp_flags_u32, blocks_for(U, 256), 256, s.valive, U, vflag);
  S_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, vflag, vscan, (int)U, ctx->stream));
  S_LAUNCH(k_simp_flags_u32, blocks_for(T, 256), 256, s.falive, T, fflag);
  tb = tmpb;
  S_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, fflag, fscan, (int)T, ctx->stream));
  ctx->launches += 4;
  uint32_t last[4];
  S_TRY(small_d2h(ctx, &last[0], vscan + (U - 1), 4));
  S_TRY(small_d2h(ctx, &last[1], vflag + (U - 1), 4));
  S_TRY(small_d2h(ctx, &last[2], fscan + (T - 1), 4));
  S_TRY(small_d2h(ctx, &last[3], fflag + (T - 1), 4));
  S_TRY(small_sync(ctx));
  const uint32_t U2 = last[0] + last[1], T2 = last[2] + last[3];
  S_LAUNCH(k_simp_new_offsets, blocks_for(K + 2, 256), 256, d_vert_off, vscan, (uint32_t)(K + 2), U, U2,
           d_new_vert_off);
  S_LAUNCH(k_simp_new_offsets, blocks_for(K + 2, 256), 256, d_tri_off, fscan, (uint32_t)(K + 2), T, T2,
           d_new_tri_off);
  // results overwrite the pooled mesher buffers (inputs were copied into the arena)
  float* pos_f = (float*)m->d_uniq_vkeys;
  S_LAUNCH(k_simp_compact_verts, blocks_for(U, 256), 256, s, vscan, pos_f);
  S_LAUNCH(k_simp_compact_faces, blocks_for(T, 256), 256, s, vscan, fscan, d_new_vert_off, m->d_faces);
  S_TRY(small_d2h(ctx, m->tri_off.data(), d_new_tri_off, (K + 2) * 4));
  S_TRY(small_d2h(ctx, m->vert_off.data(), d_new_vert_off, (K + 2) * 4));
  S_TRY(small_sync(ctx));
  m->U = U2;
  m->T = T2;
  m->d_pos_f = pos_f;
  m->simplified = true;
  m->present.clear();
  for (uint64_t l = 1; l <= K; l++)
    if (m->tri_off[l + 1] > m->tri_off[l]) m->present.push_back(m->ids[l - 1]);
  return done(IGN_OK);
}
