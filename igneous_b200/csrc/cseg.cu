// cseg.cu -- Precomputed `compressed_segmentation` chunk codec on the device
//
// the wire format either side of the hot path.  CloudVolume encodes /
// decodes it on the host around igneous/tasks/image/image.py:57-100 (every mip a
// DownsampleTask uploads) and igneous/tasks/image/ccl.py:346-356 (RelabelCCLTask's output;
// the CLI's default CCL encoding is compresso, `igneous_cli/cli.py:750`, with
// compressed_segmentation as the other segmentation codec).  Encoding where the labels
// already are shrinks the D2H of a label chunk by the compression ratio.
//
// Format (Neuroglancer): per channel [2 x u32 header per 8x8x8 block | per block: packed
// indices, then -- unless an identical table was already emitted by an earlier block of the
// channel -- the sorted lookup table]; header = (table offset : 24 | bits << 24), offset of the
// packed indices; bits in {0,1,2,4,8,16,32}; offsets in u32 words from the channel start.
// The emission order of oracle/igneous_oracle.c::orc_cseg_encode_* (block raster order, a
// table is emitted by the FIRST block that uses it) is reproduced exactly, so the streams are
// byte-identical:
//   1  k_cseg_scan<T, false>  one warp per block: the distinct values are extracted in
//      ascending order (repeated warp minimum) -> n, bits, 64-bit hash of the table
//   2  radix sort of (hash, block) -> the first block of every group of identical tables owns it
//   3  exclusive scan of the per-block sizes -> offsets
//   4  k_cseg_scan<T, true>   the same extraction again, now writing indices, tables, headers
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <vector>

#include "common.cuh"

namespace ign {

constexpr unsigned CS_FULL = 0xFFFFFFFFu;
constexpr int CS_MAX_BVOX = 1024;  // voxels per block (8x8x8 = 512 is the standard)

struct CsegDims {
  uint32_t sx, sy, sz, bx, by, bz, gx, gy, gz, bvox;
};

__device__ __forceinline__ uint64_t cs_shfl_xor(uint64_t v, int m) {
  return ((uint64_t)__shfl_xor_sync(CS_FULL, (uint32_t)(v >> 32), m) << 32) | __shfl_xor_sync(CS_FULL, (uint32_t)v, m);
}
__device__ __forceinline__ uint32_t cs_bits(uint32_t n) {
  if (n <= 1) return 0;
  uint32_t b = 1;
  while ((1u << b) < n) b *= 2;
  return b;
}

// One warp per block.  WRITE = false: info[b] = {n, bits}, hash[b].  WRITE = true: the stream.
template <typename T, bool WRITE, int CS_PER_LANE>  // CS_PER_LANE * 32 >= voxels per block
__global__ void __launch_bounds__(128)
    k_cseg_scan(const T* __restrict__ in, CsegDims d, uint64_t nblock, uint32_t* __restrict__ info_n,
                unsigned long long* __restrict__ hash, const uint32_t* __restrict__ enc_off,
                const uint32_t* __restrict__ tab_off, const uint32_t* __restrict__ owner, uint32_t* __restrict__ out) {
  constexpr int WORDS = sizeof(T) / 4;
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t b = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  if (b >= nblock) return;
  const uint32_t gxx = (uint32_t)(b % d.gx), gyy = (uint32_t)((b / d.gx) % d.gy), gzz = (uint32_t)(b / ((uint64_t)d.gx * d.gy));
  const uint32_t x0 = gxx * d.bx, y0 = gyy * d.by, z0 = gzz * d.bz;
  // lane l holds block positions l, l+32, ... (position p = (z*by + y)*bx + x)
  uint64_t val[CS_PER_LANE];
  uint32_t idx[CS_PER_LANE];
  uint32_t have = 0, todo = 0;  // bit k: slot k is inside the volume / not classified yet
#pragma unroll
  for (int k = 0; k < CS_PER_LANE; k++) {
    const uint32_t p = lane + 32 * k;
    val[k] = 0;
    idx[k] = 0;
    if (p < d.bvox) {
      const uint32_t x = p % d.bx, y = (p / d.bx) % d.by, z = p / (d.bx * d.by);
      if (x0 + x < d.sx && y0 + y < d.sy && z0 + z < d.sz) {
        val[k] = (uint64_t)in[(x0 + x) + (uint64_t)d.sx * ((y0 + y) + (uint64_t)d.sy * (z0 + z))];
        have |= 1u << k;
      }
    }
  }
  todo = have;
  uint32_t n = 0;
  uint64_t h = 0x9E3779B97F4A7C15ull;
  const uint32_t toff = WRITE ? tab_off[b] : 0u;
  const bool own = WRITE ? (owner[b] == (uint32_t)b) : false;
  while (__any_sync(CS_FULL, todo != 0)) {
    uint64_t m = ~0ull;
#pragma unroll
    for (int k = 0; k < CS_PER_LANE; k++)
      if ((todo >> k) & 1u) m = val[k] < m ? val[k] : m;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      const uint64_t o = cs_shfl_xor(m, s);
      m = o < m ? o : m;
    }
#pragma unroll
    for (int k = 0; k < CS_PER_LANE; k++)
      if (((todo >> k) & 1u) && val[k] == m) {
        idx[k] = n;
        todo &= ~(1u << k);
      }
    if (WRITE) {
      if (own && lane == 0) {
        out[toff + n * WORDS] = (uint32_t)m;
        if (WORDS == 2) out[toff + n * WORDS + 1] = (uint32_t)(m >> 32);
      }
    } else {
      h ^= m + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    }
    n++;
  }
  const uint32_t bits = cs_bits(n);
  if (!WRITE) {
    if (lane == 0) {
      info_n[b] = n;
      hash[b] = (h ^ n) * 0xBF58476D1CE4E5B9ull;
    }
    return;
  }
  // ---- packed indices: word w of the block holds positions [w*32/bits, (w+1)*32/bits)
  const uint32_t eoff = enc_off[b];
  if (bits) {
    const uint32_t per = 32 / bits;             // values per word
    const uint32_t nwords = (bits * d.bvox + 31) / 32;
    // every lane contributes its values with atomicOr-free packing: values of one word sit in
    // `per` consecutive positions, i.e. in `per` consecutive lanes (or the same lane for per > 32)
#pragma unroll
    for (int k = 0; k < CS_PER_LANE; k++) {
      const uint32_t p = lane + 32 * k;
      if (32 * k >= d.bvox) break;
      const uint32_t v = ((have >> k) & 1u) ? idx[k] : 0u;
      uint32_t word = v << ((p % per) * bits);
      // OR-reduce over the aligned group of `per` lanes (per is a power of two <= 32)
      for (uint32_t s = 1; s < per; s <<= 1) word |= __shfl_xor_sync(CS_FULL, word, s);
      if (p < d.bvox && (p % per) == 0 && p / per < nwords) out[eoff + p / per] = word;
    }
  }
  if (lane == 0) {
    // header: blocks are in raster order at the start of the channel
    out[2 * b] = toff | (bits << 24);
    out[2 * b + 1] = eoff;
  }
}

__global__ void __launch_bounds__(256) k_iota32(uint32_t* p, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = i;
}

// sorted (hash, block): the head of every run of equal hashes owns the table (the sort is stable
// and the blocks entered it in ascending order, so the head is the smallest block of the run)
__global__ void __launch_bounds__(256)
    k_cseg_heads(const unsigned long long* __restrict__ shash, uint32_t n, uint32_t* __restrict__ headpos) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) headpos[i] = (i == 0 || shash[i - 1] != shash[i]) ? i : 0u;
}
__global__ void __launch_bounds__(256)
    k_cseg_owner(const uint32_t* __restrict__ headpos, const uint32_t* __restrict__ sblock, uint32_t n,
                 uint32_t* __restrict__ owner) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) owner[sblock[i]] = sblock[headpos[i]];  // headpos: inclusive max-scan of the head positions
}

template <int WORDS>
__global__ void __launch_bounds__(256)
    k_cseg_sizes(const uint32_t* __restrict__ n, const uint32_t* __restrict__ owner, uint32_t nblock, uint32_t bvox,
                 uint32_t* __restrict__ size) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblock) return;
  const uint32_t bits = cs_bits(n[b]);
  size[b] = (bits * bvox + 31) / 32 + (owner[b] == b ? n[b] * WORDS : 0u);
}

// offsets from the channel start: indices at 2*nblock + scan[b]; own tables right after them
__global__ void __launch_bounds__(256)
    k_cseg_offsets(const uint32_t* __restrict__ n, const uint32_t* __restrict__ owner, const uint32_t* __restrict__ scan,
                   uint32_t nblock, uint32_t bvox, uint32_t* __restrict__ enc_off, uint32_t* __restrict__ tab_off) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nblock) return;
  enc_off[b] = 2 * nblock + scan[b];
  const uint32_t o = owner[b];
  const uint32_t obits = cs_bits(n[o]);
  tab_off[b] = 2 * nblock + scan[o] + (obits * bvox + 31) / 32;
}

template <typename T>
__global__ void __launch_bounds__(256)
    k_cseg_decode(const uint32_t* __restrict__ in, uint64_t nwords, CsegDims d, T* __restrict__ out, uint32_t* err) {
  constexpr int WORDS = sizeof(T) / 4;
  const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t n = (uint64_t)d.sx * d.sy * d.sz;
  if (i >= n) return;
  const uint32_t x = (uint32_t)(i % d.sx), y = (uint32_t)((i / d.sx) % d.sy), z = (uint32_t)(i / ((uint64_t)d.sx * d.sy));
  const uint64_t b = (x / d.bx) + (uint64_t)d.gx * ((y / d.by) + (uint64_t)d.gy * (z / d.bz));
  if (2 * b + 1 >= nwords) { *err = 1; return; }
  const uint32_t h0 = in[2 * b], h1 = in[2 * b + 1];
  const uint32_t bits = h0 >> 24;
  const uint64_t toff = h0 & 0xFFFFFFu, voff = h1;
  if (!(bits == 0 || bits == 1 || bits == 2 || bits == 4 || bits == 8 || bits == 16 || bits == 32)) { *err = 1; return; }
  uint64_t idx = 0;
  if (bits) {
    const uint64_t bitpos = (uint64_t)(((z % d.bz) * d.by + (y % d.by)) * d.bx + (x % d.bx)) * bits;
    const uint64_t w = voff + bitpos / 32;
    if (w >= nwords) { *err = 1; return; }
    idx = (in[w] >> (bitpos % 32)) & (bits == 32 ? 0xFFFFFFFFu : ((1u << bits) - 1u));
  }
  const uint64_t tw = toff + idx * WORDS;
  if (tw + WORDS > nwords) { *err = 1; return; }
  uint64_t v = in[tw];
  if (WORDS == 2) v |= (uint64_t)in[tw + 1] << 32;
  out[i] = (T)v;
}

static int cseg_dims(uint64_t sx, uint64_t sy, uint64_t sz, uint32_t bx, uint32_t by, uint32_t bz, CsegDims* d) {
  IGN_REQUIRE(sx && sy && sz && bx && by && bz, IGN_ERR_INVALID, "cseg: empty chunk or block");
  IGN_REQUIRE((uint64_t)bx * by * bz <= CS_MAX_BVOX, IGN_ERR_UNSUPPORTED, "cseg: blocks of more than %d voxels are not supported", CS_MAX_BVOX);
  IGN_REQUIRE(sx < (1u << 20) && sy < (1u << 20) && sz < (1u << 20), IGN_ERR_OVERFLOW, "cseg: chunk extent too large");
  d->sx = (uint32_t)sx; d->sy = (uint32_t)sy; d->sz = (uint32_t)sz;
  d->bx = bx; d->by = by; d->bz = bz;
  d->gx = (uint32_t)((sx + bx - 1) / bx); d->gy = (uint32_t)((sy + by - 1) / by); d->gz = (uint32_t)((sz + bz - 1) / bz);
  d->bvox = bx * by * bz;
  IGN_REQUIRE((uint64_t)d->gx * d->gy * d->gz < (1u << 23), IGN_ERR_OVERFLOW,
              "cseg: %llu blocks exceed the format's 24-bit table offsets; encode Precomputed chunks, not whole volumes",
              (unsigned long long)((uint64_t)d->gx * d->gy * d->gz));
  return IGN_OK;
}

// one channel; out_dev may be NULL (size query).  *n_words = words of the channel stream.
template <typename T>
static int cseg_encode_channel(ign_ctx* ctx, const T* in, const CsegDims& d, uint32_t* out_dev, uint64_t cap_words,
                               uint64_t* n_words) {
  constexpr int WORDS = sizeof(T) / 4;
  const uint32_t nb = d.gx * d.gy * d.gz;
  const size_t keep = ctx->scratch_used;
  const bool own = keep == 0;
  size_t sortb = 0, scanb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sortb, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                  (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)nb);
  cub::DeviceScan::ExclusiveSum(nullptr, scanb, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)nb + 1);
  {
    size_t mb = 0;
    cub::DeviceScan::InclusiveScan(nullptr, mb, (const uint32_t*)nullptr, (uint32_t*)nullptr, cub::Max(), (int)nb);
    if (mb > scanb) scanb = mb;
  }
  const size_t tmpb = (sortb > scanb ? sortb : scanb) + 256;
  if (own) IGN_TRY(scratch_reserve(ctx, 2 * align_up((size_t)nb * 8, 256) + 8 * align_up(((size_t)nb + 1) * 4, 256) + tmpb + 4096));
  auto fail = [&](int rc) {
    ctx->scratch_used = keep;
    return rc;
  };
  unsigned long long* hash = (unsigned long long*)scratch_take(ctx, (size_t)nb * 8);
  unsigned long long* shash = (unsigned long long*)scratch_take(ctx, (size_t)nb * 8);
  uint32_t* n = (uint32_t*)scratch_take(ctx, ((size_t)nb + 1) * 4);
  uint32_t* blk = (uint32_t*)scratch_take(ctx, ((size_t)nb + 1) * 4);
  uint32_t* sblk = (uint32_t*)scratch_take(ctx, ((size_t)nb + 1) * 4);
  uint32_t* owner = (uint32_t*)scratch_take(ctx, ((size_t)nb + 1) * 4);
  uint32_t* size = (uint32_t*)scratch_take(ctx, ((size_t)nb + 1) * 4);
  uint32_t* scan = (uint32_t*)scratch_take(ctx, ((size_t)nb + 1) * 4);
  uint32_t* enc_off = (uint32_t*)scratch_take(ctx, ((size_t)nb + 1) * 4);
  uint32_t* tab_off = (uint32_t*)scratch_take(ctx, ((size_t)nb + 1) * 4);
  void* tmp = scratch_take(ctx, tmpb);
  if (!hash || !shash || !n || !blk || !sblk || !owner || !size || !scan || !enc_off || !tab_off || !tmp) {
    set_error("scratch arena too small (cseg encode)");
    return fail(IGN_ERR_NOMEM);
  }
#define CS_CUDA(call)                                                                  \
  do {                                                                                 \
    cudaError_t _e = (call);                                                           \
    if (_e != cudaSuccess) {                                                           \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return fail(IGN_ERR_CUDA);                                                       \
    }                                                                                  \
  } while (0)
#define CS_LAUNCH(kernel, g, b, ...)                  \
  do {                                                \
    kernel<<<(g), (b), 0, ctx->stream>>>(__VA_ARGS__); \
    ctx->launches++;                                  \
    CS_CUDA(cudaGetLastError());                      \
  } while (0)
  const unsigned gw = blocks_for((uint64_t)nb * 32, 128);
  if (d.bvox <= 512)
    CS_LAUNCH((k_cseg_scan<T, false, 16>), gw, 128, in, d, (uint64_t)nb, n, hash, (const uint32_t*)nullptr,
              (const uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr);
  else
    CS_LAUNCH((k_cseg_scan<T, false, 32>), gw, 128, in, d, (uint64_t)nb, n, hash, (const uint32_t*)nullptr,
              (const uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr);
  CS_LAUNCH(k_iota32, blocks_for(nb, 256), 256, blk, nb);
  {
    size_t tb = tmpb;
    CS_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, hash, shash, blk, sblk, (int)nb, 0, 64, ctx->stream));
    ctx->launches += 9;
  }
  CS_LAUNCH(k_cseg_heads, blocks_for(nb, 256), 256, shash, nb, enc_off);  // enc_off / tab_off: free until the offsets pass
  {
    size_t tb = tmpb;
    CS_CUDA(cub::DeviceScan::InclusiveScan(tmp, tb, enc_off, tab_off, cub::Max(), (int)nb, ctx->stream));
    ctx->launches += 2;
  }
  CS_LAUNCH(k_cseg_owner, blocks_for(nb, 256), 256, tab_off, sblk, nb, owner);
  CS_LAUNCH((k_cseg_sizes<WORDS>), blocks_for(nb, 256), 256, n, owner, nb, d.bvox, size);
  CS_CUDA(cudaMemsetAsync(size + nb, 0, 4, ctx->stream));
  {
    size_t tb = tmpb;
    CS_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, size, scan, (int)nb + 1, ctx->stream));
    ctx->launches += 2;
  }
  uint32_t total = 0;
  {
    const int rc = small_d2h(ctx, &total, scan + nb, 4);
    if (rc != IGN_OK) return fail(rc);
    const int rc2 = small_sync(ctx);
    if (rc2 != IGN_OK) return fail(rc2);
  }
  const uint64_t words = 2ull * nb + total;
  *n_words = words;
  if (words > 0xFFFFFFull + 1024) {
    set_error("cseg: the encoded chunk (%llu words) exceeds the format's 24-bit table offsets", (unsigned long long)words);
    return fail(IGN_ERR_OVERFLOW);
  }
  if (out_dev != nullptr && words <= cap_words) {
    CS_LAUNCH(k_cseg_offsets, blocks_for(nb, 256), 256, n, owner, scan, nb, d.bvox, enc_off, tab_off);
    if (d.bvox <= 512)
      CS_LAUNCH((k_cseg_scan<T, true, 16>), gw, 128, in, d, (uint64_t)nb, (uint32_t*)nullptr, (unsigned long long*)nullptr,
                enc_off, tab_off, owner, out_dev);
    else
      CS_LAUNCH((k_cseg_scan<T, true, 32>), gw, 128, in, d, (uint64_t)nb, (uint32_t*)nullptr, (unsigned long long*)nullptr,
                enc_off, tab_off, owner, out_dev);
  }
  ctx->scratch_used = keep;
  return IGN_OK;
#undef CS_CUDA
#undef CS_LAUNCH
}

}  // namespace ign

using namespace ign;

extern "C" {

int ign_cseg_encode_dev(ign_ctx* ctx, const void* labels, int dtype, uint64_t sx, uint64_t sy, uint64_t sz,
                        uint64_t sc, uint32_t bx, uint32_t by, uint32_t bz, uint32_t* out, uint64_t cap_words,
                        uint64_t* n_words) {
  IGN_TRY(activate(ctx));
  IGN_REQUIRE(labels && n_words && sc >= 1, IGN_ERR_INVALID, "null argument");
  IGN_REQUIRE(dtype == IGN_U32 || dtype == IGN_U64, IGN_ERR_UNSUPPORTED, "compressed_segmentation holds uint32 / uint64 labels");
  CsegDims d;
  IGN_TRY(cseg_dims(sx, sy, sz, bx, by, bz, &d));
  const uint64_t n = sx * sy * sz;
  uint64_t at = sc;  // the channel offset table comes first
  std::vector<uint32_t> chan_off(sc, 0);
  for (uint64_t c = 0; c < sc; c++) {
    chan_off[c] = (uint32_t)at;
    uint64_t w = 0;
    uint32_t* dst = (out && at < cap_words) ? out + at : nullptr;
    const uint64_t room = (out && at < cap_words) ? cap_words - at : 0;
    if (dtype == IGN_U32) IGN_TRY(cseg_encode_channel<uint32_t>(ctx, (const uint32_t*)labels + c * n, d, dst, room, &w));
    else IGN_TRY(cseg_encode_channel<uint64_t>(ctx, (const uint64_t*)labels + c * n, d, dst, room, &w));
    at += w;
  }
  *n_words = at;
  if (out && at <= cap_words) IGN_TRY(small_h2d(ctx, out, chan_off.data(), sc * 4));
  if (out) IGN_CUDA(cudaStreamSynchronize(ctx->stream));
  return IGN_OK;
}

int ign_cseg_decode_dev(ign_ctx* ctx, const uint32_t* in, uint64_t n_words, int dtype, uint64_t sx, uint64_t sy,
                        uint64_t sz, uint64_t sc, uint32_t bx, uint32_t by, uint32_t bz, void* out) {
  IGN_TRY(activate(ctx));
  IGN_REQUIRE(in && out && sc >= 1 && n_words >= sc, IGN_ERR_INVALID, "bad argument");
  IGN_REQUIRE(dtype == IGN_U32 || dtype == IGN_U64, IGN_ERR_UNSUPPORTED, "compressed_segmentation holds uint32 / uint64 labels");
  CsegDims d;
  IGN_TRY(cseg_dims(sx, sy, sz, bx, by, bz, &d));
  const uint64_t n = sx * sy * sz;
  std::vector<uint32_t> chan_off(sc, 0);
  IGN_CUDA(cudaMemcpyAsync(chan_off.data(), in, sc * 4, cudaMemcpyDeviceToHost, ctx->stream));
  IGN_CUDA(cudaStreamSynchronize(ctx->stream));
  const size_t keep = ctx->scratch_used;
  if (keep == 0) IGN_TRY(scratch_reserve(ctx, 4096));
  uint32_t* err = (uint32_t*)scratch_take(ctx, 256);
  IGN_REQUIRE(err != nullptr, IGN_ERR_NOMEM, "scratch arena too small (cseg decode)");
  IGN_CUDA(cudaMemsetAsync(err, 0, 4, ctx->stream));
  for (uint64_t c = 0; c < sc; c++) {
    const uint64_t base = chan_off[c];
    if (base > n_words) {
      ctx->scratch_used = keep;
      set_error("cseg: channel offset outside the stream");
      return IGN_ERR_INVALID;
    }
    if (dtype == IGN_U32)
      IGN_LAUNCH(ctx, (k_cseg_decode<uint32_t>), blocks_for(n, 256), 256, 0, in + base, n_words - base, d, (uint32_t*)out + c * n, err);
    else
      IGN_LAUNCH(ctx, (k_cseg_decode<uint64_t>), blocks_for(n, 256), 256, 0, in + base, n_words - base, d, (uint64_t*)out + c * n, err);
  }
  uint32_t herr = 0;
  IGN_CUDA(cudaMemcpyAsync(&herr, err, 4, cudaMemcpyDeviceToHost, ctx->stream));
  IGN_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->scratch_used = keep;
  IGN_REQUIRE(herr == 0, IGN_ERR_INVALID, "cseg: malformed stream");
  return IGN_OK;
}

// host-buffer wrappers: encode returns the words needed in *n_words (call with out == NULL first, or
// with a capacity of sc + 2*blocks + 3*voxels words, the worst case)
int ign_cseg_encode(ign_ctx* ctx, const void* labels, int dtype, uint64_t sx, uint64_t sy, uint64_t sz, uint64_t sc,
                    uint32_t bx, uint32_t by, uint32_t bz, uint32_t* out, uint64_t cap_words, uint64_t* n_words) {
  IGN_TRY(activate(ctx));
  IGN_REQUIRE(labels && n_words, IGN_ERR_INVALID, "null argument");
  const int es = dtype_size(dtype);
  IGN_REQUIRE(dtype == IGN_U32 || dtype == IGN_U64, IGN_ERR_UNSUPPORTED, "compressed_segmentation holds uint32 / uint64 labels");
  const uint64_t n = sx * sy * sz * sc;
  scratch_reset(ctx);
  IGN_TRY(scratch_reserve(ctx, align_up(n * es, 256) + align_up((out ? cap_words : 0) * 4, 256) + (64ull << 20)));
  void* d_in = scratch_take(ctx, n * es);
  uint32_t* d_out = out ? (uint32_t*)scratch_take(ctx, cap_words * 4) : nullptr;
  IGN_REQUIRE(d_in && (!out || d_out), IGN_ERR_NOMEM, "scratch arena too small (cseg)");
  IGN_CUDA(cudaMemcpyAsync(d_in, labels, n * es, cudaMemcpyHostToDevice, ctx->stream));
  int rc = ign_cseg_encode_dev(ctx, d_in, dtype, sx, sy, sz, sc, bx, by, bz, d_out, cap_words, n_words);
  if (rc == IGN_OK && out && *n_words <= cap_words) {
    cudaError_t e = cudaMemcpyAsync(out, d_out, *n_words * 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      set_error("cseg D2H: %s", cudaGetErrorString(e));
      rc = IGN_ERR_CUDA;
    }
  }
  scratch_reset(ctx);
  return rc;
}

int ign_cseg_decode(ign_ctx* ctx, const uint32_t* in, uint64_t n_words, int dtype, uint64_t sx, uint64_t sy, uint64_t sz,
                    uint64_t sc, uint32_t bx, uint32_t by, uint32_t bz, void* out) {
  IGN_TRY(activate(ctx));
  IGN_REQUIRE(in && out, IGN_ERR_INVALID, "null argument");
  const int es = dtype_size(dtype);
  IGN_REQUIRE(dtype == IGN_U32 || dtype == IGN_U64, IGN_ERR_UNSUPPORTED, "compressed_segmentation holds uint32 / uint64 labels");
  const uint64_t n = sx * sy * sz * sc;
  scratch_reset(ctx);
  IGN_TRY(scratch_reserve(ctx, align_up(n * es, 256) + align_up(n_words * 4, 256) + (1 << 20)));
  uint32_t* d_in = (uint32_t*)scratch_take(ctx, n_words * 4);
  void* d_out = scratch_take(ctx, n * es);
  IGN_REQUIRE(d_in && d_out, IGN_ERR_NOMEM, "scratch arena too small (cseg)");
  IGN_CUDA(cudaMemcpyAsync(d_in, in, n_words * 4, cudaMemcpyHostToDevice, ctx->stream));
  int rc = ign_cseg_decode_dev(ctx, d_in, n_words, dtype, sx, sy, sz, sc, bx, by, bz, d_out);
  if (rc == IGN_OK) {
    cudaError_t e = cudaMemcpyAsync(out, d_out, n * es, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      set_error("cseg D2H: %s", cudaGetErrorString(e));
      rc = IGN_ERR_CUDA;
    }
  }
  scratch_reset(ctx);
  return rc;
}

}  // extern "C"
