// ctx.cu -- context, memory, timers, error plumbing of libigneous_b200
#include <stdarg.h>

#include "common.cuh"

namespace ign {

static thread_local std::string g_err;

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

int activate(ign_ctx* ctx) {
  IGN_REQUIRE(ctx != nullptr, IGN_ERR_INVALID, "null ign_ctx");
  IGN_CUDA(cudaSetDevice(ctx->device));
  return IGN_OK;
}

void scratch_reset(ign_ctx* ctx) { ctx->scratch_used = 0; }

int scratch_reserve(ign_ctx* ctx, size_t total) {
  total = align_up(total + 4096, 1 << 20);
  if (total <= ctx->scratch_bytes) return IGN_OK;
  IGN_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->scratch) IGN_CUDA(cudaFree(ctx->scratch));
  ctx->scratch = nullptr;
  ctx->scratch_bytes = 0;
  cudaError_t e = cudaMalloc((void**)&ctx->scratch, total);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("scratch arena: cudaMalloc(%zu) failed: %s", total, cudaGetErrorString(e));
    return IGN_ERR_NOMEM;
  }
  ctx->scratch_bytes = total;
  return IGN_OK;
}

void* scratch_take(ign_ctx* ctx, size_t bytes) {
  size_t off = align_up(ctx->scratch_used, 256);
  if (off + bytes > ctx->scratch_bytes) return nullptr;
  ctx->scratch_used = off + bytes;
  return ctx->scratch + off;
}

}  // namespace ign

using namespace ign;

extern "C" {

int ign_version(void) { return 100; }

const char* ign_last_error(void) { return g_err.c_str(); }

int ign_device_count(int* n) {
  IGN_REQUIRE(n, IGN_ERR_INVALID, "null out pointer");
  cudaError_t e = cudaGetDeviceCount(n);
  if (e != cudaSuccess) {
    *n = 0;
    set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return IGN_ERR_CUDA;
  }
  return IGN_OK;
}

int ign_init(int device, ign_ctx** out) {
  IGN_REQUIRE(out, IGN_ERR_INVALID, "null out pointer");
  *out = nullptr;
  int n = 0;
  IGN_TRY(ign_device_count(&n));
  IGN_REQUIRE(n > 0, IGN_ERR_CUDA, "no CUDA device visible (this library has no CPU fallback)");
  IGN_REQUIRE(device >= 0 && device < n, IGN_ERR_INVALID, "device %d out of range [0,%d)", device, n);
  IGN_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  IGN_CUDA(cudaGetDeviceProperties(&prop, device));
  IGN_REQUIRE(prop.major >= 10, IGN_ERR_UNSUPPORTED,
              "device %d is sm_%d%d; libigneous_b200 is built for sm_100a only", device,
              prop.major, prop.minor);
  ign_ctx* ctx = new ign_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->scratch = nullptr;
  ctx->scratch_bytes = ctx->scratch_used = 0;
  ctx->launches = 0;
  IGN_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  IGN_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  ctx->pinned_bytes = 1 << 20;
  IGN_CUDA(cudaHostAlloc((void**)&ctx->pinned, ctx->pinned_bytes, cudaHostAllocDefault));
  for (int i = 0; i < 16; i++) {
    IGN_CUDA(cudaEventCreate(&ctx->timers[i][0]));
    IGN_CUDA(cudaEventCreate(&ctx->timers[i][1]));
  }
  *out = ctx;
  return IGN_OK;
}

int ign_destroy(ign_ctx* ctx) {
  if (!ctx) return IGN_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  for (int i = 0; i < 16; i++) {
    cudaEventDestroy(ctx->timers[i][0]);
    cudaEventDestroy(ctx->timers[i][1]);
  }
  cudaStreamDestroy(ctx->stream);
  cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
  return IGN_OK;
}

int ign_sync(ign_ctx* ctx) {
  IGN_TRY(activate(ctx));
  IGN_CUDA(cudaStreamSynchronize(ctx->stream));
  return IGN_OK;
}

int ign_launch_count(ign_ctx* ctx, uint64_t* n) {
  IGN_REQUIRE(ctx && n, IGN_ERR_INVALID, "null argument");
  *n = ctx->launches;
  return IGN_OK;
}

int ign_stream(ign_ctx* ctx, void** stream) {
  IGN_REQUIRE(ctx && stream, IGN_ERR_INVALID, "null argument");
  *stream = (void*)ctx->stream;
  return IGN_OK;
}

int ign_dev_alloc(ign_ctx* ctx, uint64_t bytes, void** dptr) {
  IGN_TRY(activate(ctx));
  IGN_REQUIRE(dptr, IGN_ERR_INVALID, "null out pointer");
  *dptr = nullptr;
  cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 1);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("cudaMalloc(%llu) failed: %s", (unsigned long long)bytes, cudaGetErrorString(e));
    return IGN_ERR_NOMEM;
  }
  return IGN_OK;
}

int ign_dev_free(ign_ctx* ctx, void* dptr) {
  IGN_TRY(activate(ctx));
  IGN_CUDA(cudaStreamSynchronize(ctx->stream));
  IGN_CUDA(cudaFree(dptr));
  return IGN_OK;
}

int ign_host_alloc(ign_ctx* ctx, uint64_t bytes, void** hptr) {
  IGN_TRY(activate(ctx));
  IGN_REQUIRE(hptr, IGN_ERR_INVALID, "null out pointer");
  cudaError_t e = cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error("cudaHostAlloc(%llu) failed: %s", (unsigned long long)bytes, cudaGetErrorString(e));
    return IGN_ERR_NOMEM;
  }
  return IGN_OK;
}

int ign_host_free(ign_ctx* ctx, void* hptr) {
  IGN_TRY(activate(ctx));
  IGN_CUDA(cudaFreeHost(hptr));
  return IGN_OK;
}

int ign_h2d(ign_ctx* ctx, void* dst, const void* src, uint64_t bytes) {
  IGN_TRY(activate(ctx));
  IGN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return IGN_OK;
}

int ign_d2h(ign_ctx* ctx, void* dst, const void* src, uint64_t bytes) {
  IGN_TRY(activate(ctx));
  IGN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return IGN_OK;
}

int ign_d2d(ign_ctx* ctx, void* dst, const void* src, uint64_t bytes) {
  IGN_TRY(activate(ctx));
  IGN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return IGN_OK;
}

int ign_memset(ign_ctx* ctx, void* dst, int byte, uint64_t bytes) {
  IGN_TRY(activate(ctx));
  IGN_CUDA(cudaMemsetAsync(dst, byte, bytes, ctx->stream));
  return IGN_OK;
}

int ign_timer_start(ign_ctx* ctx, int slot) {
  IGN_TRY(activate(ctx));
  IGN_REQUIRE(slot >= 0 && slot < 16, IGN_ERR_INVALID, "timer slot %d out of range", slot);
  IGN_CUDA(cudaEventRecord(ctx->timers[slot][0], ctx->stream));
  return IGN_OK;
}

int ign_timer_stop(ign_ctx* ctx, int slot) {
  IGN_TRY(activate(ctx));
  IGN_REQUIRE(slot >= 0 && slot < 16, IGN_ERR_INVALID, "timer slot %d out of range", slot);
  IGN_CUDA(cudaEventRecord(ctx->timers[slot][1], ctx->stream));
  return IGN_OK;
}

int ign_timer_ms(ign_ctx* ctx, int slot, float* ms) {
  IGN_TRY(activate(ctx));
  IGN_REQUIRE(slot >= 0 && slot < 16 && ms, IGN_ERR_INVALID, "bad timer argument");
  IGN_CUDA(cudaEventSynchronize(ctx->timers[slot][1]));
  IGN_CUDA(cudaEventElapsedTime(ms, ctx->timers[slot][0], ctx->timers[slot][1]));
  return IGN_OK;
}

}  // extern "C"
