// Context, memory and timing entry points of libfradb200 (include/fradb200.h).
#include "common.cuh"

static std::string g_create_err;

void frad_plan_free(frad_plan* p);  // transform.cu

extern "C" int frad_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

extern "C" int frad_device_pci_bus_id(int device, char* out, int cap) {
  if (!out || cap < 16) return FRAD_E_ARG;
  if (cudaDeviceGetPCIBusId(out, cap, device) != cudaSuccess) { cudaGetLastError(); return FRAD_E_CUDA; }
  return FRAD_OK;
}

extern "C" int frad_ctx_create(int device, frad_ctx** out) {
  if (!out) return FRAD_E_ARG;
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    g_create_err = std::string("no usable CUDA device: ") + cudaGetErrorString(e);
    return FRAD_E_CUDA;   // no CPU fallback by design
  }
  if (device < 0 || device >= n) { g_create_err = "device index out of range"; return FRAD_E_ARG; }
  frad_ctx* ctx = new frad_ctx();
  ctx->device = device;
  if ((e = cudaSetDevice(device)) != cudaSuccess ||
      (e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) {
    g_create_err = std::string("cudaSetDevice/cudaStreamCreate: ") + cudaGetErrorString(e);
    delete ctx;
    return FRAD_E_CUDA;
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  for (int i = 0; i < FRAD_NEVENTS; i++) cudaEventCreate(&ctx->events[i]);
  cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking);
  *out = ctx;
  return FRAD_OK;
}

extern "C" void frad_ctx_destroy(frad_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& kv : ctx->plans) frad_plan_free(kv.second);
  for (auto& kv : ctx->rs_tables) cudaFree(kv.second);
  for (auto& kv : ctx->band_tables) cudaFree(kv.second);
  if (ctx->crc_tables) cudaFree(ctx->crc_tables);
  if (ctx->dequant_table) cudaFree(ctx->dequant_table);
  for (int i = 0; i < 8; i++) if (ctx->scratch[i]) cudaFree(ctx->scratch[i]);
  for (int i = 0; i < FRAD_NEVENTS; i++) if (ctx->events[i]) cudaEventDestroy(ctx->events[i]);
  for (auto e : ctx->pipe_events) cudaEventDestroy(e);
  if (ctx->s_in) { cudaStreamSynchronize(ctx->s_in); cudaStreamDestroy(ctx->s_in); }
  if (ctx->s_out) { cudaStreamSynchronize(ctx->s_out); cudaStreamDestroy(ctx->s_out); }
  cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* frad_last_error(frad_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

extern "C" int frad_sync(frad_ctx* ctx) {
  if (!ctx) return FRAD_E_ARG;
  FRAD_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return FRAD_OK;
}

extern "C" void* frad_stream(frad_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" int frad_sm_count(frad_ctx* ctx) { return ctx ? ctx->sm_count : 0; }
extern "C" uint64_t frad_kernel_launches(frad_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" void* frad_host_alloc(frad_ctx* ctx, size_t bytes) {
  if (!ctx) return nullptr;
  void* p = nullptr;
  cudaSetDevice(ctx->device);
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
    frad_fail(ctx, FRAD_E_NOMEM, "cudaHostAlloc failed");
    return nullptr;
  }
  return p;
}
extern "C" void frad_host_free(frad_ctx* ctx, void* p) { (void)ctx; if (p) cudaFreeHost(p); }

// Page-lock memory the caller already owns (e.g. a shared-memory mapping the ranks of one box assemble their output in), so
// that copies to and from it are asynchronous like those of frad_host_alloc memory.
extern "C" int frad_host_register(frad_ctx* ctx, void* p, size_t bytes) {
  if (!ctx || !p || !bytes) return FRAD_E_ARG;
  cudaSetDevice(ctx->device);
  FRAD_CUDA(ctx, cudaHostRegister(p, bytes, cudaHostRegisterPortable));
  return FRAD_OK;
}
extern "C" int frad_host_unregister(frad_ctx* ctx, void* p) {
  if (!ctx || !p) return FRAD_E_ARG;
  FRAD_CUDA(ctx, cudaHostUnregister(p));
  return FRAD_OK;
}

extern "C" void* frad_dev_alloc(frad_ctx* ctx, size_t bytes) {
  if (!ctx) return nullptr;
  void* p = nullptr;
  cudaSetDevice(ctx->device);
  if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();
    frad_fail(ctx, FRAD_E_NOMEM, "cudaMalloc failed");
    return nullptr;
  }
  return p;
}
extern "C" void frad_dev_free(frad_ctx* ctx, void* p) {
  if (!ctx || !p) return;
  cudaStreamSynchronize(ctx->stream);
  cudaFree(p);
}

extern "C" int frad_h2d(frad_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (!ctx) return FRAD_E_ARG;
  if (bytes) FRAD_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return FRAD_OK;
}
extern "C" int frad_d2h(frad_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (!ctx) return FRAD_E_ARG;
  if (bytes) FRAD_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return FRAD_OK;
}
extern "C" int frad_memset_dev(frad_ctx* ctx, void* dst, int value, size_t bytes) {
  if (!ctx) return FRAD_E_ARG;
  if (bytes) FRAD_CUDA(ctx, cudaMemsetAsync(dst, value, bytes, ctx->stream));
  return FRAD_OK;
}

extern "C" int frad_event_record(frad_ctx* ctx, int slot) {
  if (!ctx || slot < 0 || slot >= FRAD_NEVENTS) return FRAD_E_ARG;
  FRAD_CUDA(ctx, cudaEventRecord(ctx->events[slot], ctx->stream));
  return FRAD_OK;
}
extern "C" int frad_event_elapsed_ms(frad_ctx* ctx, int a, int b, float* ms) {
  if (!ctx || !ms || a < 0 || b < 0 || a >= FRAD_NEVENTS || b >= FRAD_NEVENTS) return FRAD_E_ARG;
  FRAD_CUDA(ctx, cudaEventSynchronize(ctx->events[b]));
  FRAD_CUDA(ctx, cudaEventElapsedTime(ms, ctx->events[a], ctx->events[b]));
  return FRAD_OK;
}

void* frad_scratch(frad_ctx* ctx, int slot, size_t bytes) {
  if (bytes <= ctx->scratch_cap[slot] && ctx->scratch[slot]) return ctx->scratch[slot];
  if (ctx->scratch[slot]) {
    cudaStreamSynchronize(ctx->stream);
    if (ctx->s_in) cudaStreamSynchronize(ctx->s_in);
    if (ctx->s_out) cudaStreamSynchronize(ctx->s_out);
    cudaFree(ctx->scratch[slot]);
    ctx->scratch[slot] = nullptr;
    ctx->scratch_cap[slot] = 0;
  }
  size_t cap = bytes + bytes / 8 + 256;
  void* p = nullptr;
  if (cudaMalloc(&p, cap) != cudaSuccess) {
    cudaGetLastError();
    frad_fail(ctx, FRAD_E_NOMEM, "scratch cudaMalloc failed");
    return nullptr;
  }
  ctx->scratch[slot] = p;
  ctx->scratch_cap[slot] = cap;
  return p;
}

extern "C" int frad_set_option(frad_ctx* ctx, int option, int value) {
  if (!ctx) return FRAD_E_ARG;
  switch (option) {
    case FRAD_OPT_FORCE_GENERIC: ctx->opt_force_generic = value; return FRAD_OK;
    case FRAD_OPT_EXACT_QUANT: ctx->opt_exact_quant = value; return FRAD_OK;
    default: return frad_fail(ctx, FRAD_E_ARG, "unknown option");
  }
}
