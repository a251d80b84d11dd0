// Shared host/device plumbing for libfradb200 (context, error handling, plan cache).
#pragma once
#define FRAD_DQ_LEN 65536           // entries of the global |q|^(4/3) table (frad_ensure_dq64): 512 KiB, L2-resident.  Loud low bins of
                                    // a 16-bit stream reach five digits, and one such lane sends its whole warp through pow()
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include "../../include/fradb200.h"

#define FRAD_MAX_RADIX_PASSES 20
#define FRAD_MAX_FRAME 28672          // largest compact frame length (fourier/profiles.py:14-23)
#define FRAD_NEVENTS 64

struct frad_plan;  // transform plan, see plan.cuh

struct frad_ctx {
  int device = 0;
  int sm_count = 0;
  size_t smem_optin = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t s_in = nullptr, s_out = nullptr;   // copy streams of the pipelined host-pointer entry points
  std::vector<cudaEvent_t> pipe_events;
  std::string err;
  uint64_t launches = 0;
  cudaEvent_t events[FRAD_NEVENTS] = {};
  // grow-only device scratch arenas used by the host-pointer entry points
  void* scratch[8] = {};
  size_t scratch_cap[8] = {};
  std::map<uint64_t, frad_plan*> plans;      // key: N | prec<<40
  std::map<uint32_t, void*> rs_tables;       // key: dsize<<8 | csize -> device tables
  std::map<uint64_t, void*> band_tables;     // key: N | srate<<32
  void* crc_tables = nullptr;                // device CRC tables (crc.cu)
  void* dequant_table = nullptr;             // |q|^(4/3) for small |q| (p1 decode)
  int opt_force_generic = 0;                 // FRAD_OPT_FORCE_GENERIC: bypass the N = 2048 fast path (tests)
  int opt_exact_quant = 0;                   // FRAD_OPT_EXACT_QUANT: always take the f64 quantiser branch (tests)
};

static inline int frad_fail(frad_ctx* ctx, int code, const char* what, const char* detail = nullptr) {
  if (ctx) {
    ctx->err = what;
    if (detail) { ctx->err += ": "; ctx->err += detail; }
  }
  return code;
}

// Table uploads (once per table, never on the timed path).  A cudaMemcpy from pageable memory may return as soon as the source
// has been staged: the DMA itself is ordered on the legacy stream, which the library's non-blocking streams do not wait for.
// A kernel launched right after the first use of a table could read it before it had arrived (seen as a rare wrong
// Reed-Solomon word in the first test run on a cold box), so the device is synchronised before the table is published.
static inline cudaError_t frad_upload(void* d, const void* h, size_t n) {
  cudaError_t e = cudaMemcpy(d, h, n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  return e;
}

#define FRAD_CUDA(ctx, call)                                                      \
  do {                                                                            \
    cudaError_t _e = (call);                                                      \
    if (_e != cudaSuccess) { cudaGetLastError(); /* reported here: a later launch check must not see it again */ \
      return frad_fail((ctx), FRAD_E_CUDA, #call, cudaGetErrorString(_e)); }        \
  } while (0)

#define FRAD_LAUNCH_CHECK(ctx)                                                    \
  do {                                                                            \
    (ctx)->launches++;                                                            \
    cudaError_t _e = cudaGetLastError();                                          \
    if (_e != cudaSuccess) return frad_fail((ctx), FRAD_E_CUDA, "kernel launch", cudaGetErrorString(_e)); \
  } while (0)

// grow-only scratch arena `slot`; returns nullptr on failure (ctx->err set)
void* frad_scratch(frad_ctx* ctx, int slot, size_t bytes);

static inline int frad_dtype_size(int dt) {
  switch (dt) {
    case FRAD_F32: case FRAD_S32: case FRAD_U32: return 4;
    case FRAD_F64: case FRAD_S64: case FRAD_U64: return 8;
    case FRAD_S16: case FRAD_U16: case FRAD_F16: return 2;
    case FRAD_S8: case FRAD_U8: return 1;
    default: return 0;
  }
}

int frad_ensure_dq64(struct frad_ctx* ctx);   // transform.cu
