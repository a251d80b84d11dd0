// Transform plans: twiddle tables and radix schedules for the DCT-II/III kernels.
//
// DCT-II of even length N is computed through one complex FFT of length M = N/2
// (Makhoul reordering + real-FFT split), see DESIGN.md "Transform".  Tables (T = float|double):
//   tw[t]  = exp(-2 pi i t / M)                       t in [0, M)     (FFT passes; conjugated for the inverse)
//   cf[k]  = exp(-i pi k / 2N) / N                    k in [0, M]     forward post-twiddle
//   df[k]  = cf[k] * exp(-2 pi i k / N)               k in [0, M]
//   ci[k]  = exp(+i pi k / 2N)                        k in [0, M]     inverse pre-twiddle
//   ei[k]  = exp(+2 pi i k / N)                       k in [0, M]
// Lengths that are odd, or whose M has a prime factor > 7, use the direct O(N^2) kernel with
//   ct[p]  = cos(pi p / 2N)                            p in [0, 4N)   (always double)
#pragma once
#include "common.cuh"

struct frad_plan {
  int N = 0, M = 0;
  int prec = 0;          // FRAD_PREC_F32 / FRAD_PREC_F64
  int direct = 0;        // 1: O(N^2) path
  int npass = 0;
  int radix[FRAD_MAX_RADIX_PASSES] = {};
  void *tw = nullptr, *cf = nullptr, *df = nullptr, *ci = nullptr, *ei = nullptr;
  void *cfh = nullptr, *dfh = nullptr;  // warp-FFT lengths (N == 2048, frad_warp_r1) only: cf / 2, df / 2 (pairwise post-processing)
  void* tw32 = nullptr;  // warp-FFT lengths only: [i][lane] = exp(-2 pi i * i * (lane mod R1) / M), R1 = M / 32 (N == 2048: lane)
  double* ct = nullptr;
};

// device-side view passed by value to kernels
struct plan_dev {
  int N, M, npass;
  int radix[FRAD_MAX_RADIX_PASSES];
  const void *tw, *cf, *df, *ci, *ei, *cfh, *dfh;
  const double* ct;
};

static inline plan_dev plan_view(const frad_plan* p) {
  plan_dev d;
  d.N = p->N; d.M = p->M; d.npass = p->npass;
  for (int i = 0; i < FRAD_MAX_RADIX_PASSES; i++) d.radix[i] = p->radix[i];
  d.tw = p->tw; d.cf = p->cf; d.df = p->df; d.ci = p->ci; d.ei = p->ei; d.ct = p->ct; d.cfh = p->cfh; d.dfh = p->dfh;
  return d;
}

// Frame lengths N = 64 R1 < 2048 of the multi-sequence warp-FFT kernels (enc2048.cu: encs_kernel, decs.cu): R1 a power of two
// (N = 128 ... 1024) or one of fftmix.cuh's mixed radices (N = 192 ... 1792).  Returns R1, or 0 for every other length.
static inline int frad_warp_r1(int N) {
  if (N <= 0 || (N % 64) != 0) return 0;
  switch (N / 64) {
    case 2: case 4: case 8: case 16:
    case 3: case 5: case 6: case 7: case 10: case 12: case 14: case 20: case 24: case 28: return N / 64;
    default: return 0;
  }
}
static inline bool frad_warp_r1_mixed(int r1) { return r1 != 0 && (r1 & (r1 - 1)) != 0; }

// returns nullptr on failure (ctx->err set)
frad_plan* frad_get_plan(frad_ctx* ctx, int N, int prec);
