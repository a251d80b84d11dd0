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
  void *cfh = nullptr, *dfh = nullptr;  // N == 2048 only: cf / 2, df / 2 (pairwise post-processing)
  void* tw32 = nullptr;  // N == 2048 only: [i][lane] = exp(-2 pi i * i * lane / 1024) for the 32 x 32 warp FFT
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

// returns nullptr on failure (ctx->err set)
frad_plan* frad_get_plan(frad_ctx* ctx, int N, int prec);
