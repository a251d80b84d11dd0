// Profile pro/epilogues fused around the transforms (device code shared by transform.cu and
// dct2048.cu).  A "sequence" is one channel of one frame; X is its spectrum in shared or global
// memory, type T (float or double).
//
// Restates (relative to src/libfrad/): fourier/profile0.py:24-42,51-66 (packing, overflow peak,
// NaN scrub); fourier/tools/p1tools.py:15-44 (band edges, masking thresholds, divisor ramps, power-law
// quantiser); fourier/profile1.py:26-40,65-74; fourier/tools/p2tools.py:8-111 and
// fourier/profile2.py:32,85 (TNS).
#pragma once
#include "devmath.cuh"

struct band_table {
  int edge[FRAD_NBANDS + 1];   // unclipped bin edges, Python round() of N/(srate/2)*Hz
  double ath[FRAD_NBANDS];     // min(absolute hearing threshold at band centre, 1)
  double log_base;             // ln(e/2)
  double base;                 // e/2
  // decoded thresholds (e/2)^(sign(t) |t|^0.75) for the threshold integers |t| <= FRAD_THR_TAB (profile1.py:66), computed
  // on the host with libm's pow exactly in the reference's order; larger |t| are evaluated on the device
  double thr_pow[2 * 255 + 1];
  // ramp band of every bin (what ramp_value's search returns; 255 from the start of the last band on), host-built
  uint8_t bidx[FRAD_MAX_FRAME];
};
#define FRAD_THR_TAB 255
__device__ __forceinline__ double decode_threshold(const band_table* bt, int t) {
  if (t >= -FRAD_THR_TAB && t <= FRAD_THR_TAB) return __ldg(&bt->thr_pow[t + FRAD_THR_TAB]);
  const double v = (double)t;
  return pow(bt->base, copysign(pow(fabs(v), 0.75), v));
}

struct band_work {             // per-CTA shared scratch
  double thr[FRAD_NBANDS + 1];
  double step[FRAD_NBANDS + 1];
  double sum[FRAD_NBANDS + 1];
  int lo[FRAD_NBANDS + 1];
  double red[33];
  double redm[FRAD_LPC_LEN * 32];   // block_sum_multi: [value][warp]
  double r[FRAD_LPC_LEN + 3];
  double a[FRAD_LPC_LEN + 3];
  int aq[FRAD_LPC_LEN + 3];
  int flag;
  int pad_;
};

struct enc_params {
  const void* pcm; int pcm_dtype; int64_t hop;
  int B, N, C, profile, bits, le, fast_ok;
  // profile 0
  uint8_t* payload; size_t payload_stride; double* peak; uint16_t* half_scratch;
  // profiles 1 / 2
  double scale, loss; const band_table* bt; int32_t* q; int32_t* tq; int32_t* lpc;
  void* gscratch; double* tns_scratch;
};

struct dec_params {
  int B, N, C, profile, bits, le, fast_ok;
  const uint8_t* payload; size_t payload_stride; const uint16_t* half_scratch;
  double scale; const band_table* bt; const int32_t* q; const int32_t* tq; const int32_t* lpc;
  void* pcm; int pcm_dtype;
  void* gscratch; double* tns_scratch;
  const double* spec;          // profile 2: dequantised spectrum after TNS synthesis (tns_synth_kernel's layout)
  const double* dq64; int dq_len;   // |q|^(4/3) for |q| < dq_len (host libm), generic-size decode kernels
};

// ---- block reductions (all threads get the result) ---------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < nw; i++) t += red[i];
  return t;
}
__device__ __forceinline__ double block_max(double v, double* red) {
  v = warp_max(v);
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[w] = v;
  __syncthreads();
  double t = red[0];
  for (int i = 1; i < nw; i++) t = fmax(t, red[i]);
  return t;
}

// K sums in one round (two barriers instead of 2 K).  Every value goes through exactly the additions of block_sum.
template <int K>
__device__ __forceinline__ void block_sum_multi(double (&v)[K], double* red /* K * 32 doubles */) {
  const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int j = 0; j < K; j++) v[j] = warp_sum(v[j]);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int j = 0; j < K; j++) red[j * 32 + w] = v[j];
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < K; j++) {
    double t = 0.0;
    for (int i = 0; i < nw; i++) t += red[j * 32 + i];
    v[j] = t;
  }
}

// ---- profile 0 ---------------------------------------------------------------------------------
template <typename T>
__device__ void p0_epilogue(const enc_params& P, const T* X, int N, int b, int c, band_work* bw) {
  const int C = P.C, bits = P.bits;
  const bool le = P.le != 0;
  uint8_t* frame = P.payload + (size_t)b * P.payload_stride;
  double mx = 0.0;
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    const T v = X[k];
    mx = fmax(mx, fabs((double)v));
    const size_t vi = (size_t)k * C + c;
    if (bits == 12) P.half_scratch[(size_t)b * N * C + vi] = (uint16_t)(f2h_bits(v) >> 4);
    else p0_store_value<T>(frame + vi * (size_t)(bits >> 3), v, bits, le);
  }
  if (P.peak) {
    mx = block_max(mx, bw->red);
    if (threadIdx.x == 0) atomicMax((unsigned long long*)&P.peak[b], (unsigned long long)__double_as_longlong(mx));
  }
}

template <typename T>
__device__ void p0_prologue(const dec_params& P, T* X, int N, int b, int c) {
  const int C = P.C, bits = P.bits;
  const bool le = P.le != 0;
  const uint8_t* frame = P.payload + (size_t)b * P.payload_stride;
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    const size_t vi = (size_t)k * C + c;
    double v;
    if (bits == 12) {
      v = (double)__half2float(__ushort_as_half(P.half_scratch[(size_t)b * N * C + vi]));
      if (!isfinite(v)) v = 0.0;
    } else {
      v = p0_load_value(frame + vi * (size_t)(bits >> 3), bits, le);
    }
    X[k] = (T)v;
  }
}

// ---- profile 1 helpers -------------------------------------------------------------------------
// |y|^0.75 with <= 1 ulp error: sqrt(|y| * sqrt(|y|))
__device__ __forceinline__ double pow075(double a) { return sqrt(a * sqrt(a)); }

// divisor ramp value for bin k (p1tools.py:35-41): linspace(t[b], t[b+1], len, endpoint=False)
__device__ __forceinline__ double ramp_value(const band_work* bw, int k) {
  if (k >= bw->lo[FRAD_NBANDS - 1]) return 0.0;    // bins past the start of the last band stay 0
  int b = 0;
#pragma unroll 1
  for (int s = 16; s > 0; s >>= 1) { if (b + s <= FRAD_NBANDS - 2 && bw->lo[b + s] <= k) b += s; }
  // lo[b] <= k < lo[b+1]
  return __dadd_rn(__dmul_rn((double)(k - bw->lo[b]), bw->step[b]), bw->thr[b]);
}

__device__ __forceinline__ void ramp_setup(band_work* bw, const band_table* bt, int N) {
  const int t = threadIdx.x;
  if (t < FRAD_NBANDS - 1) {
    const int n = bw->lo[t + 1] - bw->lo[t];
    bw->step[t] = (n > 0) ? (bw->thr[t + 1] - bw->thr[t]) / (double)n : 0.0;
  }
}

// f(k, divisor) for every bin k of the sequence (a flat, unrollable loop over the CTA's threads): what ramp_value(bw, k)
// returns, with the band of the bin read from the host-built table instead of a binary search per bin (12-15 % of the generic
// kernels' instructions in the round-2 capture).  (Striding through band after band was measured first: 27 short loops per
// thread whose loads do not overlap -- slower than the search.)
template <typename F>
__device__ __forceinline__ void for_each_bin_ramp(const band_work* bw, const band_table* bt, int N, F f) {
  const int t = threadIdx.x, nt = blockDim.x;
#pragma unroll 4
  for (int k = t; k < N; k += nt) {
    const int bnd = __ldg(&bt->bidx[k]);
    const double div = (bnd == 255) ? 0.0 : __dadd_rn(__dmul_rn((double)(k - bw->lo[bnd]), bw->step[bnd]), bw->thr[bnd]);
    f(k, div);
  }
}

__device__ __forceinline__ int32_t sat_i32(double v) {
  if (!(v == v)) return INT32_MIN;
  if (v >= 2147483647.0) return INT32_MAX;
  if (v <= -2147483648.0) return INT32_MIN;
  return (int32_t)v;
}

// numpy pairwise summation order for short vectors (n < 128), used where the reference sums tiny arrays
__device__ __forceinline__ double np_sum_small(const double* a, int n) {
  if (n < 8) { double r = 0.0; for (int i = 0; i < n; i++) r = __dadd_rn(r, a[i]); return r; }
  double r[8];
  for (int j = 0; j < 8; j++) r[j] = a[j];
  int i = 8;
  for (; i < n - (n % 8); i += 8) for (int j = 0; j < 8; j++) r[j] = __dadd_rn(r[j], a[i + j]);
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                         __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (; i < n; i++) res = __dadd_rn(res, a[i]);
  return res;
}

template <int NN>
__device__ __forceinline__ double np_sum_small_t(const double (&a)[FRAD_LPC_LEN]) {     // np_sum_small(a, NN), unrolled
  if constexpr (NN < 8) {
    double r = 0.0;
#pragma unroll
    for (int i = 0; i < NN; i++) r = __dadd_rn(r, a[i]);
    return r;
  } else {
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = a[j];
    constexpr int full = NN - (NN % 8);
    if constexpr (full > 8) {                         // (never with 13 LPC coefficients; kept for other orders)
#pragma unroll
      for (int i = 8; i < full; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = __dadd_rn(r[j], a[i + j]);
      }
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
#pragma unroll
    for (int i = full; i < NN; i++) res = __dadd_rn(res, a[i]);
    return res;
  }
}

// one Levinson-Durbin step (p2tools.py:22-32) with every index a compile-time constant: the arrays stay in registers
template <int I>
__device__ __forceinline__ void levinson_step(double (&a)[FRAD_LPC_LEN], const double (&r)[FRAD_LPC_LEN], double& err, bool& go) {
  if (go) {
    double prod[FRAD_LPC_LEN], tmp[FRAD_LPC_LEN];
#pragma unroll
    for (int j = 0; j < FRAD_LPC_LEN; j++) { prod[j] = (j < I) ? __dmul_rn(a[j], r[(I - j) > 0 ? (I - j) : 0]) : 0.0; tmp[j] = a[j]; }
    double k = -np_sum_small_t<I>(prod) / err;
    if (fabs(k) >= 0.96) k = (k > 0.0) ? 0.96 : -0.96;
    a[I] = k;
#pragma unroll
    for (int j = 1; j < I; j++) a[j] = __dadd_rn(a[j], __dmul_rn(k, tmp[I - j]));
    err = __dmul_rn(err, __dsub_rn(1.0, __dmul_rn(k, k)));
    if (err <= 1e-12) go = false;
  }
}

// round(sign(y) |y|^0.75) (p1tools.py:43): f32 first, the f64 formula only next to a rounding tie.  f32 budget: operand
// rounding 6e-8, two correctly rounded square roots and a product < 2e-7 relative in total; guard 1e-6 v.
__device__ __forceinline__ int32_t quant075(double y) {
  const float af = (float)fabs(y);
  const float v = sqrtf(af * sqrtf(af));
  const float r = rintf(v);
  if (fabsf(fabsf(v - r) - 0.5f) > 1.0e-6f * v + 1.0e-7f && v < 4.0e6f) { const int m = (int)r; return (y < 0.0) ? -m : m; }
  return sat_i32(rint(copysign(pow075(fabs(y)), y)));
}

// ---- TNS (p2tools.py) --------------------------------------------------------------------------
// x: masked spectrum (double, N), work: N doubles.  On return bw->flag != 0 means the residual in
// `work` replaces x and bw->aq holds the quantised LPC; otherwise x stands and aq is all zero.
static __device__ void tns_analysis_block(const double* x, double* work, int N, band_work* bw) {
  const int t = threadIdx.x;
  if (t < FRAD_LPC_LEN) bw->aq[t] = 0;
  if (t == 0) bw->flag = 0;
  __syncthreads();
  if (N < 2 * (FRAD_LPC_LEN - 1)) return;
  // lpc_cond (p2tools.py:108-111): spectral flatness exp(mean log(|x| + 1e-10)) / (mean |x| + 1e-10) < 0.5.  The logarithms
  // only feed this yes/no decision, so they are taken in f32 first (relative error of the ratio < 1e-5); the f64 sum of
  // logs is only formed when the ratio lands within 1e-3 of the limit.  The other sums are f64 as before.
  double s4[4] = {0.0, 0.0, 0.0, 0.0};                 // sum log(|x| + 1e-10), sum |x|, sum x^2, sum x
  float slf = 0.0f;
  for (int n = t; n < N; n += blockDim.x) {
    const double v = x[n], m = fabs(v);
    slf += logf((float)(m + 1e-10)); s4[1] += m; s4[2] += v * v; s4[3] += v;
  }
  s4[0] = (double)slf;
  block_sum_multi<4>(s4, bw->redm);
  const double sa = s4[1], se = s4[2], sx = s4[3];
  const double arith = sa / N;
  double flat = exp(s4[0] / N) / (arith + 1e-10);
  if (!(fabs(flat - 0.5) > 1e-3)) {                    // too close to call (or NaN): exact sum, same for every thread
    double sl = 0.0;
    for (int n = t; n < N; n += blockDim.x) sl += log(fabs(x[n]) + 1e-10);
    sl = block_sum(sl, bw->red);
    flat = exp(sl / N) / (arith + 1e-10);
  }
  if (!(flat < 0.5)) return;
  if (se < 1e-10) return;
  // autocorrelation of the mean-removed, unit-norm signal (p2tools.py:8-15)
  const double mean = sx / N;
  double ss = 0.0;
  for (int n = t; n < N; n += blockDim.x) { const double s = x[n] - mean; work[n] = s; ss += s * s; }
  ss = block_sum(ss, bw->red);
  const double nrm = sqrt(ss);
  const double inv_applied = (nrm > 1e-6) ? 1.0 : 0.0;
  __syncthreads();
  if (inv_applied != 0.0) {
    // x / nrm with one division per thread: q0 = x r, q = q0 + (x - q0 nrm) r with r = RN(1 / nrm) is the correctly rounded
    // quotient (Markstein), except possibly when the significand of nrm is all ones -- that case divides
    const double rinv = 1.0 / nrm;
    const bool plain = ((unsigned long long)__double_as_longlong(nrm) & 0xFFFFFFFFFFFFFull) == 0xFFFFFFFFFFFFFull;
    for (int n = t; n < N; n += blockDim.x) {
      const double w = work[n];
      const double q0 = w * rinv;
      const double q1 = fma(fma(-q0, nrm, w), rinv, q0);
      work[n] = plain ? w / nrm : q1;
    }
  }
  __syncthreads();
  double acc[FRAD_LPC_LEN];
#pragma unroll
  for (int l = 0; l < FRAD_LPC_LEN; l++) acc[l] = 0.0;
  for (int n = t; n < N; n += blockDim.x) {
    const double s0 = work[n];
    if (n + FRAD_LPC_LEN <= N) {
#pragma unroll
      for (int l = 0; l < FRAD_LPC_LEN; l++) acc[l] += s0 * work[n + l];
    } else {
#pragma unroll
      for (int l = 0; l < FRAD_LPC_LEN; l++) if (n + l < N) acc[l] += s0 * work[n + l];
    }
  }
  // the 13 lag sums are only needed by the thread that runs the recursion: warp totals to shared memory, thread l adds up
  // lag l (the additions of block_sum, in its order) and applies the lag window (p2tools.py:14)
  {
    const int w = t >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int l = 0; l < FRAD_LPC_LEN; l++) acc[l] = warp_sum(acc[l]);
    __syncthreads();
    if ((t & 31) == 0) {
#pragma unroll
      for (int l = 0; l < FRAD_LPC_LEN; l++) bw->redm[l * 32 + w] = acc[l];
    }
    __syncthreads();
    if (t < FRAD_LPC_LEN) {
      double v = 0.0;
      for (int i = 0; i < nw; i++) v += bw->redm[t * 32 + i];
      const double wl = 0.01 * t;
      bw->r[t] = v * exp(-0.5 * (wl * wl));
    }
    __syncthreads();
  }
  if (t == 0) {
    // Levinson-Durbin (p2tools.py:17-34)
    double r[FRAD_LPC_LEN];
#pragma unroll
    for (int l = 0; l < FRAD_LPC_LEN; l++) r[l] = bw->r[l];
    double a[FRAD_LPC_LEN];
#pragma unroll
    for (int i = 0; i < FRAD_LPC_LEN; i++) a[i] = 0.0;
    a[0] = 1.0;
    double err = r[0];
    bool go = err > 1e-10;
    levinson_step<1>(a, r, err, go); levinson_step<2>(a, r, err, go); levinson_step<3>(a, r, err, go);
    levinson_step<4>(a, r, err, go); levinson_step<5>(a, r, err, go); levinson_step<6>(a, r, err, go);
    levinson_step<7>(a, r, err, go); levinson_step<8>(a, r, err, go); levinson_step<9>(a, r, err, go);
    levinson_step<10>(a, r, err, go); levinson_step<11>(a, r, err, go); levinson_step<12>(a, r, err, go);
    double prod[FRAD_LPC_LEN];
#pragma unroll
    for (int j = 0; j < FRAD_LPC_LEN - 1; j++) prod[j] = fabs(a[j + 1]);
    prod[FRAD_LPC_LEN - 1] = 0.0;
    const bool ok = !(np_sum_small_t<FRAD_LPC_LEN - 1>(prod) < 0.01);
    int any = 0;
    int aqv[FRAD_LPC_LEN];
    aqv[0] = 0;
#pragma unroll
    for (int j = 1; j < FRAD_LPC_LEN; j++) {
      double v = a[j] * 15.0;
      v = fmin(fmax(v, -15.0), 14.0);
      aqv[j] = (int)rint(v);
      any |= (aqv[j] != 0);
    }
    if (ok && any) {
#pragma unroll
      for (int j = 0; j < FRAD_LPC_LEN; j++) { bw->aq[j] = aqv[j]; bw->a[j] = (j == 0) ? 1.0 : (double)aqv[j] / 15.0; }
      bw->flag = 1;
    } else {
#pragma unroll
      for (int j = 0; j < FRAD_LPC_LEN; j++) bw->aq[j] = 0;
    }
  }
  __syncthreads();
  if (!bw->flag) return;
  // FIR residual, lfilter(lpc_deq, [1], x): b0 x[n] + (b1 x[n-1] + (... + b12 x[n-12]))
  double cf[FRAD_LPC_LEN];
#pragma unroll
  for (int i = 0; i < FRAD_LPC_LEN; i++) cf[i] = bw->a[i];
  double mxv = 0.0;
  double s2[2] = {0.0, 0.0};                           // count of non-finite outputs, sum of the residual
  for (int n = t; n < N; n += blockDim.x) {
    double y;
    if (n >= FRAD_LPC_LEN - 1) {                        // all taps present: the same sums without the per-tap tests
      double a12 = __dmul_rn(cf[FRAD_LPC_LEN - 1], x[n - (FRAD_LPC_LEN - 1)]);
#pragma unroll
      for (int i = FRAD_LPC_LEN - 2; i >= 1; i--) a12 = __dadd_rn(a12, __dmul_rn(cf[i], x[n - i]));
      y = __dadd_rn(a12, __dmul_rn(cf[0], x[n]));
    } else {
      double a12 = 0.0;
      bool first = true;
#pragma unroll
      for (int i = FRAD_LPC_LEN - 1; i >= 1; i--) {
        if (n - i >= 0) {
          const double p = __dmul_rn(cf[i], x[n - i]);
          a12 = first ? p : __dadd_rn(a12, p);
          first = false;
        }
      }
      y = first ? __dmul_rn(cf[0], x[n]) : __dadd_rn(a12, __dmul_rn(cf[0], x[n]));
    }
    work[n] = y;
    if (!isfinite(y)) s2[0] = 1.0;
    mxv = fmax(mxv, fabs(y));
    s2[1] += y;
  }
  mxv = block_max(mxv, bw->red);
  block_sum_multi<2>(s2, bw->redm);
  const double sr = s2[1];
  int accept = !(s2[0] > 0.0 || mxv > 1e6);
  // prediction gain (p2tools.py:57-70)
  const double mo = sx / N, mr = sr / N;
  double e2[2] = {0.0, 0.0};
  for (int n = t; n < N; n += blockDim.x) {
    const double d0 = x[n] - mo, d1 = work[n] - mr;
    e2[0] += d0 * d0; e2[1] += d1 * d1;
  }
  block_sum_multi<2>(e2, bw->redm);
  const double eo = e2[0], er = e2[1];
  double gain = 0.0;
  if (!(eo < 1e-10 || er < 1e-10 || er >= eo)) gain = 20.0 * log10(eo / er);
  if (gain < 0.030102999566398118) accept = 0;       // log10(2) / 10
  __syncthreads();
  if (t == 0 && !accept) { bw->flag = 0; for (int j = 0; j < FRAD_LPC_LEN; j++) bw->aq[j] = 0; }
  __syncthreads();
}

// ---- profile 1 / 2 encode epilogue --------------------------------------------------------------
template <typename T>
__device__ void p1_epilogue(const enc_params& P, const T* X, int N, int b, int c, band_work* bw, int cta) {
  const band_table* bt = P.bt;
  const int t = threadIdx.x, C = P.C;
  if (t <= FRAD_NBANDS) { const int e = bt->edge[t]; bw->lo[t] = (e < N) ? e : N; }
  __syncthreads();
  // band energies of the scaled spectrum (p1tools.py:22,32)
  const int w = t >> 5, lane = t & 31, nw = blockDim.x >> 5;
  for (int band = w; band < FRAD_NBANDS; band += nw) {
    double s = 0.0;
    for (int k = bw->lo[band] + lane; k < bw->lo[band + 1]; k += 32) { const double v = (double)X[k] * P.scale; s += v * v; }
    s = warp_sum(s);
    if (lane == 0) bw->sum[band] = s;
  }
  __syncthreads();
  if (t < FRAD_NBANDS) {
    const int n = bw->lo[t + 1] - bw->lo[t];
    double thr = 0.0;
    if (n > 0) {
      const double level = pow(sqrt(bw->sum[t] / (double)n), 0.8);
      thr = fmax(level, bt->ath[t]) * P.loss;
    }
    bw->thr[t] = thr;
  }
  __syncthreads();
  if (t == 0) {   // the reference stops at the first empty band, leaving zeros (p1tools.py:23)
    bool dead = false;
    for (int i = 0; i < FRAD_NBANDS; i++) {
      if (bw->lo[i + 1] - bw->lo[i] <= 0) dead = true;
      if (dead) bw->thr[i] = 0.0;
    }
    bw->thr[FRAD_NBANDS] = 0.0;
  }
  __syncthreads();
  if (t < FRAD_NBANDS) {
    // profile1.py:38-40: round(dequant(ln(max(thr,1)) / ln(e/2)))
    const double v = log(fmax(bw->thr[t], 1.0)) / bt->log_base;
    const double d = (v > 0.0) ? pow(v, 1.0 / 0.75) : 0.0;
    P.tq[((size_t)b * FRAD_NBANDS + t) * C + c] = sat_i32(rint(d));
  }
  ramp_setup(bw, bt, N);
  __syncthreads();
  int32_t* qrow = P.q + (size_t)b * N * C;
  if (P.profile == 1) {
    for_each_bin_ramp(bw, bt, N, [&](int k, double div) {
      const double y = (div == 0.0) ? 0.0 : ((double)X[k] / div) * P.scale;
      qrow[(size_t)k * C + c] = quant075(y);               // f32 first, the f64 formula next to a rounding tie
    });
    return;
  }
  // profile 2: TNS over the masked spectrum (profile2.py:32)
  double* masked = P.tns_scratch + (size_t)cta * 2 * N;
  double* work = masked + N;
  for_each_bin_ramp(bw, bt, N, [&](int k, double div) { masked[k] = (div == 0.0) ? 0.0 : (double)X[k] / div; });
  __syncthreads();
  tns_analysis_block(masked, work, N, bw);
  const double* src = bw->flag ? work : masked;
  if (t < FRAD_LPC_LEN) P.lpc[((size_t)b * FRAD_LPC_LEN + t) * C + c] = bw->aq[t];
  for (int k = t; k < N; k += blockDim.x) {
    qrow[(size_t)k * C + c] = quant075(src[k] * P.scale);
  }
}

// ---- profile 1 / 2 decode prologue --------------------------------------------------------------
template <typename T>
__device__ void p1_prologue(const dec_params& P, T* X, int N, int b, int c, band_work* bw, int cta) {
  const band_table* bt = P.bt;
  const int t = threadIdx.x, C = P.C;
  if (t <= FRAD_NBANDS) { const int e = bt->edge[t]; bw->lo[t] = (e < N) ? e : N; }
  if (t < FRAD_NBANDS) {
    // profile1.py:66: (e/2) ** (sign(tq) |tq|^0.75)
    bw->thr[t] = decode_threshold(bt, P.tq[((size_t)b * FRAD_NBANDS + t) * C + c]);
  }
  if (t == 0) bw->thr[FRAD_NBANDS] = 0.0;
  __syncthreads();
  ramp_setup(bw, bt, N);
  __syncthreads();
  const int32_t* qrow = P.q + (size_t)b * N * C;
  if (P.profile == 1) {
    // profile1.py:65: sign(q) |q|^(4/3) / 2^(bits-1); the magnitudes come from the host-computed table (libm's pow, like the
    // reference) -- the in-line pow() was 40 % of this kernel's instructions
    const double inv_scale = 1.0 / P.scale;                   // a power of two: the product is the quotient, bit for bit
    for_each_bin_ramp(bw, bt, N, [&](int k, double div) {
      const int qi = qrow[(size_t)k * C + c];
      const uint32_t qa = (qi < 0) ? 0u - (uint32_t)qi : (uint32_t)qi;
      const double mag = (P.dq64 && qa < (uint32_t)P.dq_len) ? __ldg(P.dq64 + qa) : pow((double)qa, 1.0 / 0.75);
      const double coef = ((qi < 0) ? -mag : mag) * inv_scale;
      X[k] = (T)(coef * div);
    });
    return;
  }
  // profile 2: the dequantiser and the all-pole TNS synthesis (profile2.py:80-85, p2tools.py:97-106) have already run for
  // every sequence of the batch (tns_synth_kernel, one thread per sequence); only the ramp is left
  const int sq = b * C + c;                                           // layout of tns_synth_kernel: bin k of four sequences per sector
  const double* src = P.spec + (size_t)(sq >> 2) * N * 4 + (sq & 3);
  for_each_bin_ramp(bw, bt, N, [&](int k, double div) { X[k] = (T)(src[(size_t)k * 4] * div); });
}
