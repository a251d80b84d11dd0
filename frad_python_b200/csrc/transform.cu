// Generic-size transform path: one CTA per (frame, channel) sequence, persistent grid.
//
// Covers every even frame length whose half factors into {2,3,5,7} (all 32 compact sizes of
// fourier/profiles.py:14-23 and most lossless sizes) with a shared-memory Stockham FFT, and every
// other length with a direct O(N^2) cosine sum.  The profile pro/epilogues (profile-0 packing,
// profile-1/2 masking + quantisation, TNS) are fused into the same kernel.  The N = 2048 fast path
// lives in dct2048.cu and reuses the device functions declared here.
//
// Reference lines restated (relative to src/libfrad/): fourier/profile0.py:14-69,
// fourier/profile1.py:15-77, fourier/profile2.py:15-91, fourier/tools/p1tools.py:15-44,
// fourier/tools/p2tools.py:8-111.
#include <cmath>
#include "plan.cuh"
#include "devmath.cuh"
#include "profiles.cuh"

// ------------------------------------------------------------------------------------------------
// plans
// ------------------------------------------------------------------------------------------------
static bool factor_half(int M, int* radix, int* npass) {
  int n = 0, m = M;
  while (m % 8 == 0) { radix[n++] = 8; m /= 8; }          // round 2: radix 8 first (half as many passes and barriers as radix 4)
  while (m % 4 == 0) { radix[n++] = 4; m /= 4; }
  while (m % 2 == 0) { radix[n++] = 2; m /= 2; }
  while (m % 3 == 0) { radix[n++] = 3; m /= 3; }
  while (m % 5 == 0) { radix[n++] = 5; m /= 5; }
  while (m % 7 == 0) { radix[n++] = 7; m /= 7; }
  *npass = n;
  return m == 1 && n <= FRAD_MAX_RADIX_PASSES;
}

template <typename T>
static void* upload_cpx(frad_ctx* ctx, const std::vector<long double>& re, const std::vector<long double>& im) {
  size_t n = re.size();
  std::vector<T> h(2 * n);
  for (size_t i = 0; i < n; i++) { h[2 * i] = (T)re[i]; h[2 * i + 1] = (T)im[i]; }
  void* d = nullptr;
  if (cudaMalloc(&d, 2 * n * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  // tables are uploaded synchronously once per (N, precision); not on the timed path
  if (frad_upload(d, h.data(), 2 * n * sizeof(T)) != cudaSuccess) { cudaFree(d); return nullptr; }
  return d;
}

// exact-octant evaluation of exp(i * pi * num / den)
static void unit_root(long long num, long long den, long double* c, long double* s) {
  const long double PI = 3.14159265358979323846264338327950288L;
  long long period = 2 * den;
  num %= period; if (num < 0) num += period;
  // reduce to [0, den/2] using symmetries so that cosl/sinl see small arguments
  int quad = (int)((2 * num) / den);           // 0..3 quarter turns of pi/2
  long long r = 2 * num - (long long)quad * den;  // angle = quad*pi/2 + r*pi/(2 den), r in [0, den)
  long double a = PI * (long double)r / (2.0L * (long double)den);
  long double cc = cosl(a), ss = sinl(a);
  switch (quad & 3) {
    case 0: *c = cc; *s = ss; break;
    case 1: *c = -ss; *s = cc; break;
    case 2: *c = -cc; *s = -ss; break;
    default: *c = ss; *s = -cc; break;
  }
}

void frad_plan_free(frad_plan* p) {
  if (!p) return;
  cudaFree(p->tw); cudaFree(p->cf); cudaFree(p->df); cudaFree(p->ci); cudaFree(p->ei); cudaFree(p->ct); cudaFree(p->tw32); cudaFree(p->cfh); cudaFree(p->dfh);
  delete p;
}

template <typename T>
static bool build_tables(frad_ctx* ctx, frad_plan* p) {
  const int N = p->N, M = p->M;
  std::vector<long double> re(M), im(M);
  for (int t = 0; t < M; t++) { unit_root(-2LL * t, M, &re[t], &im[t]); }
  p->tw = upload_cpx<T>(ctx, re, im);
  re.resize(M + 1); im.resize(M + 1);
  std::vector<long double> re2(M + 1), im2(M + 1);
  for (int k = 0; k <= M; k++) {
    unit_root(-(long long)k, 2LL * N, &re[k], &im[k]);          // exp(-i pi k / 2N)
    unit_root(-5LL * k, 2LL * N, &re2[k], &im2[k]);             // * exp(-2 pi i k / N) = exp(-i pi 5k / 2N)
    re[k] /= N; im[k] /= N; re2[k] /= N; im2[k] /= N;
  }
  p->cf = upload_cpx<T>(ctx, re, im);
  p->df = upload_cpx<T>(ctx, re2, im2);
  const bool warp_fft = (N == 2048 || frad_warp_r1(N) != 0);   // frame lengths of the warp-FFT kernels: N = 64 R1
  if (warp_fft) {
    for (int k = 0; k <= M; k++) { re[k] /= 2; im[k] /= 2; re2[k] /= 2; im2[k] /= 2; }
    p->cfh = upload_cpx<T>(ctx, re, im);
    p->dfh = upload_cpx<T>(ctx, re2, im2);
    if (!p->cfh || !p->dfh) return false;
  }
  for (int k = 0; k <= M; k++) {
    unit_root((long long)k, 2LL * N, &re[k], &im[k]);           // exp(+i pi k / 2N)
    unit_root(4LL * k, 2LL * N, &re2[k], &im2[k]);              // exp(+2 pi i k / N)
  }
  p->ci = upload_cpx<T>(ctx, re, im);
  p->ei = upload_cpx<T>(ctx, re2, im2);
  if (warp_fft) {
    // cross twiddles of the warp FFT, [l][L] = exp(-2 pi i l (L mod R1) / M), R1 = M / 32 (fft32.cuh: fft_multi_warp)
    const int R1 = M / 32;
    re.resize(1024); im.resize(1024);
    for (int i = 0; i < 32; i++) for (int l = 0; l < 32; l++) unit_root(-2LL * i * (l % R1), M, &re[i * 32 + l], &im[i * 32 + l]);
    p->tw32 = upload_cpx<T>(ctx, re, im);
    if (!p->tw32) return false;
  }
  return p->tw && p->cf && p->df && p->ci && p->ei;
}

frad_plan* frad_get_plan(frad_ctx* ctx, int N, int prec) {
  uint64_t key = (uint64_t)(uint32_t)N | ((uint64_t)prec << 40);
  auto it = ctx->plans.find(key);
  if (it != ctx->plans.end()) return it->second;
  frad_plan* p = new frad_plan();
  p->N = N; p->M = N / 2; p->prec = prec;
  bool ok = true;
  if ((N & 1) || N < 4 || !factor_half(p->M, p->radix, &p->npass)) {
    p->direct = 1;
    std::vector<double> h((size_t)4 * N);
    for (long long q = 0; q < 4LL * N; q++) { long double c, s; unit_root(q, 2LL * N, &c, &s); h[q] = (double)c; }
    if (cudaMalloc(&p->ct, h.size() * sizeof(double)) != cudaSuccess ||
        frad_upload(p->ct, h.data(), h.size() * sizeof(double)) != cudaSuccess) ok = false;
  } else {
    ok = (prec == FRAD_PREC_F32) ? build_tables<float>(ctx, p) : build_tables<double>(ctx, p);
  }
  if (!ok) {
    cudaGetLastError();
    frad_plan_free(p);
    frad_fail(ctx, FRAD_E_NOMEM, "plan table allocation failed");
    return nullptr;
  }
  ctx->plans[key] = p;
  return p;
}

// ------------------------------------------------------------------------------------------------
// FFT passes in shared memory (Stockham autosort, out of place)
// ------------------------------------------------------------------------------------------------
template <typename T, int R, bool INV>
__device__ __forceinline__ void fft_pass(const typename vec2<T>::type* __restrict__ in, typename vec2<T>::type* __restrict__ out,
                                         int M, int Ns, const typename vec2<T>::type* __restrict__ tw) {
  typedef typename vec2<T>::type V;
  const int L = M / R;
  const int tstep = M / (Ns * R);
  const bool ns_pow2 = (Ns & (Ns - 1)) == 0;           // always, until an odd radix has run (odd radices come last)
  for (int j = threadIdx.x; j < L; j += blockDim.x) {
    const int k = ns_pow2 ? (j & (Ns - 1)) : (j % Ns);
    V v[R];
#pragma unroll
    for (int i = 0; i < R; i++) v[i] = in[j + i * L];
    if (Ns > 1) {
#pragma unroll
      for (int i = 1; i < R; i++) {
        V w = tw[(size_t)i * k * tstep];
        if (INV) w.y = -w.y;
        v[i] = cmul(v[i], w);
      }
    }
    small_dft<T, R, INV>::run(v);
    const int base = (j - k) * R + k;
#pragma unroll
    for (int i = 0; i < R; i++) out[base + i * Ns] = v[i];
  }
}

// runs all passes; returns the buffer that holds the result
template <typename T, bool INV>
__device__ typename vec2<T>::type* fft_run(typename vec2<T>::type* a, typename vec2<T>::type* b, const plan_dev& pl) {
  typedef typename vec2<T>::type V;
  const V* tw = (const V*)pl.tw;
  int Ns = 1;
  for (int p = 0; p < pl.npass; p++) {
    const int R = pl.radix[p];
    switch (R) {
      case 8: fft_pass<T, 8, INV>(a, b, pl.M, Ns, tw); break;
      case 4: fft_pass<T, 4, INV>(a, b, pl.M, Ns, tw); break;
      case 2: fft_pass<T, 2, INV>(a, b, pl.M, Ns, tw); break;
      case 3: fft_pass<T, 3, INV>(a, b, pl.M, Ns, tw); break;
      case 5: fft_pass<T, 5, INV>(a, b, pl.M, Ns, tw); break;
      default: fft_pass<T, 7, INV>(a, b, pl.M, Ns, tw); break;
    }
    Ns *= R;
    __syncthreads();
    V* t = a; a = b; b = t;
  }
  return a;
}

// Makhoul index: position j of the reordered sequence v reads sample n
__device__ __forceinline__ int makhoul_src(int j, int N) { return (j < (N + 1) / 2) ? 2 * j : 2 * (N - 1 - j) + 1; }

// ------------------------------------------------------------------------------------------------
// forward: PCM sequence -> X[0..N) in shared/global buffer (type T)
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ T* forward_sequence(const enc_params& P, const plan_dev& pl, int b, int c,
                               typename vec2<T>::type* bufA, typename vec2<T>::type* bufB) {
  typedef typename vec2<T>::type V;
  const int N = pl.N, M = pl.M, C = P.C;
  const int64_t base = (int64_t)b * P.hop;
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    const int n0 = makhoul_src(2 * m, N), n1 = makhoul_src(2 * m + 1, N);
    T x0 = load_pcm<T>(P.pcm, P.pcm_dtype, (base + n0) * C + c);
    T x1 = load_pcm<T>(P.pcm, P.pcm_dtype, (base + n1) * C + c);
    bufA[m] = mk2<T>(x0, x1);
  }
  __syncthreads();
  V* Z = fft_run<T, false>(bufA, bufB, pl);
  V* O = (Z == bufA) ? bufB : bufA;
  T* X = (T*)O;
  const V* cf = (const V*)pl.cf;
  const V* df = (const V*)pl.df;
  for (int k = threadIdx.x; k <= M / 2; k += blockDim.x) {
    const int km = M - k;
    V zk = Z[k], zm = cconj(Z[km % M]);
    V A = mk2<T>(T(0.5) * (zk.x + zm.x), T(0.5) * (zk.y + zm.y));
    V D = csub(zk, zm);                       // B = D / (2i) = (D.y - i D.x) / 2
    V Bv = mk2<T>(T(0.5) * D.y, T(-0.5) * D.x);
    V y = cadd(cmul(cf[k], A), cmul(df[k], Bv));
    V y2 = cadd(cmul(cf[km], cconj(A)), cmul(df[km], cconj(Bv)));
    X[k] = y.x;
    if (k >= 1) X[N - k] = -y.y;
    X[km] = y2.x;                             // km == M for k == 0
    if (k >= 1) X[N - km] = -y2.y;
  }
  __syncthreads();
  return X;
}

// direct O(N^2) forward for awkward lengths; x staged in xs (double), result to X (double)
__device__ void forward_direct(const enc_params& P, const plan_dev& pl, int b, int c, double* xs, double* X) {
  const int N = pl.N, C = P.C;
  const int64_t base = (int64_t)b * P.hop;
  for (int n = threadIdx.x; n < N; n += blockDim.x) xs[n] = load_pcm<double>(P.pcm, P.pcm_dtype, (base + n) * C + c);
  __syncthreads();
  const long long period = 4LL * N;
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    double acc = 0.0;
    long long ph = k % period;               // k*(2n+1) mod 4N, advanced by 2k per sample
    const long long step = (2LL * k) % period;
    for (int n = 0; n < N; n++) {
      acc += xs[n] * pl.ct[ph];
      ph += step; if (ph >= period) ph -= period;
    }
    X[k] = acc / N;
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// inverse: X[0..N) (type T, in bufA viewed as T*) -> PCM
// ------------------------------------------------------------------------------------------------
template <typename T>
__device__ void inverse_sequence(const dec_params& P, const plan_dev& pl, int b, int c,
                                 typename vec2<T>::type* bufA, typename vec2<T>::type* bufB) {
  typedef typename vec2<T>::type V;
  const int N = pl.N, M = pl.M, C = P.C;
  const T* X = (const T*)bufA;
  const V* ci = (const V*)pl.ci;
  const V* ei = (const V*)pl.ei;
  for (int k = threadIdx.x; k <= M / 2; k += blockDim.x) {
    const int km = M - k;
    // V_k = (X[k] - i X[N-k]) * ci[k], with X[N] := 0
    V vk = cmul(mk2<T>(X[k], (k == 0) ? T(0) : -X[N - k]), ci[k]);
    V vm = cmul(mk2<T>(X[km], (k == 0) ? -T(0) * X[0] : -X[N - km]), ci[km]);
    if (k == 0) vm = cmul(mk2<T>(X[M], -X[M]), ci[M]);      // N - M == M
    V vmc = cconj(vm), vkc = cconj(vk);
    // Z[k] = (V_k + conj V_{M-k}) + i e_k (V_k - conj V_{M-k})
    V s = cadd(vk, vmc), d = cmul(ei[k], csub(vk, vmc));
    bufB[k] = mk2<T>(s.x - d.y, s.y + d.x);
    if (k >= 1 && km < M) {
      V s2 = cadd(vm, vkc), d2 = cmul(ei[km], csub(vm, vkc));
      bufB[km] = mk2<T>(s2.x - d2.y, s2.y + d2.x);
    }
  }
  __syncthreads();
  V* z = fft_run<T, true>(bufB, bufA, pl);
  const int64_t base = (int64_t)b * N;
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    V v = z[m];
    const int n0 = makhoul_src(2 * m, N), n1 = makhoul_src(2 * m + 1, N);
    store_pcm<T>(P.pcm, P.pcm_dtype, (base + n0) * C + c, v.x);
    store_pcm<T>(P.pcm, P.pcm_dtype, (base + n1) * C + c, v.y);
  }
  __syncthreads();
}

__device__ void inverse_direct(const dec_params& P, const plan_dev& pl, int b, int c, const double* X) {
  const int N = pl.N, C = P.C;
  const long long period = 4LL * N;
  const int64_t base = (int64_t)b * N;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    double acc = 0.0;
    const long long step = (2LL * n + 1) % period;
    long long ph = step % period;             // k*(2n+1) mod 4N for k = 1
    for (int k = 1; k < N; k++) {
      acc += X[k] * pl.ct[ph];
      ph += step; if (ph >= period) ph -= period;
    }
    store_pcm<double>(P.pcm, P.pcm_dtype, (base + n) * C + c, X[0] + 2.0 * acc);
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
// MAXT: the launch bound.  Frames of up to 4096 samples run 256 threads and may use the whole register file (radix-8 butterflies
// in f64 spill under the 64-register cap of a 1024-thread bound).
template <typename T, int MAXT>
__global__ void __launch_bounds__(MAXT) seq_encode_kernel(enc_params P, plan_dev pl, int b_in_smem) {
  typedef typename vec2<T>::type V;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  band_work* bw = (band_work*)smem_raw;
  V* bufA = (V*)(smem_raw + sizeof(band_work));
  V* bufB = b_in_smem ? bufA + pl.M : ((V*)P.gscratch) + (size_t)blockIdx.x * pl.M;
  const int nseq = P.B * P.C;
  for (int s = blockIdx.x; s < nseq; s += gridDim.x) {
    const int b = s / P.C, c = s % P.C;
    {   // pull the frame of this CTA's next sequence towards L2 (its samples are fetched one scalar load at a time)
      const int s2 = s + gridDim.x;
      if (s2 < nseq && (s2 / P.C) != b) {
        const size_t es = (size_t)pcm_size(P.pcm_dtype);
        const char* nxt = (const char*)P.pcm + (size_t)(s2 / P.C) * P.hop * P.C * es;
        const size_t bytes = (size_t)pl.N * P.C * es;
        for (size_t o = (size_t)threadIdx.x * 128; o < bytes; o += (size_t)blockDim.x * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + o));
      }
    }
    T* X = forward_sequence<T>(P, pl, b, c, bufA, bufB);
    if (P.profile == 0) p0_epilogue<T>(P, X, pl.N, b, c, bw);
    else p1_epilogue<T>(P, X, pl.N, b, c, bw, blockIdx.x);
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) seq_encode_direct_kernel(enc_params P, plan_dev pl) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  band_work* bw = (band_work*)smem_raw;
  double* xs = (double*)P.gscratch + (size_t)blockIdx.x * 2 * pl.N;
  double* X = xs + pl.N;
  const int nseq = P.B * P.C;
  for (int s = blockIdx.x; s < nseq; s += gridDim.x) {
    const int b = s / P.C, c = s % P.C;
    forward_direct(P, pl, b, c, xs, X);
    if (P.profile == 0) p0_epilogue<double>(P, X, pl.N, b, c, bw);
    else p1_epilogue<double>(P, X, pl.N, b, c, bw, blockIdx.x);
    __syncthreads();
  }
}

template <typename T, int MAXT>
__global__ void __launch_bounds__(MAXT) seq_decode_kernel(dec_params P, plan_dev pl, int b_in_smem) {
  typedef typename vec2<T>::type V;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  band_work* bw = (band_work*)smem_raw;
  V* bufA = (V*)(smem_raw + sizeof(band_work));
  V* bufB = b_in_smem ? bufA + pl.M : ((V*)P.gscratch) + (size_t)blockIdx.x * pl.M;
  const int nseq = P.B * P.C;
  for (int s = blockIdx.x; s < nseq; s += gridDim.x) {
    const int b = s / P.C, c = s % P.C;
    {   // pull the coded frame of this CTA's next sequence towards L2
      const int s2 = s + gridDim.x;
      if (s2 < nseq && (s2 / P.C) != b) {
        const char* nxt = (P.profile == 0) ? (const char*)(P.payload + (size_t)(s2 / P.C) * P.payload_stride)
                                           : (const char*)(P.q + (size_t)(s2 / P.C) * pl.N * P.C);
        const size_t bytes = (P.profile == 0) ? (size_t)pl.N * P.C * P.bits / 8 : (size_t)pl.N * P.C * 4;
        for (size_t o = (size_t)threadIdx.x * 128; o < bytes; o += (size_t)blockDim.x * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + o));
      }
    }
    T* X = (T*)bufA;
    if (P.profile == 0) p0_prologue<T>(P, X, pl.N, b, c);
    else p1_prologue<T>(P, X, pl.N, b, c, bw, blockIdx.x);
    __syncthreads();
    inverse_sequence<T>(P, pl, b, c, bufA, bufB);
  }
}

__global__ void __launch_bounds__(256) seq_decode_direct_kernel(dec_params P, plan_dev pl) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  band_work* bw = (band_work*)smem_raw;
  double* X = (double*)P.gscratch + (size_t)blockIdx.x * pl.N;
  const int nseq = P.B * P.C;
  for (int s = blockIdx.x; s < nseq; s += gridDim.x) {
    const int b = s / P.C, c = s % P.C;
    if (P.profile == 0) p0_prologue<double>(P, X, pl.N, b, c);
    else p1_prologue<double>(P, X, pl.N, b, c, bw, blockIdx.x);
    __syncthreads();
    inverse_direct(P, pl, b, c, X);
  }
}

// Profile-2 encode, second step after the N = 2048 transform kernel (which leaves the spectrum as float bits in the q array,
// [B][N][C], and the sequence's thresholds in [B][C][28] doubles).  Round 2: three kernels instead of one.
//
//   tns_stats_kernel   one CTA of 128 threads per (frame, channel) sequence: masked spectrum X / div in f64
//                      (profile2.py:28-31), the sums of lpc_cond (p2tools.py:108-111) and, where the condition holds, the 13
//                      windowed autocorrelation lags of the mean-removed unit-norm signal (p2tools.py:8-15)
//   tns_lpc_kernel     one THREAD per sequence: Levinson-Durbin + 4-bit LPC quantiser (p2tools.py:17-34, 82-88).  In the
//                      one-kernel version this serial recursion ran on thread 0 while 127 threads waited at a barrier (12 % of
//                      the kernel's stall samples); here 30 000 of them run side by side
//   tns_apply_kernel   one CTA per sequence again: masked spectrum recomputed from the same operands (bit-identical), FIR
//                      residual lfilter(lpc_deq, [1], x), the finite / 1e6 check and the prediction gain (p2tools.py:57-70, 89-95),
//                      then the quantiser over the residual -- or the masked spectrum where TNS is rejected (profile2.py:34-46)
//
// Thread t owns the 16 consecutive bins [16 t, 16 t + 16) in registers; the only values that cross threads are the 12-bin halos
// of the two sliding windows (autocorrelation looks ahead, the FIR looks back), fetched once from shared memory -- the first
// version re-read every operand of both passes from shared memory (13 loads per bin and pass, 415 instructions per bin).
// Shared-memory rows are padded by one slot per 16 (`tpad`), so per-thread runs of 4- and 8-byte values are conflict-free;
// global I/O stays coalesced through a staging row.  Re-reading the spectrum costs 245 MB of HBM traffic (0.04 ms); keeping
// the two passes in separate kernels halves the live registers of each.  The block sums group differently from the generic
// path's (last bits of the f64 reductions), the LPC integers and the per-output operation order of the FIR (lfilter's) do not.
#define FRAD_TNS_THREADS 128
#define FRAD_TNS_BPT 16
#ifndef FRAD_TNS_MINB_STATS
#define FRAD_TNS_MINB_STATS 4
#endif
#ifndef FRAD_TNS_MINB_APPLY
#define FRAD_TNS_MINB_APPLY 4
#endif
#define FRAD_TNS_ROW (2048 + 2048 / 16)
__device__ __forceinline__ int tpad(int k) { return k + (k >> 4); }

// block reductions of the 4-warp TNS CTAs: block_sum / block_max / block_sum_multi of profiles.cuh (same additions in the same
// order) with the warp count a compile-time constant
// Sums K (a power of two <= 16) values over the warp with K - 1 + (5 - log2 K) exchanges instead of 5 K: every step halves the
// values a lane still carries (the lane keeps the half its bit of the lane number selects and adds the partner's copy of it).
// On return v[0] of lane l is the warp total of value wrs_index<K>(l); every total sits in 32 / K lanes.
template <int K> __device__ __forceinline__ int wrs_index(int lane) {
  int idx = 0;
#pragma unroll
  for (int n = K, o = 16; n > 1; n >>= 1, o >>= 1) idx |= (lane & o) ? (n >> 1) : 0;
  return idx;
}
template <int K>
__device__ __forceinline__ void warp_reduce_scatter(double (&v)[K], int lane) {
  int o = 16;
#pragma unroll
  for (int n = K; n > 1; n >>= 1, o >>= 1) {
    const int half = n >> 1;
    const bool upper = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < half; i++) {
      const double send = upper ? v[i] : v[i + half];
      const double keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
#pragma unroll
  for (; o > 0; o >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], o);
}
template <int K>
__device__ __forceinline__ void tns_sum_multi(double (&v)[K], double* red /* K * 32 doubles */) {
  constexpr int NW = FRAD_TNS_THREADS / 32;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  warp_reduce_scatter<K>(v, lane);
  __syncthreads();
  if ((lane & (32 / K - 1)) == 0) red[wrs_index<K>(lane) * 32 + w] = v[0];
  __syncthreads();
#pragma unroll
  for (int j = 0; j < K; j++) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < NW; i++) t += red[j * 32 + i];
    v[j] = t;
  }
}
__device__ __forceinline__ double tns_sum(double v, double* red) { double a[1] = {v}; tns_sum_multi<1>(a, red); return a[0]; }
__device__ __forceinline__ double tns_max(double v, double* red) {
  constexpr int NW = FRAD_TNS_THREADS / 32;
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = red[0];
#pragma unroll
  for (int i = 1; i < NW; i++) t = fmax(t, red[i]);
  return t;
}

// quant075 (profiles.cuh) with approximate square roots (<= 1 ulp each: 3e-7 relative in total, inside the 1e-6 guard) and the
// f64 formula next to a rounding tie out of line -- the inlined slow path made the unrolled quantiser loop miss the instruction cache
__device__ __noinline__ double tns_div(double a, double b) { return a / b; }
__device__ __noinline__ int32_t quant075_slow(double y) { return sat_i32(rint(copysign(pow075(fabs(y)), y))); }
__device__ __forceinline__ int32_t quant075_fast(double y) {
  const float af = (float)fabs(y);
  float s1, v;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s1) : "f"(af));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(v) : "f"(af * s1));
  const float r = rintf(v);
  if (fabsf(fabsf(v - r) - 0.5f) > 1.0e-6f * v + 1.0e-7f && v < 4.0e6f) { const int m = (int)r; return (y < 0.0) ? -m : m; }
  return quant075_slow(y);
}

struct tns_ctl {                   // per-sequence hand-over between the three kernels
  double r[FRAD_LPC_LEN];          // stats -> lpc: windowed lags
  double sx;                       // stats -> apply: sum of the masked spectrum
  int aq[FRAD_LPC_LEN];            // lpc -> apply: quantised coefficients (aq[0] unused)
  int cond;                        // stats: lpc_cond holds and the lags are valid
  int flag;                        // lpc: coefficients accepted
  int pad_;
};

// band of every bin (same for all sequences; 255 = past the ramps) + clipped band starts; one call per CTA
__device__ __forceinline__ void tns_cta_setup(const enc_params& P, band_work* bw, uint8_t* bidx) {
  const int t = threadIdx.x, N = P.N;
  if (t <= FRAD_NBANDS) { const int e = P.bt->edge[t]; bw->lo[t] = (e < N) ? e : N; }
  __syncthreads();
  for (int k = t; k < N; k += blockDim.x) {
    int b = 0;
#pragma unroll 1
    for (int sh = 16; sh > 0; sh >>= 1) { if (b + sh <= FRAD_NBANDS - 2 && bw->lo[b + sh] <= k) b += sh; }
    bidx[k] = (k >= bw->lo[FRAD_NBANDS - 1]) ? 255 : (uint8_t)b;     // exactly ramp_value's search
  }
}

// pulls the sequence this CTA takes after `s` towards L2
__device__ __forceinline__ void tns_prefetch_next(const enc_params& P, int s, int nseq) {
  const int s2 = s + gridDim.x;
  if (s2 < nseq && (s2 / P.C) != (s / P.C)) {
    const char* nxt = (const char*)(P.q + (size_t)(s2 / P.C) * P.N * P.C);
    const size_t bytes = (size_t)P.N * P.C * 4;
    for (size_t o = (size_t)threadIdx.x * 128; o < bytes; o += (size_t)blockDim.x * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + o));
  }
}

// masked spectrum of the own block (profile2.py:28-31); with SUMS also this thread's share of
// sum log(|x| + 1e-10) (f32 first), sum |x|, sum x^2, sum x.  Contains barriers: every thread of the CTA calls it.
template <bool SUMS>
__device__ __forceinline__ void tns_masked_own(const enc_params& P, const double* __restrict__ thr_in, band_work* bw, const uint8_t* bidx,
                                               int32_t* stage, int s, double (&x)[FRAD_TNS_BPT], double (&s4)[4]) {
  constexpr int BPT = FRAD_TNS_BPT;
  const int t = threadIdx.x, N = P.N, C = P.C;
  const int b = s / C, c = s - b * C;
  const int own = (BPT + 1) * t;                         // tpad(16 t)
  __syncthreads();                                       // the previous sequence's staging row and band parameters are free
  const int32_t* qrow = P.q + (size_t)b * N * C + c;
  {
    const int32_t* qp = qrow + (size_t)t * C;
    const size_t qstep = (size_t)FRAD_TNS_THREADS * C;
    int32_t* sp = stage + tpad(t);                       // tpad(t + 128 j) = tpad(t) + 136 j
#pragma unroll
    for (int j = 0; j < BPT; j++) sp[(FRAD_TNS_THREADS + FRAD_TNS_THREADS / 16) * j] = __ldg(qp + qstep * j);
  }
  if (t <= FRAD_NBANDS) bw->thr[t] = thr_in[(size_t)s * (FRAD_NBANDS + 1) + t];
  __syncthreads();
  ramp_setup(bw, P.bt, N);
  __syncthreads();
  const uint4 bq = *(const uint4*)(bidx + BPT * t);
  const uint32_t bwd[4] = {bq.x, bq.y, bq.z, bq.w};
  int cur = -1, lo = 0; double step = 0.0, thr0 = 0.0;
  float slf = 0.0f;
#pragma unroll
  for (int i = 0; i < BPT; i++) {
    const int bnd = (int)((bwd[i >> 2] >> (8 * (i & 3))) & 255u);
    if (bnd != cur) { cur = bnd; if (bnd != 255) { lo = bw->lo[bnd]; step = bw->step[bnd]; thr0 = bw->thr[bnd]; } }
    const double xv = (double)__int_as_float(stage[own + i]);
    const double div = (bnd == 255) ? 0.0 : __dadd_rn(__dmul_rn((double)(BPT * t + i - lo), step), thr0);
    const double v = (div == 0.0) ? 0.0 : xv / div;
    x[i] = v;
    if (SUMS) {
      const double m = fabs(v);
      // log2 from the SFU (absolute error 2^-22): the sum only feeds the flatness test, which is re-done in f64 inside +-1e-3
      slf += __log2f((float)(m + 1e-10)); s4[1] += m; s4[2] += v * v; s4[3] += v;
    }
  }
  if (SUMS) s4[0] = (double)slf * 0.69314718055994530942;
}

__global__ void __launch_bounds__(FRAD_TNS_THREADS, FRAD_TNS_MINB_STATS)
tns_stats_kernel(enc_params P, const double* __restrict__ thr_in, tns_ctl* __restrict__ ctl, double* __restrict__ xbuf) {
  constexpr int BPT = FRAD_TNS_BPT, L = FRAD_LPC_LEN, H = FRAD_LPC_LEN - 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  band_work* bw = (band_work*)smem_raw;
  double* ws = (double*)(smem_raw + ((sizeof(band_work) + 15) & ~(size_t)15));     // normalised signal (autocorrelation halos)
  int32_t* stage = (int32_t*)ws;                                                   // input staging row (before ws is live)
  uint8_t* bidx = (uint8_t*)(ws + FRAD_TNS_ROW);
  const int t = threadIdx.x, N = P.N;
  const int nseq = P.B * P.C;
  const int own = (BPT + 1) * t;
  tns_cta_setup(P, bw, bidx);
  for (int s = blockIdx.x; s < nseq; s += gridDim.x) {
    tns_prefetch_next(P, s, nseq);
    double x[BPT];
    double s4[4] = {0.0, 0.0, 0.0, 0.0};
    tns_masked_own<true>(P, thr_in, bw, bidx, stage, s, x, s4);
    {                                                    // hand the masked spectrum to tns_apply_kernel: 128 contiguous bytes per thread
      double2* xo = (double2*)(xbuf + (size_t)s * N + BPT * t);
#pragma unroll
      for (int i = 0; i < BPT / 2; i++) xo[i] = make_double2(x[2 * i], x[2 * i + 1]);
    }
    tns_sum_multi<4>(s4, bw->redm);                      // (its barriers also retire the staging row)
    const double sa = s4[1], se = s4[2], sx = s4[3];
    const double arith = sa / N;
    double flat = exp(s4[0] / N) / (arith + 1e-10);
    if (!(fabs(flat - 0.5) > 1e-3)) {                    // too close to call (or NaN): exact f64 sum of logs, same for every thread
      double sl = 0.0;
#pragma unroll
      for (int i = 0; i < BPT; i++) sl += log(fabs(x[i]) + 1e-10);
      sl = tns_sum(sl, bw->redm);
      flat = exp(sl / N) / (arith + 1e-10);
    }
    const bool cond = (flat < 0.5) && !(se < 1e-10);     // uniform over the CTA: every thread holds the same sums
    if (cond) {
      // ---- autocorrelation of the mean-removed, unit-norm signal (p2tools.py:8-15) ----
      const double mean = sx / N;
      double sg[BPT + H];
      double ss = 0.0;
#pragma unroll
      for (int i = 0; i < BPT; i++) { sg[i] = x[i] - mean; ss += sg[i] * sg[i]; }
      ss = tns_sum(ss, bw->redm);
      const double nrm = sqrt(ss);
      if (nrm > 1e-6) {
        // x / nrm with one division per thread: q0 = x r, q = q0 + (x - q0 nrm) r with r = RN(1 / nrm) is the correctly
        // rounded quotient (Markstein), except possibly when the significand of nrm is all ones -- that case divides
        const double rinv = 1.0 / nrm;
        const bool plain = ((unsigned long long)__double_as_longlong(nrm) & 0xFFFFFFFFFFFFFull) == 0xFFFFFFFFFFFFFull;
        if (!plain) {
#pragma unroll
          for (int i = 0; i < BPT; i++) {
            const double w = sg[i];
            const double q0 = w * rinv;
            sg[i] = fma(fma(-q0, nrm, w), rinv, q0);
          }
        } else {
#pragma unroll
          for (int i = 0; i < BPT; i++) sg[i] = tns_div(sg[i], nrm);
        }
      }
#pragma unroll
      for (int i = 0; i < BPT; i++) ws[own + i] = sg[i];
      __syncthreads();
#pragma unroll
      for (int j = 0; j < H; j++) sg[BPT + j] = (t + 1 < FRAD_TNS_THREADS) ? ws[own + (BPT + 1) + j] : 0.0;   // next block's first 12
      double acc[L];
#pragma unroll
      for (int l = 0; l < L; l++) acc[l] = 0.0;
#pragma unroll
      for (int i = 0; i < BPT; i++) {
#pragma unroll
        for (int l = 0; l < L; l++) acc[l] += sg[i] * sg[i + l];
      }
      // warp totals to shared memory (16 slots, 13 used); thread l adds up lag l, applies the lag window (p2tools.py:14) and
      // hands it over
      const int w = t >> 5, nw = FRAD_TNS_THREADS >> 5;
      {
        double a16[16];
#pragma unroll
        for (int l = 0; l < 16; l++) a16[l] = (l < L) ? acc[l] : 0.0;
        warp_reduce_scatter<16>(a16, t & 31);
        const int idx = wrs_index<16>(t & 31);
        if ((t & 1) == 0 && idx < L) bw->redm[idx * 32 + w] = a16[0];
      }
      __syncthreads();
      if (t < L) {
        double v = 0.0;
        for (int i = 0; i < nw; i++) v += bw->redm[t * 32 + i];
        const double wl = 0.01 * t;
        ctl[s].r[t] = v * exp(-0.5 * (wl * wl));
      }
    }
    if (t == 0) { ctl[s].sx = sx; ctl[s].cond = cond ? 1 : 0; }
  }
}

// Levinson-Durbin + 4-bit LPC quantiser (p2tools.py:17-34, 82-88) on the windowed lags; one thread per sequence.
__global__ void __launch_bounds__(64) tns_lpc_kernel(tns_ctl* __restrict__ ctl, int nseq) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseq) return;
  int flag = 0;
  int aqv[FRAD_LPC_LEN];
#pragma unroll
  for (int j = 0; j < FRAD_LPC_LEN; j++) aqv[j] = 0;
  if (ctl[s].cond) {
    double r[FRAD_LPC_LEN];
#pragma unroll
    for (int l = 0; l < FRAD_LPC_LEN; l++) r[l] = ctl[s].r[l];
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
#pragma unroll
    for (int j = 1; j < FRAD_LPC_LEN; j++) {
      double v = a[j] * 15.0;
      v = fmin(fmax(v, -15.0), 14.0);
      aqv[j] = (int)rint(v);
      any |= (aqv[j] != 0);
    }
    flag = (ok && any) ? 1 : 0;
  }
#pragma unroll
  for (int j = 0; j < FRAD_LPC_LEN; j++) ctl[s].aq[j] = flag ? aqv[j] : 0;
  ctl[s].flag = flag;
}

__global__ void __launch_bounds__(FRAD_TNS_THREADS, FRAD_TNS_MINB_APPLY)
tns_apply_kernel(enc_params P, const tns_ctl* __restrict__ ctl, const double* __restrict__ xbuf) {
  constexpr int BPT = FRAD_TNS_BPT, L = FRAD_LPC_LEN, H = FRAD_LPC_LEN - 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  band_work* bw = (band_work*)smem_raw;
  double* xs = (double*)(smem_raw + ((sizeof(band_work) + 15) & ~(size_t)15));     // masked spectrum, padded rows (FIR halos)
  int32_t* stage = (int32_t*)(xs + FRAD_TNS_ROW);                                  // output staging row
  const int t = threadIdx.x, N = P.N, C = P.C;
  const int nseq = P.B * C;
  const int own = (BPT + 1) * t;
  for (int s = blockIdx.x; s < nseq; s += gridDim.x) {
    const int b = s / C, c = s - b * C;
    double x[BPT];
    {
      const double2* xi = (const double2*)(xbuf + (size_t)s * N + BPT * t);
#pragma unroll
      for (int i = 0; i < BPT / 2; i++) { const double2 v = __ldg(xi + i); x[2 * i] = v.x; x[2 * i + 1] = v.y; }
    }
    if (s + gridDim.x < nseq) {                          // pull this CTA's next sequence towards L2 (one 128-byte line per thread)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(xbuf + (size_t)(s + gridDim.x) * N + BPT * t));
      if (t == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(ctl + s + gridDim.x));
    }
    // the sequence's hand-over record is requested together with its spectrum, ahead of the barrier (three dependent loads
    // behind it were 11 % of the kernel's stall samples)
    const int flag_s = ctl[s].flag;
    const int aq_ld = (t < L) ? ctl[s].aq[t] : 0;
    const double sx = ctl[s].sx;
    __syncthreads();                                     // the previous sequence's rows and coefficients are free
    bool accept = flag_s != 0;                           // uniform over the CTA
    int aq_t = 0;
    double y[BPT];
    if (accept) {
      if (t < L) { aq_t = aq_ld; bw->a[t] = (t == 0) ? 1.0 : (double)aq_t / 15.0; }            // dequantise_lpc, p2tools.py:49-55
#pragma unroll
      for (int i = 0; i < BPT; i++) xs[own + i] = x[i];
      __syncthreads();
      // ---- FIR residual, lfilter(lpc_deq, [1], x): b0 x[n] + (b1 x[n-1] + (... + b12 x[n-12])), tap by tap over the block ----
      double xh[H + BPT];                                // xh[H + i] = x[i], xh[0 .. H) = the previous block's last 12 bins
#pragma unroll
      for (int j = 0; j < H; j++) xh[j] = (t > 0) ? xs[own - (BPT + 1) + (BPT - H) + j] : 0.0;
#pragma unroll
      for (int i = 0; i < BPT; i++) xh[H + i] = x[i];
      if (t > 0) {                                       // all taps present
        const double c12 = bw->a[L - 1];
#pragma unroll
        for (int i = 0; i < BPT; i++) y[i] = __dmul_rn(c12, xh[i]);
#pragma unroll
        for (int k = L - 2; k >= 0; k--) {
          const double ck = bw->a[k];
#pragma unroll
          for (int i = 0; i < BPT; i++) y[i] = __dadd_rn(y[i], __dmul_rn(ck, xh[H + i - k]));
        }
      } else {                                           // the first bins of the sequence: only the taps that exist (lfilter's start-up)
#pragma unroll
        for (int i = 0; i < BPT; i++) {
          double a12 = 0.0;
          bool first = true;
#pragma unroll
          for (int k = L - 1; k >= 1; k--) {
            if (i - k >= 0) {
              const double pp = __dmul_rn(bw->a[k], xh[H + i - k]);
              a12 = first ? pp : __dadd_rn(a12, pp);
              first = false;
            }
          }
          const double p0 = __dmul_rn(bw->a[0], xh[H + i]);
          y[i] = first ? p0 : __dadd_rn(a12, p0);
        }
      }
      double mxv = 0.0;
      double s2[2] = {0.0, 0.0};                         // count of non-finite outputs, sum of the residual
#pragma unroll
      for (int i = 0; i < BPT; i++) {
        if (!isfinite(y[i])) s2[0] = 1.0;
        mxv = fmax(mxv, fabs(y[i]));
        s2[1] += y[i];
      }
      mxv = tns_max(mxv, bw->red);
      tns_sum_multi<2>(s2, bw->redm);
      const double sr = s2[1];
      accept = !(s2[0] > 0.0 || mxv > 1e6);
      // prediction gain (p2tools.py:57-70)
      const double mo = sx / N, mr = sr / N;
      double e2[2] = {0.0, 0.0};
#pragma unroll
      for (int i = 0; i < BPT; i++) {
        const double d0 = x[i] - mo, d1 = y[i] - mr;
        e2[0] += d0 * d0; e2[1] += d1 * d1;
      }
      tns_sum_multi<2>(e2, bw->redm);
      const double eo = e2[0], er = e2[1];
      double gain = 0.0;
      if (!(eo < 1e-10 || er < 1e-10 || er >= eo)) gain = 20.0 * log10(eo / er);
      if (gain < 0.030102999566398118) accept = false;    // log10(2) / 10
    }
    // ---- quantise the residual (or the masked spectrum) and store through the staging row ----
#pragma unroll
    for (int i = 0; i < BPT; i++) stage[own + i] = quant075_fast((accept ? y[i] : x[i]) * P.scale);
    if (t < L) P.lpc[((size_t)b * L + t) * C + c] = accept ? aq_t : 0;
    __syncthreads();
    int32_t* qrow = P.q + (size_t)b * N * C + c;
    {
      int32_t* qp = qrow + (size_t)t * C;
      const size_t qstep = (size_t)FRAD_TNS_THREADS * C;
      const int32_t* sp = stage + tpad(t);
#pragma unroll
      for (int j = 0; j < BPT; j++) qp[qstep * j] = sp[(FRAD_TNS_THREADS + FRAD_TNS_THREADS / 16) * j];
    }
  }
}

__device__ __noinline__ double dequant_slow(int qv, const double* __restrict__ dq64, int dq_len, double inv_scale) {
  const uint32_t qa = (qv < 0) ? 0u - (uint32_t)qv : (uint32_t)qv;
  const double mag = ((qa < (uint32_t)dq_len) ? __ldg(dq64 + qa) : pow((double)qa, 1.0 / 0.75)) * inv_scale;
  return (qv < 0) ? -mag : mag;
}

// Profile-2 decode, first steps (profile2.py:80-85): dequantise the coefficients of every (frame, channel) sequence
// (sign(q) |q|^(4/3) / 2^(bits-1)) and run the all-pole TNS synthesis lfilter([1], lpc / 15, x) over them (p2tools.py:97-106).
//
// The synthesis is a recurrence along the bins, so its parallelism is across sequences: one thread per sequence, filter
// state in registers, SciPy's transposed direct form II order of operations.  Round 2, second version: the two steps are
// separate kernels and meet in a layout made for the recurrence,
//     spec[((s >> 2) N + n) 4 + (s & 3)],   s = frame * C + channel,
// i.e. bin n of four neighbouring sequences is one 32-byte sector -- the four sequences are exactly one tile of the N = 2048
// inverse-transform kernel, which therefore reads its 64 KiB as one contiguous run.  tns_dequant_kernel is a plain pass
// (every bin in parallel, table look-ups off anybody's critical path); tns_synth_kernel then loads x[n] and stores y[n] as
// whole sectors (eight per warp request) -- the first version gathered 4-byte integers at a 16 KiB stride, chained a table
// look-up behind each of them and scattered 8-byte results over half-used sectors.  (With all 32 sequences of a warp in one
// 256-byte line the recurrence is faster still, 0.196 ms, but the transform kernel then reads 16 bytes out of every line:
// 0.27 -> 0.45 ms, measured.)  A sequence whose LPC integers are all zero needs nothing (its coefficients are already in
// place); one whose output is not finite / exceeds 1e6 is dequantised again.
// q: [B][N][C]; lpc: [B][13][C]; dq64: |q|^(4/3) for |q| < 4096 (host libm).
__device__ __forceinline__ size_t spec_index(int s, int n, int N) { return ((size_t)(s >> 2) * N + n) * 4 + (s & 3); }

// One thread per (group of four sequences, bin): 32 bytes out.  Stereo: the group is two frames, an 8-byte load from each.
// The four table look-ups are issued without a branch between them; magnitudes beyond the table (rare) are redone afterwards.
__global__ void __launch_bounds__(256) tns_dequant_kernel(const int32_t* __restrict__ q, int B, int N, int C, double inv_scale,
                                                          const double* __restrict__ dq64, int dq_len, double* __restrict__ spec) {
  const int nseq = B * C;
  const long long total = (long long)((nseq + 3) / 4) * N;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int g = (int)((unsigned long long)e / (unsigned)N), n = (int)(e - (long long)g * N);
    int qv[4];
    if (C == 2) {
      const int b = 2 * g;
      const int2 q0 = __ldg((const int2*)(q + ((size_t)b * N + n) * 2));
      const int2 q1 = (b + 1 < B) ? __ldg((const int2*)(q + ((size_t)(b + 1) * N + n) * 2)) : make_int2(0, 0);
      qv[0] = q0.x; qv[1] = q0.y; qv[2] = q1.x; qv[3] = q1.y;
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int sq = 4 * g + j;
        const int b = sq / C, c = sq - b * C;
        qv[j] = (sq < nseq) ? __ldg(q + ((size_t)b * N + n) * C + c) : 0;
      }
    }
    double v[4];
    uint32_t big = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t qa = (qv[j] < 0) ? 0u - (uint32_t)qv[j] : (uint32_t)qv[j];
      big |= qa;
      const double mag = __ldg(dq64 + min(qa, (uint32_t)(dq_len - 1))) * inv_scale;
      v[j] = (qv[j] < 0) ? -mag : mag;
    }
    if (big >= (uint32_t)dq_len) {
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = dequant_slow(qv[j], dq64, dq_len, inv_scale);
    }
    double2* o = (double2*)(spec + (size_t)e * 4);
    o[0] = make_double2(v[0], v[1]);
    o[1] = make_double2(v[2], v[3]);
  }
}

__global__ void __launch_bounds__(32) tns_synth_kernel(const int32_t* __restrict__ q, const int32_t* __restrict__ lpc, int B, int N, int C,
                                                        double inv_scale, const double* __restrict__ dq64, int dq_len,
                                                        double* __restrict__ spec) {
  const int lane = threadIdx.x;
  const int s0 = blockIdx.x * 32 + lane;
  const bool live = s0 < B * C;                                        // (no early return: the warp works together at the end)
  const int s = live ? s0 : B * C - 1;
  const int b = s / C, c = s - b * C;
  double* col = spec + (size_t)(s0 >> 2) * N * 4 + (s0 & 3);           // bin n of this thread's sequence: col[4 n]
  double a[FRAD_LPC_LEN];
  int any = 0;
#pragma unroll
  for (int i = 0; i < FRAD_LPC_LEN; i++) {
    const int v = lpc[((size_t)b * FRAD_LPC_LEN + i) * C + c];
    any |= (v != 0);
    a[i] = (i == 0) ? 1.0 : (double)v / 15.0;                          // dequantise_lpc, p2tools.py:49-55
  }
  const bool run = any && live;
  bool bad = false;
  double mx = 0.0;
  if (__any_sync(0xffffffffu, run)) {
    // Lanes without TNS walk along with a = (1, 0, ..., 0): y = x, stored over itself -- cheaper than diverging the warp.
    if (!run) {
#pragma unroll
      for (int i = 1; i < FRAD_LPC_LEN; i++) a[i] = 0.0;
    }
    double z[FRAD_LPC_LEN - 1];
#pragma unroll
    for (int i = 0; i < FRAD_LPC_LEN - 1; i++) z[i] = 0.0;
    // The recurrence is serial along n; the loads are not: the block after the current one is in flight while it runs.
    constexpr int BLK = 16;
    double xcur[BLK], xnext[BLK];
    auto fetch = [&](int n0, double (&xv)[BLK]) {
#pragma unroll
      for (int j = 0; j < BLK; j++) xv[j] = (n0 + j < N) ? col[(size_t)(n0 + j) * 4] : 0.0;
    };
    fetch(0, xcur);
    for (int n0 = 0; n0 < N; n0 += BLK) {
      fetch(n0 + BLK, xnext);
#pragma unroll
      for (int j = 0; j < BLK; j++) {
        if (n0 + j < N) {
          const double yn = __dadd_rn(z[0], xcur[j]);
          // SciPy also adds b[i] x[n] with b[i] = 0 to every state; x[n] is finite here (|q| < 2^31), so those terms are
          // +-0 and could only flip the sign of a state that is exactly zero -- left out (half of the arithmetic)
#pragma unroll
          for (int i = 0; i < FRAD_LPC_LEN - 2; i++) z[i] = __dsub_rn(z[i + 1], __dmul_rn(yn, a[i + 1]));
          z[FRAD_LPC_LEN - 2] = -__dmul_rn(yn, a[FRAD_LPC_LEN - 1]);
          col[(size_t)(n0 + j) * 4] = yn;
          bad |= !isfinite(yn);
          mx = fmax(mx, fabs(yn));
        }
      }
#pragma unroll
      for (int j = 0; j < BLK; j++) xcur[j] = xnext[j];
    }
  }
  // a rejected synthesis keeps the unfiltered coefficients: the warp dequantises those sequences again, lanes along the bins
  unsigned todo = __ballot_sync(0xffffffffu, run && (bad || mx > 1e6));
  while (todo) {
    const int l = __ffs(todo) - 1;
    todo &= todo - 1;
    const int sl = __shfl_sync(0xffffffffu, s, l);
    const int bl = sl / C, cl = sl - bl * C;
    const int32_t* ql = q + (size_t)bl * N * C + cl;
    double* ol = spec + (size_t)(sl >> 2) * N * 4 + (sl & 3);
    for (int n = lane; n < N; n += 32) ol[(size_t)n * 4] = dequant_slow(__ldg(ql + (size_t)n * C), dq64, dq_len, inv_scale);
  }
}

// ---- profile 4 (fourier/profile4.py): the PCM itself, no transform -- the value packing of profile 0 applied to the
// samples in their interleaved order (profile4.py:28-41), and back (profile4.py:46-63).  One CTA per frame on the way
// in (per-frame peak for the depth bump of profile4.py:23-25), one thread per value on the way out.
// four values -> their packed images (whole-word depths 16 / 32 / 64), one vector store
__device__ __forceinline__ void p4_store4(uint8_t* dst, const double* v, int bits, bool le) {
  if (bits == 32) {
    uint32_t u[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { u[j] = f2f_bits(v[j]); if (!le) u[j] = __byte_perm(u[j], 0, 0x0123); }
    *(uint4*)dst = make_uint4(u[0], u[1], u[2], u[3]);
  } else if (bits == 16) {
    uint32_t h[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { h[j] = f2h_bits(v[j]); if (!le) h[j] = ((h[j] >> 8) | (h[j] << 8)) & 0xFFFFu; }
    *(uint2*)dst = make_uint2(h[0] | (h[1] << 16), h[2] | (h[3] << 16));
  } else {
    uint32_t w[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint64_t u = f2d_bits(v[j]);
      const uint32_t lo = (uint32_t)u, hi = (uint32_t)(u >> 32);
      w[2 * j] = le ? lo : __byte_perm(hi, 0, 0x0123);
      w[2 * j + 1] = le ? hi : __byte_perm(lo, 0, 0x0123);
    }
    ((uint4*)dst)[0] = make_uint4(w[0], w[1], w[2], w[3]);
    ((uint4*)dst)[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
}

__global__ void __launch_bounds__(256) p4_encode_kernel(enc_params P, int vec) {
  __shared__ double red[33];
  const size_t nv = (size_t)P.N * P.C;
  for (int b = blockIdx.x; b < P.B; b += gridDim.x) {
    uint8_t* frame = P.payload + (size_t)b * P.payload_stride;
    const int64_t base = (int64_t)b * P.hop * P.C;
    double mx = 0.0;
    if (vec) {
      // float PCM, whole-word depth, everything 16-byte aligned: four samples per thread and step
      const int ob = 4 * (P.bits >> 3);
      for (size_t e4 = threadIdx.x; e4 < nv / 4; e4 += blockDim.x) {
        double v[4];
        if (P.pcm_dtype == FRAD_F32) {
          const float4 x = __ldg((const float4*)((const float*)P.pcm + base) + e4);
          v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
        } else {
          const double2* src = (const double2*)((const double*)P.pcm + base) + 2 * e4;
          const double2 x0 = __ldg(src), x1 = __ldg(src + 1);
          v[0] = x0.x; v[1] = x0.y; v[2] = x1.x; v[3] = x1.y;
        }
        mx = fmax(fmax(mx, fmax(fabs(v[0]), fabs(v[1]))), fmax(fabs(v[2]), fabs(v[3])));
        p4_store4(frame + e4 * ob, v, P.bits, P.le != 0);
      }
    } else {
      for (size_t e = threadIdx.x; e < nv; e += blockDim.x) {
        const double v = load_pcm<double>(P.pcm, P.pcm_dtype, base + (int64_t)e);
        mx = fmax(mx, fabs(v));
        if (P.bits == 12) P.half_scratch[(size_t)b * nv + e] = (uint16_t)(f2h_bits(v) >> 4);
        else p0_store_value<double>(frame + e * (size_t)(P.bits >> 3), v, P.bits, P.le != 0);
      }
    }
    if (P.peak) {
      mx = block_max(mx, red);
      if (threadIdx.x == 0) P.peak[b] = mx;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) p4_decode_kernel(dec_params P, int vec) {
  const size_t nv = (size_t)P.N * P.C, total = nv * (size_t)P.B;
  if (vec) {
    // 32-bit values, payload rows and output 16-byte aligned, nv % 4 == 0: four values per thread and step
    const size_t q4 = nv / 4, total4 = q4 * (size_t)P.B;
    const int dt = P.pcm_dtype;
    for (size_t i4 = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i4 < total4; i4 += (size_t)gridDim.x * blockDim.x) {
      const size_t b = i4 / q4, e4 = i4 - b * q4;
      const uint4 x = __ldg((const uint4*)(P.payload + b * P.payload_stride) + e4);
      const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
      uint64_t o[4];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t u = P.le ? xs[j] : __byte_perm(xs[j], 0, 0x0123);
        const double v = ((u & 0x7f800000u) == 0x7f800000u) ? 0.0 : (double)__uint_as_float(u);    // NaN / Inf -> 0
        o[j] = pcm_image(v, dt);
      }
      const size_t at = b * nv + 4 * e4;
      switch (pcm_size(dt)) {
        case 8: {
          uint4* dst = (uint4*)((uint64_t*)P.pcm + at);
          dst[0] = make_uint4((uint32_t)o[0], (uint32_t)(o[0] >> 32), (uint32_t)o[1], (uint32_t)(o[1] >> 32));
          dst[1] = make_uint4((uint32_t)o[2], (uint32_t)(o[2] >> 32), (uint32_t)o[3], (uint32_t)(o[3] >> 32));
          break;
        }
        case 4: *(uint4*)((uint32_t*)P.pcm + at) = make_uint4((uint32_t)o[0], (uint32_t)o[1], (uint32_t)o[2], (uint32_t)o[3]); break;
        case 2: *(uint2*)((uint16_t*)P.pcm + at) = make_uint2((uint32_t)o[0] | ((uint32_t)o[1] << 16), (uint32_t)o[2] | ((uint32_t)o[3] << 16)); break;
        default: *(uint32_t*)((uint8_t*)P.pcm + at) = (uint32_t)o[0] | ((uint32_t)o[1] << 8) | ((uint32_t)o[2] << 16) | ((uint32_t)o[3] << 24); break;
      }
    }
    return;
  }
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / nv, e = i - b * nv;
    double v;
    if (P.bits == 12) {
      v = (double)__half2float(__ushort_as_half(P.half_scratch[i]));
      if (!isfinite(v)) v = 0.0;
    } else {
      v = p0_load_value(P.payload + b * P.payload_stride + e * (size_t)(P.bits >> 3), P.bits, P.le != 0);   // NaN / Inf -> 0
    }
    store_pcm<double>(P.pcm, P.pcm_dtype, (int64_t)i, v);
  }
}

// 12-bit packing helpers (profile0.py:37-41, 54-58): 16-bit scratch <-> nibble-packed payload
__global__ void pack12_kernel(const uint16_t* __restrict__ half_top12, size_t nvals, int B, uint8_t* payload, size_t stride) {
  const size_t npairs = (nvals + 1) / 2;
  const size_t total = npairs * (size_t)B;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / npairs, p = i % npairs;
    const uint16_t* src = half_top12 + b * nvals;
    uint32_t a = src[2 * p];
    const bool has_b = (2 * p + 1) < nvals;
    uint32_t bb = has_b ? src[2 * p + 1] : 0;
    uint8_t* dst = payload + b * stride + 3 * p;
    dst[0] = (uint8_t)(a >> 4);
    dst[1] = (uint8_t)(((a & 0xF) << 4) | (bb >> 8));
    if (has_b) dst[2] = (uint8_t)bb;
  }
}

__global__ void unpack12_kernel(const uint8_t* __restrict__ payload, size_t stride, size_t nvals, int B, uint16_t* half_bits) {
  const size_t total = nvals * (size_t)B;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / nvals, v = i % nvals;
    const uint8_t* src = payload + b * stride + 3 * (v / 2);
    uint32_t t;
    if ((v & 1) == 0) t = ((uint32_t)src[0] << 4) | (src[1] >> 4);
    else t = (((uint32_t)src[1] & 0xF) << 8) | src[2];
    half_bits[i] = (uint16_t)(t << 4);
  }
}

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
#ifndef FRAD_GEN_TPB_DIV
#define FRAD_GEN_TPB_DIV 8            // one radix-8 butterfly per thread and pass (measured: 4 -> 8 gains 10-20 % with radix 8)
#endif
static int pick_threads(int M) {
  int t = (M + FRAD_GEN_TPB_DIV - 1) / FRAD_GEN_TPB_DIV;
  t = ((t + 31) / 32) * 32;
  if (t < 64) t = 64;
  if (t > 1024) t = 1024;
  return t;
}

struct launch_cfg { int threads, grid, b_in_smem; size_t smem; size_t gscratch; };

template <typename T>
static int configure(frad_ctx* ctx, const frad_plan* pl, int nseq, const void* kernel, launch_cfg* cfg) {
  const size_t one = (size_t)pl->M * 2 * sizeof(T);
  const size_t fixed = sizeof(band_work);
  const size_t limit = ctx->smem_optin;
  cfg->threads = pick_threads(pl->M);
  if (fixed + 2 * one <= limit) { cfg->b_in_smem = 1; cfg->smem = fixed + 2 * one; }
  else if (fixed + one <= limit) { cfg->b_in_smem = 0; cfg->smem = fixed + one; }
  else return frad_fail(ctx, FRAD_E_UNSUPPORTED, "frame too long for the shared-memory transform");
  FRAD_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg->smem));
  int per_sm = 0;
  FRAD_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, cfg->threads, cfg->smem));
  if (per_sm < 1) per_sm = 1;
  long long g = (long long)per_sm * ctx->sm_count;
  cfg->grid = (int)((nseq < g) ? nseq : g);
  if (cfg->grid < 1) cfg->grid = 1;
  cfg->gscratch = cfg->b_in_smem ? 0 : (size_t)cfg->grid * one;
  return FRAD_OK;
}

static int resolve_prec_fwd(int prec, int pcm_dtype) {
  if (prec == FRAD_PREC_F32 || prec == FRAD_PREC_F64) return prec;
  return (pcm_dtype == FRAD_F32 || pcm_dtype == FRAD_F16) ? FRAD_PREC_F32 : FRAD_PREC_F64;
}
static int resolve_prec_inv(int prec) { return (prec == FRAD_PREC_F32) ? FRAD_PREC_F32 : FRAD_PREC_F64; }

int frad_launch_encode_2048r(frad_ctx* ctx, const enc_params& P, const frad_plan* pl, int prec, int force_exact);     // enc2048.cu
static int frad_launch_encode_2048(frad_ctx* ctx, const enc_params& P, const frad_plan* pl, int prec, int force_exact) {
  return frad_launch_encode_2048r(ctx, P, pl, prec, force_exact);
}
int frad_launch_decode_2048(frad_ctx* ctx, const dec_params& P, const frad_plan* pl, int prec);
bool frad_fast2048_encode_ok(const enc_params& P, int prec);
bool frad_fast_small_encode_ok(const enc_params& P, int prec);                                                       // enc2048.cu
int frad_launch_encode_small(frad_ctx* ctx, const enc_params& P, const frad_plan* pl, int prec, int force_exact);
bool frad_fast2048_decode_ok(const dec_params& P, int prec);
bool frad_fast_small_decode_ok(const dec_params& P, int prec);                                                       // decs.cu
int frad_launch_decode_small(frad_ctx* ctx, const dec_params& P, const frad_plan* pl, int prec);

static int launch_encode(frad_ctx* ctx, enc_params& P, int prec) {
  if (P.B <= 0) return FRAD_OK;
  frad_plan* pl = frad_get_plan(ctx, P.N, prec);
  if (!pl) return FRAD_E_NOMEM;
  const int nseq = P.B * P.C;
  if (P.profile == 2) {   // TNS work arrays: masked + residual per resident CTA
    // sized below once the grid is known
  }
  if (pl->direct) {
    int per_sm = 4;
    int grid = per_sm * ctx->sm_count; if (grid > nseq) grid = nseq;
    size_t need = (size_t)grid * (2 * P.N) * sizeof(double);
    size_t tns = (P.profile == 2) ? (size_t)grid * 2 * P.N * sizeof(double) : 0;
    char* g = (char*)frad_scratch(ctx, 0, need + tns);
    if (!g) return FRAD_E_NOMEM;
    P.gscratch = g; P.tns_scratch = (double*)(g + need);
    seq_encode_direct_kernel<<<grid, 256, sizeof(band_work), ctx->stream>>>(P, plan_view(pl));
    FRAD_LAUNCH_CHECK(ctx);
    return FRAD_OK;
  }
#ifdef FRAD_HAVE_FAST2048
  if (!ctx->opt_force_generic && frad_fast_small_encode_ok(P, prec)) {
    P.fast_ok = 2;
    return frad_launch_encode_small(ctx, P, pl, prec, ctx->opt_exact_quant);
  }
  if (!ctx->opt_force_generic && frad_fast2048_encode_ok(P, prec)) {
    P.fast_ok = 2;                                   // tells the caller that 12-bit payloads are already nibble-packed
    if (P.profile != 2) return frad_launch_encode_2048(ctx, P, pl, prec, ctx->opt_exact_quant);
    // profile 2: transform + thresholds, then the three TNS kernels (statistics + lags, Levinson per sequence, residual + quantiser)
    const size_t thr_bytes = (sizeof(double) * (size_t)nseq * (FRAD_NBANDS + 1) + 15) & ~(size_t)15;
    const size_t ctl_bytes = (sizeof(tns_ctl) * (size_t)nseq + 15) & ~(size_t)15;
    char* sc = (char*)frad_scratch(ctx, 4, thr_bytes + ctl_bytes + sizeof(double) * (size_t)nseq * P.N);
    if (!sc) return FRAD_E_NOMEM;
    double* thr = (double*)sc;
    tns_ctl* ctl = (tns_ctl*)(sc + thr_bytes);
    double* xbuf = (double*)(sc + thr_bytes + ctl_bytes);       // masked spectrum [sequence][N] f64, stats -> apply
    P.tns_scratch = thr;
    int rc = frad_launch_encode_2048(ctx, P, pl, prec, ctx->opt_exact_quant);
    if (rc) return rc;
    const size_t head = (sizeof(band_work) + 15) & ~(size_t)15;
    const size_t smem_stats = head + (size_t)FRAD_TNS_ROW * sizeof(double) + (size_t)P.N;
    const size_t smem_apply = head + (size_t)FRAD_TNS_ROW * (sizeof(double) + sizeof(int32_t));
    FRAD_CUDA(ctx, cudaFuncSetAttribute(tns_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_stats));
    FRAD_CUDA(ctx, cudaFuncSetAttribute(tns_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_apply));
    int per_sm = 1;
    FRAD_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tns_stats_kernel, FRAD_TNS_THREADS, smem_stats));
    if (per_sm < 1) per_sm = 1;
    int grid = per_sm * ctx->sm_count; if (grid > nseq) grid = nseq;
    tns_stats_kernel<<<grid, FRAD_TNS_THREADS, smem_stats, ctx->stream>>>(P, thr, ctl, xbuf);
    FRAD_LAUNCH_CHECK(ctx);
    tns_lpc_kernel<<<(nseq + 63) / 64, 64, 0, ctx->stream>>>(ctl, nseq);
    FRAD_LAUNCH_CHECK(ctx);
    FRAD_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tns_apply_kernel, FRAD_TNS_THREADS, smem_apply));
    if (per_sm < 1) per_sm = 1;
    grid = per_sm * ctx->sm_count; if (grid > nseq) grid = nseq;
    tns_apply_kernel<<<grid, FRAD_TNS_THREADS, smem_apply, ctx->stream>>>(P, ctl, xbuf);
    FRAD_LAUNCH_CHECK(ctx);
    return FRAD_OK;
  }
#endif
  launch_cfg cfg;
  // (the 256-thread bound costs the encoder occupancy: 114-118 registers, 2 CTAs/SM, p1 encode at N = 4096 0.71 -> 1.06 ms; the
  //  decoder gains from it: 0.86 -> 0.76 ms.  So only the decoder uses it.)
  const bool small_cta = false;
  const void* k = small_cta ? ((prec == FRAD_PREC_F32) ? (const void*)seq_encode_kernel<float, 256> : (const void*)seq_encode_kernel<double, 256>)
                            : ((prec == FRAD_PREC_F32) ? (const void*)seq_encode_kernel<float, 1024> : (const void*)seq_encode_kernel<double, 1024>);
  int rc = (prec == FRAD_PREC_F32) ? configure<float>(ctx, pl, nseq, k, &cfg) : configure<double>(ctx, pl, nseq, k, &cfg);
  if (rc) return rc;
  size_t tns = (P.profile == 2) ? (size_t)cfg.grid * 2 * P.N * sizeof(double) : 0;
  if (cfg.gscratch + tns) {
    char* g = (char*)frad_scratch(ctx, 0, cfg.gscratch + tns);
    if (!g) return FRAD_E_NOMEM;
    P.gscratch = g; P.tns_scratch = (double*)(g + cfg.gscratch);
  }
  if (small_cta) {
    if (prec == FRAD_PREC_F32) seq_encode_kernel<float, 256><<<cfg.grid, cfg.threads, cfg.smem, ctx->stream>>>(P, plan_view(pl), cfg.b_in_smem);
    else seq_encode_kernel<double, 256><<<cfg.grid, cfg.threads, cfg.smem, ctx->stream>>>(P, plan_view(pl), cfg.b_in_smem);
  } else {
    if (prec == FRAD_PREC_F32) seq_encode_kernel<float, 1024><<<cfg.grid, cfg.threads, cfg.smem, ctx->stream>>>(P, plan_view(pl), cfg.b_in_smem);
    else seq_encode_kernel<double, 1024><<<cfg.grid, cfg.threads, cfg.smem, ctx->stream>>>(P, plan_view(pl), cfg.b_in_smem);
  }
  FRAD_LAUNCH_CHECK(ctx);
  return FRAD_OK;
}

// |q|^(4/3) for |q| < FRAD_DQ_LEN, computed on the host with libm's pow exactly like the reference (p1tools.py:44)
int frad_ensure_dq64(frad_ctx* ctx) {
  if (ctx->dequant_table) return FRAD_OK;
  std::vector<double> h(FRAD_DQ_LEN);
  for (int i = 0; i < FRAD_DQ_LEN; i++) h[i] = pow((double)i, 1.0 / 0.75);
  void* d = nullptr;
  if (cudaMalloc(&d, sizeof(double) * FRAD_DQ_LEN) != cudaSuccess ||
      frad_upload(d, h.data(), sizeof(double) * FRAD_DQ_LEN) != cudaSuccess) {
    cudaGetLastError();
    return frad_fail(ctx, FRAD_E_NOMEM, "dequantiser table upload failed");
  }
  ctx->dequant_table = d;
  return FRAD_OK;
}

// profile 2: dequantiser + TNS synthesis of the whole batch into ctx scratch slot 4; sets P.spec
static int run_tns_synthesis(frad_ctx* ctx, dec_params& P) {
  int rc = frad_ensure_dq64(ctx);
  if (rc) return rc;
  const int nseq = P.B * P.C;
  const int nwarps = (nseq + 31) / 32;
  double* spec = (double*)frad_scratch(ctx, 4, sizeof(double) * (size_t)nwarps * 32 * P.N);        // layout: spec_index()
  if (!spec) return FRAD_E_NOMEM;
  // (A variant that stages TNS_BLK-bin blocks through shared memory inside the recurrence was measured in round 2:
  //  1.01 ms against 0.72 ms for the pair of decode kernels; not kept.)
  const long long want = ((long long)nwarps * 8 * P.N + 255) / 256;
  const long long cap = (long long)ctx->sm_count * 16;
  tns_dequant_kernel<<<(int)(want < cap ? want : cap), 256, 0, ctx->stream>>>(P.q, P.B, P.N, P.C, 1.0 / P.scale,
                                                                             (const double*)ctx->dequant_table, FRAD_DQ_LEN, spec);
  FRAD_LAUNCH_CHECK(ctx);
  tns_synth_kernel<<<nwarps, 32, 0, ctx->stream>>>(P.q, P.lpc, P.B, P.N, P.C, 1.0 / P.scale, (const double*)ctx->dequant_table,
                                                   FRAD_DQ_LEN, spec);
  FRAD_LAUNCH_CHECK(ctx);
  P.spec = spec;
  return FRAD_OK;
}

static int launch_decode(frad_ctx* ctx, dec_params& P, int prec) {
  if (P.B <= 0) return FRAD_OK;
  frad_plan* pl = frad_get_plan(ctx, P.N, prec);
  if (!pl) return FRAD_E_NOMEM;
  const int nseq = P.B * P.C;
  if (P.profile == 2) {
    int rc = run_tns_synthesis(ctx, P);
    if (rc) return rc;
  }
  if (P.profile == 1 || P.profile == 2) {                    // the generic prologue's dequantiser table
    int rc = frad_ensure_dq64(ctx);
    if (rc) return rc;
    P.dq64 = (const double*)ctx->dequant_table; P.dq_len = FRAD_DQ_LEN;
  }
  if (pl->direct) {
    int grid = 4 * ctx->sm_count; if (grid > nseq) grid = nseq;
    size_t need = (size_t)grid * P.N * sizeof(double);
    size_t tns = 0;
    char* g = (char*)frad_scratch(ctx, 0, need + tns);
    if (!g) return FRAD_E_NOMEM;
    P.gscratch = g; P.tns_scratch = (double*)(g + need);
    seq_decode_direct_kernel<<<grid, 256, sizeof(band_work), ctx->stream>>>(P, plan_view(pl));
    FRAD_LAUNCH_CHECK(ctx);
    return FRAD_OK;
  }
#ifdef FRAD_HAVE_FAST2048
  if (!ctx->opt_force_generic && frad_fast_small_decode_ok(P, prec)) return frad_launch_decode_small(ctx, P, pl, prec);
  if (!ctx->opt_force_generic && frad_fast2048_decode_ok(P, prec)) return frad_launch_decode_2048(ctx, P, pl, prec);
#endif
  launch_cfg cfg;
  const bool small_cta = pick_threads(pl->M) <= 256;
  const void* k = small_cta ? ((prec == FRAD_PREC_F32) ? (const void*)seq_decode_kernel<float, 256> : (const void*)seq_decode_kernel<double, 256>)
                            : ((prec == FRAD_PREC_F32) ? (const void*)seq_decode_kernel<float, 1024> : (const void*)seq_decode_kernel<double, 1024>);
  int rc = (prec == FRAD_PREC_F32) ? configure<float>(ctx, pl, nseq, k, &cfg) : configure<double>(ctx, pl, nseq, k, &cfg);
  if (rc) return rc;
  size_t tns = 0;
  if (cfg.gscratch + tns) {
    char* g = (char*)frad_scratch(ctx, 0, cfg.gscratch + tns);
    if (!g) return FRAD_E_NOMEM;
    P.gscratch = g; P.tns_scratch = (double*)(g + cfg.gscratch);
  }
  if (small_cta) {
    if (prec == FRAD_PREC_F32) seq_decode_kernel<float, 256><<<cfg.grid, cfg.threads, cfg.smem, ctx->stream>>>(P, plan_view(pl), cfg.b_in_smem);
    else seq_decode_kernel<double, 256><<<cfg.grid, cfg.threads, cfg.smem, ctx->stream>>>(P, plan_view(pl), cfg.b_in_smem);
  } else {
    if (prec == FRAD_PREC_F32) seq_decode_kernel<float, 1024><<<cfg.grid, cfg.threads, cfg.smem, ctx->stream>>>(P, plan_view(pl), cfg.b_in_smem);
    else seq_decode_kernel<double, 1024><<<cfg.grid, cfg.threads, cfg.smem, ctx->stream>>>(P, plan_view(pl), cfg.b_in_smem);
  }
  FRAD_LAUNCH_CHECK(ctx);
  return FRAD_OK;
}

// ------------------------------------------------------------------------------------------------
// C ABI: profile 0
// ------------------------------------------------------------------------------------------------
static bool p0_bits_ok(int bits) { return bits == 12 || bits == 16 || bits == 24 || bits == 32 || bits == 48 || bits == 64; }

extern "C" int frad_frame_length_class(int N) {
#ifdef FRAD_HAVE_FAST2048
  if (N == 2048) return 2;
  if (frad_warp_r1(N)) return 1;
#endif
  return 0;
}

extern "C" size_t frad_p0_payload_bytes(int N, int C, int bits) {
  size_t nv = (size_t)N * (size_t)C;
  if (bits == 12) return (nv * 3 + 1) / 2;
  return nv * (size_t)(bits / 8);
}

static int lossless_encode_dev(frad_ctx* ctx, int profile, const void* pcm, int pcm_dtype, int64_t hop, int B, int N, int C,
                               int bits, int little_endian, int prec,
                               uint8_t* payload, size_t payload_stride, double* peak) {
  if (!ctx) return FRAD_E_ARG;
  if (B < 0 || N < 1 || C < 1 || !p0_bits_ok(bits) || !frad_dtype_size(pcm_dtype) || !pcm || !payload)
    return frad_fail(ctx, FRAD_E_ARG, "frad_p0_encode_dev: bad argument");
  if (payload_stride < frad_p0_payload_bytes(N, C, bits)) return frad_fail(ctx, FRAD_E_ARG, "payload_stride too small");
  FRAD_CUDA(ctx, cudaSetDevice(ctx->device));
  enc_params P = {};
  P.pcm = pcm; P.pcm_dtype = pcm_dtype; P.hop = hop; P.B = B; P.N = N; P.C = C;
  P.profile = profile; P.bits = bits; P.le = (little_endian && bits % 8 == 0) ? 1 : 0;
  P.payload = payload; P.payload_stride = payload_stride; P.peak = peak;
  P.fast_ok = 1;
  const size_t nvals = (size_t)N * C;
  if (bits == 12) {
    P.half_scratch = (uint16_t*)frad_scratch(ctx, 1, nvals * (size_t)B * sizeof(uint16_t));
    if (!P.half_scratch) return FRAD_E_NOMEM;
  }
  if (peak) FRAD_CUDA(ctx, cudaMemsetAsync(peak, 0, sizeof(double) * (size_t)B, ctx->stream));
  int rc = FRAD_OK;
  if (profile == 4) {
    if (B > 0) {
      const int grid = B < ctx->sm_count * 8 ? B : ctx->sm_count * 8;
      const int es = frad_dtype_size(pcm_dtype);
      const int vec = (pcm_dtype == FRAD_F32 || pcm_dtype == FRAD_F64) && (bits == 16 || bits == 32 || bits == 64) && (nvals % 4) == 0 &&
                      ((uintptr_t)pcm % 16) == 0 && (((size_t)hop * C * es) % 16) == 0 && ((uintptr_t)payload % 16) == 0 &&
                      (payload_stride % 16) == 0;
      p4_encode_kernel<<<grid, 256, 0, ctx->stream>>>(P, vec);
      FRAD_LAUNCH_CHECK(ctx);
    }
  } else {
    rc = launch_encode(ctx, P, resolve_prec_fwd(prec, pcm_dtype));
  }
  if (rc) return rc;
  if (bits == 12 && B > 0 && P.fast_ok != 2) {
    size_t total = ((nvals + 1) / 2) * (size_t)B;
    int grid = (int)((total + 255) / 256); if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
    pack12_kernel<<<grid, 256, 0, ctx->stream>>>(P.half_scratch, nvals, B, payload, payload_stride);
    FRAD_LAUNCH_CHECK(ctx);
  }
  return FRAD_OK;
}

extern "C" int frad_p0_encode_dev(frad_ctx* ctx, const void* pcm, int pcm_dtype, int64_t hop, int B, int N, int C,
                                  int bits, int little_endian, int prec,
                                  uint8_t* payload, size_t payload_stride, double* peak) {
  return lossless_encode_dev(ctx, 0, pcm, pcm_dtype, hop, B, N, C, bits, little_endian, prec, payload, payload_stride, peak);
}
extern "C" int frad_p4_encode_dev(frad_ctx* ctx, const void* pcm, int pcm_dtype, int64_t hop, int B, int N, int C,
                                  int bits, int little_endian, uint8_t* payload, size_t payload_stride, double* peak) {
  return lossless_encode_dev(ctx, 4, pcm, pcm_dtype, hop, B, N, C, bits, little_endian, FRAD_PREC_REF, payload, payload_stride, peak);
}

static int lossless_decode_dev(frad_ctx* ctx, int profile, const uint8_t* payload, size_t payload_stride, int B, int N, int C,
                               int bits, int little_endian, int prec, void* pcm, int pcm_dtype) {
  if (!ctx) return FRAD_E_ARG;
  if (B < 0 || N < 1 || C < 1 || !p0_bits_ok(bits) || !pcm || !payload || !frad_dtype_size(pcm_dtype))
    return frad_fail(ctx, FRAD_E_ARG, "frad_p0_decode_dev: bad argument");
  if (payload_stride < frad_p0_payload_bytes(N, C, bits)) return frad_fail(ctx, FRAD_E_ARG, "payload_stride too small");
  FRAD_CUDA(ctx, cudaSetDevice(ctx->device));
  dec_params P = {};
  P.B = B; P.N = N; P.C = C; P.profile = profile; P.bits = bits; P.le = (little_endian && bits % 8 == 0) ? 1 : 0;
  P.payload = payload; P.payload_stride = payload_stride; P.pcm = pcm; P.pcm_dtype = pcm_dtype;
  P.fast_ok = 1;
  const size_t nvals = (size_t)N * C;
  const bool fast12 = profile == 0 && !ctx->opt_force_generic && (frad_fast2048_decode_ok(P, prec) || frad_fast_small_decode_ok(P, prec));   // unpacks the nibbles itself
  if (bits == 12 && B > 0 && !fast12) {
    uint16_t* hs = (uint16_t*)frad_scratch(ctx, 1, nvals * (size_t)B * sizeof(uint16_t));
    if (!hs) return FRAD_E_NOMEM;
    size_t total = nvals * (size_t)B;
    int grid = (int)((total + 255) / 256); if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
    unpack12_kernel<<<grid, 256, 0, ctx->stream>>>(payload, payload_stride, nvals, B, hs);
    FRAD_LAUNCH_CHECK(ctx);
    P.half_scratch = hs;
  }
  if (profile == 4) {
    if (B > 0) {
      size_t total = nvals * (size_t)B;
      int grid = (int)((total + 255) / 256); if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
      const int vec = bits == 32 && (nvals % 4) == 0 && ((uintptr_t)payload % 16) == 0 && (payload_stride % 16) == 0 &&
                      ((uintptr_t)pcm % 16) == 0;
      if (vec) { total = (nvals / 4) * (size_t)B; grid = (int)((total + 255) / 256); if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16; }
      p4_decode_kernel<<<grid, 256, 0, ctx->stream>>>(P, vec);
      FRAD_LAUNCH_CHECK(ctx);
    }
    return FRAD_OK;
  }
  return launch_decode(ctx, P, resolve_prec_inv(prec));
}
extern "C" int frad_p0_decode_dev(frad_ctx* ctx, const uint8_t* payload, size_t payload_stride, int B, int N, int C,
                                  int bits, int little_endian, int prec, void* pcm, int pcm_dtype) {
  return lossless_decode_dev(ctx, 0, payload, payload_stride, B, N, C, bits, little_endian, prec, pcm, pcm_dtype);
}
extern "C" int frad_p4_decode_dev(frad_ctx* ctx, const uint8_t* payload, size_t payload_stride, int B, int N, int C,
                                  int bits, int little_endian, void* pcm, int pcm_dtype) {
  return lossless_decode_dev(ctx, 4, payload, payload_stride, B, N, C, bits, little_endian, FRAD_PREC_REF, pcm, pcm_dtype);
}

// ------------------------------------------------------------------------------------------------
// C ABI: profiles 1 and 2
// ------------------------------------------------------------------------------------------------
static const int k_compact_srates[12] = {96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000};
static bool compact_size_ok(int N) {
  for (int s = 0; s < 8; s++) for (int b = 128; b <= 224; b += 32) if (N == (b << s)) return true;
  return false;
}
static bool compact_srate_ok(int sr) { for (int i = 0; i < 12; i++) if (k_compact_srates[i] == sr) return true; return false; }

static const band_table* get_band_table(frad_ctx* ctx, int N, int srate) {
  uint64_t key = (uint64_t)(uint32_t)N | ((uint64_t)(uint32_t)srate << 32);
  auto it = ctx->band_tables.find(key);
  if (it != ctx->band_tables.end()) return (const band_table*)it->second;
  static const double hz[28] = {0, 200, 400, 600, 800, 1000, 1200, 1400, 1600, 2000, 2400, 2800, 3200, 4000, 4800, 5600,
                                6800, 8000, 9600, 12000, 15600, 20000, 24000, 28800, 34400, 40800, 48000, 4294967295.0};
  std::vector<band_table> hv(1);                           // (28 KiB: not on the stack)
  band_table& h = hv[0];
  for (int i = 0; i < 28; i++) {
    // p1tools.py:16  round(dlen/(srate/2)*Hz), half-to-even
    double v = (double)N / ((double)srate / 2.0) * hz[i];
    double r = nearbyint(v);
    h.edge[i] = (r > 2147483647.0) ? 2147483647 : (int)r;
  }
  for (int b = 0; b < FRAD_NBANDS; b++) {
    // p1tools.py:25-30
    double f = (hz[b] + hz[b + 1]) / 2.0 / 1000.0;
    double e = (3.64 * pow(f, -0.8) - 6.5 * exp(-0.6 * pow(f - 3.3, 2.0)) + 1e-3 * pow(f, 4.0)) / 20.0;
    double ath = pow(10.0, e);
    h.ath[b] = (ath < 1.0) ? ath : 1.0;
  }
  h.log_base = log(M_E / 2.0);      // np.log(np.e / 2)
  h.base = M_E / 2.0;
  for (int t = -FRAD_THR_TAB; t <= FRAD_THR_TAB; t++) {
    const double v = (double)t;      // np.power(np.e / 2, sign(t) * |t| ** 0.75), profile1.py:66
    h.thr_pow[t + FRAD_THR_TAB] = pow(M_E / 2.0, copysign(pow(fabs(v), 0.75), v));
  }
  {   // ramp band of every bin: ramp_value's search over the clipped band starts (profiles.cuh)
    int lo[FRAD_NBANDS + 1];
    for (int i = 0; i <= FRAD_NBANDS; i++) lo[i] = (h.edge[i] < N) ? h.edge[i] : N;
    for (int k = 0; k < FRAD_MAX_FRAME; k++) {
      int b = 0;
      for (int s2 = 16; s2 > 0; s2 >>= 1) { if (b + s2 <= FRAD_NBANDS - 2 && lo[b + s2] <= k) b += s2; }
      h.bidx[k] = (k >= N || k >= lo[FRAD_NBANDS - 1]) ? 255 : (uint8_t)b;
    }
  }
  void* d = nullptr;
  if (cudaMalloc(&d, sizeof(band_table)) != cudaSuccess ||
      frad_upload(d, &h, sizeof(band_table)) != cudaSuccess) {
    cudaGetLastError();
    frad_fail(ctx, FRAD_E_NOMEM, "band table allocation failed");
    return nullptr;
  }
  ctx->band_tables[key] = d;
  return (const band_table*)d;
}

static bool p12_bits_ok(int bits, bool p2) {
  static const int d1[7] = {8, 12, 16, 24, 32, 48, 64}, d2[7] = {8, 10, 12, 14, 16, 20, 24};
  for (int i = 0; i < 7; i++) if ((p2 ? d2[i] : d1[i]) == bits) return true;
  return false;
}

extern "C" int frad_p1_encode_dev(frad_ctx* ctx, const void* pcm, int pcm_dtype, int64_t hop, int B, int N, int C,
                                  int bits, int srate, double loss_level, int prec,
                                  int32_t* q, int32_t* tq, int32_t* lpc) {
  if (!ctx) return FRAD_E_ARG;
  if (B < 0 || C < 1 || !pcm || !q || !tq || !frad_dtype_size(pcm_dtype) || !compact_size_ok(N) ||
      !compact_srate_ok(srate) || !p12_bits_ok(bits, lpc != nullptr))
    return frad_fail(ctx, FRAD_E_ARG, "frad_p1_encode_dev: bad argument");
  FRAD_CUDA(ctx, cudaSetDevice(ctx->device));
  enc_params P = {};
  P.pcm = pcm; P.pcm_dtype = pcm_dtype; P.hop = hop; P.B = B; P.N = N; P.C = C;
  P.profile = lpc ? 2 : 1; P.bits = bits;
  P.scale = ldexp(1.0, bits - 1);
  loss_level = fabs(loss_level); if (loss_level < 0.125) loss_level = 0.125;    // profile1.py:20
  P.loss = loss_level;
  P.bt = get_band_table(ctx, N, srate);
  if (!P.bt) return FRAD_E_NOMEM;
  P.q = q; P.tq = tq; P.lpc = lpc;
  P.fast_ok = 1;
  return launch_encode(ctx, P, resolve_prec_fwd(prec, pcm_dtype));
}

extern "C" int frad_p1_decode_dev(frad_ctx* ctx, const int32_t* q, const int32_t* tq, const int32_t* lpc,
                                  int B, int N, int C, int bits, int srate, int prec, void* pcm, int pcm_dtype) {
  if (!ctx) return FRAD_E_ARG;
  if (B < 0 || C < 1 || !pcm || !q || !tq || !compact_size_ok(N) || !compact_srate_ok(srate) ||
      !p12_bits_ok(bits, lpc != nullptr) || !frad_dtype_size(pcm_dtype))
    return frad_fail(ctx, FRAD_E_ARG, "frad_p1_decode_dev: bad argument");
  FRAD_CUDA(ctx, cudaSetDevice(ctx->device));
  dec_params P = {};
  P.B = B; P.N = N; P.C = C; P.profile = lpc ? 2 : 1; P.bits = bits;
  P.scale = ldexp(1.0, bits - 1);
  P.bt = get_band_table(ctx, N, srate);
  if (!P.bt) return FRAD_E_NOMEM;
  P.q = q; P.tq = tq; P.lpc = lpc; P.pcm = pcm; P.pcm_dtype = pcm_dtype;
  P.fast_ok = 1;
  return launch_decode(ctx, P, resolve_prec_inv(prec));
}
