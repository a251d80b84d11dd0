// Register-resident 32-point complex FFT (radix 4 x 8) and the warp-level 1024-point FFT built
// from two of them (one shared-memory exchange, no block barrier).  All register indices are
// compile-time constants after unrolling.
#pragma once
#include "devmath.cuh"

// exp(-2 pi i t / 32): cos / sin tables; folded to immediates after unrolling
__device__ __forceinline__ double w32_cos(int t) {
  const double c[32] = {
      1.0, 0.9807852804032304, 0.9238795325112867, 0.8314696123025452, 0.7071067811865476,
      0.5555702330196023, 0.38268343236508984, 0.19509032201612833, 0.0, -0.1950903220161282,
      -0.3826834323650897, -0.555570233019602, -0.7071067811865475, -0.8314696123025453,
      -0.9238795325112867, -0.9807852804032304, -1.0, -0.9807852804032304, -0.9238795325112868,
      -0.8314696123025455, -0.7071067811865477, -0.5555702330196022, -0.38268343236509034,
      -0.19509032201612866, 0.0, 0.1950903220161283, 0.38268343236509, 0.5555702330196018,
      0.7071067811865474, 0.8314696123025452, 0.9238795325112865, 0.9807852804032303};
  return c[t & 31];
}
__device__ __forceinline__ double w32_sin(int t) {   // sin(2 pi t / 32) (positive angle)
  const double s[32] = {
      0.0, 0.19509032201612825, 0.3826834323650898, 0.5555702330196022, 0.7071067811865475,
      0.8314696123025452, 0.9238795325112867, 0.9807852804032304, 1.0, 0.9807852804032304,
      0.9238795325112867, 0.8314696123025455, 0.7071067811865476, 0.5555702330196022, 0.3826834323650899,
      0.1950903220161286, 0.0, -0.19509032201612836, -0.38268343236508967, -0.555570233019602,
      -0.7071067811865475, -0.8314696123025452, -0.9238795325112865, -0.9807852804032303, -1.0,
      -0.9807852804032304, -0.9238795325112866, -0.8314696123025455, -0.7071067811865477,
      -0.5555702330196022, -0.3826834323650904, -0.19509032201612872};
  return s[t & 31];
}

// exp(+i pi o / 128) and exp(+i pi o / 32), o = 0..16 (correctly rounded doubles; folded to immediates after unrolling)
__device__ __forceinline__ double g128_cos(int o) {
  const double t[17] = {
      1.0, 0.9996988186962042, 0.9987954562051724, 0.9972904566786902,
      0.9951847266721969, 0.99247953459871, 0.989176509964781, 0.9852776423889412,
      0.9807852804032304, 0.9757021300385286, 0.970031253194544, 0.9637760657954398,
      0.9569403357322088, 0.9495281805930367, 0.9415440651830208, 0.9329927988347388,
      0.9238795325112867};
  return t[o];
}
__device__ __forceinline__ double g128_sin(int o) {
  const double t[17] = {
      0.0, 0.024541228522912288, 0.049067674327418015, 0.07356456359966743,
      0.0980171403295606, 0.1224106751992162, 0.14673047445536175, 0.17096188876030122,
      0.19509032201612828, 0.2191012401568698, 0.2429801799032639, 0.26671275747489837,
      0.2902846772544624, 0.31368174039889146, 0.33688985339222005, 0.35989503653498817,
      0.3826834323650898};
  return t[o];
}
__device__ __forceinline__ double g32_cos(int o) {
  const double t[17] = {
      1.0, 0.9951847266721969, 0.9807852804032304, 0.9569403357322088,
      0.9238795325112867, 0.881921264348355, 0.8314696123025452, 0.773010453362737,
      0.7071067811865476, 0.6343932841636455, 0.5555702330196022, 0.47139673682599764,
      0.3826834323650898, 0.2902846772544624, 0.19509032201612828, 0.0980171403295606,
      -5.053478687678417e-52};
  return t[o];
}
__device__ __forceinline__ double g32_sin(int o) {
  const double t[17] = {
      0.0, 0.0980171403295606, 0.19509032201612828, 0.2902846772544624,
      0.3826834323650898, 0.47139673682599764, 0.5555702330196022, 0.6343932841636455,
      0.7071067811865476, 0.773010453362737, 0.8314696123025452, 0.881921264348355,
      0.9238795325112867, 0.9569403357322088, 0.9807852804032304, 0.9951847266721969,
      1.0};
  return t[o];
}

// exp(+i pi 5 o / 128) = exp(+i pi o / 128) exp(+i pi o / 32), o = 0..16
__device__ __forceinline__ double g160_cos(int o) {
  const double t[17] = {1.0, 0.99247953459871, 0.970031253194544, 0.932992798834739, 0.881921264348355, 0.8175848131515837, 0.7409511253549591, 0.6531728429537768, 0.5555702330196023, 0.4496113296546066, 0.33688985339222005, 0.21910124015686977, 0.09801714032956077, -0.024541228522912142, -0.14673047445536164, -0.2667127574748983, -0.3826834323650897};
  return t[o];
}
__device__ __forceinline__ double g160_sin(int o) {
  const double t[17] = {0.0, 0.1224106751992162, 0.24298017990326387, 0.3598950365349881, 0.47139673682599764, 0.5758081914178453, 0.6715589548470183, 0.7572088465064845, 0.8314696123025452, 0.8932243011955153, 0.9415440651830208, 0.9757021300385286, 0.9951847266721968, 0.9996988186962042, 0.989176509964781, 0.9637760657954398, 0.9238795325112867};
  return t[o];
}

// v * (c + i s)
__device__ __forceinline__ float2 crot(float2 v, float c, float s) { return cmul_cs(v, c, s); }
__device__ __forceinline__ double2 crot(double2 v, double c, double s) {
  double2 r;
  r.x = v.x * c - v.y * s;
  r.y = v.x * s + v.y * c;
  return r;
}

// v *= exp(-+ 2 pi i t / 32), t compile-time after unrolling
template <typename T, bool INV, typename V>
__device__ __forceinline__ V mul_w32(V v, int t) {
  t &= 31;
  if (t == 0) return v;
  if (t == 8) return mul_mi<INV>(v);                  // -i (forward) / +i (inverse)
  if (t == 16) { v.x = -v.x; v.y = -v.y; return v; }
  if (t == 24) return mul_mi<!INV>(v);
  const T c = (T)w32_cos(t);
  const T s = (T)(INV ? w32_sin(t) : -w32_sin(t));
  return crot(v, c, s);
}

#include "fftmix.cuh"

template <typename T, bool INV, typename V>
__device__ __forceinline__ void fft4_inplace(V& a, V& b, V& c, V& d) {
  V s0 = cadd(a, c), s1 = csub(a, c), s2 = cadd(b, d), s3 = mul_mi<INV>(csub(b, d));
  a = cadd(s0, s2); c = csub(s0, s2); b = cadd(s1, s3); d = csub(s1, s3);
}

// 8-point DFT in place; output index of position p is perm8(p) = (p >> 2) + 2 * (p & 3)
template <typename T, bool INV, typename V>
__device__ __forceinline__ void fft8_perm(V* v) {
#pragma unroll
  for (int n2 = 0; n2 < 4; n2++) {
    V a = v[n2], b = v[n2 + 4];
    v[n2] = cadd(a, b);
    v[n2 + 4] = mul_w32<T, INV>(csub(a, b), 4 * n2);   // W8^{n2}
  }
  fft4_inplace<T, INV>(v[0], v[1], v[2], v[3]);
  fft4_inplace<T, INV>(v[4], v[5], v[6], v[7]);
}

__host__ __device__ constexpr int perm8(int p) { return (p >> 2) + 2 * (p & 3); }
// position P of fft32_perm's output holds frequency index perm32(P)
__host__ __device__ constexpr int perm32(int P) { return (P >> 3) + 4 * perm8(P & 7); }
// position that holds frequency index o
__host__ __device__ constexpr int inv32(int o) { return 8 * (o & 3) + 4 * ((o >> 2) & 1) + (o >> 3); }

// 32-point DFT in place, natural input order, permuted output (see perm32)
template <typename T, bool INV, typename V>
__device__ __forceinline__ void fft32_perm(V* v) {
#pragma unroll
  for (int n2 = 0; n2 < 8; n2++) {
    fft4_inplace<T, INV>(v[n2], v[n2 + 8], v[n2 + 16], v[n2 + 24]);
#pragma unroll
    for (int k1 = 1; k1 < 4; k1++) v[8 * k1 + n2] = mul_w32<T, INV>(v[8 * k1 + n2], n2 * k1);
  }
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) fft8_perm<T, INV>(v + 8 * k1);
}

// 16-point DFT in place (4 x 4); position P of the output holds frequency index perm16(P)
__host__ __device__ constexpr int perm16(int P) { return (P >> 2) + 4 * (P & 3); }
template <typename T, bool INV, typename V>
__device__ __forceinline__ void fft16_perm(V* v) {
#pragma unroll
  for (int n2 = 0; n2 < 4; n2++) {
    fft4_inplace<T, INV>(v[n2], v[n2 + 4], v[n2 + 8], v[n2 + 12]);
#pragma unroll
    for (int k1 = 1; k1 < 4; k1++) v[4 * k1 + n2] = mul_w32<T, INV>(v[4 * k1 + n2], 2 * n2 * k1);   // W16^(n2 k1)
  }
#pragma unroll
  for (int k1 = 0; k1 < 4; k1++) fft4_inplace<T, INV>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
}
// R-point DFT in place for R in {2, 4, 8, 16, 32}; position p of the output holds frequency index perm_r<R>(p)
template <int R> __host__ __device__ constexpr int perm_r(int p) {
  return mix_radix<R>() ? mix_perm<R>(p) : (R == 32) ? perm32(p) : (R == 16) ? perm16(p) : (R == 8) ? perm8(p) : p;
}
template <typename T, bool INV, int R, typename V>
__device__ __forceinline__ void fft_r_perm(V* v) {
  if constexpr (mix_radix<R>()) mix_fft<T, INV, R>(v);
  else if constexpr (R == 32) fft32_perm<T, INV>(v);
  else if constexpr (R == 16) fft16_perm<T, INV>(v);
  else if constexpr (R == 8) fft8_perm<T, INV>(v);
  else if constexpr (R == 4) fft4_inplace<T, INV>(v[0], v[1], v[2], v[3]);
  else { const V a = v[0], b = v[1]; v[0] = cadd(a, b); v[1] = csub(a, b); }
}

// padded complex index inside a per-warp exchange buffer: one pad element per 32
__device__ __forceinline__ int padc(int a) { return a + (a >> 5); }

// Warp-level 1024-point FFT.  In: v[i] = z[lane + 32 i].  Out: v[o] = Z[lane + 32 o].
// buf: this warp's exchange buffer (>= 1056 complex).  tw32: table [i][lane] = exp(-2 pi i * i * lane / 1024).
// f64 only: `wlane` (optional) = tw32[32 + lane] already fetched by the caller, so its latency is not exposed here.
template <typename T, bool INV>
__device__ __forceinline__ void fft1024_warp(typename vec2<T>::type* v, typename vec2<T>::type* buf, int lane,
                                             const typename vec2<T>::type* __restrict__ tw32,
                                             const typename vec2<T>::type* wlane = nullptr) {
  typedef typename vec2<T>::type V;
  fft32_perm<T, INV>(v);
#pragma unroll
  for (int P = 0; P < 32; P++) buf[33 * lane + perm32(P)] = v[P];
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = buf[lane + 33 * i];
  __syncwarp();
  if constexpr (sizeof(T) == 8) {
    // f64: the 31 cross twiddles w^(i lane) are powers of one per-lane value, generated in registers (four independent
    // chains of step w^4: <= 9 roundings, ~1e-15) instead of 31 16-byte table loads per lane -- the 16 KiB table does not
    // stay in the little L1 that is left beside the shared memory and the loads stall on L2 latency
    V b = wlane ? *wlane : __ldg(&tw32[32 + lane]);
    if (INV) b.y = -b.y;
    const V b2 = cmul(b, b), b4 = cmul(b2, b2);
    V w0 = b, w1 = b2, w2 = cmul(b2, b), w3 = b4;            // w^1 .. w^4
#pragma unroll
    for (int i = 1; i < 32; i += 4) {
      v[i] = cmul(v[i], w0);
      if (i + 1 < 32) v[i + 1] = cmul(v[i + 1], w1);
      if (i + 2 < 32) v[i + 2] = cmul(v[i + 2], w2);
      if (i + 3 < 32) v[i + 3] = cmul(v[i + 3], w3);
      if (i + 4 < 32) { w0 = cmul(w0, b4); w1 = cmul(w1, b4); w2 = cmul(w2, b4); w3 = cmul(w3, b4); }
    }
  } else {
#pragma unroll
    for (int i = 1; i < 32; i++) {
      V w = __ldg(&tw32[i * 32 + lane]);
      if (INV) w.y = -w.y;
      v[i] = cmul(v[i], w);
    }
  }
  fft32_perm<T, INV>(v);
  V t[32];
#pragma unroll
  for (int o = 0; o < 32; o++) t[o] = v[inv32(o)];
#pragma unroll
  for (int o = 0; o < 32; o++) v[o] = t[o];
}

// Warp-level FFTs of S = 32 / R1 sequences of M = 32 R1 points each (R1 = 32: fft1024_warp).
// In: v[s R1 + i] = z_s[lane + 32 i].  Out: lane L = s R1 + k1 holds v[o] = Z_s[k1 + R1 o].
// M = 32 R1 with input index l + 32 i and output index k1 + R1 o:  Z[k] = sum_l w_32^(l o) w_M^(l k1) sum_i z[l + 32 i] w_R1^(i k1):
// an R1-point DFT per lane and sequence, the cross twiddles w_M^(l k1) = tw32[l][L] (a 32 x 32 table per frame length), then
// the 32-point DFT over l after the transposing exchange.
template <typename T, bool INV, int R1>
__device__ __forceinline__ void fft_multi_warp(typename vec2<T>::type* v, typename vec2<T>::type* buf, int lane,
                                               const typename vec2<T>::type* __restrict__ tw32,
                                               const typename vec2<T>::type* wlane = nullptr) {
  typedef typename vec2<T>::type V;
  if constexpr (R1 == 32) { fft1024_warp<T, INV>(v, buf, lane, tw32, wlane); return; }
  constexpr int S = 32 / R1;                             // R1 not a power of two: S R1 < 32 registers (then lanes) carry data
#pragma unroll
  for (int s = 0; s < S; s++) fft_r_perm<T, INV, R1>(v + s * R1);
#pragma unroll
  for (int P = 0; P < S * R1; P++) buf[33 * lane + (P / R1) * R1 + perm_r<R1>(P % R1)] = v[P];
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = buf[lane + 33 * i];
  __syncwarp();
  if constexpr (sizeof(T) == 8) {
    V b = wlane ? *wlane : __ldg(&tw32[32 + lane]);
    if (INV) b.y = -b.y;
    const V b2 = cmul(b, b), b4 = cmul(b2, b2);
    V w0 = b, w1 = b2, w2 = cmul(b2, b), w3 = b4;
#pragma unroll
    for (int i = 1; i < 32; i += 4) {
      v[i] = cmul(v[i], w0);
      if (i + 1 < 32) v[i + 1] = cmul(v[i + 1], w1);
      if (i + 2 < 32) v[i + 2] = cmul(v[i + 2], w2);
      if (i + 3 < 32) v[i + 3] = cmul(v[i + 3], w3);
      if (i + 4 < 32) { w0 = cmul(w0, b4); w1 = cmul(w1, b4); w2 = cmul(w2, b4); w3 = cmul(w3, b4); }
    }
  } else {
#pragma unroll
    for (int i = 1; i < 32; i++) {
      V w = __ldg(&tw32[i * 32 + lane]);
      if (INV) w.y = -w.y;
      v[i] = cmul(v[i], w);
    }
  }
  fft32_perm<T, INV>(v);
  V t[32];
#pragma unroll
  for (int o = 0; o < 32; o++) t[o] = v[inv32(o)];
#pragma unroll
  for (int o = 0; o < 32; o++) v[o] = t[o];
}

// Inverse counterpart for the small frames: the mirror image of fft_multi_warp.
// In: lane L = s R1 + k1 holds v[o] = Z_s[k1 + R1 o].  Out: v[s R1 + p] = z_s[lane + 32 perm_r<R1>(p)] (unnormalised inverse).
// z[n2 + 32 n1] = sum_k1 w_R1^(-k1 n1) w_M^(-k1 n2) sum_o Z[k1 + R1 o] w_32^(-o n2): the 32-point DFT over o in registers, the
// cross twiddles conj(tw32[n2][L]), the transposing exchange, then an R1-point DFT per sequence.
template <typename T, int R1>
__device__ __forceinline__ void ifft_multi_warp(typename vec2<T>::type* v, typename vec2<T>::type* buf, int lane,
                                                const typename vec2<T>::type* __restrict__ tw32) {
  typedef typename vec2<T>::type V;
  constexpr int S = 32 / R1;
  fft32_perm<T, true>(v);                                  // position P holds n2 = perm32(P)
  if constexpr (sizeof(T) == 8) {
    V b = __ldg(&tw32[32 + lane]);                          // w_M^(k1); conjugated for the inverse
    b.y = -b.y;
    const V b2 = cmul(b, b), b4 = cmul(b2, b2);
    V w0 = b, w1 = b2, w2 = cmul(b2, b), w3 = b4;            // w^1 .. w^4
#pragma unroll
    for (int i = 1; i < 32; i += 4) {
      v[inv32(i)] = cmul(v[inv32(i)], w0);
      if (i + 1 < 32) v[inv32(i + 1)] = cmul(v[inv32(i + 1)], w1);
      if (i + 2 < 32) v[inv32(i + 2)] = cmul(v[inv32(i + 2)], w2);
      if (i + 3 < 32) v[inv32(i + 3)] = cmul(v[inv32(i + 3)], w3);
      if (i + 4 < 32) { w0 = cmul(w0, b4); w1 = cmul(w1, b4); w2 = cmul(w2, b4); w3 = cmul(w3, b4); }
    }
  } else {
#pragma unroll
    for (int i = 1; i < 32; i++) {
      V w = __ldg(&tw32[i * 32 + lane]);
      w.y = -w.y;
      v[inv32(i)] = cmul(v[inv32(i)], w);
    }
  }
#pragma unroll
  for (int P = 0; P < 32; P++) buf[lane + 33 * perm32(P)] = v[P];
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 32; i++) v[i] = buf[i + 33 * lane];  // new lane = n2, register i = old lane (s, k1)
  __syncwarp();
#pragma unroll
  for (int s = 0; s < S; s++) fft_r_perm<T, true, R1>(v + s * R1);
}
