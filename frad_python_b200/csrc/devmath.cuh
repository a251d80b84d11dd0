// Device helpers shared by the transform kernels: complex arithmetic, small DFTs, PCM sample
// conversion, profile-0 value packing.
#pragma once
#include <cuda_fp16.h>
#include <cstdint>
#include "../../include/fradb200.h"

template <typename T> struct vec2;
template <> struct vec2<float> { typedef float2 type; };
template <> struct vec2<double> { typedef double2 type; };

template <typename T> __device__ __forceinline__ typename vec2<T>::type mk2(T x, T y) {
  typename vec2<T>::type r; r.x = x; r.y = y; return r;
}
template <typename V> __device__ __forceinline__ V cmul(V a, V b) {
  V r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}
template <typename V> __device__ __forceinline__ V cadd(V a, V b) { V r; r.x = a.x + b.x; r.y = a.y + b.y; return r; }
template <typename V> __device__ __forceinline__ V csub(V a, V b) { V r; r.x = a.x - b.x; r.y = a.y - b.y; return r; }
template <typename V> __device__ __forceinline__ V cconj(V a) { a.y = -a.y; return a; }
// multiply by -i (forward) or +i (inverse)
template <bool INV, typename V> __device__ __forceinline__ V mul_mi(V a) {
  V r;
  if (INV) { r.x = -a.y; r.y = a.x; } else { r.x = a.y; r.y = -a.x; }
  return r;
}

// ---- sm_100 packed FP32x2 forms (FADD2 / FMUL2 / FFMA2): one issue slot per complex add, two per
// complex multiply; the (-y, x) swap-negate and scalar broadcasts are operand modifiers in SASS ----
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
  return __ffma2_rn(make_float2(-a.y, a.x), make_float2(w.y, w.y), __fmul2_rn(a, make_float2(w.x, w.x)));
}
// a * w.x + (-a.y, a.x) * w.y with w given as two scalars (compile-time twiddles)
__device__ __forceinline__ float2 cmul_cs(float2 a, float c, float s) {
  return __ffma2_rn(make_float2(-a.y, a.x), make_float2(s, s), __fmul2_rn(a, make_float2(c, c)));
}
__device__ __forceinline__ float2 cscale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
// acc + v * c (c real): one FFMA2 for float2
__device__ __forceinline__ float2 caxpy(float2 acc, float2 v, float c) { return __ffma2_rn(v, make_float2(c, c), acc); }
__device__ __forceinline__ double2 caxpy(double2 acc, double2 v, double c) { acc.x = fma(v.x, c, acc.x); acc.y = fma(v.y, c, acc.y); return acc; }

// ---- small DFTs, natural order in place; forward kernel exp(-2 pi i pq/R), inverse conjugated ----
template <typename T, int R, bool INV> struct small_dft;

template <typename T, bool INV> struct small_dft<T, 2, INV> {
  typedef typename vec2<T>::type V;
  static __device__ __forceinline__ void run(V* v) {
    V a = v[0], b = v[1];
    v[0] = cadd(a, b); v[1] = csub(a, b);
  }
};
template <typename T, bool INV> struct small_dft<T, 4, INV> {
  typedef typename vec2<T>::type V;
  static __device__ __forceinline__ void run(V* v) {
    V s02 = cadd(v[0], v[2]), d02 = csub(v[0], v[2]);
    V s13 = cadd(v[1], v[3]), d13 = mul_mi<INV>(csub(v[1], v[3]));
    v[0] = cadd(s02, s13); v[2] = csub(s02, s13);
    v[1] = cadd(d02, d13); v[3] = csub(d02, d13);
  }
};
// radix 8 = two radix-4 butterflies over the even / odd inputs and one twiddled radix-2 stage (W8^k = e^{-+ 2 pi i k / 8})
template <typename T, bool INV> struct small_dft<T, 8, INV> {
  typedef typename vec2<T>::type V;
  static __device__ __forceinline__ void run(V* v) {
    V e[4] = {v[0], v[2], v[4], v[6]}, o[4] = {v[1], v[3], v[5], v[7]};
    small_dft<T, 4, INV>::run(e);
    small_dft<T, 4, INV>::run(o);
    const T h = (T)0.70710678118654752440;
    // forward: W8 = (1 - i)/sqrt2, W8^2 = -i, W8^3 = (-1 - i)/sqrt2; the inverse conjugates
    const V o1 = INV ? mk2<T>((o[1].x - o[1].y) * h, (o[1].x + o[1].y) * h) : mk2<T>((o[1].x + o[1].y) * h, (o[1].y - o[1].x) * h);
    const V o2 = mul_mi<INV>(o[2]);
    const V o3 = INV ? mk2<T>(-(o[3].x + o[3].y) * h, (o[3].x - o[3].y) * h) : mk2<T>((o[3].y - o[3].x) * h, -(o[3].x + o[3].y) * h);
    v[0] = cadd(e[0], o[0]); v[4] = csub(e[0], o[0]);
    v[1] = cadd(e[1], o1);   v[5] = csub(e[1], o1);
    v[2] = cadd(e[2], o2);   v[6] = csub(e[2], o2);
    v[3] = cadd(e[3], o3);   v[7] = csub(e[3], o3);
  }
};
template <int R> struct odd_roots;
template <> struct odd_roots<3> {
  static __device__ __forceinline__ double c(int i) { const double t[3] = {1.0, -0.5, -0.5}; return t[i]; }
  static __device__ __forceinline__ double s(int i) { const double t[3] = {0.0, 0.8660254037844386, -0.8660254037844386}; return t[i]; }
};
template <> struct odd_roots<5> {
  static __device__ __forceinline__ double c(int i) {
    const double t[5] = {1.0, 0.30901699437494745, -0.8090169943749475, -0.8090169943749475, 0.30901699437494745}; return t[i]; }
  static __device__ __forceinline__ double s(int i) {
    const double t[5] = {0.0, 0.9510565162951535, 0.5877852522924731, -0.5877852522924731, -0.9510565162951535}; return t[i]; }
};
template <> struct odd_roots<7> {
  static __device__ __forceinline__ double c(int i) {
    const double t[7] = {1.0, 0.6234898018587336, -0.2225209339563144, -0.9009688679024191,
                         -0.9009688679024191, -0.2225209339563144, 0.6234898018587336}; return t[i]; }
  static __device__ __forceinline__ double s(int i) {
    const double t[7] = {0.0, 0.7818314824680298, 0.9749279121818236, 0.4338837391175581,
                         -0.4338837391175581, -0.9749279121818236, -0.7818314824680298}; return t[i]; }
};
// odd radix: y[p] = a_p -+ i b_p with a_p = v0 + sum_q cos(2 pi pq/R)(v_q + v_{R-q}), b_p = sum_q sin(2 pi pq/R)(v_q - v_{R-q})
template <typename T, int R, bool INV> struct small_dft {
  typedef typename vec2<T>::type V;
  static __device__ __forceinline__ void run(V* v) {
    constexpr int H = (R - 1) / 2;
    V sp[H], sm[H];
#pragma unroll
    for (int q = 1; q <= H; q++) { sp[q - 1] = cadd(v[q], v[R - q]); sm[q - 1] = csub(v[q], v[R - q]); }
    V y0 = v[0];
#pragma unroll
    for (int q = 0; q < H; q++) y0 = cadd(y0, sp[q]);
    V out[R];
    out[0] = y0;
#pragma unroll
    for (int p = 1; p <= H; p++) {
      V a = v[0];
      V b = mk2<T>(T(0), T(0));
#pragma unroll
      for (int q = 1; q <= H; q++) {
        const T cc = (T)odd_roots<R>::c((p * q) % R);
        const T ss = (T)odd_roots<R>::s((p * q) % R);
        a = caxpy(a, sp[q - 1], cc);
        b = caxpy(b, sm[q - 1], ss);
      }
      // forward: y[p] = a - i b ; y[R-p] = a + i b   (inverse swaps)
      V mib = mul_mi<INV>(b);
      out[p] = cadd(a, mib);
      out[R - p] = csub(a, mib);
    }
#pragma unroll
    for (int i = 0; i < R; i++) v[i] = out[i];
  }
};

// ---- PCM sample formats (backend/pcmformat.py:34-62) ------------------------------------------
template <typename T> __device__ __forceinline__ T load_pcm(const void* p, int dt, int64_t i) {
  switch (dt) {
    case FRAD_F32: return (T)((const float*)p)[i];
    case FRAD_F64: return (T)((const double*)p)[i];
    case FRAD_F16: return (T)__half2float(((const __half*)p)[i]);
    case FRAD_S8:  return (T)((double)((const int8_t*)p)[i] / 128.0);
    case FRAD_S16: return (T)((double)((const int16_t*)p)[i] / 32768.0);
    case FRAD_S32: return (T)((double)((const int32_t*)p)[i] / 2147483648.0);
    case FRAD_S64: return (T)((double)((const long long*)p)[i] / 9223372036854775808.0);
    case FRAD_U8:  return (T)((double)((const uint8_t*)p)[i] / 128.0 - 1.0);
    case FRAD_U16: return (T)((double)((const uint16_t*)p)[i] / 32768.0 - 1.0);
    case FRAD_U32: return (T)((double)((const uint32_t*)p)[i] / 2147483648.0 - 1.0);
    case FRAD_U64: return (T)((double)((const unsigned long long*)p)[i] / 9223372036854775808.0 - 1.0);
    default: return T(0);
  }
}

// f64 -> sample format, what `from_f64(pcm, fmt).astype(fmt)` yields (pcmformat.py:49-62, src/decoder.py:23): floats
// round to nearest even; integers scale by 2^(bits-1) (unsigned: after + 1) and TRUNCATE towards zero.  Out-of-range
// values follow NumPy on x86-64 (a C cast: cvttsd2si gives the "integer indefinite" 0x80..0, then the low bytes are
// kept), so e.g. +1.0 wraps to the most negative 16-bit sample exactly like the reference.
__device__ __forceinline__ int32_t cvtt32_x86(double v) {
  return (v > -2147483649.0 && v < 2147483648.0) ? __double2int_rz(v) : (int32_t)0x80000000;
}
__device__ __forceinline__ long long cvtt64_x86(double v) {
  return (v >= -9223372036854775808.0 && v < 9223372036854775808.0) ? __double2ll_rz(v) : (long long)0x8000000000000000ull;
}
// little-endian image of one sample in the low frad_dtype_size(dt) bytes
__device__ __forceinline__ uint64_t pcm_image(double v, int dt) {
  switch (dt) {
    case FRAD_F64: return (uint64_t)__double_as_longlong(v);
    case FRAD_F32: return (uint64_t)__float_as_uint(__double2float_rn(v));
    case FRAD_F16: return (uint64_t)__half_as_ushort(__double2half(v));
    case FRAD_S8:  return (uint64_t)(uint8_t)cvtt32_x86(v * 128.0);
    case FRAD_S16: return (uint64_t)(uint16_t)cvtt32_x86(v * 32768.0);
    case FRAD_S32: return (uint64_t)(uint32_t)cvtt32_x86(v * 2147483648.0);
    case FRAD_S64: return (uint64_t)cvtt64_x86(v * 9223372036854775808.0);
    case FRAD_U8:  return (uint64_t)(uint8_t)cvtt32_x86((v + 1.0) * 128.0);
    case FRAD_U16: return (uint64_t)(uint16_t)cvtt32_x86((v + 1.0) * 32768.0);
    case FRAD_U32: return (uint64_t)(uint32_t)cvtt64_x86((v + 1.0) * 2147483648.0);
    case FRAD_U64: {
      const double x = (v + 1.0) * 9223372036854775808.0;
      // C cast double -> uint64 as gcc emits it: values >= 2^63 go through (x - 2^63) with the top bit set afterwards
      return (x >= 9223372036854775808.0) ? ((uint64_t)cvtt64_x86(x - 9223372036854775808.0) ^ 0x8000000000000000ull)
                                          : (uint64_t)cvtt64_x86(x);
    }
    default: return 0;
  }
}
__device__ __forceinline__ int pcm_size(int dt) {
  switch (dt) {
    case FRAD_F64: case FRAD_S64: case FRAD_U64: return 8;
    case FRAD_F32: case FRAD_S32: case FRAD_U32: return 4;
    case FRAD_F16: case FRAD_S16: case FRAD_U16: return 2;
    default: return 1;
  }
}

template <typename T> __device__ __forceinline__ void store_pcm(void* p, int dt, int64_t i, T v) {
  const uint64_t u = pcm_image((double)v, dt);
  switch (pcm_size(dt)) {
    case 8: ((uint64_t*)p)[i] = u; break;
    case 4: ((uint32_t*)p)[i] = (uint32_t)u; break;
    case 2: ((uint16_t*)p)[i] = (uint16_t)u; break;
    default: ((uint8_t*)p)[i] = (uint8_t)u; break;
  }
}

// ---- profile-0 value <-> bytes (profile0.py:28-42, 51-66) -------------------------------------
__device__ __forceinline__ uint16_t f2h_bits(float x) { return __half_as_ushort(__float2half_rn(x)); }
__device__ __forceinline__ uint16_t f2h_bits(double x) { return __half_as_ushort(__double2half(x)); }
__device__ __forceinline__ uint32_t f2f_bits(float x) { return __float_as_uint(x); }
__device__ __forceinline__ uint32_t f2f_bits(double x) { return __float_as_uint(__double2float_rn(x)); }
__device__ __forceinline__ uint64_t f2d_bits(float x) { return (uint64_t)__double_as_longlong((double)x); }
__device__ __forceinline__ uint64_t f2d_bits(double x) { return (uint64_t)__double_as_longlong(x); }

// writes the `bits/8`-byte image of value x at dst (bits in {16,24,32,48,64})
template <typename T> __device__ __forceinline__ void p0_store_value(uint8_t* dst, T x, int bits, bool le) {
  if (bits == 32 || bits == 24) {
    uint32_t u = f2f_bits(x);
    if (bits == 32) {
      if (!le) u = __byte_perm(u, 0, 0x0123);
      if ((((uintptr_t)dst) & 3) == 0) *(uint32_t*)dst = u;
      else { dst[0] = u; dst[1] = u >> 8; dst[2] = u >> 16; dst[3] = u >> 24; }
    } else {
      if (!le) { dst[0] = u >> 24; dst[1] = u >> 16; dst[2] = u >> 8; }   // top three bytes, MSB first
      else { dst[0] = u >> 8; dst[1] = u >> 16; dst[2] = u >> 24; }       // drop the LSB, keep LE order
    }
  } else if (bits == 64 || bits == 48) {
    uint64_t u = f2d_bits(x);
    if (bits == 64) {
      if (!le) {
        uint32_t lo = (uint32_t)u, hi = (uint32_t)(u >> 32);
        u = ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
      }
      if ((((uintptr_t)dst) & 7) == 0) *(uint64_t*)dst = u;
      else { for (int i = 0; i < 8; i++) dst[i] = (uint8_t)(u >> (8 * i)); }
    } else {
      if (!le) { for (int i = 0; i < 6; i++) dst[i] = (uint8_t)(u >> (56 - 8 * i)); }
      else { for (int i = 0; i < 6; i++) dst[i] = (uint8_t)(u >> (16 + 8 * i)); }
    }
  } else {  // 16
    uint16_t h = f2h_bits(x);
    if (!le) h = (uint16_t)((h >> 8) | (h << 8));
    if ((((uintptr_t)dst) & 1) == 0) *(uint16_t*)dst = h;
    else { dst[0] = (uint8_t)h; dst[1] = (uint8_t)(h >> 8); }
  }
}

// reads one value; NaN / Inf become 0 (profile0.py:66)
__device__ __forceinline__ double p0_load_value(const uint8_t* src, int bits, bool le) {
  double v;
  if (bits == 32 || bits == 24) {
    uint32_t u;
    if (bits == 32) {
      u = (uint32_t)src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16) | ((uint32_t)src[3] << 24);
      if (!le) u = __byte_perm(u, 0, 0x0123);
    } else {
      if (!le) u = ((uint32_t)src[0] << 24) | ((uint32_t)src[1] << 16) | ((uint32_t)src[2] << 8);
      else u = ((uint32_t)src[0] << 8) | ((uint32_t)src[1] << 16) | ((uint32_t)src[2] << 24);
    }
    v = (double)__uint_as_float(u);
  } else if (bits == 64 || bits == 48) {
    uint64_t u = 0;
    if (bits == 64) {
      if (!le) { for (int i = 0; i < 8; i++) u = (u << 8) | src[i]; }
      else { for (int i = 0; i < 8; i++) u |= (uint64_t)src[i] << (8 * i); }
    } else {
      if (!le) { for (int i = 0; i < 6; i++) u |= (uint64_t)src[i] << (56 - 8 * i); }
      else { for (int i = 0; i < 6; i++) u |= (uint64_t)src[i] << (16 + 8 * i); }
    }
    v = __longlong_as_double((long long)u);
  } else {
    uint16_t h = le ? (uint16_t)(src[0] | (src[1] << 8)) : (uint16_t)((src[0] << 8) | src[1]);
    v = (double)__half2float(__ushort_as_half(h));
  }
  return isfinite(v) ? v : 0.0;
}
