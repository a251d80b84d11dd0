// Register-resident DFTs of the lengths R = 3, 5, 6, 7, 10, 12, 14, 20, 24, 28: the first pass of the warp FFTs for the frame
// lengths N = 64 R of the 160 / 192 / 224 families (fourier/profiles.py:14-23; N = 192 ... 1792).  R = RA x RB with RA the odd
// prime and RB in {1, 2, 4, 8}: RA-point DFTs over the stride-RB subsequences, the twiddles W_R^(n2 k1), then RB-point DFTs over
// each block (small_dft of devmath.cuh, natural order).  Position p of the result holds frequency index mix_perm<R>(p).
// All indices are compile-time constants after unrolling; the twiddles fold to immediates (fftmix_tables.cuh).
// Included by fft32.cuh (needs its crot).
#pragma once
#include "fftmix_tables.cuh"

template <int R> struct mixfac {
  static constexpr int RA = (R % 3 == 0) ? 3 : (R % 5 == 0) ? 5 : (R % 7 == 0) ? 7 : R;
  static constexpr int RB = R / RA;
};
template <int R> __host__ __device__ constexpr bool mix_radix() { return R == 3 || R == 5 || R == 6 || R == 7 || R == 10 || R == 12 || R == 14 || R == 20 || R == 24 || R == 28; }
template <int R> __host__ __device__ constexpr int mix_perm(int p) { return (p / mixfac<R>::RB) + mixfac<R>::RA * (p % mixfac<R>::RB); }

template <typename T, bool INV, int R, typename V>
__device__ __forceinline__ void mix_fft(V* v) {
  constexpr int RA = mixfac<R>::RA, RB = mixfac<R>::RB;
  if constexpr (RB == 1) {
    small_dft<T, R, INV>::run(v);
  } else {
#pragma unroll
    for (int n2 = 0; n2 < RB; n2++) {
      V t[RA];
#pragma unroll
      for (int n1 = 0; n1 < RA; n1++) t[n1] = v[n2 + RB * n1];
      small_dft<T, RA, INV>::run(t);
#pragma unroll
      for (int k1 = 0; k1 < RA; k1++) {
        const int tw = (n2 * k1) % R;
        V y = t[k1];
        if (tw != 0) {
          if (4 * tw == R) y = mul_mi<INV>(y);
          else if (2 * tw == R) { y.x = -y.x; y.y = -y.y; }
          else if (4 * tw == 3 * R) y = mul_mi<!INV>(y);
          else y = crot(y, (T)mixw_cos(R, tw), (T)(INV ? mixw_sin(R, tw) : -mixw_sin(R, tw)));
        }
        v[n2 + RB * k1] = y;
      }
    }
#pragma unroll
    for (int k1 = 0; k1 < RA; k1++) small_dft<T, RB, INV>::run(v + RB * k1);
  }
}
