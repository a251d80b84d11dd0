// Fast path for the default frame length N = 2048 (src/tools/cli.py:26, every BASELINE config).
//
// One warp per (frame, channel) sequence: the 1024-point complex FFT behind the DCT runs entirely
// in registers as 32 x 32 with a single padded shared-memory exchange (fft32.cuh), so the only
// block-wide barriers are the two around the cooperative, fully coalesced 16-byte global loads /
// stores of the interleaved [n][c] / [k][c] layouts.  The profile pro/epilogues run per warp on the
// sequence while it sits in shared memory:
//   decode  payload / quantised ints -> dequantise -> DCT-III -> PCM (f64 or f32 out)        (this file)
//   encode  PCM f32 -> DCT-II -> profile 0 payload / profile 1 quantised ints               (enc2048.cu)
// The real-FFT split and the DCT twiddles are applied pairwise (k, M-k): one pair yields
// X[k], X[N-k], X[M-k], X[M+k] from two complex multiplies (see DESIGN.md "Transform").
// Reference lines: fourier/profile0.py:21-30,63-69; fourier/profile1.py:21-40,65-77;
// fourier/tools/p1tools.py:18-44.
#include <cstdio>
#include <cstdlib>
#include "tile2048.cuh"

namespace {


// Decoder tails.  PAIR (stereo tiling, CT == 2): instead of a band index per bin the CTA keeps one word per pair of
// neighbouring bins (2p, 2p + 1) -- j0 | band(2p) << 16 | band(2p + 1) << 24 with j0 = 2p - lo[band(2p)] -- because the
// bins a thread converts are the same for every tile: one conflict-free 4-byte load replaces the bidx / lo look-ups.
template <typename T, bool PAIR> struct dec_tail {    // f32 flavour
  int lo[FRAD_NBANDS + 1];
  uint8_t bidx[PAIR ? 16 : FN];
  uint32_t ptab[PAIR ? FM : 4];
  float4 bp[FW][FRAD_NBANDS + 1];                     // t0 / n, t1 / n, lo, hi
  float aux[DQ_TAB];                                  // |q|^(4/3) / scale for |q| < DQ_TAB
};
template <bool PAIR> struct dec_tail<double, PAIR> {  // f64 flavour: exact np.linspace ramp
  int lo[FRAD_NBANDS + 1];
  uint8_t bidx[PAIR ? 16 : FN];
  uint32_t ptab[PAIR ? FM : 4];
  double2 rp[FW][FRAD_NBANDS + 1];                    // start, step
  double dq[DQ_SMEM64];                               // |q|^(4/3) for the most frequent small |q|
};

// band index table shared by all warps of the CTA (same N and sample rate)
template <typename TAIL>
__device__ __forceinline__ void build_bidx(TAIL* tail, const band_table* bt, int tid) {
  if (tid <= FRAD_NBANDS) { const int e = bt->edge[tid]; tail->lo[tid] = (e < FN) ? e : FN; }
  __syncthreads();
  for (int k = tid; k < FN; k += FW * 32) {
    int b = 0;
#pragma unroll 1
    for (int s = 16; s > 0; s >>= 1) { if (b + s <= FRAD_NBANDS - 1 && tail->lo[b + s] <= k) b += s; }
    tail->bidx[k] = (uint8_t)b;                       // largest b <= 26 with lo[b] <= k
  }
}

// pair table of the PAIR tails (see dec_tail)
template <typename TAIL>
__device__ __forceinline__ void build_ptab(TAIL* tail, const band_table* bt, int tid) {
  if (tid <= FRAD_NBANDS) { const int e = bt->edge[tid]; tail->lo[tid] = (e < FN) ? e : FN; }
  __syncthreads();
  for (int p = tid; p < FM; p += FW * 32) {
    int bb[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int k = 2 * p + h;
      int b = 0;
#pragma unroll 1
      for (int s = 16; s > 0; s >>= 1) { if (b + s <= FRAD_NBANDS - 1 && tail->lo[b + s] <= k) b += s; }
      bb[h] = b;
    }
    tail->ptab[p] = (uint32_t)(2 * p - tail->lo[bb[0]]) | ((uint32_t)bb[0] << 16) | ((uint32_t)bb[1] << 24);
  }
}
// ------------------------------------------------------------------------------------------------
// decode
// ------------------------------------------------------------------------------------------------
// rare large magnitudes: kept out of line so the hot loop stays small
__device__ __noinline__ double dequant_big(const double* __restrict__ dq64, int qa) {
  return (qa < DQ_TAB64) ? __ldg(dq64 + qa) : pow((double)qa, 1.0 / 0.75);
}
__device__ __noinline__ float dequant_big_f32(int qa, float inv_scale) {
  const float af = (float)qa;
  return af * cbrtf(af) * inv_scale;
}

template <typename T> struct dec_layout {
  static constexpr int SEQ_BYTES = SEQ_CPX * 2 * (int)sizeof(T);
  static constexpr int WARP_BYTES = SEQ_BYTES;
};

// Store phase for the integer / f16 sample formats of pcmformat.py:49-62 (converted in registers: 4 samples per thread
// and store).  Out of line so that its 64-bit conversion code does not set the register budget of the kernel's main loop.
template <typename T, int CT>
__device__ __noinline__ void store_tile_formats(const T* seq_all, void* pcm, int dt, int C, int nf, int b0, int grp, int words_per_frame,
                                                int tid) {
  constexpr int STRIDE_T = dec_layout<T>::WARP_BYTES / (int)sizeof(T);
  for (int f = 0; f < nf; f++) {
    const size_t fbase = (size_t)(b0 + f) * words_per_frame;
    for (int e4 = tid; e4 < vec_count<CT>(C); e4 += FW * 32) {
      const size_t woff = fbase + (size_t)vec_word<CT>(e4, C, grp);
      uint64_t o4[4];
      gather4<CT, STRIDE_T>(seq_all, C, f, e4, [](int n) { return swzT<T>(n); }, [dt](T x) { return pcm_image((double)x, dt); }, o4);
      switch (pcm_size(dt)) {
        case 8: {
          uint4* dst = (uint4*)((uint64_t*)pcm + woff);
          dst[0] = make_uint4((uint32_t)o4[0], (uint32_t)(o4[0] >> 32), (uint32_t)o4[1], (uint32_t)(o4[1] >> 32));
          dst[1] = make_uint4((uint32_t)o4[2], (uint32_t)(o4[2] >> 32), (uint32_t)o4[3], (uint32_t)(o4[3] >> 32));
          break;
        }
        case 4: *(uint4*)((uint32_t*)pcm + woff) = make_uint4((uint32_t)o4[0], (uint32_t)o4[1], (uint32_t)o4[2], (uint32_t)o4[3]); break;
        case 2: *(uint2*)((uint16_t*)pcm + woff) = make_uint2((uint32_t)o4[0] | ((uint32_t)o4[1] << 16), (uint32_t)o4[2] | ((uint32_t)o4[3] << 16)); break;
        default: *(uint32_t*)((uint8_t*)pcm + woff) = (uint32_t)o4[0] | ((uint32_t)o4[1] << 8) | ((uint32_t)o4[2] << 16) | ((uint32_t)o4[3] << 24); break;
      }
    }
  }
}

// Second pass of the pair staging: magnitudes outside the shared-memory dequantiser table.  `mask` holds the steps i of
// this thread that met one; their vectors are read again (L2 hits) and only the large values are redone -- global table
// up to DQ_TAB64, pow() beyond (profile1.py:65) -- with the same ramp arithmetic as the first pass.
template <typename T, typename TAIL>
__device__ __noinline__ void stage_pairs_fix(uint32_t mask, const uint32_t* f0, const uint32_t* f1, int nf, const TAIL* tail, T* seq_base,
                                             int stride_t, const double* __restrict__ dq64, double inv_scale, int tid) {
  constexpr uint32_t LIM = (sizeof(T) == 8) ? DQ_SMEM64 : DQ_TAB;
  for (; mask; mask &= mask - 1) {
    const int i = __ffs(mask) - 1;
    const uint32_t pw = tail->ptab[tid + FW * 32 * i];
    const int bnd0 = (pw >> 16) & 0xff, bnd1 = pw >> 24;
    const int k0 = 2 * (tid + FW * 32 * i);
    for (int f = 0; f < nf; f++) {
      const uint4 xv = __ldg((const uint4*)((f ? f1 : f0) + 4 * FW * 32 * i));
      const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w};
      for (int j = 0; j < 4; j++) {
        const int qi = (int)xs[j];
        const uint32_t qa = (qi < 0) ? 0u - (uint32_t)qi : (uint32_t)qi;
        if (qa < LIM) continue;
        const int h = j >> 1, sq = 2 * f + (j & 1), k = k0 + h;
        const int bnd = h ? bnd1 : bnd0;
        T val;
        if constexpr (sizeof(T) == 8) {
          const int jj = (h && bnd1 != bnd0) ? 0 : (int)(pw & 0xffffu) + h;
          const double mag = ((qa < (uint32_t)DQ_TAB64) ? __ldg(dq64 + qa) : pow((double)qa, 1.0 / 0.75)) * inv_scale;
          const double2 rp = tail->rp[sq][bnd];
          const double div = __dadd_rn(__dmul_rn((double)jj, rp.y), rp.x);
          val = ((qi < 0) ? -mag : mag) * div;
        } else {
          const float af = (float)qa;
          const float mag = af * cbrtf(af) * (float)inv_scale;
          const float4 bpv = tail->bp[sq][bnd];
          const float kf = (float)k;
          const float dv = fmaf(kf - bpv.z, bpv.y, (bpv.w - kf) * bpv.x);
          val = ((qi < 0) ? -mag : mag) * dv;
        }
        seq_base[(size_t)sq * stride_t + k] = val;
      }
    }
  }
}

// 16-byte shared-memory load that only happens where p is set (v keeps its value elsewhere): a predicated LDS, never a branch
__device__ __forceinline__ void lds128_if(bool p, uint32_t saddr, uint4& v) {
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %4, 0;\n\t@q ld.shared.v4.u32 {%0, %1, %2, %3}, [%5];\n\t}"
               : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w) : "r"((uint32_t)p), "r"(saddr));
}
__device__ __forceinline__ uint4 lds128(uint32_t saddr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr));
  return v;
}

template <typename T, int PROFILE, int CT>
__global__ void __launch_bounds__(FW * 32, (sizeof(T) == 8) ? 3 : FRAD_MINB)
dec2048_kernel(dec_params P, plan_dev pl, const typename vec2<T>::type* __restrict__ tw32, const double* __restrict__ dq64) {
  typedef typename vec2<T>::type V;
  typedef dec_layout<T> LY;
  constexpr bool PAIR = (PROFILE == 1 && CT == 2);
  typedef dec_tail<T, PAIR> TAIL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TAIL* tail = (TAIL*)(smem_raw + (size_t)FW * LY::WARP_BYTES);
  const int C = (CT == 1 || CT == 2) ? CT : P.C;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const tile_map tm = make_tiles<CT>(P.B, C);
  const V* ci = (const V*)pl.ci;
  const V* ei = (const V*)pl.ei;
  const double inv_scale = 1.0 / P.scale;
  if (PROFILE != 0) {
    if constexpr (PAIR) build_ptab(tail, P.bt, tid); else build_bidx(tail, P.bt, tid);
    if constexpr (sizeof(T) == 4) {
      for (int i = tid; i < DQ_TAB; i += FW * 32) { const float a = (float)i; tail->aux[i] = a * cbrtf(a) * (float)inv_scale; }
    } else {
      for (int i = tid; i < DQ_SMEM64; i += FW * 32) tail->dq[i] = __ldg(dq64 + i) * inv_scale;
    }
  }
  const int words_per_frame = FN * C;

  // ---- per-band ramp parameters of this warp's sequence from its threshold integers (profile1.py:66, p1tools.py:35-41).
  //      They are only read while a tile is staged, so the warp prepares those of its NEXT tile while it transforms the
  //      current one: the threshold integers are fetched before the FFT, the table is written after it ----
  auto fetch_tq = [&](int t) -> int {
    if (PROFILE == 0 || t >= tm.ntiles) return 0;
    const int tb0 = (CT == 4) ? t / tm.groups : t * tm.fpc;
    const int tgrp = (CT == 4) ? t - tb0 * tm.groups : 0;
    const int tnseq = (CT == 4) ? 4 : min(tm.fpc, P.B - tb0) * C;
    if (w >= tnseq || lane >= FRAD_NBANDS) return 0;
    const int f = (CT == 4) ? 0 : w / C, c = (CT == 4) ? 4 * tgrp + w : w - f * C;
    return __ldg(P.tq + ((size_t)(tb0 + f) * FRAD_NBANDS + lane) * C + c);
  };
  auto put_ramps = [&](int tqv) {
    if (PROFILE == 0) return;
    double thr = 0.0;
    if (lane < FRAD_NBANDS) thr = decode_threshold(P.bt, tqv);                                      // profile1.py:66
    const double thr_next = __shfl_down_sync(0xffffffffu, thr, 1);
    if (lane <= FRAD_NBANDS) {
      const int n = (lane < FRAD_NBANDS - 1) ? tail->lo[lane + 1] - tail->lo[lane] : 0;
      if constexpr (sizeof(T) == 8) {
        tail->rp[w][lane] = (n > 0) ? make_double2(thr, (thr_next - thr) / (double)n) : make_double2(0.0, 0.0);
      } else {
        tail->bp[w][lane] = (n > 0) ? make_float4((float)(thr / n), (float)(thr_next / n), (float)tail->lo[lane],
                                                  (float)tail->lo[lane + 1])
                                    : make_float4(0.0f, 0.0f, 0.0f, 1.0f);
      }
    }
  };
  if (PROFILE != 0) {
    __syncthreads();                                        // lo[] of the tail is complete
    put_ramps(fetch_tq(blockIdx.x));
  }

  // the tile's 32-bit words travel through registers: those of tile t + gridDim.x are requested before tile t is stored
  uint4 x[tile_regs<CT>::NX];
  const bool p0_any = (PROFILE == 0) && P.bits != 32;      // profile 0 at the other depths: values unpacked from bytes
  auto request_tile = [&](int t) {
    if (CT == 0 || PROFILE == 2 || p0_any) return;
    if (t >= tm.ntiles) t = tm.ntiles - 1;      // always overwrite x (no stale values kept alive across the transform): re-read the last tile
    const int tb0 = (CT == 4) ? t / tm.groups : t * tm.fpc;
    const int tgrp = (CT == 4) ? t - tb0 * tm.groups : 0;
    load_tile<CT>(
        [&](int f) {
          return (PROFILE == 0) ? (const uint32_t*)(P.payload + (size_t)(tb0 + f) * P.payload_stride)
                                : (const uint32_t*)(P.q + (size_t)(tb0 + f) * words_per_frame);
        },
        C, min(tm.fpc, P.B - tb0), tgrp, tid, x);
  };
  // Measured on B200 (48 kHz stereo, 15 000 frames): requesting early gains 7 % on the profile-1 stereo kernels, whose
  // staging is short; the other tilings lose 3-5 % (their stores queue behind the loads), so they request at the top
  constexpr bool EARLY = PAIR;
  if (EARLY) request_tile(blockIdx.x);

  for (int tile = blockIdx.x; tile < tm.ntiles; tile += gridDim.x) {
    const int b0 = (CT == 4) ? tile / tm.groups : tile * tm.fpc;
    const int grp = (CT == 4) ? tile - b0 * tm.groups : 0;
    const int nf = min(tm.fpc, P.B - b0);
    const int nseq = (CT == 4) ? 4 : nf * C;
    __syncthreads();                                        // previous tile stored, ramp tables visible
    if (!EARLY) request_tile(tile);
    if constexpr (PROFILE == 2) {
      // ---- profile 2: the spectrum arrives dequantised and TNS-synthesised (tns_synth_kernel) as f64, four sequences per sector;
      //      what is left of profile2.py:85-91 is the ramp multiplication ----
      constexpr int STRIDE_T = LY::WARP_BYTES / (int)sizeof(T);
      T* seq_base = (T*)smem_raw;
      const int Cg = (CT == 4) ? 4 : C, c0 = (CT == 4) ? 4 * grp : 0;
      auto ramp_mul = [&](double v, int sq, int k) -> T {
        const int bnd = tail->bidx[k];
        if constexpr (sizeof(T) == 8) {
          const double2 rp = tail->rp[sq][bnd];
          return v * __dadd_rn(__dmul_rn((double)(k - tail->lo[bnd]), rp.y), rp.x);        // j * step + start
        } else {
          const float4 bpv = tail->bp[sq][bnd];
          const float kf = (float)k;
          return (float)v * fmaf(kf - bpv.z, bpv.y, (bpv.w - kf) * bpv.x);
        }
      };
      if constexpr (CT == 1 || CT == 2 || CT == 4) {
        // the tile's four sequences are one group of tns_synth_kernel's layout (their first index is a multiple of four):
        // 64 KiB in one run, a thread takes two sequences of a bin with one 16-byte load
        const int s4 = (CT == 4) ? b0 * C + 4 * grp : b0 * C;
        const double2* base = (const double2*)(P.spec + (size_t)(s4 >> 2) * FN * 4);
#pragma unroll 4
        for (int idx = tid; idx < FN * 2; idx += FW * 32) {
          const int k = idx >> 1, h = idx & 1;
          const double2 v2 = __ldg(base + idx);
          // both sequences share the bin: one band look-up, both products formed before either is stored
          const int bnd = tail->bidx[k];
          T r0, r1;
          if constexpr (sizeof(T) == 8) {
            const double jd = (double)(k - tail->lo[bnd]);
            const double2 p0 = tail->rp[2 * h][bnd], p1 = tail->rp[2 * h + 1][bnd];
            r0 = v2.x * __dadd_rn(__dmul_rn(jd, p0.y), p0.x);                              // j * step + start
            r1 = v2.y * __dadd_rn(__dmul_rn(jd, p1.y), p1.x);
          } else {
            const float4 b0v = tail->bp[2 * h][bnd], b1v = tail->bp[2 * h + 1][bnd];
            const float kf = (float)k;
            r0 = (float)v2.x * fmaf(kf - b0v.z, b0v.y, (b0v.w - kf) * b0v.x);
            r1 = (float)v2.y * fmaf(kf - b1v.z, b1v.y, (b1v.w - kf) * b1v.x);
          }
          seq_base[(size_t)(2 * h) * STRIDE_T + k] = r0;
          seq_base[(size_t)(2 * h + 1) * STRIDE_T + k] = r1;
        }
      } else {
      for (int f = 0; f < nf; f++) {
        const int s_first = (b0 + f) * C + c0;                // layout of tns_synth_kernel: spec[((s >> 2) N + k) 4 + (s & 3)]
#pragma unroll 4
        for (int idx = tid; idx < FN * Cg; idx += FW * 32) {
          const int k = idx / C, cc = idx - k * C;
          const int sg = s_first + cc;
          const double v = __ldg(P.spec + ((size_t)(sg >> 2) * FN + k) * 4 + (sg & 3));
          const int sq = f * C + cc;
          seq_base[(size_t)sq * STRIDE_T + k] = ramp_mul(v, sq, k);
        }
      }
      }
    } else if constexpr (PAIR) {
      // ---- stereo profile-1 staging: thread `tid` converts the bin pairs p = tid + 128 i of all four sequences ----
      constexpr int VPF = FN * 2 / 4 / (FW * 32);
      constexpr int STRIDE_T = LY::WARP_BYTES / (int)sizeof(T);
      T* seq_base = (T*)smem_raw;
      const uint32_t* f0 = (const uint32_t*)(P.q + (size_t)b0 * words_per_frame) + 4 * tid;
      const uint32_t* f1 = (nf > 1) ? f0 + words_per_frame : f0;
      uint32_t bigmask = 0;
      uint32_t ramp_s;
      if constexpr (sizeof(T) == 8) ramp_s = (uint32_t)__cvta_generic_to_shared(&tail->rp[0][0]);
      else ramp_s = (uint32_t)__cvta_generic_to_shared(&tail->bp[0][0]);
      constexpr uint32_t LIM = (sizeof(T) == 8) ? DQ_SMEM64 : DQ_TAB;
#pragma unroll
      for (int i = 0; i < VPF; i++) {
        const uint32_t pw = tail->ptab[tid + FW * 32 * i];
        const uint32_t off0 = (pw >> 12) & 0x1f0u, off1 = (pw >> 20) & 0x1f0u;      // band * 16 bytes
        const bool same = (off0 == off1);
        const int k0 = 2 * (tid + FW * 32 * i);
        double jd0 = 0.0, jd1 = 0.0;
        float kf0 = 0.0f, kf1 = 0.0f;
        if constexpr (sizeof(T) == 8) {
          jd0 = __hiloint2double(0x43300000, (int)(pw & 0xffffu)) - 4503599627370496.0;   // (double)j0, j0 < 2^16
          jd1 = same ? jd0 + 1.0 : 0.0;                                                    // a new band starts at j = 0
        } else {
          kf0 = (float)k0;
          kf1 = kf0 + 1.0f;
        }
        uint32_t big = 0;
#pragma unroll
        for (int f = 0; f < 2; f++) {
          if (f < nf) {
            const uint4 xv = x[f * VPF + i];
#pragma unroll
            for (int ch = 0; ch < 2; ch++) {
              const int sq = 2 * f + ch;
              const int q0 = (int)(ch ? xv.y : xv.x), q1 = (int)(ch ? xv.w : xv.z);
              const uint32_t a0 = (uint32_t)abs(q0), a1 = (uint32_t)abs(q1);
              big |= a0 | a1;
              T* dst = seq_base + (size_t)sq * STRIDE_T + k0;
              // ramp parameters of the band of bin k0; those of bin k0 + 1 are only fetched where a band starts there
              const uint32_t rbase = ramp_s + (uint32_t)sq * (uint32_t)((FRAD_NBANDS + 1) * 16);
              const uint4 w0 = lds128(rbase + off0);
              uint4 w1 = w0;
              lds128_if(!same, rbase + off1, w1);
              if constexpr (sizeof(T) == 8) {
                const double2 r0 = make_double2(__hiloint2double((int)w0.y, (int)w0.x), __hiloint2double((int)w0.w, (int)w0.z));
                const double2 r1 = make_double2(__hiloint2double((int)w1.y, (int)w1.x), __hiloint2double((int)w1.w, (int)w1.z));
                // j * step + start: np.linspace rounds the product first; the fused form differs by < 1 ulp of the divisor
                // (1e-16 relative, four orders below the stated 1e-12) and saves an FP64 issue slot per coefficient
                const double d0 = fma(jd0, r0.y, r0.x);
                const double d1 = fma(jd1, r1.y, r1.x);
                const double m0 = tail->dq[min(a0, LIM - 1)];
                const double m1 = tail->dq[min(a1, LIM - 1)];
                const double s0 = __hiloint2double(__double2hiint(m0) ^ (q0 & 0x80000000), __double2loint(m0));
                const double s1 = __hiloint2double(__double2hiint(m1) ^ (q1 & 0x80000000), __double2loint(m1));
                *(double2*)dst = make_double2(s0 * d0, s1 * d1);
              } else {
                const float4 p0 = make_float4(__uint_as_float(w0.x), __uint_as_float(w0.y), __uint_as_float(w0.z), __uint_as_float(w0.w));
                const float4 p1 = make_float4(__uint_as_float(w1.x), __uint_as_float(w1.y), __uint_as_float(w1.z), __uint_as_float(w1.w));
                const float d0 = fmaf(kf0 - p0.z, p0.y, (p0.w - kf0) * p0.x);
                const float d1 = fmaf(kf1 - p1.z, p1.y, (p1.w - kf1) * p1.x);
                const float m0 = tail->aux[min(a0, LIM - 1)];
                const float m1 = tail->aux[min(a1, LIM - 1)];
                const float s0 = __uint_as_float(__float_as_uint(m0) ^ ((uint32_t)q0 & 0x80000000u));
                const float s1 = __uint_as_float(__float_as_uint(m1) ^ ((uint32_t)q1 & 0x80000000u));
                *(float2*)dst = make_float2(s0 * d0, s1 * d1);
              }
            }
          }
        }
        if (big >= LIM) bigmask |= 1u << i;
      }
      if (bigmask) stage_pairs_fix<T, TAIL>(bigmask, f0, f1, nf, tail, seq_base, STRIDE_T, dq64, inv_scale, tid);
    } else {
    // value X[k] of sequence slot sq from the 32-bit word: profile 0 unpack + scrub, profile 1 dequantiser x ramp.
    // The dequantiser table already carries the 1 / 2^(bits-1) of profile1.py:65 (a power of two: exact).
    auto coef = [&](uint32_t u0, int k, int sq) -> T {
      if (PROFILE == 0) {
        const uint32_t u = P.le ? u0 : __byte_perm(u0, 0, 0x0123);
        const float fv = __uint_as_float(u);
        return (T)(((u & 0x7f800000u) == 0x7f800000u) ? 0.0f : fv);          // NaN / Inf -> 0 (profile0.py:66)
      } else {
        const int qi = (int)u0;
        const int qa = abs(qi);
        const int bnd = tail->bidx[k];
        if constexpr (sizeof(T) == 8) {
          double mag;
          if (qa < DQ_SMEM64) mag = tail->dq[qa];
          else mag = dequant_big(dq64, qa) * inv_scale;
          const double2 rp = tail->rp[sq][bnd];
          const double div = __dadd_rn(__dmul_rn((double)(k - tail->lo[bnd]), rp.y), rp.x);   // j * step + start
          return (T)(((qi < 0) ? -mag : mag) * div);
        } else {
          float mag;
          if (qa < DQ_TAB) mag = tail->aux[qa];
          else mag = dequant_big_f32(qa, (float)inv_scale);
          const float4 bpv = tail->bp[sq][bnd];
          const float kf = (float)k;
          const float dv = fmaf(kf - bpv.z, bpv.y, (bpv.w - kf) * bpv.x);
          return (T)(((qi < 0) ? -mag : mag) * dv);
        }
      }
    };
    auto coef4 = [&](const uint32_t* u, const int* ks, const int* sqs, T* v) {
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = coef(u[j], ks[j], sqs[j]);
    };
    if (p0_any) {
      // profile 0 at 12 / 16 / 24 / 48 / 64 bits (profile0.py:51-66): every value from its bytes, NaN / Inf scrubbed;
      // consecutive threads take consecutive values of the tile's channels
      if constexpr (PROFILE == 0) {
        constexpr int STRIDE_T = LY::WARP_BYTES / (int)sizeof(T);
        T* seq_base = (T*)smem_raw;
        const int Cg = (CT == 4) ? 4 : C, c0 = (CT == 4) ? 4 * grp : 0;
        const int bits = P.bits;
        const bool le = P.le != 0;
        for (int f = 0; f < nf; f++) {
          const uint8_t* frame = P.payload + (size_t)(b0 + f) * P.payload_stride;
          if constexpr (CT == 1 || CT == 2) {             // eight consecutive values per thread and step, word loads
            for (int e8 = tid; e8 < FN * C / 8; e8 += FW * 32) {
              T v8[8];
              p0_load8<T>(frame + (size_t)e8 * bits, bits, le, v8);
              if constexpr (CT == 2) {                    // bins 4 e8 .. 4 e8 + 3 of both channels
                T* d0 = seq_base + (size_t)(2 * f) * STRIDE_T + 4 * e8;
                T* d1 = d0 + STRIDE_T;
#pragma unroll
                for (int j = 0; j < 4; j++) { d0[j] = v8[2 * j]; d1[j] = v8[2 * j + 1]; }
              } else {
                T* d0 = seq_base + (size_t)f * STRIDE_T + 8 * e8;
#pragma unroll
                for (int j = 0; j < 8; j++) d0[j] = v8[j];
              }
            }
            continue;
          }
          for (int idx = tid; idx < FN * Cg; idx += FW * 32) {
            int k, cc;
            if (CT == 1) { k = idx; cc = 0; } else if (CT == 2) { k = idx >> 1; cc = idx & 1; }
            else if (CT == 4) { k = idx >> 2; cc = idx & 3; } else { k = idx / C; cc = idx - k * C; }
            const size_t vi = (size_t)k * C + c0 + cc;
            double v;
            if (bits == 12) {                              // two values in three bytes, big-endian nibbles
              const uint8_t* p3 = frame + (vi >> 1) * 3;
              const uint32_t h12 = (vi & 1) ? (((uint32_t)(p3[1] & 15u) << 8) | p3[2]) : (((uint32_t)p3[0] << 4) | (p3[1] >> 4));
              v = (double)__half2float(__ushort_as_half((unsigned short)(h12 << 4)));
              if (!isfinite(v)) v = 0.0;
            } else {
              v = p0_load_value(frame + vi * (size_t)(bits >> 3), bits, le);
            }
            seq_base[(size_t)((CT == 4) ? cc : f * C + cc) * STRIDE_T + k] = (T)v;
          }
        }
      }
    } else if constexpr (CT == 0) {
      stage_convert_any<LY::WARP_BYTES / (int)sizeof(T), T>(
          [&](int f) {
            return (PROFILE == 0) ? (const uint32_t*)(P.payload + (size_t)(b0 + f) * P.payload_stride)
                                  : (const uint32_t*)(P.q + (size_t)(b0 + f) * words_per_frame);
          },
          (T*)smem_raw, C, nf, tid, coef);
    } else {
      convert_tile<CT, LY::WARP_BYTES / (int)sizeof(T), T>(x, (T*)smem_raw, C, nf, tid, coef4);
    }
    }
    // f64: the three per-lane twiddle bases (cross twiddle, ci, ei) are fetched before the barrier, not at their first use
    V ci_lane = mk2<T>(T(0), T(0)), ei_lane = ci_lane, tw_lane = ci_lane;
    if constexpr (sizeof(T) == 8 && FRAD_DEC_PRELOAD) { ci_lane = __ldg(ci + lane); ei_lane = __ldg(ei + lane); tw_lane = __ldg(&tw32[32 + lane]); }
    __syncthreads();
    {   // pull the next tile of this CTA towards L2 while the current one is transformed
      const int nt = tile + gridDim.x;
      if (nt < tm.ntiles) {
        const int nb0 = (CT == 4) ? nt / tm.groups : nt * tm.fpc;
        const int nnf = min(tm.fpc, P.B - nb0);
        const char* base = (PROFILE == 0) ? (const char*)(P.payload + (size_t)nb0 * P.payload_stride)
                                          : (const char*)(P.q + (size_t)nb0 * words_per_frame);
        size_t bytes = (PROFILE == 0) ? (size_t)nnf * P.payload_stride : (size_t)nnf * words_per_frame * 4;
        if constexpr (PROFILE == 2 && (CT == 1 || CT == 2 || CT == 4)) {      // the tile's group of tns_synth_kernel's layout: 64 KiB
          const int ngrp = (CT == 4) ? nt - nb0 * tm.groups : 0;
          const int ns4 = (CT == 4) ? nb0 * C + 4 * ngrp : nb0 * C;
          base = (const char*)(P.spec + (size_t)(ns4 >> 2) * FN * 4);
          bytes = (size_t)FN * 4 * sizeof(double);
        }
        for (size_t o = (size_t)tid * 128; o < bytes; o += (size_t)FW * 32 * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + o));
      }
    }
    const int tq_next = fetch_tq(tile + gridDim.x);        // consumed after the transform
    if (w < nseq) {
      unsigned char* mine = smem_raw + (size_t)w * LY::WARP_BYTES;
      V* buf = (V*)mine;
      T* X = (T*)mine;
      // ---- pairwise pre-processing: R1 = X[k] - i X[N-k], R2 = (X[M-k] + i X[M+k]) e^{-i pi/4},
      //      S = ci (R1 + R2), G = ei ci (R1 - R2), Z[k] = S + iG, Z[M-k] = conj(S - iG) ----
      V z[32];
      V zm_keep[17];
      const T r2 = (T)FRAD_RSQRT2;
      // twiddles of bin k = lane + 32 o.  f64: ci[k] = ci[lane] e^{i pi o / 128}, ei[k] = ei[lane] e^{i pi o / 32} -- one
      // per-lane table value times a compile-time constant instead of two 16-byte loads per pair (see fft1024_warp)
      if constexpr (sizeof(T) == 8 && !FRAD_DEC_PRELOAD) { ci_lane = __ldg(ci + lane); ei_lane = __ldg(ei + lane); }
      V eci_lane = ci_lane;
      if constexpr (sizeof(T) == 8) eci_lane = cmul(ei_lane, ci_lane);
      auto pre_pair = [&](int k, int o, V& zk, V& zmk) {
        const T xk = X[k], xnk = (k == 0) ? T(0) : X[FN - k];
        const T a = X[FM - k], bb = X[FM + k];
        const V R1 = mk2<T>(xk, -xnk);
        const V R2 = mk2<T>((a + bb) * r2, (bb - a) * r2);
        V S, G;
        if constexpr (sizeof(T) == 8) {
          // ci[k] and the product ei[k] ci[k] from per-lane bases times compile-time constants: four complex multiplies per pair
          const V cik = cmul(ci_lane, mk2<T>((T)g128_cos(o), (T)g128_sin(o)));
          const V ecik = cmul(eci_lane, mk2<T>((T)g160_cos(o), (T)g160_sin(o)));
          S = cmul(cik, cadd(R1, R2));
          G = cmul(ecik, csub(R1, R2));
        } else {
          const V cik = __ldg(ci + k), eik = __ldg(ei + k);
          S = cmul(cik, cadd(R1, R2));
          G = cmul(eik, cmul(cik, csub(R1, R2)));
        }
        zk = mk2<T>(S.x - G.y, S.y + G.x);
        zmk = mk2<T>(S.x + G.y, -(S.y - G.x));          // conj(S - iG)
      };
#pragma unroll
      for (int o = 0; o < 16; o++) pre_pair(lane + 32 * o, o, z[o], zm_keep[o]);
      zm_keep[16] = mk2<T>(T(0), T(0));
      if (lane == 0) { V dummy; pre_pair(FM / 2, 16, zm_keep[16], dummy); }      // Z[512] (self pair): lane 0, o = 16
      __syncwarp();
      const int partner = (32 - lane) & 31;
#pragma unroll
      for (int o = 0; o < 16; o++) {                     // recv[o] = partner's Z[M - k'] = my Z[lane + 32 (31 - o)]
        V rcv;
        rcv.x = shfl_t<T>(zm_keep[o].x, partner);
        rcv.y = shfl_t<T>(zm_keep[o].y, partner);
        zm_keep[o] = rcv;
      }
      // general lanes: z[31 - o] = recv[o]; lane 0 (its own partner): z[32 - o] = recv[o] (o >= 1), z[16] = self pair
#pragma unroll
      for (int j = 16; j < 32; j++) {
        const V g = zm_keep[31 - j];
        const V l0 = (j == 16) ? zm_keep[16] : zm_keep[32 - j];
        z[j] = (lane == 0) ? l0 : g;
      }
      fft1024_warp<T, true>(z, buf, lane, tw32, (sizeof(T) == 8 && FRAD_DEC_PRELOAD) ? &tw_lane : nullptr);
      // ---- undo the Makhoul reordering into the swizzled time-domain staging.  Output m = lane + 32 o goes to samples
      //      n = 4m, 4m + 2 (o < 16) or 2N - 1 - 4m, 2N - 3 - 4m (o >= 16); swzT(n) = per-lane base + 128 o, where the
      //      base only depends on the parity of o (f64: 16-element rows rotated by the row number) or not at all (f32:
      //      rows rotated by row & 3) -- the stores take immediate offsets ----
      if constexpr (sizeof(T) == 8) {
        int lo_a[2], lo_b[2], hi_a[2], hi_b[2];
#pragma unroll
        for (int par = 0; par < 2; par++) {
          lo_a[par] = swzT<T>(4 * lane + 128 * par) - 128 * par;
          lo_b[par] = swzT<T>(4 * lane + 2 + 128 * par) - 128 * par;
          hi_a[par] = swzT<T>(2 * FN - 1 - 4 * lane - 128 * (16 + par)) + 128 * (16 + par);
          hi_b[par] = swzT<T>(2 * FN - 3 - 4 * lane - 128 * (16 + par)) + 128 * (16 + par);
        }
#pragma unroll
        for (int o = 0; o < 32; o++) {
          if (o < 16) { X[lo_a[o & 1] + 128 * o] = z[o].x; X[lo_b[o & 1] + 128 * o] = z[o].y; }
          else { X[hi_a[o & 1] - 128 * o] = z[o].x; X[hi_b[o & 1] - 128 * o] = z[o].y; }
        }
      } else {
        const int lo_a = swzT<T>(4 * lane), lo_b = swzT<T>(4 * lane + 2);
        const int hi_a = swzT<T>(2 * FN - 1 - 4 * lane - 128 * 16) + 128 * 16, hi_b = swzT<T>(2 * FN - 3 - 4 * lane - 128 * 16) + 128 * 16;
#pragma unroll
        for (int o = 0; o < 32; o++) {
          if (o < 16) { X[lo_a + 128 * o] = z[o].x; X[lo_b + 128 * o] = z[o].y; }
          else { X[hi_a - 128 * o] = z[o].x; X[hi_b - 128 * o] = z[o].y; }
        }
      }
    }
    if (tile + gridDim.x < tm.ntiles) put_ramps(tq_next);   // every warp has left the staging of this tile (barrier above)
    __syncthreads();
    if (EARLY) request_tile(tile + gridDim.x);              // in flight while this tile is stored
    // ---- cooperative, coalesced store of PCM [n][c].  f64 / f32 output are lean loops (the next tile's words are
    //      waiting in registers); the other sample formats go through an out-of-line routine ----
    if (P.pcm_dtype == FRAD_F64 || P.pcm_dtype == FRAD_F32) {
      const bool out64 = (P.pcm_dtype == FRAD_F64);
      const bool wide64 = (((uintptr_t)P.pcm) % 32) == 0;
      for (int f = 0; f < nf; f++) {
        const size_t fbase = (size_t)(b0 + f) * words_per_frame;
#pragma unroll 2
        for (int e4 = tid; e4 < vec_count<CT>(C); e4 += FW * 32) {
          const size_t woff = fbase + (size_t)vec_word<CT>(e4, C, grp);
          T o4[4];
          gather4<CT, LY::WARP_BYTES / (int)sizeof(T)>((const T*)smem_raw, C, f, e4, [](int n) { return swzT<T>(n); },
                                                       [](T v) { return v; }, o4);
          if (out64) {
            if (wide64) {                                  // one 32-byte store
              w8 o;
#pragma unroll
              for (int j = 0; j < 4; j++) {
                const long long u = __double_as_longlong((double)o4[j]);
                o.v[2 * j] = (uint32_t)u; o.v[2 * j + 1] = (uint32_t)((unsigned long long)u >> 32);
              }
              stg256((double*)P.pcm + woff, o);
            } else {
              double2* dst = (double2*)((double*)P.pcm + woff);
              dst[0] = make_double2((double)o4[0], (double)o4[1]);
              dst[1] = make_double2((double)o4[2], (double)o4[3]);
            }
          } else {
            *(float4*)((float*)P.pcm + woff) = make_float4((float)o4[0], (float)o4[1], (float)o4[2], (float)o4[3]);
          }
        }
      }
    } else {
      store_tile_formats<T, CT>((const T*)smem_raw, P.pcm, P.pcm_dtype, C, nf, b0, grp, words_per_frame, tid);
    }
  }
}

}  // namespace

// eligibility checks shared with transform.cu
bool frad_fast2048_decode_ok(const dec_params& P, int prec) {
  (void)prec;
  if (P.N != FN || (P.C > FW && (P.C % 4) != 0)) return false;
  if (P.profile == 0 && ((P.payload_stride % 16) != 0 || ((uintptr_t)P.payload % 16) != 0)) return false;
  if (P.profile == 2 && !P.spec) return false;
  if (P.profile == 1 && ((uintptr_t)P.q % 16) != 0) return false;
  if (((uintptr_t)P.pcm % 16) != 0) return false;
  return true;
}

#define DEC_DISPATCH(TT, PROF, CT)                                                                       \
  do {                                                                                                   \
    auto k = dec2048_kernel<TT, PROF, CT>;                                                               \
    const tile_map tm = make_tiles<CT>(P.B, P.C);                                                        \
    int grid;                                                                                            \
    const size_t smem = (size_t)FW * dec_layout<TT>::WARP_BYTES + ((PROF == 0) ? 16 : sizeof(dec_tail<TT, (PROF == 1 && CT == 2)>)); \
    int rc = launch_cfg(ctx, k, smem, tm.ntiles, &grid);                                                 \
    if (rc) return rc;                                                                                   \
    k<<<grid, FW * 32, smem, ctx->stream>>>(P, plan_view(pl), (const typename vec2<TT>::type*)pl->tw32,  \
                                            (const double*)ctx->dequant_table);                          \
  } while (0)

#define DEC_DISPATCH_C(TT, PROF)                                                                       \
  do {                                                                                                 \
    if (P.C == 2) DEC_DISPATCH(TT, PROF, 2); else if (P.C == 1) DEC_DISPATCH(TT, PROF, 1);             \
    else if (P.C % 4 == 0) DEC_DISPATCH(TT, PROF, 4); else DEC_DISPATCH(TT, PROF, 0);                  \
  } while (0)

int frad_launch_decode_2048(frad_ctx* ctx, const dec_params& P, const frad_plan* pl, int prec) {
  int rc = frad_ensure_dq64(ctx);
  if (rc) return rc;
  if (prec == FRAD_PREC_F32) {
    if (P.profile == 0) DEC_DISPATCH_C(float, 0); else if (P.profile == 1) DEC_DISPATCH_C(float, 1); else DEC_DISPATCH_C(float, 2);
  } else {
    if (P.profile == 0) DEC_DISPATCH_C(double, 0); else if (P.profile == 1) DEC_DISPATCH_C(double, 1); else DEC_DISPATCH_C(double, 2);
  }
  FRAD_LAUNCH_CHECK(ctx);
  return FRAD_OK;
}
