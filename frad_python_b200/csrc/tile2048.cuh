// Shared pieces of the N = 2048 fast kernels (dct2048.cu: decoders, enc2048.cu: encoders): frame constants, the z / time-domain
// staging layouts and the cooperative, coalesced tile staging helpers.
#pragma once
#include <cstdio>
#include <cstdlib>
#include "plan.cuh"
#include "fft32.cuh"
#include "profiles.cuh"

namespace {

#ifndef FRAD_FW
#define FRAD_FW 4
#endif
#ifndef FRAD_MINB
#define FRAD_MINB 4
#endif
#ifndef FRAD_ENC_EARLY
#define FRAD_ENC_EARLY 0          // measured: requesting the next tile before the store phase costs the encoder 3 %
#endif
#ifndef FRAD_DEC_PRELOAD
#define FRAD_DEC_PRELOAD 1
#endif
#ifndef FRAD_ENC_F32THR
#define FRAD_ENC_F32THR 1
#endif
constexpr int FN = 2048, FM = 1024, FW = FRAD_FW;          // frame, half, warps per CTA
constexpr int SEQ_CPX = 1056;                        // padded complex slots per warp buffer
constexpr int DQ_TAB = 1024;                         // f32 dequantiser table entries (|q| < DQ_TAB)
constexpr int DQ_TAB64 = FRAD_DQ_LEN;                // f64 table (global memory, frad_ensure_dq64)
constexpr int DQ_SMEM64 = 256;                       // head of the f64 table kept in shared memory
#define FRAD_RSQRT2 0.70710678118654752440

__device__ __forceinline__ int zword(int n) {        // word index (units of T) of sample n inside the z layout
  const int nn = (n & 1) ? (2 * FN - 1 - n) : n;
  return 2 * padc(nn >> 2) + ((nn >> 1) & 1);
}
// f32 time-domain staging: 32-element rows rotated by (row & 3) -- enough to separate the four rows a stride-4 scatter
// touches, and unchanged by the 128 o offsets of the FFT output order (immediate-offset stores)
__device__ __forceinline__ int swz(int n) { return (n & ~31) | ((n + ((n >> 5) & 3)) & 31); }
// time-domain staging swizzle by element size: rows of 128 bytes (32 floats / 16 doubles) rotated by the row number, so
// that the stride-4 scatter after the FFT and the stride-1/2/4 gathers of the store are both conflict-free per wavefront
template <typename T> __device__ __forceinline__ int swzT(int n) {
  if constexpr (sizeof(T) == 8) return (n & ~15) | ((n + (n >> 4)) & 15);
  else return swz(n);
}

template <typename T> __device__ __forceinline__ T shfl_t(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// ---- cooperative staging helpers ----------------------------------------------------------------
// Frame words (32-bit) are interleaved [i][c].  A tile is a set of whole sequences that fit the CTA's
// FW warps; CT selects how it is cut (all variants move 16 bytes per thread per step, fully coalesced):
//   CT == 1  mono: FW frames per tile, a float4 is 4 consecutive samples of the one channel
//   CT == 2  stereo: FW/2 frames per tile, a float4 is 2 samples x 2 channels
//   CT == 4  C % 4 == 0: one frame x one group of 4 channels per tile, a float4 is 1 sample x 4 channels
//   CT == 0  any C <= FW: FW/C frames per tile, element index split by a run-time division
struct tile_map { int fpc, groups, ntiles; };

template <int CT>
__host__ __device__ inline tile_map make_tiles(int B, int C) {
  tile_map t;
  if (CT == 4) { t.fpc = 1; t.groups = C / 4; t.ntiles = B * t.groups; return t; }
  t.fpc = FW / C; if (t.fpc < 1) t.fpc = 1;
  t.groups = 1;
  t.ntiles = (B + t.fpc - 1) / t.fpc;
  return t;
}

template <int CT> __device__ __forceinline__ int vec_count(int C) { return (CT == 4) ? FN : FN * C / 4; }
// 32-bit word offset inside the frame of vector step e4 (group g of 4 channels for CT == 4)
template <int CT> __device__ __forceinline__ int vec_word(int e4, int C, int g) { return (CT == 4) ? e4 * C + 4 * g : 4 * e4; }
// element j (0..3) of vector step e4 -> sample / bin index i and sequence slot s inside the tile
template <int CT> __device__ __forceinline__ void vec_elem(int e4, int j, int C, int f, int& i, int& s) {
  if (CT == 2) { i = 2 * e4 + (j >> 1); s = 2 * f + (j & 1); }
  else if (CT == 1) { i = 4 * e4 + j; s = f; }
  else if (CT == 4) { i = e4; s = j; }
  else { const int e = 4 * e4 + j; i = e / C; s = f * C + (e - i * C); }
}

template <int CT, int STRIDE_WORDS, typename POS>
__device__ __forceinline__ void scatter_frame(const uint32_t* __restrict__ frame, uint32_t* seq_base, int C, int f, int g,
                                              int tid, POS pos) {
  const int cnt = vec_count<CT>(C);
#pragma unroll 4
  for (int e4 = tid; e4 < cnt; e4 += FW * 32) {
    const uint4 x = __ldg((const uint4*)(frame + vec_word<CT>(e4, C, g)));
    const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int i, sq;
      vec_elem<CT>(e4, j, C, f, i, sq);
      (seq_base + (size_t)sq * STRIDE_WORDS)[pos(i)] = xs[j];
    }
  }
}

// streaming 16-byte load: read-only path, no L1 allocation (the small L1 left beside ~210 KB of shared memory is kept
// for the twiddle tables, which every sequence re-reads)
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

// Whole-tile staging: every 16-byte load of the tile (16 per thread for FW = 4) is issued before the first
// scatter, so one global-memory latency is exposed per tile instead of one per unrolled group of four.
// src(f) returns the first 32-bit word of frame f of the tile.
template <int CT, int STRIDE_WORDS, typename SRC, typename POS>
__device__ __forceinline__ void stage_tile(SRC src, uint32_t* seq_base, int C, int nf, int g, int tid, POS pos) {
  if constexpr (CT == 0) {
    for (int f = 0; f < nf; f++) scatter_frame<CT, STRIDE_WORDS>(src(f), seq_base, C, f, g, tid, pos);
  } else {
    constexpr int FPT = (CT == 4) ? 1 : FW / CT;                                  // frames per tile
    constexpr int VPF = ((CT == 4) ? FN : FN * CT / 4) / (FW * 32);               // vector steps per thread and frame
    static_assert(VPF * FW * 32 == ((CT == 4) ? FN : FN * CT / 4), "tile does not divide evenly");
    uint4 x[FPT * VPF];
#pragma unroll
    for (int f = 0; f < FPT; f++) {
      const uint32_t* frame = src(f < nf ? f : nf - 1) + vec_word<CT>(tid, C, g);   // a short tile re-reads its last frame
#pragma unroll
      for (int i = 0; i < VPF; i++) x[f * VPF + i] = ldg_stream((const uint4*)(frame + vec_word<CT>(i * FW * 32, C, 0)));
    }
#pragma unroll
    for (int f = 0; f < FPT; f++) {
      if (f < nf) {
#pragma unroll
        for (int i = 0; i < VPF; i++) {
          const uint4 xv = x[f * VPF + i];
          const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int j = 0; j < 4; j++) {
            int ii, sq;
            vec_elem<CT>(tid + i * FW * 32, j, C, f, ii, sq);
            (seq_base + (size_t)sq * STRIDE_WORDS)[pos(ii)] = xs[j];
          }
        }
      }
    }
  }
}

// Decoder staging.  The 16-byte loads of a tile and their conversion are two steps, so that the kernel can issue the
// loads of its NEXT tile before it stores the current one (their latency is hidden behind the store phase); the words wait
// in registers.  Every word is converted on the way into shared memory (profile-0 unpack or profile-1 dequantiser x ramp,
// `cvt4`) and stored as T at index k of its sequence -- the raw words never touch shared memory.  Neighbouring bins of one
// sequence are stored as one 2 x T vector (conflict-free).
template <int CT> struct tile_regs {
  static constexpr int FPT = (CT == 4) ? 1 : ((CT == 0) ? 1 : FW / CT);
  static constexpr int VPF = (CT == 0) ? 1 : ((CT == 4) ? FN : FN * CT / 4) / (FW * 32);
  static constexpr int NX = FPT * VPF;
};
template <int CT, typename SRC>
__device__ __forceinline__ void load_tile(SRC src, int C, int nf, int g, int tid, uint4* x) {
  constexpr int FPT = tile_regs<CT>::FPT, VPF = tile_regs<CT>::VPF;
#pragma unroll
  for (int f = 0; f < FPT; f++) {
    const uint32_t* frame = src(f < nf ? f : nf - 1) + vec_word<CT>(tid, C, g);       // a short tile re-reads its last frame
#pragma unroll
    for (int i = 0; i < VPF; i++) x[f * VPF + i] = ldg_stream((const uint4*)(frame + vec_word<CT>(i * FW * 32, C, 0)));
  }
}
// encoder side: the loaded words go to position pos(i) of their sequence unchanged
template <int CT, int STRIDE_WORDS, typename POS>
__device__ __forceinline__ void scatter_tile(const uint4* x, uint32_t* seq_base, int C, int nf, int tid, POS pos) {
  constexpr int FPT = tile_regs<CT>::FPT, VPF = tile_regs<CT>::VPF;
#pragma unroll
  for (int f = 0; f < FPT; f++) {
    if (f < nf) {
#pragma unroll
      for (int i = 0; i < VPF; i++) {
        const uint4 xv = x[f * VPF + i];
        const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          int ii, sq;
          vec_elem<CT>(tid + i * FW * 32, j, C, f, ii, sq);
          (seq_base + (size_t)sq * STRIDE_WORDS)[pos(ii)] = xs[j];
        }
      }
    }
  }
}
template <int CT, int STRIDE_T, typename T, typename CVT4>
__device__ __forceinline__ void convert_tile(const uint4* x, T* seq_base, int C, int nf, int tid, CVT4 cvt4) {
  typedef typename vec2<T>::type V2;
  constexpr int FPT = tile_regs<CT>::FPT, VPF = tile_regs<CT>::VPF;
#pragma unroll
  for (int f = 0; f < FPT; f++) {
    if (f < nf) {
#pragma unroll
      for (int i = 0; i < VPF; i++) {
        const uint4 xv = x[f * VPF + i];
        const uint32_t xs[4] = {xv.x, xv.y, xv.z, xv.w};
        const int e4 = tid + i * FW * 32;
        T v[4];
        int ks[4], sqs[4];
#pragma unroll
        for (int j = 0; j < 4; j++) vec_elem<CT>(e4, j, C, f, ks[j], sqs[j]);
        cvt4(xs, ks, sqs, v);
        if constexpr (CT == 2) {            // (k, c0) (k, c1) (k+1, c0) (k+1, c1): one pair of bins per channel
          *(V2*)(seq_base + (size_t)sqs[0] * STRIDE_T + ks[0]) = mk2<T>(v[0], v[2]);
          *(V2*)(seq_base + (size_t)sqs[1] * STRIDE_T + ks[1]) = mk2<T>(v[1], v[3]);
        } else if constexpr (CT == 1) {     // four consecutive bins of the one channel
          *(V2*)(seq_base + (size_t)sqs[0] * STRIDE_T + ks[0]) = mk2<T>(v[0], v[1]);
          *(V2*)(seq_base + (size_t)sqs[2] * STRIDE_T + ks[2]) = mk2<T>(v[2], v[3]);
        } else {
#pragma unroll
          for (int j = 0; j < 4; j++) (seq_base + (size_t)sqs[j] * STRIDE_T)[ks[j]] = v[j];
        }
      }
    }
  }
}
// run-time channel count (CT == 0): plain loop, loads and conversion together
template <int STRIDE_T, typename T, typename SRC, typename CVT>
__device__ __forceinline__ void stage_convert_any(SRC src, T* seq_base, int C, int nf, int tid, CVT cvt) {
  for (int f = 0; f < nf; f++) {
    const uint32_t* frame = src(f);
    const int cnt = vec_count<0>(C);
#pragma unroll 4
    for (int e4 = tid; e4 < cnt; e4 += FW * 32) {
      const uint4 x = ldg_stream((const uint4*)(frame + vec_word<0>(e4, C, 0)));
      const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int j = 0; j < 4; j++) {
        int k, sq;
        vec_elem<0>(e4, j, C, f, k, sq);
        (seq_base + (size_t)sq * STRIDE_T)[k] = cvt(xs[j], k, sq);
      }
    }
  }
}

template <int CT, int STRIDE_WORDS, typename W, typename O, typename POS, typename CVT>
__device__ __forceinline__ void gather4(const W* seq_base, int C, int f, int e4, POS pos, CVT cvt, O* o4) {
#pragma unroll
  for (int j = 0; j < 4; j++) {
    int i, sq;
    vec_elem<CT>(e4, j, C, f, i, sq);
    o4[j] = cvt((seq_base + (size_t)sq * STRIDE_WORDS)[pos(i)]);
  }
}


// Eight consecutive profile-0 values from their bytes (profile0.py:51-66), word loads only: src is 4-byte aligned because
// eight values take 12 / 16 / 24 / 48 / 64 bytes.  NaN / Inf become 0.
__device__ __forceinline__ uint32_t byte_of(const uint32_t* w, int i) { return (w[i >> 2] >> (8 * (i & 3))) & 255u; }
template <typename T> __device__ __forceinline__ void p0_load8(const uint8_t* src, int bits, bool le, T* out) {
  const uint32_t* g = (const uint32_t*)src;
  if (bits == 16 || bits == 12) {
    uint32_t w[4];
    if (bits == 16) { const uint4 a = __ldg((const uint4*)g); w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; }
    else { w[0] = __ldg(g); w[1] = __ldg(g + 1); w[2] = __ldg(g + 2); w[3] = 0; }
#pragma unroll
    for (int j = 0; j < 8; j++) {
      uint32_t h;
      if (bits == 16) {
        h = (w[j >> 1] >> (16 * (j & 1))) & 0xffffu;
        if (!le) h = ((h & 255u) << 8) | (h >> 8);
      } else {
        const int p = 3 * (j >> 1);
        h = (j & 1) ? (((byte_of(w, p + 1) & 15u) << 8) | byte_of(w, p + 2)) : ((byte_of(w, p) << 4) | (byte_of(w, p + 1) >> 4));
        h <<= 4;
      }
      const float f = __half2float(__ushort_as_half((unsigned short)h));
      out[j] = (T)(((h & 0x7c00u) == 0x7c00u) ? 0.0f : f);
    }
  } else if (bits == 24) {
    uint32_t w[6];
#pragma unroll
    for (int i = 0; i < 6; i++) w[i] = __ldg(g + i);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const uint32_t b0 = byte_of(w, 3 * j), b1 = byte_of(w, 3 * j + 1), b2 = byte_of(w, 3 * j + 2);
      const uint32_t u = le ? ((b0 << 8) | (b1 << 16) | (b2 << 24)) : ((b0 << 24) | (b1 << 16) | (b2 << 8));
      out[j] = (T)(((u & 0x7f800000u) == 0x7f800000u) ? 0.0f : __uint_as_float(u));
    }
  } else {
    uint32_t w[16];
    const int nw = (bits == 48) ? 12 : 16;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
      if (i < nw) { const uint4 a = __ldg((const uint4*)(g + i)); w[i] = a.x; w[i + 1] = a.y; w[i + 2] = a.z; w[i + 3] = a.w; }
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
      uint64_t u = 0;
      if (bits == 64) {
        u = le ? (((uint64_t)w[2 * j + 1] << 32) | w[2 * j])
               : (((uint64_t)__byte_perm(w[2 * j], 0, 0x0123) << 32) | __byte_perm(w[2 * j + 1], 0, 0x0123));
      } else {
#pragma unroll
        for (int i = 0; i < 6; i++) u |= (uint64_t)byte_of(w, 6 * j + i) << (le ? (16 + 8 * i) : (56 - 8 * i));
      }
      const double d = __longlong_as_double((long long)u);
      out[j] = (T)((((u >> 52) & 0x7ffull) == 0x7ffull) ? 0.0 : d);
    }
  }
}

// 32-byte global accesses (LDG.256 / STG.256 on sm_100)
struct w8 { uint32_t v[8]; };
__device__ __forceinline__ w8 ldg256_stream(const void* p) {
  w8 r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg256(void* p, const w8& r) {
  asm volatile("st.global.v8.u32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]),
               "r"(r.v[3]), "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7])
               : "memory");
}

// e = i * C + c  ->  (i, c); shifts for the common channel counts, a division otherwise
__device__ __forceinline__ void split_c(int e, int C, int& i, int& c) {
  if (C == 2) { i = e >> 1; c = e & 1; }
  else if (C == 1) { i = e; c = 0; }
  else if (C == 8) { i = e >> 3; c = e & 7; }
  else if (C == 4) { i = e >> 2; c = e & 3; }
  else { i = e / C; c = e - i * C; }
}

template <typename K>
int launch_cfg(frad_ctx* ctx, K kernel, size_t smem, int ntiles, int* grid) {
  FRAD_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  FRAD_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, FW * 32, smem));
  if (per_sm < 1) return frad_fail(ctx, FRAD_E_UNSUPPORTED, "fast path does not fit on this device");
  if (getenv("FRAD_DEBUG_LAUNCH")) fprintf(stderr, "[fradb200] fast2048 launch: smem %zu B, %d CTAs/SM, %d tiles\n", smem, per_sm, ntiles);
  long long g = (long long)per_sm * ctx->sm_count;
  *grid = (int)(ntiles < g ? ntiles : g);
  if (*grid < 1) *grid = 1;
  return FRAD_OK;
}


}  // namespace
