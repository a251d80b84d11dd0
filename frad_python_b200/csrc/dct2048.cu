// Fast path for the default frame length N = 2048 (src/tools/cli.py:26, every BASELINE config).
//
// One warp per (frame, channel) sequence: the 1024-point complex FFT behind the DCT runs entirely
// in registers as 32 x 32 with a single padded shared-memory exchange (fft32.cuh), so the only
// block-wide barriers are the two around the cooperative, fully coalesced 16-byte global loads /
// stores of the interleaved [n][c] / [k][c] layouts.  The profile pro/epilogues run per warp on the
// sequence while it sits in shared memory:
//   encode  PCM f32 -> DCT-II -> profile 0: big/little-endian f32 payload + per-frame peak
//                             -> profile 1: band energies (register partials + transposed warp
//                                reduction), thresholds, divisor ramps, power-law quantiser (f32 fast
//                                path with an exact f64 re-evaluation next to rounding boundaries)
//   decode  payload / quantised ints -> dequantise -> DCT-III -> PCM (f64 or f32 out)
// Reference lines: fourier/profile0.py:21-30,63-69; fourier/profile1.py:21-40,65-77;
// fourier/tools/p1tools.py:18-44.
#include "plan.cuh"
#include "fft32.cuh"
#include "profiles.cuh"

namespace {

constexpr int FN = 2048, FM = 1024, FW = 8;          // frame, half, warps per CTA
constexpr int SEQ_CPX = 1056;                        // padded complex slots per warp buffer

__device__ __forceinline__ int zword(int n) {        // word index (in units of T) of sample n inside the z layout
  const int nn = (n & 1) ? (2 * FN - 1 - n) : n;
  return 2 * padc(nn >> 2) + ((nn >> 1) & 1);
}
__device__ __forceinline__ int swz(int n) { return (n & ~31) | ((n + (n >> 5)) & 31); }

template <typename T> __device__ __forceinline__ T shfl_t(T v, int src);
template <> __device__ __forceinline__ float shfl_t<float>(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <> __device__ __forceinline__ double shfl_t<double>(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// 32 per-lane partial sums (index = band) -> lane b holds the warp total of band b.  31 shuffles.
__device__ __forceinline__ float transpose_reduce32(float* part, int lane) {
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const bool up = (lane & s) != 0;
#pragma unroll
    for (int j = 0; j < s; j++) {
      const float send = up ? part[j] : part[j + s];
      const float keep = up ? part[j + s] : part[j];
      part[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return part[0];
}

struct tile_map { int fpc, active, ntiles; };

__host__ __device__ inline tile_map make_tiles(int B, int C) {
  tile_map t;
  t.fpc = FW / C; if (t.fpc < 1) t.fpc = 1;
  t.active = t.fpc * C;
  t.ntiles = (B + t.fpc - 1) / t.fpc;
  return t;
}

// ------------------------------------------------------------------------------------------------
// encode
// ------------------------------------------------------------------------------------------------
struct enc_smem_tail {
  int lo[FRAD_NBANDS + 1];
  double thr[FW][FRAD_NBANDS + 1];
};

// quantiser for one coefficient: rint(sign(x) |x * scale / div|^0.75), exact w.r.t. the f64 formula
__device__ __forceinline__ int quant_one(float x, float a, float nb, float t0f, float t1f, float invn, float scalef,
                                         double t0, double step, double scale, bool force_exact) {
  // ramp value as a convex combination (no cancellation): (t0 (n - j) + t1 j) / n
  const float divf = (t0f * nb + t1f * a) * invn;
  const float y = fabsf(x) * scalef * __frcp_rn(divf);
  float s, v;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(y));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(v) : "f"(y * s));
  const float big = v + 12582912.0f;                 // round-to-nearest-even to an integer for v < 2^22
  const float r = big - 12582912.0f;
  const float dist = fabsf(v - r);
  const bool fast_ok = (dist < 0.5f - (8e-7f * v + 1e-7f)) && (v < 2.0e6f) && (divf > 0.0f) && !force_exact;
  int m;
  if (fast_ok) {
    m = __float_as_int(big) - 0x4B400000;
  } else {
    const double div = __dadd_rn(__dmul_rn((double)a, step), t0);       // np.linspace: j * step + start
    double rr = 0.0;
    if (div != 0.0) {
      const double yy = fabs(((double)x / div) * scale);
      rr = rint(sqrt(yy * sqrt(yy)));
    }
    m = (rr >= 2147483647.0) ? 2147483647 : (int)rr;
    if (!(rr == rr)) m = 0;
  }
  return (x < 0.0f) ? -m : m;
}

template <int PROFILE, int CT>
__global__ void __launch_bounds__(FW * 32, 2) enc2048_kernel(enc_params P, plan_dev pl, const float2* __restrict__ tw32,
                                                             int force_exact) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* seq_all = (float2*)smem_raw;
  enc_smem_tail* tail = (enc_smem_tail*)(smem_raw + (size_t)FW * SEQ_CPX * sizeof(float2));
  const int C = CT ? CT : P.C;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const tile_map tm = make_tiles(P.B, C);
  const float* pcm = (const float*)P.pcm;
  const float2* cf = (const float2*)pl.cf;
  const float2* df = (const float2*)pl.df;
  if (PROFILE == 1 && tid <= FRAD_NBANDS) { const int e = P.bt->edge[tid]; tail->lo[tid] = (e < FN) ? e : FN; }
  const int words_per_frame = FN * C;                     // 32-bit words of input and of output per frame

  for (int tile = blockIdx.x; tile < tm.ntiles; tile += gridDim.x) {
    const int b0 = tile * tm.fpc;
    const int nf = min(tm.fpc, P.B - b0);
    __syncthreads();                                      // previous tile fully stored before restaging
    // ---- cooperative, coalesced load of nf frames; scatter into the per-sequence z layout ----
    for (int f = 0; f < nf; f++) {
      const float4* src = (const float4*)(pcm + (size_t)(b0 + f) * P.hop * C);
      for (int e4 = tid; e4 < words_per_frame / 4; e4 += FW * 32) {
        const float4 x = __ldg(src + e4);
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int e = 4 * e4 + j;
          const int n = e / C, c = e - n * C;
          ((float*)(seq_all + (size_t)(f * C + c) * SEQ_CPX))[zword(n)] = xs[j];
        }
      }
    }
    __syncthreads();
    if (w < nf * C) {
      const int f = w / C, c = w - f * C, b = b0 + f;
      float2* buf = seq_all + (size_t)w * SEQ_CPX;
      float* X = (float*)buf;
      float2 v[32];
#pragma unroll
      for (int i = 0; i < 32; i++) v[i] = buf[lane + 33 * i];
      __syncwarp();
      fft1024_warp<float, false>(v, buf, lane, tw32);
      // ---- real-FFT split + DCT post-twiddle: X[k], X[N-k] for k = lane + 32 o ----
      const int partner = (32 - lane) & 31;
      float mx = 0.0f;
#pragma unroll
      for (int o = 0; o < 32; o++) {
        const float2 snd = (lane == 0) ? v[(32 - o) & 31] : v[31 - o];
        float2 zm;
        zm.x = __shfl_sync(0xffffffffu, snd.x, partner);
        zm.y = -__shfl_sync(0xffffffffu, snd.y, partner);       // conj(Z[M-k])
        const float2 zk = v[o];
        const float2 A = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y + zm.y));
        const float2 Bv = make_float2(0.5f * (zk.y - zm.y), -0.5f * (zk.x - zm.x));
        const int k = lane + 32 * o;
        const float2 ck = __ldg(cf + k), dk = __ldg(df + k);
        const float2 y = cadd(cmul(ck, A), cmul(dk, Bv));
        float xk = y.x, xnk = -y.y;
        if (PROFILE == 0) {
          mx = fmaxf(mx, fmaxf(fabsf(xk), (k >= 1) ? fabsf(xnk) : 0.0f));
          if (!P.le) { xk = __uint_as_float(__byte_perm(__float_as_uint(xk), 0, 0x0123)); xnk = __uint_as_float(__byte_perm(__float_as_uint(xnk), 0, 0x0123)); }
        }
        X[k] = xk;
        if (k >= 1) X[FN - k] = xnk;
        if (o == 0 && lane == 0) {
          const float2 cm = __ldg(cf + FM), dm = __ldg(df + FM);
          const float2 y2 = cadd(cmul(cm, cconj(A)), cmul(dm, cconj(Bv)));
          float xm = y2.x;
          if (PROFILE == 0) {
            mx = fmaxf(mx, fabsf(xm));
            if (!P.le) xm = __uint_as_float(__byte_perm(__float_as_uint(xm), 0, 0x0123));
          }
          X[FM] = xm;
        }
      }
      if (PROFILE == 0) {
        if (P.peak) {
#pragma unroll
          for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
          if (lane == 0) atomicMax((unsigned long long*)&P.peak[b], (unsigned long long)__double_as_longlong((double)mx));
        }
      } else {
        __syncwarp();
        // ---- band energies of the scaled spectrum: per-lane partials, one transposed reduction ----
        const float scalef = (float)P.scale;
        float part[32];
#pragma unroll
        for (int bnd = 0; bnd < 32; bnd++) {
          float s = 0.0f;
          if (bnd < FRAD_NBANDS) {
            const int hi = tail->lo[bnd + 1];
            for (int k = tail->lo[bnd] + lane; k < hi; k += 32) { const float xv = X[k] * scalef; s = fmaf(xv, xv, s); }
          }
          part[bnd] = s;
        }
        const float tot = transpose_reduce32(part, lane);
        int cnt = 1;
        if (lane < FRAD_NBANDS) cnt = tail->lo[lane + 1] - tail->lo[lane];
        const unsigned empty = __ballot_sync(0xffffffffu, cnt <= 0);
        const int first_empty = empty ? (__ffs(empty) - 1) : 32;
        double thr = 0.0;
        if (lane < FRAD_NBANDS && lane < first_empty) {
          const double level = pow(sqrt((double)tot / (double)cnt), 0.8);
          thr = fmax(level, P.bt->ath[lane]) * P.loss;
        }
        if (lane <= FRAD_NBANDS) tail->thr[w][lane] = thr;             // entry 27 = 0
        if (lane < FRAD_NBANDS) {
          const double vv = log(fmax(thr, 1.0)) / P.bt->log_base;
          const double d = (vv > 0.0) ? pow(vv, 1.0 / 0.75) : 0.0;
          P.tq[((size_t)b * FRAD_NBANDS + lane) * C + c] = sat_i32(rint(d));
        }
        __syncwarp();
        // ---- quantise band by band, in place (float -> int32) ----
        int* Q = (int*)X;
        for (int bnd = 0; bnd < FRAD_NBANDS - 1; bnd++) {
          const int lo = tail->lo[bnd], hi = tail->lo[bnd + 1];
          const int n = hi - lo;
          if (n <= 0) continue;
          const double t0 = tail->thr[w][bnd], t1 = tail->thr[w][bnd + 1];
          const double step = (t1 - t0) / (double)n;
          const float t0f = (float)t0, t1f = (float)t1, nfl = (float)n, invn = 1.0f / (float)n;
          for (int k = lo + lane; k < hi; k += 32) {
            const float a = (float)(k - lo);
            Q[k] = quant_one(X[k], a, nfl - a, t0f, t1f, invn, scalef, t0, step, P.scale, force_exact != 0);
          }
        }
        for (int k = tail->lo[FRAD_NBANDS - 1] + lane; k < FN; k += 32) Q[k] = 0;
      }
    }
    __syncthreads();
    // ---- cooperative, coalesced store: word e = k*C + c of frame f <- seq[f*C + c][k] ----
    for (int f = 0; f < nf; f++) {
      uint4* dst = (PROFILE == 0) ? (uint4*)(P.payload + (size_t)(b0 + f) * P.payload_stride)
                                  : (uint4*)(P.q + (size_t)(b0 + f) * words_per_frame);
      for (int e4 = tid; e4 < words_per_frame / 4; e4 += FW * 32) {
        uint32_t o4[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int e = 4 * e4 + j;
          const int k = e / C, c = e - k * C;
          o4[j] = ((const uint32_t*)(seq_all + (size_t)(f * C + c) * SEQ_CPX))[k];
        }
        dst[e4] = make_uint4(o4[0], o4[1], o4[2], o4[3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// decode
// ------------------------------------------------------------------------------------------------
template <typename T> struct dec_layout {
  static constexpr int SEQ_BYTES = SEQ_CPX * 2 * (int)sizeof(T);
  static constexpr int RAW_OFF = (sizeof(T) == 8) ? SEQ_BYTES : 0;       // f64: raw words live beside X
  static constexpr int WARP_BYTES = SEQ_BYTES + ((sizeof(T) == 8) ? FN * 4 : 0);
};

struct dec_smem_tail {
  int lo[FRAD_NBANDS + 1];
  double thr[FW][FRAD_NBANDS + 1];
};

template <typename T, int PROFILE, int CT>
__global__ void __launch_bounds__(FW * 32, (sizeof(T) == 8) ? 1 : 2)
dec2048_kernel(dec_params P, plan_dev pl, const typename vec2<T>::type* __restrict__ tw32) {
  typedef typename vec2<T>::type V;
  typedef dec_layout<T> LY;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  dec_smem_tail* tail = (dec_smem_tail*)(smem_raw + (size_t)FW * LY::WARP_BYTES);
  const int C = CT ? CT : P.C;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const tile_map tm = make_tiles(P.B, C);
  const V* ci = (const V*)pl.ci;
  const V* ei = (const V*)pl.ei;
  if (PROFILE == 1 && tid <= FRAD_NBANDS) { const int e = P.bt->edge[tid]; tail->lo[tid] = (e < FN) ? e : FN; }
  const int words_per_frame = FN * C;

  for (int tile = blockIdx.x; tile < tm.ntiles; tile += gridDim.x) {
    const int b0 = tile * tm.fpc;
    const int nf = min(tm.fpc, P.B - b0);
    __syncthreads();
    // ---- cooperative load of the 32-bit input words (payload or q), de-interleave per sequence ----
    for (int f = 0; f < nf; f++) {
      const uint4* src = (PROFILE == 0) ? (const uint4*)(P.payload + (size_t)(b0 + f) * P.payload_stride)
                                        : (const uint4*)(P.q + (size_t)(b0 + f) * words_per_frame);
      for (int e4 = tid; e4 < words_per_frame / 4; e4 += FW * 32) {
        const uint4 x = __ldg(src + e4);
        const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int e = 4 * e4 + j;
          const int k = e / C, c = e - k * C;
          ((uint32_t*)(smem_raw + (size_t)(f * C + c) * LY::WARP_BYTES + LY::RAW_OFF))[k] = xs[j];
        }
      }
    }
    __syncthreads();
    if (w < nf * C) {
      const int f = w / C, c = w - f * C, b = b0 + f;
      unsigned char* mine = smem_raw + (size_t)w * LY::WARP_BYTES;
      V* buf = (V*)mine;
      T* X = (T*)mine;
      const uint32_t* raw = (const uint32_t*)(mine + LY::RAW_OFF);
      // ---- dequantise / unpack into X[k] (type T) ----
      if (PROFILE == 0) {
        for (int k = lane; k < FN; k += 32) {
          uint32_t u = raw[k];
          if (!P.le) u = __byte_perm(u, 0, 0x0123);
          float fv = __uint_as_float(u);
          if (!isfinite(fv)) fv = 0.0f;
          X[k] = (T)fv;
        }
      } else {
        double thr = 0.0;
        if (lane < FRAD_NBANDS) {
          const double tv = (double)P.tq[((size_t)b * FRAD_NBANDS + lane) * C + c];
          thr = pow(P.bt->base, copysign(pow(fabs(tv), 0.75), tv));        // profile1.py:66
        }
        if (lane <= FRAD_NBANDS) tail->thr[w][lane] = thr;
        __syncwarp();
        const int last = tail->lo[FRAD_NBANDS - 1];
        for (int bnd = 0; bnd < FRAD_NBANDS - 1; bnd++) {
          const int lo = tail->lo[bnd], hi = tail->lo[bnd + 1];
          const int n = hi - lo;
          if (n <= 0) continue;
          const double t0 = tail->thr[w][bnd], t1 = tail->thr[w][bnd + 1];
          const double step = (t1 - t0) / (double)n;
          for (int k = lo + lane; k < hi; k += 32) {
            const int qi = (int)raw[k];
            if (sizeof(T) == 8) {
              const double qv = (double)qi;
              const double coef = copysign(pow(fabs(qv), 1.0 / 0.75), qv) / P.scale;
              X[k] = (T)(coef * __dadd_rn(__dmul_rn((double)(k - lo), step), t0));
            } else {
              const float qa = fabsf((float)qi);
              const float mag = qa * cbrtf(qa);                         // |q|^(4/3)
              const float div = (float)(((double)(k - lo)) * step + t0);
              X[k] = (T)(copysignf(mag, (float)qi) * (float)(1.0 / P.scale) * div);
            }
          }
        }
        for (int k = last + lane; k < FN; k += 32) X[k] = (T)0;
      }
      __syncwarp();
      // ---- V_k = (X[k] - i X[N-k]) ci[k]; Z[k] = (V_k + conj V_{M-k}) + i e_k (V_k - conj V_{M-k}) ----
      V v[32];
#pragma unroll
      for (int o = 0; o < 32; o++) {
        const int k = lane + 32 * o;
        const T xk = X[k];
        const T xnk = (k == 0) ? T(0) : X[FN - k];
        v[o] = cmul(mk2<T>(xk, -xnk), __ldg(ci + k));
      }
      const T xm = X[FM];
      const V vM = cmul(mk2<T>(xm, -xm), __ldg(ci + FM));
      __syncwarp();
      const int partner = (32 - lane) & 31;
      V z[32];
#pragma unroll
      for (int o = 0; o < 32; o++) {
        const V snd = (lane == 0) ? v[(32 - o) & 31] : v[31 - o];
        V vm;
        vm.x = shfl_t<T>(snd.x, partner);
        vm.y = shfl_t<T>(snd.y, partner);
        if (o == 0 && lane == 0) vm = vM;
        vm.y = -vm.y;                                                   // conj(V_{M-k})
        const int k = lane + 32 * o;
        const V s = cadd(v[o], vm);
        const V d = cmul(__ldg(ei + k), csub(v[o], vm));
        z[o] = mk2<T>(s.x - d.y, s.y + d.x);
      }
      fft1024_warp<T, true>(z, buf, lane, tw32);
      // ---- undo the Makhoul reordering into the swizzled time-domain staging ----
#pragma unroll
      for (int o = 0; o < 32; o++) {
        const int m = lane + 32 * o;
        const int n0 = (o < 16) ? 4 * m : (2 * FN - 1 - 4 * m);
        const int n1 = (o < 16) ? 4 * m + 2 : (2 * FN - 3 - 4 * m);
        X[swz(n0)] = z[o].x;
        X[swz(n1)] = z[o].y;
      }
    }
    __syncthreads();
    // ---- cooperative, coalesced store of PCM [n][c] ----
    for (int f = 0; f < nf; f++) {
      const size_t fbase = (size_t)(b0 + f) * words_per_frame;
      for (int e4 = tid; e4 < words_per_frame / 4; e4 += FW * 32) {
        T o4[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int e = 4 * e4 + j;
          const int n = e / C, c = e - n * C;
          o4[j] = ((const T*)(smem_raw + (size_t)(f * C + c) * LY::WARP_BYTES))[swz(n)];
        }
        if (P.pcm_dtype == FRAD_F64) {
          double2* dst = (double2*)((double*)P.pcm + fbase) + 2 * (size_t)e4;
          dst[0] = make_double2((double)o4[0], (double)o4[1]);
          dst[1] = make_double2((double)o4[2], (double)o4[3]);
        } else {
          float4* dst = (float4*)((float*)P.pcm + fbase) + e4;
          *dst = make_float4((float)o4[0], (float)o4[1], (float)o4[2], (float)o4[3]);
        }
      }
    }
  }
}

template <typename K>
int launch_cfg(frad_ctx* ctx, K kernel, size_t smem, int ntiles, int* grid) {
  FRAD_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  FRAD_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, FW * 32, smem));
  if (per_sm < 1) return frad_fail(ctx, FRAD_E_UNSUPPORTED, "fast path does not fit on this device");
  long long g = (long long)per_sm * ctx->sm_count;
  *grid = (int)(ntiles < g ? ntiles : g);
  if (*grid < 1) *grid = 1;
  return FRAD_OK;
}

}  // namespace

// eligibility checks shared with transform.cu
bool frad_fast2048_encode_ok(const enc_params& P, int prec) {
  if (P.N != FN || prec != FRAD_PREC_F32 || P.pcm_dtype != FRAD_F32) return false;
  if (P.C > FW) return false;
  if (P.profile == 0 && (P.bits != 32 || (P.payload_stride % 16) != 0 || ((uintptr_t)P.payload % 16) != 0)) return false;
  if (P.profile == 2) return false;
  if (P.profile == 1 && ((uintptr_t)P.q % 16) != 0) return false;
  if (((uintptr_t)P.pcm % 16) != 0 || ((P.hop * P.C) % 4) != 0) return false;
  return true;
}
bool frad_fast2048_decode_ok(const dec_params& P, int prec) {
  (void)prec;
  if (P.N != FN || P.C > FW) return false;
  if (P.profile == 0 && (P.bits != 32 || (P.payload_stride % 16) != 0 || ((uintptr_t)P.payload % 16) != 0)) return false;
  if (P.profile == 2) return false;
  if (P.profile == 1 && ((uintptr_t)P.q % 16) != 0) return false;
  if (((uintptr_t)P.pcm % 16) != 0) return false;
  return true;
}

#define ENC_DISPATCH(PROF, CT)                                                                         \
  do {                                                                                                 \
    auto k = enc2048_kernel<PROF, CT>;                                                                 \
    int grid;                                                                                          \
    int rc = launch_cfg(ctx, k, smem, tm.ntiles, &grid);                                               \
    if (rc) return rc;                                                                                 \
    k<<<grid, FW * 32, smem, ctx->stream>>>(P, plan_view(pl), (const float2*)pl->tw32, force_exact);   \
  } while (0)

int frad_launch_encode_2048(frad_ctx* ctx, const enc_params& P, const frad_plan* pl, int prec, int force_exact) {
  (void)prec;
  const tile_map tm = make_tiles(P.B, P.C);
  const size_t smem = (size_t)FW * SEQ_CPX * sizeof(float2) + sizeof(enc_smem_tail);
  if (P.profile == 0) {
    if (P.C == 2) ENC_DISPATCH(0, 2); else if (P.C == 1) ENC_DISPATCH(0, 1); else if (P.C == 8) ENC_DISPATCH(0, 8); else ENC_DISPATCH(0, 0);
  } else {
    if (P.C == 2) ENC_DISPATCH(1, 2); else if (P.C == 1) ENC_DISPATCH(1, 1); else if (P.C == 8) ENC_DISPATCH(1, 8); else ENC_DISPATCH(1, 0);
  }
  FRAD_LAUNCH_CHECK(ctx);
  return FRAD_OK;
}

#define DEC_DISPATCH(TT, PROF, CT)                                                                       \
  do {                                                                                                   \
    auto k = dec2048_kernel<TT, PROF, CT>;                                                               \
    int grid;                                                                                            \
    const size_t smem = (size_t)FW * dec_layout<TT>::WARP_BYTES + sizeof(dec_smem_tail);                 \
    int rc = launch_cfg(ctx, k, smem, tm.ntiles, &grid);                                                 \
    if (rc) return rc;                                                                                   \
    k<<<grid, FW * 32, smem, ctx->stream>>>(P, plan_view(pl), (const typename vec2<TT>::type*)pl->tw32); \
  } while (0)

#define DEC_DISPATCH_C(TT, PROF)                                                                       \
  do {                                                                                                 \
    if (P.C == 2) DEC_DISPATCH(TT, PROF, 2); else if (P.C == 1) DEC_DISPATCH(TT, PROF, 1);             \
    else if (P.C == 8) DEC_DISPATCH(TT, PROF, 8); else DEC_DISPATCH(TT, PROF, 0);                      \
  } while (0)

int frad_launch_decode_2048(frad_ctx* ctx, const dec_params& P, const frad_plan* pl, int prec) {
  const tile_map tm = make_tiles(P.B, P.C);
  if (prec == FRAD_PREC_F32) {
    if (P.profile == 0) DEC_DISPATCH_C(float, 0); else DEC_DISPATCH_C(float, 1);
  } else {
    if (P.profile == 0) DEC_DISPATCH_C(double, 0); else DEC_DISPATCH_C(double, 1);
  }
  FRAD_LAUNCH_CHECK(ctx);
  return FRAD_OK;
}
