// Decoder for the frame lengths N = 64 R1 < 2048 (R1 = 16, 8, 4, 2: N = 1024 ... 128 of fourier/profiles.py:14-23; the mixed radices below): S = 32 / R1
// sequences per warp, the mirror image of enc2048.cu's small-frame encoder.
//
//   staging    the tile's 32-bit words (quantised ints / profile-0 values) are copied to shared memory in memory order
//              (coalesced); lane L = s R1 + k1 then owns the 64 contiguous bins [64 k1, 64 k1 + 64) of sequence s -- in the padded
//              layout the run starts at word 68 L -- and dequantises them in place: |q|^(4/3) from a table, the band ramp
//              (profile1.py:65-74, p1tools.py:35-41) with the band parameters reloaded only where its static mask of band starts
//              says so (f64: highest group first, so the 8-byte values never overwrite integers that are still to be read)
//   transform  pairwise pre-processing (k, M-k) as in dct2048.cu, then fft32.cuh: ifft_multi_warp
//   store      time samples through a rotated staging layout, then cooperative 16-byte stores of the interleaved PCM
// R1 = 3, 5, 6, 7, 10, 12, 14, 20, 24, 28 (N = 192 ... 1792, the 160 / 192 / 224 families): S = floor(32 / R1) sequences per warp and the
// mixed-radix last pass of fftmix.cuh; the lanes at and above S R1 own no sequence before the exchange (see enc2048.cu).
// Reference lines: fourier/profile0.py:46-69; fourier/profile1.py:54-77; fourier/tools/p1tools.py:35-44.
#include "tile2048.cuh"

namespace {

constexpr int DSEQ = 1088;                            // complex slots per warp buffer (2176 T slots: padded spectrum / staging)
__device__ __forceinline__ int xposd(int k) { return k + ((k >> 6) << 2); }
template <typename T> __device__ __forceinline__ int qposd(int k) {
  if constexpr (sizeof(T) == 8) return 136 * (k >> 6) + (k & 63);
  else return xposd(k);
}
__device__ __forceinline__ int tpos(int n) { return n + (n >> 5); }     // time-domain staging: one pad slot per 32 samples

// per-CTA tables behind the sequence buffers.  Each arithmetic type keeps only its own ramp form: with both, the f64 kernel at
// S <= 2 (N >= 768) was 2.6 KB over three CTAs per SM
template <typename T, int S> struct decs_tail {       // f32
  int lo[FRAD_NBANDS + 1];
  float4 bp[FW][S][FRAD_NBANDS + 1];                  // per band: t0 / n, t1 / n, lo, hi
  T dq[1024];                                         // |q|^(4/3) / scale for small |q|
};
template <int S> struct decs_tail<double, S> {        // f64: exact np.linspace ramp
  int lo[FRAD_NBANDS + 1];
  double2 rp[FW][S][FRAD_NBANDS + 1];                 // per band: start, step
  double dq[256];
};

__device__ __noinline__ double dequant_big_s(const double* __restrict__ dq64, int qa) {
  return (qa < FRAD_DQ_LEN) ? __ldg(dq64 + qa) : pow((double)qa, 1.0 / 0.75);
}

template <typename T, int PROFILE, int R1>
__global__ void __launch_bounds__(FW * 32, (sizeof(T) == 8) ? 3 : FRAD_MINB)
decs_kernel(dec_params P, plan_dev pl, const typename vec2<T>::type* __restrict__ tw32, const double* __restrict__ dq64) {
  typedef typename vec2<T>::type V;
  constexpr bool F32 = sizeof(T) == 4;
  constexpr int NN = 64 * R1, MM = 32 * R1, S = 32 / R1, SPC = FW * S;
  constexpr int SEQ_T = 68 * R1;
  constexpr int RS = S * R1;                          // lanes that own a sequence before the exchange (32 for the powers of two)
  constexpr int DQN = F32 ? 1024 : 256;
  typedef decs_tail<T, S> TAIL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V* seq_all = (V*)smem_raw;
  TAIL* tail = (TAIL*)(smem_raw + (size_t)FW * DSEQ * sizeof(V));
  const int C = P.C;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int k1 = lane % R1, sl = (RS == 32) ? lane / R1 : min(lane / R1, S - 1);
  const bool act = (RS == 32) || lane < RS;
  const int fpt = SPC / C;
  const int ntiles = (P.B + fpt - 1) / fpt;
  const V* ci = (const V*)pl.ci;
  const V* ei = (const V*)pl.ei;
  const double inv_scale = 1.0 / P.scale;
  const bool p0_words = (PROFILE == 0) && P.bits == 32;

  int bfirst = 0;
  uint32_t fl_lo = 0, fl_hi = 0;
  if (PROFILE != 0) {
    if (tid <= FRAD_NBANDS) { const int e = P.bt->edge[tid]; tail->lo[tid] = (e < NN) ? e : NN; }
    for (int i = tid; i < DQN; i += FW * 32) {
      if constexpr (F32) { const float a = (float)i; tail->dq[i] = a * cbrtf(a) * (float)inv_scale; }
      else tail->dq[i] = __ldg(dq64 + i) * inv_scale;
    }
    __syncthreads();
#pragma unroll 1
    for (int b = 1; b < FRAD_NBANDS; b++) {
      const int d = tail->lo[b] - 64 * k1;
      if (d <= 0) bfirst = b;
      else if (d < 64 && tail->lo[b] < NN) { if (d < 32) fl_lo |= 1u << d; else fl_hi |= 1u << (d - 32); }
    }
  }

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b0 = tile * fpt;
    const int nf = min(fpt, P.B - b0);
    const int nseq = nf * C;
    __syncthreads();
    // ---- staging: the tile's frames in memory order; word of (frame f, bin k, channel c) -> sequence slot f C + c ----
    if ((PROFILE != 0 || p0_words) && C == 2) {
      // stereo: a 16-byte vector is bins k0, k0 + 1 of both channels -- two 8-byte stores, every divisor a compile-time constant
      constexpr int VPF = NN / 2;
      for (int e4 = tid; e4 < nf * VPF; e4 += FW * 32) {
        const int f = e4 / VPF, k0 = 2 * (e4 - f * VPF);
        const uint32_t* src = (PROFILE == 0) ? (const uint32_t*)(P.payload + (size_t)(b0 + f) * P.payload_stride)
                                             : (const uint32_t*)(P.q + (size_t)(b0 + f) * NN * 2);
        const uint4 x = ldg_stream((const uint4*)(src + 2 * k0));
        const int q0 = 2 * f, q1 = q0 + 1;
        uint32_t* bq0 = (uint32_t*)(seq_all + (size_t)(q0 / S) * DSEQ) + (size_t)(q0 % S) * SEQ_T * (sizeof(T) / 4);
        uint32_t* bq1 = (uint32_t*)(seq_all + (size_t)(q1 / S) * DSEQ) + (size_t)(q1 % S) * SEQ_T * (sizeof(T) / 4);
        const int pk = qposd<T>(k0);                       // even, and bin k0 + 1 is the next word
        *(uint2*)(bq0 + pk) = make_uint2(x.x, x.z);
        *(uint2*)(bq1 + pk) = make_uint2(x.y, x.w);
      }
    } else if ((PROFILE != 0 || p0_words) && C == 1) {
      // mono: a 16-byte vector is four consecutive bins of one sequence (their words stay together in the padded layouts)
      constexpr int VPF = NN / 4;
      for (int e4 = tid; e4 < nf * VPF; e4 += FW * 32) {
        const int f = e4 / VPF, k0 = 4 * (e4 - f * VPF);
        const uint32_t* src = (PROFILE == 0) ? (const uint32_t*)(P.payload + (size_t)(b0 + f) * P.payload_stride)
                                             : (const uint32_t*)(P.q + (size_t)(b0 + f) * NN);
        const uint4 x = ldg_stream((const uint4*)(src + k0));
        uint32_t* bq = (uint32_t*)(seq_all + (size_t)(f / S) * DSEQ) + (size_t)(f % S) * SEQ_T * (sizeof(T) / 4);
        *(uint4*)(bq + qposd<T>(k0)) = x;
      }
    } else if (PROFILE != 0 || p0_words) {
      for (int e4 = tid; e4 < nf * (NN * C / 4); e4 += FW * 32) {
        const int f = e4 / (NN * C / 4), vi = 4 * (e4 - f * (NN * C / 4));
        const uint32_t* src = (PROFILE == 0) ? (const uint32_t*)(P.payload + (size_t)(b0 + f) * P.payload_stride)
                                             : (const uint32_t*)(P.q + (size_t)(b0 + f) * NN * C);
        const uint4 x = ldg_stream((const uint4*)(src + vi));
        const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          int k, c; split_c(vi + j, C, k, c); const int q = f * C + c;
          uint32_t* base = (uint32_t*)(seq_all + (size_t)(q / S) * DSEQ) + (size_t)(q % S) * SEQ_T * (sizeof(T) / 4);
          if (PROFILE == 0) {
            if constexpr (F32) base[xposd(k)] = xs[j];          // converted in the run pass below
            else ((uint32_t*)base)[qposd<T>(k)] = xs[j];
          } else base[qposd<T>(k)] = xs[j];
        }
      }
    } else if constexpr (PROFILE == 0) {
      // profile 0 at the other depths: eight consecutive values from their bytes (profile0.py:51-66), stored as T
      const int bits = P.bits;
      const bool le = P.le != 0;
      for (int e8 = tid; e8 < nf * (NN * C / 8); e8 += FW * 32) {
        const int f = e8 / (NN * C / 8), vi = 8 * (e8 - f * (NN * C / 8));
        const uint8_t* frame = P.payload + (size_t)(b0 + f) * P.payload_stride;
        T v8[8];
        p0_load8<T>(frame + (size_t)(vi / 8) * bits, bits, le, v8);
#pragma unroll
        for (int j = 0; j < 8; j++) {
          int k, c; split_c(vi + j, C, k, c); const int q = f * C + c;
          ((T*)(seq_all + (size_t)(q / S) * DSEQ))[(size_t)(q % S) * SEQ_T + xposd(k)] = v8[j];
        }
      }
    }
    // ---- ramp parameters of the warp's sequences from their threshold integers (profile1.py:66), band = lane ----
    if (PROFILE != 0) {
      int tqv[S];                                          // all threshold integers first: one exposed load latency, not S
#pragma unroll
      for (int s2 = 0; s2 < S; s2++) {
        const int q2 = w * S + s2;
        int f2, c2;
        split_c(q2, C, f2, c2);
        tqv[s2] = (q2 < nseq && lane < FRAD_NBANDS) ? __ldg(P.tq + ((size_t)(b0 + f2) * FRAD_NBANDS + lane) * C + c2) : 0;
      }
#pragma unroll
      for (int s2 = 0; s2 < S; s2++) {
        if (w * S + s2 >= nseq) break;
        double thr = 0.0;
        if (lane < FRAD_NBANDS) thr = decode_threshold(P.bt, tqv[s2]);
        const double thr_next = __shfl_down_sync(0xffffffffu, thr, 1);
        if (lane <= FRAD_NBANDS) {
          const int n = (lane < FRAD_NBANDS - 1) ? tail->lo[lane + 1] - tail->lo[lane] : 0;
          if constexpr (F32) {
            tail->bp[w][s2][lane] = (n > 0) ? make_float4((float)(thr / n), (float)(thr_next / n), (float)tail->lo[lane], (float)tail->lo[lane + 1])
                                            : make_float4(0.0f, 0.0f, 0.0f, 1.0f);
          } else {
            tail->rp[w][s2][lane] = (n > 0) ? make_double2(thr, (thr_next - thr) / (double)n) : make_double2(0.0, 0.0);
          }
        }
      }
    }
    __syncthreads();
    {   // pull the next tile of this CTA towards L2 while the current one is transformed (its staging loads are one per iteration)
      const int nt = tile + gridDim.x;
      if (nt < ntiles) {
        const int nb0 = nt * fpt, nnf = min(fpt, P.B - nb0);
        const char* base;
        size_t bytes;
        if (PROFILE == 0) {
          base = (const char*)(P.payload + (size_t)nb0 * P.payload_stride);
          bytes = (size_t)(nnf - 1) * P.payload_stride + (size_t)NN * C * P.bits / 8;
        } else {
          base = (const char*)(P.q + (size_t)nb0 * NN * C);
          bytes = (size_t)nnf * NN * C * 4;
        }
        for (size_t o = (size_t)tid * 128; o < bytes; o += (size_t)FW * 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + o));
      }
    }
    if (w * S < nseq) {
      const int q_me = w * S + sl;
      V* buf = seq_all + (size_t)w * DSEQ;
      T* X = (T*)buf;
      T* Xs = X + SEQ_T * sl;
      // ---- the lane's run, converted in place: 32-bit words -> spectrum values of type T ----
      if (PROFILE != 0 || p0_words) {
        if (act) {
        const int* qrun = (const int*)X + qposd<T>(64 * lane);
        T* xrun = X + 68 * lane;
        int bnd = bfirst;
        // bands at the start of every group of eight (the f64 pass walks the groups downwards)
        int bstart[8];
#pragma unroll
        for (int g = 0; g < 8; g++) {
          bstart[g] = bnd;
          const uint32_t fb = ((g < 4) ? (fl_lo >> (8 * g)) : (fl_hi >> (8 * g - 32))) & (g == 0 ? 0xfeu : 0xffu);
          bnd += __popc(fb);
        }
#pragma unroll 1
        for (int gg = 0; gg < 8; gg++) {
          const int g = F32 ? gg : 7 - gg;
          const int4 a = ((const int4*)qrun)[2 * g], bq = ((const int4*)qrun)[2 * g + 1];
          const int qi[8] = {a.x, a.y, a.z, a.w, bq.x, bq.y, bq.z, bq.w};
          const uint32_t fb = ((g < 4) ? (fl_lo >> (8 * g)) : (fl_hi >> (8 * g - 32))) & (g == 0 ? 0xfeu : 0xffu);
          T out[8];
          if (PROFILE == 0) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
              const uint32_t u0 = (uint32_t)qi[u];
              const uint32_t uu = P.le ? u0 : __byte_perm(u0, 0, 0x0123);
              out[u] = (T)(((uu & 0x7f800000u) == 0x7f800000u) ? 0.0f : __uint_as_float(uu));      // NaN / Inf -> 0 (profile0.py:66)
            }
          } else {
            int bn = bstart[g];
            // band parameters live in registers and are reloaded only where a band starts inside the group
            float4 bpv = make_float4(0.0f, 0.0f, 0.0f, 1.0f);
            double2 rp = make_double2(0.0, 0.0);
            int blo = 0;
            if constexpr (F32) bpv = tail->bp[w][sl][bn]; else { rp = tail->rp[w][sl][bn]; blo = tail->lo[bn]; }
#pragma unroll
            for (int u = 0; u < 8; u++) {
              if ((fb >> u) & 1u) {
                bn++;
                if constexpr (F32) bpv = tail->bp[w][sl][bn]; else { rp = tail->rp[w][sl][bn]; blo = tail->lo[bn]; }
              }
              const int k = 64 * k1 + 8 * g + u;
              const int qv = qi[u];
              const int qa = abs(qv);
              if constexpr (F32) {
                const float mag = (qa < DQN) ? tail->dq[qa] : (float)qa * cbrtf((float)qa) * (float)inv_scale;
                const float kf = (float)k;
                const float dv = fmaf(kf - bpv.z, bpv.y, (bpv.w - kf) * bpv.x);
                out[u] = ((qv < 0) ? -mag : mag) * dv;
              } else {
                const double mag = (qa < DQN) ? tail->dq[qa] : dequant_big_s(dq64, qa) * inv_scale;
                const double div = __dadd_rn(__dmul_rn((double)(k - blo), rp.y), rp.x);   // j * step + start (np.linspace); 0 past the ramps
                out[u] = ((qv < 0) ? -mag : mag) * div;
              }
            }
          }
          if constexpr (F32) {
            ((float4*)xrun)[2 * g] = make_float4(out[0], out[1], out[2], out[3]);
            ((float4*)xrun)[2 * g + 1] = make_float4(out[4], out[5], out[6], out[7]);
          } else {
#pragma unroll
            for (int u2 = 0; u2 < 4; u2++) ((double2*)xrun)[4 * g + u2] = make_double2(out[2 * u2], out[2 * u2 + 1]);
          }
        }
        }
        __syncwarp();
      }
      // ---- pairwise pre-processing: R1 = X[k] - i X[N-k], R2 = (X[M-k] + i X[M+k]) e^{-i pi/4},
      //      S = ci (R1 + R2), G = ei ci (R1 - R2), Z[k] = S + iG, Z[M-k] = conj(S - iG),  k = k1 + R1 o ----
      V z[32];
      V zm_keep[17];
      const T r2 = (T)FRAD_RSQRT2;
      V ci_lane = mk2<T>(T(0), T(0)), ei_lane = ci_lane;
      if constexpr (!F32) { ci_lane = __ldg(ci + k1); ei_lane = __ldg(ei + k1); }
      auto pre_pair = [&](int k, int o, V& zk, V& zmk) {
        const T xk = Xs[xposd(k)], xnk = (k == 0) ? T(0) : Xs[xposd(NN - k)];
        const T a = Xs[xposd(MM - k)], bb = Xs[xposd(MM + k)];
        const V Ra = mk2<T>(xk, -xnk);
        const V Rb = mk2<T>((a + bb) * r2, (bb - a) * r2);
        V cik, eik;
        if constexpr (!F32) {
          cik = cmul(ci_lane, mk2<T>((T)g128_cos(o), (T)g128_sin(o)));
          eik = cmul(ei_lane, mk2<T>((T)g32_cos(o), (T)g32_sin(o)));
        } else {
          cik = __ldg(ci + k);
          eik = __ldg(ei + k);
        }
        const V Sv = cmul(cik, cadd(Ra, Rb));
        const V G = cmul(eik, cmul(cik, csub(Ra, Rb)));
        zk = mk2<T>(Sv.x - G.y, Sv.y + G.x);
        zmk = mk2<T>(Sv.x + G.y, -(Sv.y - G.x));
      };
#pragma unroll
      for (int o = 0; o < 16; o++) pre_pair(k1 + R1 * o, o, z[o], zm_keep[o]);
      zm_keep[16] = mk2<T>(T(0), T(0));
      if (k1 == 0) { V dummy; pre_pair(MM / 2, 16, zm_keep[16], dummy); }        // Z[M/2] (self pair)
      __syncwarp();
      const int partner = sl * R1 + ((R1 - k1) % R1);
#pragma unroll
      for (int o = 0; o < 16; o++) {
        V rcv;
        rcv.x = shfl_t<T>(zm_keep[o].x, partner);
        rcv.y = shfl_t<T>(zm_keep[o].y, partner);
        zm_keep[o] = rcv;
      }
#pragma unroll
      for (int j = 16; j < 32; j++) {
        const V g = zm_keep[31 - j];
        const V l0 = (j == 16) ? zm_keep[16] : zm_keep[32 - j];
        z[j] = (k1 == 0) ? l0 : g;
      }
      ifft_multi_warp<T, R1>(z, buf, lane, tw32);
      // ---- undo the Makhoul reordering into the time-domain staging: register s R1 + p holds z_s[m], m = lane + 32 perm_r(p);
      //      m < M/2: samples 4m, 4m + 2;  m >= M/2: samples 2N - 1 - 4m, 2N - 3 - 4m ----
#pragma unroll
      for (int i = 0; i < RS; i++) {
        const int s = i / R1, m = lane + 32 * perm_r<R1>(i % R1);
        T* ts = X + SEQ_T * s;
        const int na = (m < MM / 2) ? 4 * m : 2 * NN - 1 - 4 * m;
        const int nb = (m < MM / 2) ? 4 * m + 2 : 2 * NN - 3 - 4 * m;
        ts[tpos(na)] = z[i].x;
        ts[tpos(nb)] = z[i].y;
      }
    }
    __syncthreads();
    // ---- cooperative, coalesced store of PCM [n][c] ----
    const int vpf = (C == 2) ? NN / 2 : (C == 1) ? NN / 4 : NN * C / 4;         // 16-byte output vectors per frame
    for (int e4 = tid; e4 < nf * vpf; e4 += FW * 32) {
      int f, vi;
      T o4[4];
      if (C == 2) {                                          // samples n0, n0 + 1 of both channels; constant divisors
        f = e4 / (NN / 2); vi = 4 * (e4 - f * (NN / 2));
        const int q0 = 2 * f, q1 = q0 + 1;
        const T* t0 = (const T*)(seq_all + (size_t)(q0 / S) * DSEQ) + (size_t)(q0 % S) * SEQ_T;
        const T* t1 = (const T*)(seq_all + (size_t)(q1 / S) * DSEQ) + (size_t)(q1 % S) * SEQ_T;
        const int pn = tpos(vi >> 1);                        // n0 is even: n0 + 1 lies in the same group of 32
        o4[0] = t0[pn]; o4[1] = t1[pn]; o4[2] = t0[pn + 1]; o4[3] = t1[pn + 1];
      } else if (C == 1) {                                   // four consecutive samples of one sequence (same group of 32)
        f = e4 / (NN / 4); vi = 4 * (e4 - f * (NN / 4));
        const T* t0 = (const T*)(seq_all + (size_t)(f / S) * DSEQ) + (size_t)(f % S) * SEQ_T + tpos(vi);
        o4[0] = t0[0]; o4[1] = t0[1]; o4[2] = t0[2]; o4[3] = t0[3];
      } else {
        f = e4 / vpf; vi = 4 * (e4 - f * vpf);
#pragma unroll
        for (int j = 0; j < 4; j++) {
          int n, c; split_c(vi + j, C, n, c); const int q = f * C + c;
          o4[j] = ((const T*)(seq_all + (size_t)(q / S) * DSEQ))[(size_t)(q % S) * SEQ_T + tpos(n)];
        }
      }
      const size_t woff = (size_t)(b0 + f) * NN * C + vi;
      if (P.pcm_dtype == FRAD_F64) {
        if ((((uintptr_t)P.pcm) % 32) == 0) {               // one 32-byte store
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
      } else if (P.pcm_dtype == FRAD_F32) {
        *(float4*)((float*)P.pcm + woff) = make_float4((float)o4[0], (float)o4[1], (float)o4[2], (float)o4[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; j++) store_pcm<T>(P.pcm, P.pcm_dtype, (int64_t)(woff + j), o4[j]);
      }
    }
  }
}

}  // namespace

#define DECS_DISPATCH(TT, PROF, RR)                                                                          \
  do {                                                                                                       \
    auto k = decs_kernel<TT, PROF, RR>;                                                                      \
    const int fpt = FW * (32 / RR) / P.C;                                                                    \
    int grid;                                                                                                \
    const size_t smem = (size_t)FW * DSEQ * sizeof(vec2<TT>::type) + ((PROF == 0) ? 16 : sizeof(decs_tail<TT, 32 / RR>)); \
    int rc = launch_cfg(ctx, k, smem, (P.B + fpt - 1) / fpt, &grid);                                         \
    if (rc) return rc;                                                                                       \
    k<<<grid, FW * 32, smem, ctx->stream>>>(P, plan_view(pl), (const vec2<TT>::type*)pl->tw32, (const double*)ctx->dequant_table); \
  } while (0)

#ifdef FRAD_TU_MIXED
// ---- translation units decm32.cu / decm64.cu: the mixed-radix frame lengths, one arithmetic type each ----
#define DECM_BY_R(TT, PROF)                                                                             \
  do {                                                                                                  \
    switch (P.N / 64) {                                                                                 \
      case 3: DECS_DISPATCH(TT, PROF, 3); break;    case 5: DECS_DISPATCH(TT, PROF, 5); break;          \
      case 6: DECS_DISPATCH(TT, PROF, 6); break;    case 7: DECS_DISPATCH(TT, PROF, 7); break;          \
      case 10: DECS_DISPATCH(TT, PROF, 10); break;  case 12: DECS_DISPATCH(TT, PROF, 12); break;        \
      case 14: DECS_DISPATCH(TT, PROF, 14); break;  case 20: DECS_DISPATCH(TT, PROF, 20); break;        \
      case 24: DECS_DISPATCH(TT, PROF, 24); break;  case 28: DECS_DISPATCH(TT, PROF, 28); break;        \
      default: return frad_fail(ctx, FRAD_E_ARG, "mixed-radix decoder: frame length");                  \
    }                                                                                                   \
  } while (0)
#if FRAD_TU_MIXED == 32
int frad_launch_decode_mixed_f32(frad_ctx* ctx, const dec_params& P, const frad_plan* pl) {
  if (P.profile == 0) DECM_BY_R(float, 0); else DECM_BY_R(float, 1);
  FRAD_LAUNCH_CHECK(ctx);
  return FRAD_OK;
}
#else
int frad_launch_decode_mixed_f64(frad_ctx* ctx, const dec_params& P, const frad_plan* pl) {
  if (P.profile == 0) DECM_BY_R(double, 0); else DECM_BY_R(double, 1);
  FRAD_LAUNCH_CHECK(ctx);
  return FRAD_OK;
}
#endif
#else   // !FRAD_TU_MIXED

bool frad_fast_small_decode_ok(const dec_params& P, int prec) {
  (void)prec;
  const int r1 = frad_warp_r1(P.N);
  if (!r1) return false;
  const int spc = FW * (32 / r1);
  if (P.C < 1 || P.C > spc || (spc % P.C) != 0) return false;
  if (((size_t)P.N * P.C) % 8) return false;
  if (P.profile == 2 || P.profile == 4) return false;
  if (P.profile == 0 && ((P.payload_stride % 16) != 0 || ((uintptr_t)P.payload % 16) != 0)) return false;
  if (P.profile == 1 && ((uintptr_t)P.q % 16) != 0) return false;
  if (((uintptr_t)P.pcm % 16) != 0) return false;
  return true;
}

#define DECS_BY_R(TT, PROF)                                                                             \
  do {                                                                                                  \
    if (P.N == 1024) DECS_DISPATCH(TT, PROF, 16); else if (P.N == 512) DECS_DISPATCH(TT, PROF, 8);      \
    else if (P.N == 256) DECS_DISPATCH(TT, PROF, 4); else DECS_DISPATCH(TT, PROF, 2);                 \
  } while (0)

int frad_launch_decode_mixed_f32(frad_ctx* ctx, const dec_params& P, const frad_plan* pl);   // decm32.cu
int frad_launch_decode_mixed_f64(frad_ctx* ctx, const dec_params& P, const frad_plan* pl);   // decm64.cu

int frad_launch_decode_small(frad_ctx* ctx, const dec_params& P, const frad_plan* pl, int prec) {
  int rc = frad_ensure_dq64(ctx);
  if (rc) return rc;
  if (frad_warp_r1_mixed(frad_warp_r1(P.N)))
    return (prec == FRAD_PREC_F32) ? frad_launch_decode_mixed_f32(ctx, P, pl) : frad_launch_decode_mixed_f64(ctx, P, pl);
  if (prec == FRAD_PREC_F32) { if (P.profile == 0) DECS_BY_R(float, 0); else DECS_BY_R(float, 1); }
  else { if (P.profile == 0) DECS_BY_R(double, 0); else DECS_BY_R(double, 1); }
  FRAD_LAUNCH_CHECK(ctx);
  return FRAD_OK;
}
#endif  // FRAD_TU_MIXED
