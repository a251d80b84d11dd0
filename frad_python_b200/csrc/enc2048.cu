// N = 2048 encoder, "run" epilogue (round 2).  Same transform as dct2048.cu's first encoder -- one warp per (frame, channel)
// sequence, 1024-point complex FFT as 32 x 32 in registers -- but everything after the FFT works on whole runs:
//
//   * the spectrum goes to shared memory in a padded layout, bin k at word k + 4 (k >> 6): lane l then owns the 64
//     CONTIGUOUS bins [64 l, 64 l + 64) and fetches them with sixteen conflict-free 16-byte loads;
//   * which band a bin belongs to is static per kernel (N and sample rate fix the edges, p1tools.py:15-16): a lane keeps
//     the band of its first bin and a 64-bit mask of the bins where a new band starts -- no per-bin band look-ups;
//   * band energies (p1tools.py:22-32), the divisor ramp (p1tools.py:35-41) and the power-law quantiser (profile1.py:34-36)
//     run out of registers with the band parameters reloaded only where the mask says so, and the integers go back with
//     sixteen 16-byte stores;
//   * global I/O of the stereo tiling is 32 bytes per thread and instruction (LDG.256 / STG.256): four samples x two
//     channels in, four bins x two channels out.
//
// Measured against the first kernel (15 000 stereo frames, B200): see profiles/README.md.
// Reference lines: fourier/profile0.py:21-30; fourier/profile1.py:21-40; fourier/tools/p1tools.py:18-44.
#include "tile2048.cuh"

namespace {

constexpr int ESEQ = 1088;                           // complex slots per warp buffer: 2176 words = the padded spectrum layout
__device__ __forceinline__ int xpos(int k) { return k + ((k >> 6) << 2); }

struct enc_tail {
  int lo[FRAD_NBANDS + 1];
  double thr[FW][FRAD_NBANDS + 1];
  double inv_ns[FRAD_NBANDS + 1];                     // 1 / (band width x scale)
  float4 bp[FW][FRAD_NBANDS + 1];                     // per band: ramp end points (pre-scaled), lo, hi as floats
  float part[FW][64];                                 // (lane, band) partial energies at slot lane + band (a staircase)
};

// exact quantiser of one coefficient (the f64 formula of profile1.py:31-36 with np.linspace's ramp)
__device__ __noinline__ int quant_exact_r(float x, int j, double t0, double step, double scale) {
  const double div = __dadd_rn(__dmul_rn((double)j, step), t0);
  if (div == 0.0) return 0;
  const double yy = fabs(((double)x / div) * scale);
  const double rr = rint(sqrt(yy * sqrt(yy)));
  if (!(rr == rr)) return 0;
  const int m = (rr >= 2147483647.0) ? 2147483647 : (int)rr;
  return (x < 0.0f) ? -m : m;
}
__device__ __noinline__ int tq_exact_r(double thr, double log_base) {
  const double vv = log(fmax(thr, 1.0)) / log_base;
  const double d = (vv > 0.0) ? vv * cbrt(vv) : 0.0;                     // v^(4/3)
  return sat_i32(rint(d));
}

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

template <int PROFILE, int CT>
__global__ void __launch_bounds__(FW * 32, FRAD_MINB)
enc2048r_kernel(enc_params P, plan_dev pl, const float2* __restrict__ tw32, int force_exact, int wide) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* seq_all = (float2*)smem_raw;
  enc_tail* tail = (enc_tail*)(smem_raw + (size_t)FW * ESEQ * sizeof(float2));
  const int C = (CT == 1 || CT == 2) ? CT : P.C;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const tile_map tm = make_tiles<CT>(P.B, C);
  const float* pcm = (const float*)P.pcm;
  const float2* cfh = (const float2*)pl.cfh;
  const float2* dfh = (const float2*)pl.dfh;
  const int words_per_frame = FN * C;

  // ---- static run description of this lane: band of bin 64 lane, and the bins of the run where a new band starts ----
  int bfirst = 0;
  uint32_t fl_lo = 0, fl_hi = 0;
  if (PROFILE != 0) {
    if (tid <= FRAD_NBANDS) { const int e = P.bt->edge[tid]; tail->lo[tid] = (e < FN) ? e : FN; }
    __syncthreads();
    if (tid < FRAD_NBANDS) {
      const int n = tail->lo[tid + 1] - tail->lo[tid];
      tail->inv_ns[tid] = (n > 0) ? 1.0 / ((double)n * P.scale) : 0.0;
    }
#pragma unroll 1
    for (int b = 1; b < FRAD_NBANDS; b++) {
      const int d = tail->lo[b] - 64 * lane;
      if (d <= 0) bfirst = b;                            // lo[] ascends: the last b with lo[b] <= first bin of the run
      else if (d < 64 && tail->lo[b] < FN) { if (d < 32) fl_lo |= 1u << d; else fl_hi |= 1u << (d - 32); }
    }
  }
  auto flag = [&](int j) -> bool { return ((j < 32 ? (fl_lo >> j) : (fl_hi >> (j - 32))) & 1u) != 0; };

  for (int tile = blockIdx.x; tile < tm.ntiles; tile += gridDim.x) {
    const int b0 = (CT == 4) ? tile / tm.groups : tile * tm.fpc;
    const int grp = (CT == 4) ? tile - b0 * tm.groups : 0;
    const int nf = min(tm.fpc, P.B - b0);
    const int nseq = (CT == 4) ? 4 : nf * C;              // sequences (= busy warps) in this tile
    __syncthreads();                                      // previous tile fully stored before restaging
    if (CT == 2 && wide) {
      // thread t takes the sample quads m = t + 128 i of both frames: 32 bytes = samples 4m .. 4m+3 of both channels.
      // Makhoul order: z[m] = (x[4m], x[4m+2]), z[1023 - m] = (x[4m+3], x[4m+1])  (see zword)
      w8 qd[8];
#pragma unroll
      for (int f = 0; f < 2; f++) {
        const float* frame = pcm + (size_t)(b0 + (f < nf ? f : nf - 1)) * P.hop * 2;
#pragma unroll
        for (int i = 0; i < 4; i++) qd[4 * f + i] = ldg256_stream(frame + 8 * (tid + 128 * i));
      }
#pragma unroll
      for (int f = 0; f < 2; f++) {
        if (f < nf) {
          uint2* s0 = (uint2*)(seq_all + (size_t)(2 * f) * ESEQ);
          uint2* s1 = (uint2*)(seq_all + (size_t)(2 * f + 1) * ESEQ);
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int m = tid + 128 * i;
            const w8& q = qd[4 * f + i];
            s0[padc(m)] = make_uint2(q.v[0], q.v[4]);
            s0[padc(FM - 1 - m)] = make_uint2(q.v[6], q.v[2]);
            s1[padc(m)] = make_uint2(q.v[1], q.v[5]);
            s1[padc(FM - 1 - m)] = make_uint2(q.v[7], q.v[3]);
          }
        }
      }
    } else {
      stage_tile<CT, ESEQ * 2>([&](int f) { return (const uint32_t*)(pcm + (size_t)(b0 + f) * P.hop * C); }, (uint32_t*)seq_all, C, nf,
                               grp, tid, [](int n) { return zword(n); });
    }
    __syncthreads();
    {   // pull the next tile of this CTA towards L2 while the current one is transformed
      const int nt = tile + gridDim.x;
      if (nt < tm.ntiles) {
        const int nb0 = (CT == 4) ? nt / tm.groups : nt * tm.fpc;
        const int nnf = min(tm.fpc, P.B - nb0);
        const char* base = (const char*)(pcm + (size_t)nb0 * P.hop * C);
        const size_t bytes = (size_t)nnf * words_per_frame * 4;
        for (size_t o = (size_t)tid * 128; o < bytes; o += (size_t)FW * 32 * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + o));
      }
    }
    if (w < nseq) {
      const int f = (CT == 4) ? 0 : w / C, c = (CT == 4) ? 4 * grp + w : w - f * C, b = b0 + f;
      float2* buf = seq_all + (size_t)w * ESEQ;
      float* X = (float*)buf;
      float2 v[32];
#pragma unroll
      for (int i = 0; i < 32; i++) v[i] = buf[lane + 33 * i];
      __syncwarp();
      fft1024_warp<float, false>(v, buf, lane, tw32);
      __syncwarp();                                       // every lane has read its exchange rows: the buffer becomes X
      // ---- pairwise real-FFT split + DCT twiddle: (k, M-k) -> X[k], X[N-k], X[M-k], X[M+k], k = lane + 32 o.
      //      Padded positions: two steps of o move 64 bins = 68 words, so every store takes an immediate offset from one of
      //      eight per-lane bases (k, M+k ascending; N-k, M-k descending; even and odd o) ----
      const int partner = (32 - lane) & 31;
      const bool swap_out = (PROFILE == 0) && !P.le;
      float mx = 0.0f;
      const int oK[2] = {xpos(lane), xpos(lane + 32)};
      const int oMK[2] = {xpos(FM + lane), xpos(FM + lane + 32)};
      const int oNK[2] = {xpos(FN - lane), xpos(FN - lane - 32)};             // o = 0, lane 0 (bin 2048) is never stored
      const int oMmK[2] = {xpos(FM - lane), xpos(FM - lane - 32)};
      auto put = [&](int pos, float val) {
        if (PROFILE == 0) {
          mx = fmaxf(mx, fabsf(val));
          if (swap_out) val = __uint_as_float(__byte_perm(__float_as_uint(val), 0, 0x0123));
        }
        X[pos] = val;
      };
#pragma unroll
      for (int o = 0; o < 16; o++) {
        const float2 snd = (lane == 0) ? v[(32 - o) & 31] : v[31 - o];
        float2 zm;
        zm.x = __shfl_sync(0xffffffffu, snd.x, partner);
        zm.y = -__shfl_sync(0xffffffffu, snd.y, partner);                   // conj(Z[M-k])
        const int k = lane + 32 * o;
        const bool first = (o == 0) && (lane == 0);
        const float2 A = cadd(v[o], zm), D = csub(v[o], zm);
        const float2 Bv = make_float2(D.y, -D.x);
        const float2 Pk = cmul(__ldg(cfh + k), A), Qk = cmul(__ldg(dfh + k), Bv);
        const float2 y = cadd(Pk, Qk), t = csub(Pk, Qk);
        const int h = o >> 1, e = o & 1;
        put(oK[e] + 68 * h, y.x);
        if (!first) put(oNK[e] - 68 * h, -y.y);
        put(oMmK[e] - 68 * h, (t.x - t.y) * (float)FRAD_RSQRT2);
        if (!first) put(oMK[e] + 68 * h, (t.x + t.y) * (float)FRAD_RSQRT2);
      }
      if (lane == 0) {                                                     // self-paired bin k = 512 (and N - 512)
        const float2 zk = v[16], zm = cconj(v[16]);
        const float2 A = cadd(zk, zm), D = csub(zk, zm);
        const float2 Bv = make_float2(D.y, -D.x);
        const float2 y = cadd(cmul(__ldg(cfh + FM / 2), A), cmul(__ldg(dfh + FM / 2), Bv));
        put(xpos(FM / 2), y.x);
        put(xpos(FN - FM / 2), -y.y);
      }
      if (PROFILE == 0) {
        if (P.peak) {
#pragma unroll
          for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
          if (lane == 0) atomicMax((unsigned long long*)&P.peak[b], (unsigned long long)__double_as_longlong((double)mx));
        }
      } else {
        __syncwarp();
        // ---- the lane's run: bins [64 lane, 64 lane + 64), sixteen 16-byte loads (conflict-free: 68-word pitch).
        //      Band energies of the spectrum (scaled below): running sum, closed where a new band starts ----
        const float4* run = (const float4*)(X + 68 * lane);
        {
          float* slot = tail->part[w] + lane + bfirst;
          float acc = 0.0f;
#pragma unroll
          for (int i = 0; i < 16; i++) {
            const float4 t = run[i];
            const float xe[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const int j = 4 * i + e;
              if (j > 0 && flag(j)) { *slot = acc; slot++; acc = 0.0f; }
              acc = fmaf(xe[e], xe[e], acc);
            }
          }
          *slot = acc;
        }
        __syncwarp();
        const float scalef = (float)P.scale;
        const float inv_log_base_f = (float)(1.0 / 0.30685281944005469058);   // 1 / ln(e/2)
        float tot = 0.0f;
        int cnt = 1;
        if (lane < FRAD_NBANDS) {
          const int blo = tail->lo[lane], bhi = tail->lo[lane + 1];
          cnt = bhi - blo;
          if (bhi > blo) {
            const float* row = tail->part[w] + lane;
            for (int l = blo >> 6; l <= (bhi - 1) >> 6; l++) tot += row[l];
          }
          tot *= scalef * scalef;
        }
        const unsigned empty = __ballot_sync(0xffffffffu, cnt <= 0);
        const int first_empty = empty ? (__ffs(empty) - 1) : 32;
        double thr = 0.0;
        if (lane < FRAD_NBANDS && lane < first_empty) {
          // f32 like the reference for f32 spectra: np.sqrt(np.mean(x**2)) ** 0.8 on float32 (p1tools.py:32)
          const float level = powf(sqrtf(tot / (float)cnt), 0.8f);
          thr = fmax((double)level, P.bt->ath[lane]) * P.loss;
        }
        if (lane <= FRAD_NBANDS) tail->thr[w][lane] = thr;             // entry 27 = 0
        if (lane < FRAD_NBANDS) {
          // threshold integer round((ln(max(thr, 1)) / ln(e/2))^(4/3)) (profile1.py:38-40).  f32 first: the relative error of
          // logf / cbrtf and of the operand rounding is < 1e-6, so the integer is already decided unless d lies within
          // 1e-3 (1 + d) of a rounding tie (or is huge) -- only then the f64 formula runs
          int tqi = 0;
          bool exact = false;
          if (thr > 1.0) {
            const float lv = logf((float)thr) * inv_log_base_f;
            const float df = lv * cbrtf(lv);
            const float rf = rintf(df);
            exact = !(fabsf(fabsf(df - rf) - 0.5f) > 1.0e-3f * (1.0f + df) && df < 1.0e6f);
            tqi = (int)rf;
          }
          if (exact) tqi = tq_exact_r(thr, P.bt->log_base);
          P.tq[((size_t)b * FRAD_NBANDS + lane) * C + c] = tqi;
        }
        if constexpr (PROFILE == 2) {
          // profile 2: TNS works on the whole masked spectrum in f64 (profile2.py:32) -- tns_encode_kernel (transform.cu) takes
          // over from here with the spectrum (stored below as float bits in the q array) and this sequence's thresholds
          if (lane <= FRAD_NBANDS) P.tns_scratch[((size_t)b * C + c) * (FRAD_NBANDS + 1) + lane] = thr;
        } else {
          // per-band ramp parameters, one band per lane: the divisor over the scale as a convex combination
          //   div / scale = t0s (hi - k) + t1s (k - lo),  t0s = thr[b] / (n scale), t1s = thr[b+1] / (n scale)
          {
            const double thr_next = __shfl_down_sync(0xffffffffu, thr, 1);
            if (lane <= FRAD_NBANDS) {
              float4 bpv = make_float4(1.0e30f, 1.0e30f, -1.0f, 2.0f * FN);  // huge divisor: bins past the last ramp quantise to 0
              if (lane < FRAD_NBANDS - 1 && cnt > 0) {
                const double ins = tail->inv_ns[lane];                      // 1 / (n scale)
                bpv = make_float4((float)(thr * ins), (float)(thr_next * ins), (float)tail->lo[lane], (float)tail->lo[lane + 1]);
              }
              tail->bp[w][lane] = bpv;
            }
          }
          __syncwarp();
          // ---- quantise the run in registers, eight coefficients at a time (independent chains for the MUFU latencies); the
          //      rare ones next to a rounding tie are redone with the exact f64 formula ----
          const float4* bpp = tail->bp[w] + bfirst;
          float4 bpv = *bpp;
          const float kb = (float)(64 * lane);
          int4* qrun = (int4*)(X + 68 * lane);
#pragma unroll 1
          for (int g = 0; g < 8; g++) {
            int ms[8];
            bool allok = true;
            const float4 xa = run[2 * g], xb = run[2 * g + 1];
            const float xs[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            const uint32_t fb = ((g < 4) ? (fl_lo >> (8 * g)) : (fl_hi >> (8 * g - 32))) & (g == 0 ? 0xfeu : 0xffu);   // band starts in this group
            const float kg = kb + (float)(8 * g);
#pragma unroll
            for (int u = 0; u < 8; u++) {
              const int j = 8 * g + u;
              if ((fb >> u) & 1u) { bpp++; bpv = *bpp; }
              const float kf = kg + (float)u;
              const float x = xs[u];
              const float dv = fmaf(kf - bpv.z, bpv.y, (bpv.w - kf) * bpv.x);
              float rc, s, vv;
              asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(dv));
              const float y = fabsf(x) * rc;
              asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(y));
              asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(vv) : "f"(y * s));
              // signed magic-number rounding: nearest-even integer of +-v for v < 2^22, read from the mantissa
              const float big = copysignf(vv, x) + 12582912.0f;
              const float r = big - 12582912.0f;
              // error budget of the f32 branch <= 6e-7 relative: redo exactly when closer than 1e-6 v
              // to a tie.  Inf / NaN (zero divisor) and v >= 5e5 fail this test by themselves.
              allok = allok && (fmaf(1.0e-6f, vv, fabsf(fabsf(vv) - fabsf(r))) < 0.4999999f);
              ms[u] = __float_as_int(big) - 0x4B400000;
            }
            if (!allok || force_exact) {
#pragma unroll
              for (int u = 0; u < 8; u++) {
                const int j = 8 * g + u;
                // band of bin j of the run: the first one plus the band starts up to j
                const unsigned long long fl = ((unsigned long long)fl_hi << 32) | fl_lo;
                const int bn = bfirst + __popcll(fl & ((j >= 63) ? ~0ull : ((2ull << j) - 1ull)));
                bool redo = force_exact != 0;
                if (!redo) {
                  const float4 bq = tail->bp[w][bn];
                  const float kf = kg + (float)u;
                  const float dv = fmaf(kf - bq.z, bq.y, (bq.w - kf) * bq.x);
                  float rc, s, vv;
                  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(dv));
                  const float y = fabsf(xs[u]) * rc;
                  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(y));
                  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(vv) : "f"(y * s));
                  const float r = (vv + 12582912.0f) - 12582912.0f;
                  redo = !(fmaf(1.0e-6f, vv, fabsf(vv - r)) < 0.4999999f);
                }
                if (redo && bn != FRAD_NBANDS - 1) {
                  const int lo = tail->lo[bn], n = tail->lo[bn + 1] - lo;
                  const double t0 = tail->thr[w][bn], t1 = tail->thr[w][bn + 1];
                  ms[u] = quant_exact_r(xs[u], 64 * lane + j - lo, t0, (t1 - t0) / (double)n, P.scale);
                }
              }
            }
            qrun[2 * g] = make_int4(ms[0], ms[1], ms[2], ms[3]);
            qrun[2 * g + 1] = make_int4(ms[4], ms[5], ms[6], ms[7]);
          }
        }   // PROFILE == 1
      }
    }
    __syncthreads();
    // ---- cooperative, coalesced store: word e = k*C + c of frame f <- seq[f*C + c][xpos(k)] ----
    if (CT == 2 && wide) {
#pragma unroll
      for (int f = 0; f < 2; f++) {
        if (f < nf) {
          uint32_t* dst = (PROFILE == 0) ? (uint32_t*)(P.payload + (size_t)(b0 + f) * P.payload_stride)
                                         : (uint32_t*)(P.q + (size_t)(b0 + f) * words_per_frame);
          const uint32_t* X0 = (const uint32_t*)(seq_all + (size_t)(2 * f) * ESEQ);
          const uint32_t* X1 = (const uint32_t*)(seq_all + (size_t)(2 * f + 1) * ESEQ);
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const int qd = tid + 128 * i;                                 // bins 4 qd .. 4 qd + 3 of both channels
            const uint4 a = *(const uint4*)(X0 + xpos(4 * qd));
            const uint4 bq = *(const uint4*)(X1 + xpos(4 * qd));
            w8 o;
            o.v[0] = a.x; o.v[1] = bq.x; o.v[2] = a.y; o.v[3] = bq.y; o.v[4] = a.z; o.v[5] = bq.z; o.v[6] = a.w; o.v[7] = bq.w;
            stg256(dst + 8 * qd, o);
          }
        }
      }
    } else {
      for (int f = 0; f < nf; f++) {
        uint32_t* dst = (PROFILE == 0) ? (uint32_t*)(P.payload + (size_t)(b0 + f) * P.payload_stride)
                                       : (uint32_t*)(P.q + (size_t)(b0 + f) * words_per_frame);
        for (int e4 = tid; e4 < vec_count<CT>(C); e4 += FW * 32) {
          uint32_t o4[4];
          gather4<CT, ESEQ * 2>((const uint32_t*)seq_all, C, f, e4, [](int k) { return xpos(k); }, [](uint32_t u) { return u; }, o4);
          *(uint4*)(dst + vec_word<CT>(e4, C, grp)) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        }
      }
    }
  }
}

}  // namespace

#define ENCR_DISPATCH(PROF, CT)                                                                              \
  do {                                                                                                       \
    auto k = enc2048r_kernel<PROF, CT>;                                                                      \
    const tile_map tm = make_tiles<CT>(P.B, P.C);                                                            \
    int grid;                                                                                                \
    int rc = launch_cfg(ctx, k, smem, tm.ntiles, &grid);                                                     \
    if (rc) return rc;                                                                                       \
    k<<<grid, FW * 32, smem, ctx->stream>>>(P, plan_view(pl), (const float2*)pl->tw32, force_exact, wide);   \
  } while (0)

int frad_launch_encode_2048r(frad_ctx* ctx, const enc_params& P, const frad_plan* pl, int prec, int force_exact) {
  (void)prec;
  const size_t smem = (size_t)FW * ESEQ * sizeof(float2) + ((P.profile == 0) ? 16 : sizeof(enc_tail));
  // 32-byte global accesses of the stereo tiling need 32-byte aligned frames on both sides
  const void* out = (P.profile == 0) ? (const void*)P.payload : (const void*)P.q;
  const int wide = (P.C == 2 && ((uintptr_t)P.pcm % 32) == 0 && ((P.hop * 2) % 8) == 0 && ((uintptr_t)out % 32) == 0 &&
                    (P.profile != 0 || (P.payload_stride % 32) == 0)) ? 1 : 0;
#define ENCR_BY_C(PROF)                                                                                 \
  do {                                                                                                  \
    if (P.C == 2) ENCR_DISPATCH(PROF, 2); else if (P.C == 1) ENCR_DISPATCH(PROF, 1);                    \
    else if (P.C % 4 == 0) ENCR_DISPATCH(PROF, 4); else ENCR_DISPATCH(PROF, 0);                         \
  } while (0)
  if (P.profile == 0) ENCR_BY_C(0); else if (P.profile == 1) ENCR_BY_C(1); else ENCR_BY_C(2);   // 2: P.tns_scratch = thresholds out
  FRAD_LAUNCH_CHECK(ctx);
  return FRAD_OK;
}
