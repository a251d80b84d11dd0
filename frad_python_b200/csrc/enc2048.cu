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
// The kernel is a template over the arithmetic type: float for f16 / f32 PCM, double for integer and f64 PCM -- the reference's
// rule (scipy transforms in the input's precision, and to_f64 makes every integer format float64) -- and reads
// every sample format of backend/pcmformat.py:4-32.  Profile 0 packs all six depths of profile0.py:28-42 in its store phase.
// Measured against the first kernel (15 000 stereo frames, B200): see profiles/README.md.
// Reference lines: fourier/profile0.py:21-30; fourier/profile1.py:21-40; fourier/tools/p1tools.py:18-44.
#include "tile2048.cuh"

namespace {

constexpr int ESEQ = 1088;                           // complex slots per warp buffer: 2176 words = the padded spectrum layout
__device__ __forceinline__ int xpos(int k) { return k + ((k >> 6) << 2); }

// int32 position of quantised bin k inside a sequence buffer: the lane that owns the run writes it into its own run
// (f32: in place, word for word; f64: the first half of the 8-byte run, so no lane writes where another still reads)
template <typename T> __device__ __forceinline__ int qpos(int k) {
  if constexpr (sizeof(T) == 8) return 136 * (k >> 6) + (k & 63);
  else return xpos(k);
}

template <typename T> struct enc_tail {
  int lo[FRAD_NBANDS + 1];
  double thr[FW][FRAD_NBANDS + 1];
  double inv_ns[FRAD_NBANDS + 1];                     // 1 / (band width x scale)
  float4 bp[FW][FRAD_NBANDS + 1];                     // per band: ramp end points (pre-scaled), lo, hi as floats
  T part[FW][64];                                     // (lane, band) partial energies at slot lane + band (a staircase)
};

// exact quantiser of one coefficient (the f64 formula of profile1.py:31-36 with np.linspace's ramp)
__device__ __noinline__ int quant_exact_r(double x, int j, double t0, double step, double scale) {
  const double div = __dadd_rn(__dmul_rn((double)j, step), t0);
  if (div == 0.0) return 0;
  const double yy = fabs((x / div) * scale);
  const double rr = rint(sqrt(yy * sqrt(yy)));
  if (!(rr == rr)) return 0;
  const int m = (rr >= 2147483647.0) ? 2147483647 : (int)rr;
  return (x < 0.0) ? -m : m;
}
__device__ __noinline__ int tq_exact_r(double thr, double log_base) {
  const double vv = log(fmax(thr, 1.0)) / log_base;
  const double d = (vv > 0.0) ? vv * cbrt(vv) : 0.0;                     // v^(4/3)
  return sat_i32(rint(d));
}

// eight consecutive samples (index i0 .. i0 + 7 of the interleaved PCM, i0 a multiple of 8) of any format, as T
template <typename T> __device__ __forceinline__ void load_quad8(const void* pcm, int dt, int64_t i0, T* q) {
  switch (dt) {
    case FRAD_F32: {
      const uint4 a = __ldg((const uint4*)((const float*)pcm + i0)), b = __ldg((const uint4*)((const float*)pcm + i0) + 1);
      const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 8; j++) q[j] = (T)__uint_as_float(w[j]);
      break;
    }
    case FRAD_F64: {
      const double2* p = (const double2*)((const double*)pcm + i0);
#pragma unroll
      for (int j = 0; j < 4; j++) { const double2 d = __ldg(p + j); q[2 * j] = (T)d.x; q[2 * j + 1] = (T)d.y; }
      break;
    }
    case FRAD_S16: case FRAD_U16: case FRAD_F16: {
      const uint4 a = __ldg((const uint4*)((const uint16_t*)pcm + i0));
      const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int j = 0; j < 8; j++) {
        const uint32_t h = (w[j >> 1] >> (16 * (j & 1))) & 0xffffu;
        if (dt == FRAD_S16) q[j] = (T)((double)(int16_t)h / 32768.0);
        else if (dt == FRAD_U16) q[j] = (T)((double)h / 32768.0 - 1.0);
        else q[j] = (T)__half2float(__ushort_as_half((unsigned short)h));
      }
      break;
    }
    case FRAD_S32: case FRAD_U32: {
      const uint4 a = __ldg((const uint4*)((const uint32_t*)pcm + i0)), b = __ldg((const uint4*)((const uint32_t*)pcm + i0) + 1);
      const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 8; j++) q[j] = (dt == FRAD_S32) ? (T)((double)(int32_t)w[j] / 2147483648.0) : (T)((double)w[j] / 2147483648.0 - 1.0);
      break;
    }
    default:
#pragma unroll
      for (int j = 0; j < 8; j++) q[j] = load_pcm<T>(pcm, dt, i0 + j);
  }
}

// ---- profile-0 value packing of the store phase (profile0.py:28-42): four consecutive values -> 4 * bits / 8 bytes --------
// image of one value in stream order (first stream byte in the low byte)
template <typename T> __device__ __forceinline__ uint64_t p0_image(T x, int bits, bool le) {
  if (bits == 16) { const uint32_t h = f2h_bits(x); return le ? h : (uint64_t)__byte_perm(h, 0, 0x4401); }
  if (bits == 32) { const uint32_t u = f2f_bits(x); return le ? u : (uint64_t)__byte_perm(u, 0, 0x0123); }
  if (bits == 24) { const uint32_t u = f2f_bits(x); return le ? (u >> 8) : (uint64_t)__byte_perm(u, 0, 0x4123); }   // top 3 bytes
  const uint64_t u = f2d_bits(x);
  if (bits == 64 || bits == 48) {
    if (le) return bits == 64 ? u : (u >> 16);
    const uint64_t sw = ((uint64_t)__byte_perm((uint32_t)u, 0, 0x0123) << 32) | __byte_perm((uint32_t)(u >> 32), 0, 0x0123);
    return bits == 64 ? sw : (sw & 0xffffffffffffull);                        // big endian: the six most significant bytes first
  }
  return 0;
}
// dst: start of the four values' bytes (4-byte aligned for every depth because value index vi is a multiple of 4)
template <typename T> __device__ __forceinline__ void p0_store4(uint8_t* dst, const T* v, int bits, bool le) {
  if (bits == 12) {                          // top 12 bits of the big-endian half, nibble-packed (always big endian)
    const uint32_t a = f2h_bits(v[0]) >> 4, b = f2h_bits(v[1]) >> 4, c = f2h_bits(v[2]) >> 4, d = f2h_bits(v[3]) >> 4;
    const uint32_t w0 = (a >> 4) | (((a & 15u) << 4 | (b >> 8)) << 8) | ((b & 255u) << 16) | ((c >> 4) << 24);
    const uint32_t w1 = ((c & 15u) << 4 | (d >> 8)) | ((d & 255u) << 8);
    if (((uintptr_t)dst & 3) == 0) { *(uint32_t*)dst = w0; *(uint16_t*)(dst + 4) = (uint16_t)w1; }
    else { *(uint16_t*)dst = (uint16_t)w0; *(uint32_t*)(dst + 2) = (w0 >> 16) | (w1 << 16); }
    return;
  }
  const uint64_t a = p0_image(v[0], bits, le), b = p0_image(v[1], bits, le), c = p0_image(v[2], bits, le), d = p0_image(v[3], bits, le);
  switch (bits) {
    case 16: *(uint2*)dst = make_uint2((uint32_t)(a | (b << 16)), (uint32_t)(c | (d << 16))); break;
    case 24: {
      uint32_t* w = (uint32_t*)dst;
      w[0] = (uint32_t)(a | (b << 24)); w[1] = (uint32_t)((b >> 8) | (c << 16)); w[2] = (uint32_t)((c >> 16) | (d << 8));
      break;
    }
    case 32: *(uint4*)dst = make_uint4((uint32_t)a, (uint32_t)b, (uint32_t)c, (uint32_t)d); break;
    case 48: {
      uint64_t* q = (uint64_t*)dst;
      q[0] = a | (b << 48); q[1] = (b >> 16) | (c << 32); q[2] = (c >> 32) | (d << 16);
      break;
    }
    default: {
      uint4* q = (uint4*)dst;
      q[0] = make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
      q[1] = make_uint4((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)d, (uint32_t)(d >> 32));
    }
  }
}

template <typename T, int PROFILE, int CT>
__global__ void __launch_bounds__(FW * 32, (sizeof(T) == 8) ? 3 : FRAD_MINB)
enc2048r_kernel(enc_params P, plan_dev pl, const typename vec2<T>::type* __restrict__ tw32, int force_exact, int wide) {
  typedef typename vec2<T>::type V;
  typedef enc_tail<T> TAIL;
  constexpr bool F32 = sizeof(T) == 4;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V* seq_all = (V*)smem_raw;
  TAIL* tail = (TAIL*)(smem_raw + (size_t)FW * ESEQ * sizeof(V));
  const int C = (CT == 1 || CT == 2) ? CT : P.C;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const tile_map tm = make_tiles<CT>(P.B, C);
  const float* pcm = (const float*)P.pcm;
  const V* cfh = (const V*)pl.cfh;
  const V* dfh = (const V*)pl.dfh;
  const int words_per_frame = FN * C;
  const bool raw32 = F32 && P.pcm_dtype == FRAD_F32;       // f32 samples travel as 32-bit words, no conversion
  const bool p0_words = (PROFILE == 0) && F32 && P.bits == 32;   // profile 0 at 32 bits: byte order applied in registers, words copied

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
    if (F32 && CT == 2 && wide) {
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
    } else if (raw32) {
      if constexpr (F32)
        stage_tile<CT, ESEQ * 2>([&](int f) { return (const uint32_t*)(pcm + (size_t)(b0 + f) * P.hop * C); }, (uint32_t*)seq_all, C, nf,
                                 grp, tid, [](int n) { return zword(n); });
    } else if (CT == 2 && (P.hop % 4) == 0 && (((uintptr_t)P.pcm) % 16) == 0) {
      // stereo, any sample format: a thread takes whole sample quads (4 samples x 2 channels = 8 values, 8 .. 64 bytes in
      // one or two loads), converts them like to_f64 (pcmformat.py:34-47) and stores the two complex values per channel
      for (int f = 0; f < nf; f++) {
        V* s0 = seq_all + (size_t)(2 * f) * ESEQ;
        V* s1 = seq_all + (size_t)(2 * f + 1) * ESEQ;
        const int64_t fbase = (int64_t)(b0 + f) * P.hop * 2;
#pragma unroll 2
        for (int i = 0; i < 4; i++) {
          const int m = tid + 128 * i;
          T q[8];
          load_quad8<T>(P.pcm, P.pcm_dtype, fbase + 8 * (int64_t)m, q);
          s0[padc(m)] = mk2<T>(q[0], q[4]);
          s0[padc(FM - 1 - m)] = mk2<T>(q[6], q[2]);
          s1[padc(m)] = mk2<T>(q[1], q[5]);
          s1[padc(FM - 1 - m)] = mk2<T>(q[7], q[3]);
        }
      }
    } else {
      // any sample format (backend/pcmformat.py:34-47: integers become x / 2^(bits-1), unsigned ones minus 1), converted to T on
      // the way in; consecutive threads read consecutive samples of the tile's channels
      T* zall = (T*)seq_all;
      const int Cg = (CT == 4) ? 4 : C, c0 = (CT == 4) ? 4 * grp : 0;
      for (int f = 0; f < nf; f++) {
        const int64_t fbase = (int64_t)(b0 + f) * P.hop * C + c0;
        for (int e = tid; e < FN * Cg; e += FW * 32) {
          int n, cc;
          if (CT == 1) { n = e; cc = 0; } else if (CT == 2) { n = e >> 1; cc = e & 1; }
          else if (CT == 4) { n = e >> 2; cc = e & 3; } else { n = e / C; cc = e - n * C; }
          const int sq = (CT == 4) ? cc : f * C + cc;
          zall[(size_t)sq * (2 * ESEQ) + zword(n)] = load_pcm<T>(P.pcm, P.pcm_dtype, fbase + (int64_t)n * C + cc);
        }
      }
    }
    __syncthreads();
    {   // pull the next tile of this CTA towards L2 while the current one is transformed
      const int nt = tile + gridDim.x;
      if (nt < tm.ntiles) {
        const int nb0 = (CT == 4) ? nt / tm.groups : nt * tm.fpc;
        const int nnf = min(tm.fpc, P.B - nb0);
        const size_t es = (size_t)pcm_size(P.pcm_dtype);
        const char* base = (const char*)P.pcm + (size_t)nb0 * P.hop * C * es;
        const size_t bytes = (size_t)nnf * words_per_frame * es;
        for (size_t o = (size_t)tid * 128; o < bytes; o += (size_t)FW * 32 * 128)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(base + o));
      }
    }
    if (w < nseq) {
      const int f = (CT == 4) ? 0 : w / C, c = (CT == 4) ? 4 * grp + w : w - f * C, b = b0 + f;
      V* buf = seq_all + (size_t)w * ESEQ;
      T* X = (T*)buf;
      V v[32];
#pragma unroll
      for (int i = 0; i < 32; i++) v[i] = buf[lane + 33 * i];
      __syncwarp();
      fft1024_warp<T, false>(v, buf, lane, tw32);
      __syncwarp();                                       // every lane has read its exchange rows: the buffer becomes X
      // ---- pairwise real-FFT split + DCT twiddle: (k, M-k) -> X[k], X[N-k], X[M-k], X[M+k], k = lane + 32 o.
      //      Padded positions: two steps of o move 64 bins = 68 words, so every store takes an immediate offset from one of
      //      eight per-lane bases (k, M+k ascending; N-k, M-k descending; even and odd o) ----
      const int partner = (32 - lane) & 31;
      const bool swap_out = p0_words && !P.le;
      T mx = T(0);
      const int oK[2] = {xpos(lane), xpos(lane + 32)};
      const int oMK[2] = {xpos(FM + lane), xpos(FM + lane + 32)};
      const int oNK[2] = {xpos(FN - lane), xpos(FN - lane - 32)};             // o = 0, lane 0 (bin 2048) is never stored
      const int oMmK[2] = {xpos(FM - lane), xpos(FM - lane - 32)};
      auto put = [&](int pos, T val) {
        if (PROFILE == 0) {
          mx = fmax(mx, fabs(val));
          if constexpr (F32) { if (swap_out) val = __uint_as_float(__byte_perm(__float_as_uint(val), 0, 0x0123)); }
        }
        X[pos] = val;
      };
      const T r2 = (T)FRAD_RSQRT2;
      // f64: cf[k] = cf[lane] e^{-i pi o / 128}, df[k] = df[lane] e^{-i pi o / 128} e^{-i pi o / 32} for k = lane + 32 o -- one
      // per-lane table value times a compile-time constant instead of two 16-byte loads per pair (the 32 KiB of tables do not
      // stay in the little L1 that is left beside 3 x 70 KiB of shared memory; same scheme as the f64 decoder)
      V cf_lane = mk2<T>(T(0), T(0)), df_lane = cf_lane;
      if constexpr (!F32) { cf_lane = __ldg(cfh + lane); df_lane = __ldg(dfh + lane); }
      auto post_tw = [&](int k, int o, V& ck, V& dk) {
        if constexpr (F32) { ck = __ldg(cfh + k); dk = __ldg(dfh + k); }
        else {
          const V g = mk2<T>((T)g128_cos(o), (T)(-g128_sin(o)));
          const V g2 = mk2<T>((T)g32_cos(o), (T)(-g32_sin(o)));
          ck = cmul(cf_lane, g);
          dk = cmul(cmul(df_lane, g), g2);
        }
      };
#pragma unroll
      for (int o = 0; o < 16; o++) {
        const V snd = (lane == 0) ? v[(32 - o) & 31] : v[31 - o];
        V zm;
        zm.x = shfl_t<T>(snd.x, partner);
        zm.y = -shfl_t<T>(snd.y, partner);                                  // conj(Z[M-k])
        const int k = lane + 32 * o;
        const bool first = (o == 0) && (lane == 0);
        const V A = cadd(v[o], zm), D = csub(v[o], zm);
        const V Bv = mk2<T>(D.y, -D.x);
        V ck, dk;
        post_tw(k, o, ck, dk);
        const V Pk = cmul(ck, A), Qk = cmul(dk, Bv);
        const V y = cadd(Pk, Qk), t = csub(Pk, Qk);
        const int h = o >> 1, e = o & 1;
        put(oK[e] + 68 * h, y.x);
        if (!first) put(oNK[e] - 68 * h, -y.y);
        put(oMmK[e] - 68 * h, (t.x - t.y) * r2);
        if (!first) put(oMK[e] + 68 * h, (t.x + t.y) * r2);
      }
      if (lane == 0) {                                                     // self-paired bin k = 512 (and N - 512)
        const V zk = v[16], zm = cconj(v[16]);
        const V A = cadd(zk, zm), D = csub(zk, zm);
        const V Bv = mk2<T>(D.y, -D.x);
        const V y = cadd(cmul(__ldg(cfh + FM / 2), A), cmul(__ldg(dfh + FM / 2), Bv));
        put(xpos(FM / 2), y.x);
        put(xpos(FN - FM / 2), -y.y);
      }
      if (PROFILE == 0) {
        if (P.peak) {
#pragma unroll
          for (int s = 16; s > 0; s >>= 1) mx = fmax(mx, (T)__shfl_xor_sync(0xffffffffu, mx, s));
          if (lane == 0) atomicMax((unsigned long long*)&P.peak[b], (unsigned long long)__double_as_longlong((double)mx));
        }
      } else {
        __syncwarp();
        // ---- the lane's run: bins [64 lane, 64 lane + 64), sixteen 16-byte loads (conflict-free: 68-word pitch).
        //      Band energies of the spectrum (scaled below): running sum, closed where a new band starts ----
        const T* run = X + 68 * lane;
        {
          T* slot = tail->part[w] + lane + bfirst;
          T acc = T(0);
#pragma unroll
          for (int i = 0; i < 16; i++) {
            T xe[4];
            if constexpr (F32) { const float4 t = ((const float4*)run)[i]; xe[0] = t.x; xe[1] = t.y; xe[2] = t.z; xe[3] = t.w; }
            else {
              const double2 t0 = ((const double2*)run)[2 * i], t1 = ((const double2*)run)[2 * i + 1];
              xe[0] = t0.x; xe[1] = t0.y; xe[2] = t1.x; xe[3] = t1.y;
            }
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const int j = 4 * i + e;
              if (j > 0 && flag(j)) { *slot = acc; slot++; acc = T(0); }
              acc = fma(xe[e], xe[e], acc);
            }
          }
          *slot = acc;
        }
        __syncwarp();
        const T scalef = (T)P.scale;
        const float inv_log_base_f = (float)(1.0 / 0.30685281944005469058);   // 1 / ln(e/2)
        T tot = T(0);
        int cnt = 1;
        if (lane < FRAD_NBANDS) {
          const int blo = tail->lo[lane], bhi = tail->lo[lane + 1];
          cnt = bhi - blo;
          if (bhi > blo) {
            const T* row = tail->part[w] + lane;
            for (int l = blo >> 6; l <= (bhi - 1) >> 6; l++) tot += row[l];
          }
          tot *= scalef * scalef;
        }
        const unsigned empty = __ballot_sync(0xffffffffu, cnt <= 0);
        const int first_empty = empty ? (__ffs(empty) - 1) : 32;
        double thr = 0.0;
        if (lane < FRAD_NBANDS && lane < first_empty) {
          // in the spectrum's precision like the reference: np.sqrt(np.mean(x**2)) ** 0.8 on float32 / float64 (p1tools.py:32)
          double level;
          if constexpr (F32) level = (double)powf(sqrtf(tot / (float)cnt), 0.8f);
          else level = pow(sqrt(tot / (double)cnt), 0.8);
          thr = fmax(level, P.bt->ath[lane]) * P.loss;
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
          int4* qrun = (int4*)((int*)X + qpos<T>(64 * lane));
#pragma unroll 1
          for (int g = 0; g < 8; g++) {
            int ms[8];
            bool allok = true;
            T xt[8];
            if constexpr (F32) {
              const float4 xa = ((const float4*)run)[2 * g], xb = ((const float4*)run)[2 * g + 1];
              xt[0] = xa.x; xt[1] = xa.y; xt[2] = xa.z; xt[3] = xa.w; xt[4] = xb.x; xt[5] = xb.y; xt[6] = xb.z; xt[7] = xb.w;
            } else {
#pragma unroll
              for (int u2 = 0; u2 < 4; u2++) { const double2 t = ((const double2*)run)[4 * g + u2]; xt[2 * u2] = t.x; xt[2 * u2 + 1] = t.y; }
            }
            float xs[8];
#pragma unroll
            for (int u = 0; u < 8; u++) xs[u] = (float)xt[u];
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
                const int j = 8 * g + u;                 // bin of the run
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
                  ms[u] = quant_exact_r((double)xt[u], 64 * lane + j - lo, t0, (t1 - t0) / (double)n, P.scale);
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
    // ---- cooperative, coalesced store: value e = k*C + c of frame f <- seq[f*C + c][xpos(k)] ----
    const bool words = (PROFILE != 0) || p0_words;           // 32-bit words copied as they are (quantised ints / finished f32 words)
    if (F32 && CT == 2 && wide && words) {
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
    } else if (words) {
      constexpr int STRIDE_W = ESEQ * 2 * (int)(sizeof(T) / 4);         // 32-bit words per sequence buffer
      for (int f = 0; f < nf; f++) {
        uint32_t* dst = (PROFILE == 0) ? (uint32_t*)(P.payload + (size_t)(b0 + f) * P.payload_stride)
                                       : (uint32_t*)(P.q + (size_t)(b0 + f) * words_per_frame);
        for (int e4 = tid; e4 < vec_count<CT>(C); e4 += FW * 32) {
          uint32_t o4[4];
          gather4<CT, STRIDE_W>((const uint32_t*)seq_all, C, f, e4, [](int k) { return (PROFILE == 0) ? xpos(k) : qpos<T>(k); },
                                [](uint32_t u) { return u; }, o4);
          *(uint4*)(dst + vec_word<CT>(e4, C, grp)) = make_uint4(o4[0], o4[1], o4[2], o4[3]);
        }
      }
    } else {
      // profile 0 at the other depths (and from a double spectrum): four consecutive values -> their bytes (profile0.py:28-42)
      if constexpr (PROFILE == 0) {
        const bool le = P.le != 0;
        for (int f = 0; f < nf; f++) {
          uint8_t* frame = P.payload + (size_t)(b0 + f) * P.payload_stride;
          for (int e4 = tid; e4 < vec_count<CT>(C); e4 += FW * 32) {
            T o4[4];
            gather4<CT, ESEQ * 2>((const T*)seq_all, C, f, e4, [](int k) { return xpos(k); }, [](T u) { return u; }, o4);
            const size_t vi = (size_t)vec_word<CT>(e4, C, grp);          // index of the first of the four values
            p0_store4<T>(frame + (P.bits == 12 ? vi * 3 / 2 : vi * (size_t)(P.bits >> 3)), o4, P.bits, le);
          }
        }
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------------------------------
// Frame lengths N = 64 R1 < 2048 (R1 = 16, 8, 4, 2: N = 1024 ... 128 of the compact table, fourier/profiles.py:14-23; mixed radices below): the same
// kernel with S = 32 / R1 sequences per warp.  Lane L = s R1 + k1 of a warp works on sequence s: in the first FFT pass it
// transforms R1 points, after the exchange it holds Z_s[k1 + R1 o] (fft32.cuh: fft_multi_warp), and in the epilogue it owns the
// 64 contiguous bins [64 k1, 64 k1 + 64) of its sequence -- which in the padded layout start at word 68 L of the warp buffer,
// exactly as in the N = 2048 kernel, so the run epilogue is the same code with k1 in place of the lane number.
// A tile is FW S sequences = FW S / C whole frames (the host checks that C divides FW S).
//
// R1 = 3, 5, 6, 7, 10, 12, 14, 20, 24, 28 (N = 192 ... 1792: the 160 / 192 / 224 families of the compact table): the same kernel with
// the mixed-radix first pass of fftmix.cuh and S = floor(32 / R1) sequences per warp.  S R1 < 32 then: the lanes at and above
// S R1 have no sequence after the exchange -- they run along on the last sequence's addresses (every shuffle keeps its full
// mask) and store nothing.
template <typename T, int S> struct encs_tail {
  int lo[FRAD_NBANDS + 1];
  double inv_ns[FRAD_NBANDS + 1];
  double thr[FW][S][FRAD_NBANDS + 1];
  float4 bp[FW][S][FRAD_NBANDS + 1];
  T part[FW][S][64];
};

template <typename T, int PROFILE, int R1>
__global__ void __launch_bounds__(FW * 32, (sizeof(T) == 8) ? 3 : FRAD_MINB)
encs_kernel(enc_params P, plan_dev pl, const typename vec2<T>::type* __restrict__ tw32, int force_exact) {
  typedef typename vec2<T>::type V;
  constexpr bool F32 = sizeof(T) == 4;
  constexpr int NN = 64 * R1, MM = 32 * R1, S = 32 / R1, SPC = FW * S;
  constexpr int RS = S * R1;                           // lanes that own a sequence after the exchange (32 for the powers of two)
  constexpr int SEQ_T = 68 * R1;                       // T slots of one sequence in the padded spectrum layout
  typedef encs_tail<T, S> TAIL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  V* seq_all = (V*)smem_raw;
  TAIL* tail = (TAIL*)(smem_raw + (size_t)FW * ESEQ * sizeof(V));
  const int C = P.C;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int k1 = lane % R1, sl = (RS == 32) ? lane / R1 : min(lane / R1, S - 1);
  const bool act = (RS == 32) || lane < RS;
  const int fpt = SPC / C;                             // whole frames per tile
  const int ntiles = (P.B + fpt - 1) / fpt;
  const V* cfh = (const V*)pl.cfh;
  const V* dfh = (const V*)pl.dfh;
  const bool p0_words = (PROFILE == 0) && F32 && P.bits == 32;
  const bool vec_in = F32 && P.pcm_dtype == FRAD_F32 && C == 2 && (((uintptr_t)P.pcm) % 16) == 0 && ((P.hop * 2) % 4) == 0;

  int bfirst = 0;
  uint32_t fl_lo = 0, fl_hi = 0;
  if (PROFILE != 0) {
    if (tid <= FRAD_NBANDS) { const int e = P.bt->edge[tid]; tail->lo[tid] = (e < NN) ? e : NN; }
    __syncthreads();
    if (tid < FRAD_NBANDS) {
      const int n = tail->lo[tid + 1] - tail->lo[tid];
      tail->inv_ns[tid] = (n > 0) ? 1.0 / ((double)n * P.scale) : 0.0;
    }
#pragma unroll 1
    for (int b = 1; b < FRAD_NBANDS; b++) {
      const int d = tail->lo[b] - 64 * k1;
      if (d <= 0) bfirst = b;
      else if (d < 64 && tail->lo[b] < NN) { if (d < 32) fl_lo |= 1u << d; else fl_hi |= 1u << (d - 32); }
    }
  }
  auto flag = [&](int j) -> bool { return ((j < 32 ? (fl_lo >> j) : (fl_hi >> (j - 32))) & 1u) != 0; };
  // T-slot of sample n inside the z layout of sequence slot sq of its warp (Makhoul order, see zword)
  auto zslot = [&](int n, int sq) -> int {
    const int nn = (n & 1) ? (2 * NN - 1 - n) : n;
    const int m = nn >> 2;
    return 2 * (33 * R1 * sq + m + (m >> 5)) + ((nn >> 1) & 1);
  };

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b0 = tile * fpt;
    const int nf = min(fpt, P.B - b0);
    const int nseq = nf * C;
    __syncthreads();
    // ---- staging: the tile's frames in memory order (frame, sample, channel) ----
    if (vec_in) {
      // 16 bytes = samples n, n + 1 of both channels.  S >= 3 (N <= 640): four vectors per thread are requested before the first
      // is scattered (p0 encode at N = 640: 0.245 -> 0.182 ms); measured slower by 4-7 % at S <= 2, and in the decoder everywhere
      constexpr int UB = (S >= 3) ? 4 : 1;
      const int total = nf * (NN / 2);
      for (int e0 = tid; e0 < total; e0 += FW * 32 * UB) {
        uint4 x[UB];
#pragma unroll
        for (int u = 0; u < UB; u++) {
          const int e4 = e0 + u * FW * 32;
          if (e4 < total) {
            const int f = e4 / (NN / 2), n = 2 * (e4 - f * (NN / 2));
            x[u] = ldg_stream((const uint4*)((const float*)P.pcm + ((size_t)(b0 + f) * P.hop + n) * 2));
          }
        }
#pragma unroll
        for (int u = 0; u < UB; u++) {
          const int e4 = e0 + u * FW * 32;
          if (e4 < total) {
            const int f = e4 / (NN / 2), n = 2 * (e4 - f * (NN / 2));
            const int q0 = 2 * f, q1 = q0 + 1;
            uint32_t* bq0 = (uint32_t*)(seq_all + (size_t)(q0 / S) * ESEQ) + 2 * 33 * R1 * (q0 % S);
            uint32_t* bq1 = (uint32_t*)(seq_all + (size_t)(q1 / S) * ESEQ) + 2 * 33 * R1 * (q1 % S);
            const int z0 = zslot(n, 0), z1 = zslot(n + 1, 0);          // the slot inside a sequence is the same for both channels
            bq0[z0] = x[u].x; bq1[z0] = x[u].y; bq0[z1] = x[u].z; bq1[z1] = x[u].w;
          }
        }
      }
    } else {
      for (int e = tid; e < nf * NN * C; e += FW * 32) {
        const int f = e / (NN * C), r = e - f * NN * C; int n, c; split_c(r, C, n, c);
        const int q = f * C + c;
        ((T*)(seq_all + (size_t)(q / S) * ESEQ))[zslot(n, q % S)] = load_pcm<T>(P.pcm, P.pcm_dtype, ((int64_t)(b0 + f) * P.hop + n) * C + c);
      }
    }
    __syncthreads();
    {   // pull the next tile of this CTA towards L2 while the current one is transformed
      const int nt = tile + gridDim.x;
      if (nt < ntiles) {
        const int nb0 = nt * fpt, nnf = min(fpt, P.B - nb0);
        const size_t es = (size_t)pcm_size(P.pcm_dtype);
        const char* base = (const char*)P.pcm + (size_t)nb0 * P.hop * C * es;
        const size_t bytes = ((size_t)(nnf - 1) * P.hop + NN) * C * es;
        for (size_t o = (size_t)tid * 128; o < bytes; o += (size_t)FW * 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + o));
      }
    }
    if (w * S < nseq) {
      const int q_me = w * S + sl;                       // my sequence slot in the tile
      const bool live = act && q_me < nseq;              // lanes of slots beyond the batch compute on stale data and store nothing global
      const int f_me = q_me / C, c_me = q_me - f_me * C, b_me = b0 + f_me;
      V* buf = seq_all + (size_t)w * ESEQ;
      T* X = (T*)buf;
      T* Xs = X + SEQ_T * sl;                            // my sequence's spectrum
      V v[32];
#pragma unroll
      for (int i = 0; i < 32; i++) v[i] = buf[lane + 33 * i];
      __syncwarp();
      fft_multi_warp<T, false, R1>(v, buf, lane, tw32);
      __syncwarp();
      // ---- pairwise real-FFT split + DCT twiddle: (k, M-k) -> X[k], X[N-k], X[M-k], X[M+k], k = k1 + R1 o ----
      const int partner = sl * R1 + ((R1 - k1) % R1);
      const bool swap_out = p0_words && !P.le;
      T mx = T(0);
      auto put = [&](int k, T val) {
        if (!act) return;
        if (PROFILE == 0) {
          mx = fmax(mx, fabs(val));
          if constexpr (F32) { if (swap_out) val = __uint_as_float(__byte_perm(__float_as_uint(val), 0, 0x0123)); }
        }
        Xs[xpos(k)] = val;
      };
      const T r2 = (T)FRAD_RSQRT2;
      V cf_lane = mk2<T>(T(0), T(0)), df_lane = cf_lane;
      if constexpr (!F32) { cf_lane = __ldg(cfh + k1); df_lane = __ldg(dfh + k1); }
#pragma unroll
      for (int o = 0; o < 16; o++) {
        const V snd = (k1 == 0) ? v[(32 - o) & 31] : v[31 - o];
        V zm;
        zm.x = shfl_t<T>(snd.x, partner);
        zm.y = -shfl_t<T>(snd.y, partner);
        const int k = k1 + R1 * o;
        const bool first = (o == 0) && (k1 == 0);
        const V A = cadd(v[o], zm), D = csub(v[o], zm);
        const V Bv = mk2<T>(D.y, -D.x);
        V ck, dk;
        if constexpr (F32) { ck = __ldg(cfh + k); dk = __ldg(dfh + k); }
        else {   // cf[k1 + R1 o] = cf[k1] e^{-i pi o / 128}, df likewise with another e^{-i pi o / 32}: the same constants as for N = 2048
          const V g = mk2<T>((T)g128_cos(o), (T)(-g128_sin(o)));
          const V g2 = mk2<T>((T)g32_cos(o), (T)(-g32_sin(o)));
          ck = cmul(cf_lane, g);
          dk = cmul(cmul(df_lane, g), g2);
        }
        const V Pk = cmul(ck, A), Qk = cmul(dk, Bv);
        const V y = cadd(Pk, Qk), t = csub(Pk, Qk);
        put(k, y.x);
        if (!first) put(NN - k, -y.y);
        put(MM - k, (t.x - t.y) * r2);
        if (!first) put(MM + k, (t.x + t.y) * r2);
      }
      if (k1 == 0) {                                       // self-paired bin k = M / 2 (and N - M / 2)
        const V zk = v[16], zm = cconj(v[16]);
        const V A = cadd(zk, zm), D = csub(zk, zm);
        const V Bv = mk2<T>(D.y, -D.x);
        const V y = cadd(cmul(__ldg(cfh + MM / 2), A), cmul(__ldg(dfh + MM / 2), Bv));
        put(MM / 2, y.x);
        put(NN - MM / 2, -y.y);
      }
      if (PROFILE == 0) {
        if (P.peak) {
          if constexpr ((R1 & (R1 - 1)) == 0) {
#pragma unroll
            for (int s = 1; s < R1; s <<= 1) mx = fmax(mx, (T)__shfl_xor_sync(0xffffffffu, mx, s));   // over the R1 lanes of my sequence
          } else {
            const T mine = mx;
#pragma unroll
            for (int s = 1; s < R1; s++) mx = fmax(mx, shfl_t<T>(mine, sl * R1 + (k1 + s) % R1));
          }
          if (k1 == 0 && live) atomicMax((unsigned long long*)&P.peak[b_me], (unsigned long long)__double_as_longlong((double)mx));
        }
      } else {
        __syncwarp();
        const T* run = X + 68 * lane;
        if (act) {
          T* slot = tail->part[w][sl] + k1 + bfirst;
          T acc = T(0);
#pragma unroll
          for (int i = 0; i < 16; i++) {
            T xe[4];
            if constexpr (F32) { const float4 t = ((const float4*)run)[i]; xe[0] = t.x; xe[1] = t.y; xe[2] = t.z; xe[3] = t.w; }
            else {
              const double2 t0 = ((const double2*)run)[2 * i], t1 = ((const double2*)run)[2 * i + 1];
              xe[0] = t0.x; xe[1] = t0.y; xe[2] = t1.x; xe[3] = t1.y;
            }
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const int j = 4 * i + e;
              if (j > 0 && flag(j)) { *slot = acc; slot++; acc = T(0); }
              acc = fma(xe[e], xe[e], acc);
            }
          }
          *slot = acc;
        }
        __syncwarp();
        const T scalef = (T)P.scale;
        const float inv_log_base_f = (float)(1.0 / 0.30685281944005469058);
        // thresholds of the warp's sequences, one after the other: band = lane
#pragma unroll 1
        for (int s2 = 0; s2 < S; s2++) {
          const int q2 = w * S + s2;
          if (q2 >= nseq) break;                           // warp-uniform
          const int f2 = q2 / C, c2 = q2 - f2 * C, b2 = b0 + f2;
          T tot = T(0);
          int cnt = 1;
          if (lane < FRAD_NBANDS) {
            const int blo = tail->lo[lane], bhi = tail->lo[lane + 1];
            cnt = bhi - blo;
            if (bhi > blo) {
              const T* row = tail->part[w][s2] + lane;
              for (int l = blo >> 6; l <= (bhi - 1) >> 6; l++) tot += row[l];
            }
            tot *= scalef * scalef;
          }
          const unsigned empty = __ballot_sync(0xffffffffu, cnt <= 0);
          const int first_empty = empty ? (__ffs(empty) - 1) : 32;
          double thr = 0.0;
          if (lane < FRAD_NBANDS && lane < first_empty) {
            double level;
            if constexpr (F32) level = (double)powf(sqrtf(tot / (float)cnt), 0.8f);
            else level = pow(sqrt(tot / (double)cnt), 0.8);
            thr = fmax(level, P.bt->ath[lane]) * P.loss;
          }
          if (lane <= FRAD_NBANDS) tail->thr[w][s2][lane] = thr;
          if (lane < FRAD_NBANDS) {
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
            P.tq[((size_t)b2 * FRAD_NBANDS + lane) * C + c2] = tqi;
          }
          const double thr_next = __shfl_down_sync(0xffffffffu, thr, 1);
          if (lane <= FRAD_NBANDS) {
            float4 bpv = make_float4(1.0e30f, 1.0e30f, -1.0f, 2.0f * NN);
            if (lane < FRAD_NBANDS - 1 && cnt > 0) {
              const double ins = tail->inv_ns[lane];
              bpv = make_float4((float)(thr * ins), (float)(thr_next * ins), (float)tail->lo[lane], (float)tail->lo[lane + 1]);
            }
            tail->bp[w][s2][lane] = bpv;
          }
        }
        __syncwarp();
        if (act) {
          const float4* bpp = tail->bp[w][sl] + bfirst;
          float4 bpv = *bpp;
          const float kb = (float)(64 * k1);
          int4* qrun = (int4*)((int*)X + qpos<T>(64 * lane));
#pragma unroll 1
          for (int g = 0; g < 8; g++) {
            int ms[8];
            bool allok = true;
            T xt[8];
            if constexpr (F32) {
              const float4 xa = ((const float4*)run)[2 * g], xb = ((const float4*)run)[2 * g + 1];
              xt[0] = xa.x; xt[1] = xa.y; xt[2] = xa.z; xt[3] = xa.w; xt[4] = xb.x; xt[5] = xb.y; xt[6] = xb.z; xt[7] = xb.w;
            } else {
#pragma unroll
              for (int u2 = 0; u2 < 4; u2++) { const double2 t = ((const double2*)run)[4 * g + u2]; xt[2 * u2] = t.x; xt[2 * u2 + 1] = t.y; }
            }
            float xs[8];
#pragma unroll
            for (int u = 0; u < 8; u++) xs[u] = (float)xt[u];
            const uint32_t fb = ((g < 4) ? (fl_lo >> (8 * g)) : (fl_hi >> (8 * g - 32))) & (g == 0 ? 0xfeu : 0xffu);
            const float kg = kb + (float)(8 * g);
#pragma unroll
            for (int u = 0; u < 8; u++) {
              if ((fb >> u) & 1u) { bpp++; bpv = *bpp; }
              const float kf = kg + (float)u;
              const float x = xs[u];
              const float dv = fmaf(kf - bpv.z, bpv.y, (bpv.w - kf) * bpv.x);
              float rc, s, vv;
              asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(dv));
              const float y = fabsf(x) * rc;
              asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(y));
              asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(vv) : "f"(y * s));
              const float big = copysignf(vv, x) + 12582912.0f;
              const float r = big - 12582912.0f;
              allok = allok && (fmaf(1.0e-6f, vv, fabsf(fabsf(vv) - fabsf(r))) < 0.4999999f);
              ms[u] = __float_as_int(big) - 0x4B400000;
            }
            if (!allok || force_exact) {
#pragma unroll
              for (int u = 0; u < 8; u++) {
                const int j = 8 * g + u;
                const unsigned long long fl = ((unsigned long long)fl_hi << 32) | fl_lo;
                const int bn = bfirst + __popcll(fl & ((j >= 63) ? ~0ull : ((2ull << j) - 1ull)));
                bool redo = force_exact != 0;
                if (!redo) {
                  const float4 bq = tail->bp[w][sl][bn];
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
                  const double t0 = tail->thr[w][sl][bn], t1 = tail->thr[w][sl][bn + 1];
                  ms[u] = quant_exact_r((double)xt[u], 64 * k1 + j - lo, t0, (t1 - t0) / (double)n, P.scale);
                }
              }
            }
            qrun[2 * g] = make_int4(ms[0], ms[1], ms[2], ms[3]);
            qrun[2 * g + 1] = make_int4(ms[4], ms[5], ms[6], ms[7]);
          }
        }
      }
    }
    __syncthreads();
    // ---- cooperative store: value e = k*C + c of frame f <- sequence slot f*C + c, four consecutive values per thread ----
    const bool words = (PROFILE != 0) || p0_words;
    const int vpf = (C == 2) ? NN / 2 : (C == 1) ? NN / 4 : NN * C / 4;        // 16-byte output vectors per frame
    for (int e4 = tid; e4 < nf * vpf; e4 += FW * 32) {
      int f, vi;                                           // vi: first of four consecutive values of frame f
      uint32_t ow[4];
      T ov[4];
      if (C == 2 && words) {                               // bins k0, k0 + 1 of both channels: two 8-byte loads, constant divisors
        f = e4 / (NN / 2); vi = 4 * (e4 - f * (NN / 2));
        const int q0 = 2 * f, q1 = q0 + 1, k0 = vi >> 1;
        const uint32_t* bq0 = (const uint32_t*)(seq_all + (size_t)(q0 / S) * ESEQ) + (size_t)(q0 % S) * SEQ_T * (sizeof(T) / 4);
        const uint32_t* bq1 = (const uint32_t*)(seq_all + (size_t)(q1 / S) * ESEQ) + (size_t)(q1 % S) * SEQ_T * (sizeof(T) / 4);
        const int pk = (PROFILE == 0) ? xpos(k0) : qpos<T>(k0);
        const uint2 a = *(const uint2*)(bq0 + pk), bb = *(const uint2*)(bq1 + pk);
        ow[0] = a.x; ow[1] = bb.x; ow[2] = a.y; ow[3] = bb.y;
      } else if (C == 1 && words) {                        // mono: four consecutive bins, one 16-byte load
        f = e4 / (NN / 4); vi = 4 * (e4 - f * (NN / 4));
        const uint32_t* bq = (const uint32_t*)(seq_all + (size_t)(f / S) * ESEQ) + (size_t)(f % S) * SEQ_T * (sizeof(T) / 4);
        const uint4 a = *(const uint4*)(bq + ((PROFILE == 0) ? xpos(vi) : qpos<T>(vi)));
        ow[0] = a.x; ow[1] = a.y; ow[2] = a.z; ow[3] = a.w;
      } else {
      f = e4 / vpf; vi = 4 * (e4 - f * vpf);
#pragma unroll
      for (int j = 0; j < 4; j++) {
        int k, c; split_c(vi + j, C, k, c); const int q = f * C + c;
        if (words) {
          const uint32_t* base = (const uint32_t*)(seq_all + (size_t)(q / S) * ESEQ) + (size_t)(q % S) * SEQ_T * (sizeof(T) / 4);
          ow[j] = base[(PROFILE == 0) ? xpos(k) : qpos<T>(k)];
        } else {
          ov[j] = ((const T*)(seq_all + (size_t)(q / S) * ESEQ))[(size_t)(q % S) * SEQ_T + xpos(k)];
        }
      }
      }
      if (words) {
        uint32_t* dst = (PROFILE == 0) ? (uint32_t*)(P.payload + (size_t)(b0 + f) * P.payload_stride) : (uint32_t*)(P.q + (size_t)(b0 + f) * NN * C);
        *(uint4*)(dst + vi) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
      } else if constexpr (PROFILE == 0) {
        uint8_t* frame = P.payload + (size_t)(b0 + f) * P.payload_stride;
        p0_store4<T>(frame + (P.bits == 12 ? (size_t)vi * 3 / 2 : (size_t)vi * (size_t)(P.bits >> 3)), ov, P.bits, P.le != 0);
      }
    }
  }
}

}  // namespace

#define ENCS_DISPATCH(TT, PROF, RR)                                                                          \
  do {                                                                                                       \
    auto k = encs_kernel<TT, PROF, RR>;                                                                      \
    const int fpt = FW * (32 / RR) / P.C;                                                                    \
    int grid;                                                                                                \
    const size_t smem = (size_t)FW * ESEQ * sizeof(vec2<TT>::type) + ((PROF == 0) ? 16 : sizeof(encs_tail<TT, 32 / RR>)); \
    int rc = launch_cfg(ctx, k, smem, (P.B + fpt - 1) / fpt, &grid);                                         \
    if (rc) return rc;                                                                                       \
    k<<<grid, FW * 32, smem, ctx->stream>>>(P, plan_view(pl), (const vec2<TT>::type*)pl->tw32, force_exact); \
  } while (0)

#ifdef FRAD_TU_MIXED
// ---- translation units encm32.cu / encm64.cu: the mixed-radix frame lengths (N = 64 R1, R1 as in fftmix.cuh), one arithmetic
//      type each, so that their twenty kernels compile beside this file's ----
#define ENCM_BY_R(TT, PROF)                                                                             \
  do {                                                                                                  \
    switch (P.N / 64) {                                                                                 \
      case 3: ENCS_DISPATCH(TT, PROF, 3); break;    case 5: ENCS_DISPATCH(TT, PROF, 5); break;          \
      case 6: ENCS_DISPATCH(TT, PROF, 6); break;    case 7: ENCS_DISPATCH(TT, PROF, 7); break;          \
      case 10: ENCS_DISPATCH(TT, PROF, 10); break;  case 12: ENCS_DISPATCH(TT, PROF, 12); break;        \
      case 14: ENCS_DISPATCH(TT, PROF, 14); break;  case 20: ENCS_DISPATCH(TT, PROF, 20); break;        \
      case 24: ENCS_DISPATCH(TT, PROF, 24); break;  case 28: ENCS_DISPATCH(TT, PROF, 28); break;        \
      default: return frad_fail(ctx, FRAD_E_ARG, "mixed-radix encoder: frame length");                  \
    }                                                                                                   \
  } while (0)
#if FRAD_TU_MIXED == 32
int frad_launch_encode_mixed_f32(frad_ctx* ctx, const enc_params& P, const frad_plan* pl, int force_exact) {
  if (P.profile == 0) ENCM_BY_R(float, 0); else ENCM_BY_R(float, 1);
  FRAD_LAUNCH_CHECK(ctx);
  return FRAD_OK;
}
#else
int frad_launch_encode_mixed_f64(frad_ctx* ctx, const enc_params& P, const frad_plan* pl, int force_exact) {
  if (P.profile == 0) ENCM_BY_R(double, 0); else ENCM_BY_R(double, 1);
  FRAD_LAUNCH_CHECK(ctx);
  return FRAD_OK;
}
#endif
#else   // !FRAD_TU_MIXED

#define ENCR_DISPATCH(TT, PROF, CT)                                                                          \
  do {                                                                                                       \
    auto k = enc2048r_kernel<TT, PROF, CT>;                                                                  \
    const tile_map tm = make_tiles<CT>(P.B, P.C);                                                            \
    int grid;                                                                                                \
    const size_t smem = (size_t)FW * ESEQ * sizeof(vec2<TT>::type) + ((PROF == 0) ? 16 : sizeof(enc_tail<TT>)); \
    int rc = launch_cfg(ctx, k, smem, tm.ntiles, &grid);                                                     \
    if (rc) return rc;                                                                                       \
    k<<<grid, FW * 32, smem, ctx->stream>>>(P, plan_view(pl), (const vec2<TT>::type*)pl->tw32, force_exact, wide); \
  } while (0)

#define ENCR_BY_C(TT, PROF)                                                                             \
  do {                                                                                                  \
    if (P.C == 2) ENCR_DISPATCH(TT, PROF, 2); else if (P.C == 1) ENCR_DISPATCH(TT, PROF, 1);            \
    else if (P.C % 4 == 0) ENCR_DISPATCH(TT, PROF, 4); else ENCR_DISPATCH(TT, PROF, 0);                 \
  } while (0)

// eligibility (transform.cu asks before it takes this path)
bool frad_fast2048_encode_ok(const enc_params& P, int prec) {
  if (P.N != FN || (prec != FRAD_PREC_F32 && prec != FRAD_PREC_F64)) return false;
  if (P.C > FW && (P.C % 4) != 0) return false;
  if (prec == FRAD_PREC_F64 && P.profile == 2) return false;      // the TNS kernel takes an f32 spectrum from this path
  if (P.profile == 0) {
    if ((P.payload_stride % 16) != 0 || ((uintptr_t)P.payload % 16) != 0) return false;
    if (P.bits != 12 && P.bits != 16 && P.bits != 24 && P.bits != 32 && P.bits != 48 && P.bits != 64) return false;
  } else if (((uintptr_t)P.q % 16) != 0) return false;
  if (P.pcm_dtype == FRAD_F32 && prec == FRAD_PREC_F32 && ((((uintptr_t)P.pcm) % 16) != 0 || ((P.hop * P.C) % 4) != 0)) return false;
  return true;
}

#define ENCS_BY_R(TT, PROF)                                                                             \
  do {                                                                                                  \
    if (P.N == 1024) ENCS_DISPATCH(TT, PROF, 16); else if (P.N == 512) ENCS_DISPATCH(TT, PROF, 8);      \
    else if (P.N == 256) ENCS_DISPATCH(TT, PROF, 4); else ENCS_DISPATCH(TT, PROF, 2);                 \
  } while (0)

// N = 64 R1 < 2048 (frad_warp_r1): S = floor(32 / R1) sequences per warp, FW S sequences per tile; whole frames per tile
bool frad_fast_small_encode_ok(const enc_params& P, int prec) {
  const int r1 = frad_warp_r1(P.N);
  if (!r1) return false;
  if (prec != FRAD_PREC_F32 && prec != FRAD_PREC_F64) return false;
  const int spc = FW * (32 / r1);
  if (P.C < 1 || P.C > spc || (spc % P.C) != 0) return false;
  if (((size_t)P.N * P.C) % 4) return false;
  if (P.profile == 2) return false;                           // profile 2 keeps the generic path at these sizes
  if (P.profile == 0) {
    if ((P.payload_stride % 16) != 0 || ((uintptr_t)P.payload % 16) != 0) return false;
  } else if (((uintptr_t)P.q % 16) != 0) return false;
  return true;
}

int frad_launch_encode_mixed_f32(frad_ctx* ctx, const enc_params& P, const frad_plan* pl, int force_exact);   // encm32.cu
int frad_launch_encode_mixed_f64(frad_ctx* ctx, const enc_params& P, const frad_plan* pl, int force_exact);   // encm64.cu

int frad_launch_encode_small(frad_ctx* ctx, const enc_params& P, const frad_plan* pl, int prec, int force_exact) {
  if (frad_warp_r1_mixed(frad_warp_r1(P.N)))
    return (prec == FRAD_PREC_F32) ? frad_launch_encode_mixed_f32(ctx, P, pl, force_exact) : frad_launch_encode_mixed_f64(ctx, P, pl, force_exact);
  if (prec == FRAD_PREC_F32) { if (P.profile == 0) ENCS_BY_R(float, 0); else ENCS_BY_R(float, 1); }
  else { if (P.profile == 0) ENCS_BY_R(double, 0); else ENCS_BY_R(double, 1); }
  FRAD_LAUNCH_CHECK(ctx);
  return FRAD_OK;
}

int frad_launch_encode_2048r(frad_ctx* ctx, const enc_params& P, const frad_plan* pl, int prec, int force_exact) {
  // 32-byte global accesses of the stereo tiling need 32-byte aligned frames on both sides
  const void* out = (P.profile == 0) ? (const void*)P.payload : (const void*)P.q;
  const int wide = (P.C == 2 && prec == FRAD_PREC_F32 && P.pcm_dtype == FRAD_F32 && ((uintptr_t)P.pcm % 32) == 0 &&
                    ((P.hop * 2) % 8) == 0 && ((uintptr_t)out % 32) == 0 && (P.profile != 0 || (P.payload_stride % 32) == 0)) ? 1 : 0;
  if (prec == FRAD_PREC_F32) {
    if (P.profile == 0) ENCR_BY_C(float, 0); else if (P.profile == 1) ENCR_BY_C(float, 1); else ENCR_BY_C(float, 2);   // 2: P.tns_scratch = thresholds out
  } else {
    if (P.profile == 0) ENCR_BY_C(double, 0); else ENCR_BY_C(double, 1);
  }
  FRAD_LAUNCH_CHECK(ctx);
  return FRAD_OK;
}
#endif  // FRAD_TU_MIXED
