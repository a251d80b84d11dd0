// Exp-Golomb-Rice bit coder of fourier/tools/p1tools.py:46-74 on the device, and the assembly / parsing of
// the profile-1 / profile-2 frame payload that the reference deflates (profile1.py:43-46, :61-62;
// profile2.py:48-51, :66-69).  Deflate itself stays on the host (north_star).
//
// Bit format (p1tools.py:46-60): one byte k = ceil(log2(max|v|)) (0 when every value is 0), then per value
// u = 2v-1 (v > 0) else -2v, w = u + 2^k written as (bitlen(w) - k - 1) zero bits followed by w in bitlen(w)
// bits, MSB first, zero-padded to a whole byte.  Because k is derived from the maximum, bitlen(w) <= k + 2:
// every code is k+1 or k+3 bits long (<= 35 bits for 32-bit values).
//
// pack    one CTA per frame: max-reduce -> k, code lengths, block-wide exclusive scan -> bit offsets; every
//         thread assembles the codes of its 16 consecutive values in a 64-bit register and writes whole
//         32-bit words into a shared-memory bit buffer (atomicOr only on the two words it shares with its
//         neighbours); the buffer is written out as bytes, stream after stream, with the length prefixes.
// compact exclusive scan of the per-frame lengths + byte copy into one contiguous buffer (what goes over
//         PCIe: about 5 times fewer bytes than the int32 streams).
// unpack  one warp per frame.  The code boundaries form a serial chain, but the code is self-synchronising, so the
//         32 lanes decode 32 segments of the bit string speculatively and fix their entry points up afterwards
//         (see decode_stream_warp).  Per lane: a 64-bit register window over big-endian words, count-leading-
//         zeros for the unary part, one variable shift per value; four values per 16-byte store.
#include <cstdlib>
#include "common.cuh"

namespace {

#ifndef GE_MINB
#define GE_MINB 3
#endif
constexpr int GE_VPT = 16;                                  // values per thread and chunk
constexpr int GE_WARPS = 8;                                 // frames per CTA (one warp each)

// The coder is written once for a group of NT cooperating threads: NT = 32, a warp per frame (round 2: no block barrier
// anywhere, eight independent frames per CTA -- the CTA-per-frame form spent 32 % of its stall samples at barriers between its
// max / scan / emit / flush phases with three CTAs per SM to hide them), or NT = 256, the CTA-per-frame form.
template <int NT> struct ge_smem_t {
  static constexpr int CHUNK = NT * GE_VPT;                   // values per chunk
  static constexpr int WORDS = (CHUNK * 35 + 31) / 32 + 3;    // bit buffer of one chunk (+ carried partial word)
  uint32_t bits[WORDS];
  uint32_t warp_tot[NT / 32];
  uint32_t red[NT / 32];
};

template <int NT> __device__ __forceinline__ int ge_tid() { return (NT == 32) ? (int)(threadIdx.x & 31) : (int)threadIdx.x; }
template <int NT> __device__ __forceinline__ void ge_sync() { if (NT == 32) __syncwarp(); else __syncthreads(); }

template <int NT> __device__ __forceinline__ uint32_t group_max(uint32_t v, ge_smem_t<NT>* sm) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, s));
  if constexpr (NT == 32) return v;
  else {
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm->red[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < NT / 32; i++) r = max(r, sm->red[i]);
    return r;
  }
}

// exclusive prefix of `v` over the group; *total receives the group sum.  Its barriers also order what precedes it.
template <int NT> __device__ __forceinline__ uint32_t group_exscan(uint32_t v, ge_smem_t<NT>* sm, uint32_t* total) {
  const int lane = threadIdx.x & 31;
  uint32_t inc = v;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, s); if (lane >= s) inc += t; }
  if constexpr (NT == 32) {
    *total = __shfl_sync(0xffffffffu, inc, 31);
    __syncwarp();
    return inc - v;
  } else {
    const int w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 31) sm->warp_tot[w] = inc;
    __syncthreads();
    uint32_t base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < NT / 32; i++) { const uint32_t t = sm->warp_tot[i]; if (i < w) base += t; tot += t; }
    *total = tot;
    return base + inc - v;
  }
}

// this thread's 16 consecutive values [i0, i0 + 16) of the stream (0 beyond n): four 16-byte loads where the row allows them
__device__ __forceinline__ void load_vals(const int32_t* __restrict__ v, int n, int i0, int32_t (&val)[GE_VPT]) {
  if (i0 + GE_VPT <= n && (((uintptr_t)(v + i0)) & 15) == 0) {
    const int4* p = (const int4*)(v + i0);
#pragma unroll
    for (int j = 0; j < GE_VPT / 4; j++) { const int4 x = __ldg(p + j); val[4 * j] = x.x; val[4 * j + 1] = x.y; val[4 * j + 2] = x.z; val[4 * j + 3] = x.w; }
  } else {
#pragma unroll
    for (int j = 0; j < GE_VPT; j++) val[j] = (i0 + j < n) ? __ldg(v + i0 + j) : 0;
  }
}

// Bytes [skip, nbytes) of the big-endian bit buffer to o[skip .. nbytes); o is 4-byte aligned: whole words are byte-swapped and
// stored as words, the (at most 3 + 3) bytes in front of and behind them one by one.
template <int NT>
__device__ __forceinline__ void flush_bits(const uint32_t* __restrict__ bits, uint8_t* __restrict__ o, uint32_t skip, uint32_t nbytes) {
  const uint32_t w0 = (skip + 3) >> 2, w1 = nbytes >> 2;
  const uint32_t tid = (uint32_t)ge_tid<NT>();
  uint32_t* ow = (uint32_t*)o;
  for (uint32_t w = w0 + tid; w < w1; w += NT) ow[w] = __byte_perm(bits[w], 0, 0x0123);
  if (tid < 8) {
    const uint32_t j = tid & 3;
    uint32_t b; bool ok;
    if (tid < 4) { b = j; ok = (b >= skip) && (b < min(4 * w0, nbytes)); }                          // in front of the first whole word
    else { b = 4 * w1 + j; ok = (w1 >= w0) && (b >= skip) && (b < nbytes); }                         // behind the last one
    if (ok) o[b] = (uint8_t)(bits[b >> 2] >> (24 - 8 * (b & 3)));
  }
}

// Encodes the stream v[0..n) to dst (k byte + bits); returns the byte length.  All NT threads of the group take part.
// The values of a stream of up to one chunk are loaded once and stay in registers between the maximum and the code pass;
// code lengths are k + 1 or k + 3, so a thread's bit count is a population count; only the words a chunk can reach are
// cleared; the bit buffer is laid out relative to the 4-byte aligned address below dst + 1, so that it leaves as
// byte-swapped words instead of single bytes.
template <int NT>
__device__ __noinline__ uint32_t encode_stream(const int32_t* __restrict__ v, int n, uint8_t* __restrict__ dst, ge_smem_t<NT>* sm) {
  constexpr int CHUNK = ge_smem_t<NT>::CHUNK;
  const int tid = ge_tid<NT>();
  if (n <= 0) { if (tid == 0) dst[0] = 0; return 1; }
  const bool single = (n <= CHUNK);
  int32_t val[GE_VPT];
  uint32_t mx = 0;
  if (single) {
    load_vals(v, n, tid * GE_VPT, val);
#pragma unroll
    for (int j = 0; j < GE_VPT; j++) mx = max(mx, (uint32_t)((val[j] < 0) ? -(int64_t)val[j] : (int64_t)val[j]));
  } else if ((((uintptr_t)v) & 15) == 0) {
    const int4* v4 = (const int4*)v;
    for (int i = tid; i < (n >> 2); i += NT) {
      const int4 x = __ldg(v4 + i);
      const uint32_t a0 = (uint32_t)((x.x < 0) ? -(int64_t)x.x : (int64_t)x.x), a1 = (uint32_t)((x.y < 0) ? -(int64_t)x.y : (int64_t)x.y);
      const uint32_t a2 = (uint32_t)((x.z < 0) ? -(int64_t)x.z : (int64_t)x.z), a3 = (uint32_t)((x.w < 0) ? -(int64_t)x.w : (int64_t)x.w);
      mx = max(max(mx, max(a0, a1)), max(a2, a3));
    }
    for (int i = (n & ~3) + tid; i < n; i += NT) { const int32_t x = v[i]; mx = max(mx, (uint32_t)((x < 0) ? -(int64_t)x : (int64_t)x)); }
  } else {
    for (int i = tid; i < n; i += NT) { const int32_t x = v[i]; mx = max(mx, (uint32_t)((x < 0) ? -(int64_t)x : (int64_t)x)); }
  }
  mx = group_max<NT>(mx, sm);
  const int k = (mx <= 1) ? 0 : 32 - __clz((int)(mx - 1));
  if (tid == 0) dst[0] = (uint8_t)k;
  uint8_t* out = dst + 1;
  const uint32_t mis = (uint32_t)((uintptr_t)out & 3);         // the buffer's byte 0 is out - mis
  uint8_t* obase = out - mis;
  uint64_t flushed = 0;                                        // buffer bytes already written (a multiple of 4)
  uint32_t carry = 8 * mis;                                    // bits of sm->bits[0] in front of this chunk's codes
  uint32_t skip = mis;                                         // bytes at the front of the buffer that are not ours (first chunk)
  for (int c0 = 0; c0 < n; c0 += CHUNK) {
    const int nv = min(CHUNK, n - c0);
    const uint32_t reach = (carry + (uint32_t)nv * (uint32_t)(k + 3) + 31) / 32 + 1;
    for (uint32_t i = tid; i < reach; i += NT) if (i > 0 || c0 == 0) sm->bits[i] = 0;
    // this thread's values: [c0 + 16 tid, +16)
    const int i0 = c0 + tid * GE_VPT;
    if (!single) load_vals(v, n, i0, val);
    const int cnt = max(0, min(GE_VPT, n - i0));
    uint32_t longmask = 0;                                     // codes of k + 3 bits: u >= 2^k
#pragma unroll
    for (int j = 0; j < GE_VPT; j++) {
      const uint64_t a = (val[j] < 0) ? (uint64_t)(-(int64_t)val[j]) : (uint64_t)val[j];
      const uint64_t u = (val[j] > 0) ? 2 * a - 1 : 2 * a;
      if (j < cnt && (u >> k) != 0) longmask |= 1u << j;
    }
    const uint32_t mine = (uint32_t)cnt * (uint32_t)(k + 1) + 2u * (uint32_t)__popc(longmask);
    uint32_t chunk_bits;
    const uint32_t start = carry + group_exscan<NT>(mine, sm, &chunk_bits);   // bit offset inside sm->bits
    // (group_exscan's barriers also order the zeroing above before the writes below)
    if (mine) {
      uint32_t wi = start >> 5;
      int nacc = (int)(start & 31);                           // bits of word wi that belong to the threads before
      uint64_t acc = 0;
      bool first = true;
      auto emit = [&]() {
        const uint32_t word = (uint32_t)(acc >> (nacc - 32));
        if (first) { atomicOr(&sm->bits[wi], word); first = false; } else sm->bits[wi] = word;
        wi++; nacc -= 32;
      };
#pragma unroll
      for (int j = 0; j < GE_VPT; j++) {
        if (j < cnt) {
          const uint64_t a = (val[j] < 0) ? (uint64_t)(-(int64_t)val[j]) : (uint64_t)val[j];
          const uint64_t u = (val[j] > 0) ? 2 * a - 1 : 2 * a;
          uint64_t w = u + (1ull << k);
          int l = ((longmask >> j) & 1u) ? k + 3 : k + 1;
          if (l > 32) { acc = (acc << (l - 16)) | (w >> 16); nacc += l - 16; if (nacc >= 32) emit(); w &= 0xFFFFu; l = 16; }
          acc = (acc << l) | w; nacc += l; if (nacc >= 32) emit();
        }
      }
      if (nacc > 0) atomicOr(&sm->bits[wi], (uint32_t)(acc << (32 - nacc)));
    }
    ge_sync<NT>();
    // flush the complete words and keep the partial one for the next chunk
    const uint32_t filled = carry + chunk_bits;
    const bool last = (c0 + CHUNK >= n);
    const uint32_t nbytes = last ? (filled + 7) / 8 : (filled / 32) * 4;
    flush_bits<NT>(sm->bits, obase + flushed, skip, nbytes);
    if (last) return 1 + (uint32_t)(flushed + nbytes - mis);
    const uint32_t keep = sm->bits[filled / 32];
    ge_sync<NT>();
    if (tid == 0) sm->bits[0] = keep;
    flushed += (filled / 32) * 4;
    carry = filled & 31;
    skip = 0;
    ge_sync<NT>();
  }
  return 0;   // not reached
}

// One group of NT threads per frame: NT = 32, eight frames per CTA (the bulk kernel); NT = 256, one frame per CTA.
template <int NT>
__global__ void __launch_bounds__(256, GE_MINB) p1_pack_kernel(const int32_t* __restrict__ q, const int32_t* __restrict__ tq,
                                                               const int32_t* __restrict__ lpc, int B, int nq, int ntq, int nlpc,
                                                               uint8_t* __restrict__ raw, size_t stride, uint32_t* __restrict__ raw_len) {
  constexpr int GROUPS = 256 / NT;
  __shared__ ge_smem_t<NT> sm_all[GROUPS];
  ge_smem_t<NT>* sm = &sm_all[(NT == 32) ? (threadIdx.x >> 5) : 0];
  const int tid = ge_tid<NT>();
  for (int b = blockIdx.x * GROUPS + ((NT == 32) ? (int)(threadIdx.x >> 5) : 0); b < B; b += gridDim.x * GROUPS) {
    uint8_t* out = raw + (size_t)b * stride;
    uint32_t pos = 0;
    if (lpc) {                                                 // profile2.py:51  >H len || gol(lpc)
      const uint32_t L = encode_stream<NT>(lpc + (size_t)b * nlpc, nlpc, out + 2, sm);
      if (tid == 0) { out[0] = (uint8_t)(L >> 8); out[1] = (uint8_t)L; }
      pos = 2 + L;
      ge_sync<NT>();
    }
    {                                                          // profile1.py:45  >I len || gol(thresholds)
      const uint32_t L = encode_stream<NT>(tq + (size_t)b * ntq, ntq, out + pos + 4, sm);
      if (tid == 0) { out[pos] = (uint8_t)(L >> 24); out[pos + 1] = (uint8_t)(L >> 16); out[pos + 2] = (uint8_t)(L >> 8); out[pos + 3] = (uint8_t)L; }
      pos += 4 + L;
      ge_sync<NT>();
    }
    pos += encode_stream<NT>(q + (size_t)b * nq, nq, out + pos, sm);
    if (tid == 0) raw_len[b] = pos;
    ge_sync<NT>();
  }
}

// ---- compaction -----------------------------------------------------------------------------------------
// off[0] = base, off[i+1] = off[i] + len[i]; one CTA, any B.  base = *running (0 when running == NULL); *running is
// advanced to off[B] so that consecutive chunks append to one buffer.
__global__ void __launch_bounds__(1024) scan_len_kernel(const uint32_t* __restrict__ len, int B, uint64_t* __restrict__ off,
                                                        unsigned long long* running) {
  __shared__ unsigned long long wsum[32];
  __shared__ unsigned long long carry_s;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) { carry_s = running ? *running : 0; off[0] = carry_s; }
  __syncthreads();
  for (int base = 0; base < B; base += 1024) {
    const int i = base + threadIdx.x;
    const unsigned long long v = (i < B) ? len[i] : 0;
    unsigned long long inc = v;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, s); if (lane >= s) inc += t; }
    if (lane == 31) wsum[w] = inc;
    __syncthreads();
    unsigned long long pre = carry_s;
    for (int j = 0; j < w; j++) pre += wsum[j];
    if (i < B) off[i + 1] = pre + inc;
    __syncthreads();
    if (threadIdx.x == 1023) carry_s = pre + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0 && running) *running = carry_s;
}

__global__ void __launch_bounds__(128) compact_kernel(const uint8_t* __restrict__ src, size_t stride, const uint64_t* __restrict__ off,
                                                      int B, uint8_t* __restrict__ dst, uint64_t dst_cap) {
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    if (off[b + 1] > dst_cap) continue;                      // does not fit: the caller sees off[B] > capacity
    const uint8_t* s = src + (size_t)b * stride;
    uint8_t* d = dst + off[b];
    const uint32_t n = (uint32_t)(off[b + 1] - off[b]);
    // word copies where source and destination agree modulo 4, bytes otherwise
    const uint32_t head = min(n, (uint32_t)((4 - ((uintptr_t)d & 3)) & 3));
    if ((((uintptr_t)(s + head)) & 3) == 0) {
      if (threadIdx.x < head) d[threadIdx.x] = s[threadIdx.x];
      const uint32_t nw = (n - head) >> 2;
      const uint32_t* sw = (const uint32_t*)(s + head);
      uint32_t* dw = (uint32_t*)(d + head);
      for (uint32_t i = threadIdx.x; i < nw; i += blockDim.x) dw[i] = sw[i];
      for (uint32_t i = head + nw * 4 + threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
    } else {
      // misaligned pair: aligned destination words assembled from two aligned source words
      if (threadIdx.x < head) d[threadIdx.x] = s[threadIdx.x];
      const uint32_t nw = (n - head) >> 2;
      const uint8_t* sp = s + head;
      const uint32_t sh = (uint32_t)((uintptr_t)sp & 3) * 8;
      const uint32_t* sw = (const uint32_t*)((uintptr_t)sp & ~(uintptr_t)3);
      uint32_t* dw = (uint32_t*)(d + head);
      for (uint32_t i = threadIdx.x; i < nw; i += blockDim.x) dw[i] = __funnelshift_r(sw[i], sw[i + 1], sh);
      for (uint32_t i = head + nw * 4 + threadIdx.x; i < n; i += blockDim.x) d[i] = s[i];
    }
  }
}

// ---- decode ---------------------------------------------------------------------------------------------
struct bit_window {
  const uint32_t* w; uint32_t wi, nwords, last_mask;
  uint32_t nxt;                                                // word wi as loaded (0 beyond the end): the load is issued one refill
                                                               // ahead of its use, byte order and end mask are applied at the use
  uint64_t win; int navail;
  __device__ __forceinline__ uint32_t load(uint32_t i) const { return (i < nwords) ? __ldg(w + i) : 0u; }
  __device__ __forceinline__ uint32_t cook(uint32_t raw, uint32_t i) const {
    uint32_t x = __byte_perm(raw, 0, 0x0123);
    if (i == nwords - 1) x &= last_mask;
    return x;
  }
  __device__ __forceinline__ void start(uint32_t abs_bit) {    // window at absolute bit abs_bit of the word array
    wi = abs_bit >> 5;
    const int off = (int)(abs_bit & 31);
    win = ((uint64_t)(cook(load(wi), wi) << off)) << 32;
    navail = 32 - off;
    wi++;
    nxt = load(wi);
  }
  __device__ __forceinline__ void refill() {                   // navail <= 32 -> navail in (32, 64]
    if (navail <= 32) {
      win |= (uint64_t)cook(nxt, wi) << (32 - navail); navail += 32;
      wi++; nxt = load(wi);
    }
  }
  // the same where the word taken lies wholly inside the stream (callers: at least 64 bits of stream after the window start)
  __device__ __forceinline__ void refill_inside() {
    if (navail <= 32) {
      win |= (uint64_t)__byte_perm(nxt, 0, 0x0123) << (32 - navail); navail += 32;
      wi++; nxt = load(wi);
    }
  }
  __device__ __forceinline__ void drop(int n) { win = (n >= 64) ? 0 : (win << n); navail -= n; }
};

__device__ __forceinline__ uint64_t min64(uint64_t a, uint64_t b) { return a < b ? a : b; }

// ---- warp-parallel decode of one stream ---------------------------------------------------------------------
// The code boundaries form a serial chain, but the code is self-synchronising: a decoder started at a wrong bit
// almost always falls into step with the true chain within a few codes.  The bit string is cut into 32 segments, one
// per lane.  Pass 1: every lane decodes from the first bit of its segment (a guess) and reports where its chain leaves
// the segment.  Pass 2: every lane restarts where its left neighbour's chain ended; this repeats until no entry point
// changes any more (lane 0 starts at the true bit 0, so after round r lanes 0..r are final; in practice one round).
// A prefix sum of the per-lane counts gives the output positions and a last pass writes the values.
struct stream_view {
  const uint32_t* w; uint32_t nwords, last_mask, nbits; int skip, k;
};

struct lane_chain {
  uint32_t end;      // bit position where the chain leaves the segment (nbits when the stream ends inside it)
  int count;         // codes decoded
  int bad;           // stopped at a code no 32-bit value can produce
};

// Decodes codes from bit `pos` while pos < seg_hi.  WRITE: stores value i at out[obase + i] when obase + i < cap
// (four values per 16-byte store once the position is aligned).
template <bool WRITE>
__device__ __forceinline__ lane_chain run_chain(const stream_view& sv, uint32_t pos, uint32_t seg_hi, int32_t* __restrict__ out,
                                                int obase, int cap) {
  lane_chain r; r.count = 0; r.bad = 0;
  int32_t v0 = 0, v1 = 0, v2 = 0;
  int filled = 0;                                              // values held for the current aligned group of four
  const bool vec = (((uintptr_t)out) & 15) == 0;
  if (pos < seg_hi) {
    bit_window bw;
    bw.w = sv.w; bw.nwords = sv.nwords; bw.last_mask = sv.last_mask;
    bw.start((uint32_t)sv.skip + pos);
    const int64_t bias = (int64_t)1 << sv.k;
    auto put = [&](int32_t v) {                                // value number r.count of this chain
      const int at = obase + r.count;
      if (at < cap) {
        const int slot = at & 3;
        if (slot == 3 && filled == 3 && vec) { *(int4*)(out + at - 3) = make_int4(v0, v1, v2, v); filled = 0; }
        else if (slot == 3) {                                  // group started in the middle (or unaligned row): scalar stores
          if (filled >= 3) out[at - 3] = v0;
          if (filled >= 2) out[at - 2] = v1;
          if (filled >= 1) out[at - 1] = v2;
          out[at] = v; filled = 0;
        } else { v0 = v1; v1 = v2; v2 = v; filled++; }         // v2 is the most recent value
      }
    };
    // Fast walk (round 2).  What a conforming encoder writes is a '1' followed by k bits or '01' followed by k + 1 bits (k is
    // derived from the maximum, see the header), so away from the end of the stream a code is classified by its two leading
    // bits and its length is k + 1 or k + 3 -- for k <= 29 at most 32 bits, i.e. one refill and one shift per code.  Two leading
    // zeros, the last 64 bits of the stream and k >= 30 go through the general loop below, which accepts everything the
    // reference's parser accepts; both loops classify a code the same way, so where the walk stops makes no difference.
    const uint32_t lim = (sv.k <= 29 && sv.nbits > 64) ? min(seg_hi, sv.nbits - 64) : 0;
    const int l1 = sv.k + 1, l3 = sv.k + 3;
    while (pos < seg_hi) {
      const bool inside = pos < lim;
      if (inside) bw.refill_inside(); else bw.refill();        // navail in (32, 64]
      const uint32_t top = (uint32_t)(bw.win >> 62);
      if (top != 0 && inside) {
        const int l = (top & 2u) ? l1 : l3;
        if (WRITE) {
          const int32_t nv = (int32_t)((uint32_t)(bw.win >> (64 - l)) - (uint32_t)bias);      // < 2^(k+2) <= 2^31
          put((nv & 1) ? ((nv + 1) >> 1) : -(nv >> 1));
        }
        bw.win <<= l; bw.navail -= l;
        pos += (uint32_t)l;
        r.count++;
        continue;
      }
      const uint32_t remaining = sv.nbits - pos;
      if (bw.win == 0) {                                       // only padding left, or > 32 zeros in front of a code
        if (remaining > (uint32_t)bw.navail) r.bad = 1;
        pos = sv.nbits;
        break;
      }
      const int m = __clzll((long long)bw.win);
      int nbv = m + sv.k + 1;
      if (m > 32 || nbv > 62) { r.bad = 1; pos = sv.nbits; break; }
      const bool trunc = (uint32_t)(m + nbv) > remaining;      // Python slicing: a cut last code is still parsed
      if (trunc) nbv = (int)remaining - m;
      bw.drop(m);
      uint64_t x;
      if (nbv <= 32) {
        bw.refill();
        x = bw.win >> (64 - nbv);
        bw.drop(nbv);
      } else {
        bw.refill();
        const uint64_t hi = bw.win >> 32;
        bw.drop(32);
        bw.refill();
        x = (hi << (nbv - 32)) | (bw.win >> (64 - (nbv - 32)));
        bw.drop(nbv - 32);
      }
      pos = trunc ? sv.nbits : pos + (uint32_t)(m + nbv);
      if (WRITE) {
        const int64_t nv = (int64_t)x - bias;
        const int64_t val = (nv & 1) ? ((nv + 1) >> 1) : -(nv >> 1);
        put((int32_t)max((int64_t)-2147483647 - 1, min((int64_t)2147483647, val)));
      }
      r.count++;
    }
  }
  if (WRITE && filled) {                                       // flush an incomplete last group
    const int at = obase + r.count;                            // first index not written
    const int hi = min(at, cap);
    if (filled >= 3 && hi - 3 >= 0) out[hi - 3] = v0;
    if (filled >= 2) out[hi - 2] = v1;
    out[hi - 1] = v2;
  }
  r.end = pos;
  return r;
}

// all 32 lanes of a warp; returns the (warp-uniform) malformed flag
__device__ int decode_stream_warp(const uint8_t* __restrict__ d, uint64_t len, int32_t* __restrict__ out, int cap, int lane) {
  int total = 0, bad = 0;
  if (len > 1 && cap > 0) {
    stream_view sv;
    sv.k = d[0];
    if (sv.k > 32 || len > (1ull << 28)) bad = 1;
    else {
      const uint8_t* p = d + 1;
      sv.nbits = (uint32_t)((len - 1) * 8);
      sv.skip = (int)((uintptr_t)p & 3) * 8;
      sv.w = (const uint32_t*)((uintptr_t)p & ~(uintptr_t)3);
      const uint32_t span = (uint32_t)sv.skip + sv.nbits;
      sv.nwords = (span + 31) >> 5;
      const int lastbits = (int)(span - 32u * (sv.nwords - 1));
      sv.last_mask = (lastbits >= 32) ? 0xFFFFFFFFu : ~(0xFFFFFFFFu >> lastbits);
      uint32_t S = (sv.nbits + 31) / 32;
      S = (S + 31) & ~31u;
      if (S < 64) S = 64;
      const uint32_t seg_lo = min((uint32_t)lane * S, sv.nbits), seg_hi = min(seg_lo + S, sv.nbits);
      uint32_t entry = seg_lo;
      lane_chain ch = run_chain<false>(sv, entry, seg_hi, out, 0, 0);
      for (int round = 0; round < 32; round++) {
        uint32_t e = __shfl_up_sync(0xffffffffu, ch.end, 1);
        if (lane == 0) e = 0;
        e = max(e, seg_lo);                                    // a chain never ends before the next segment starts
        const bool redo = (e != entry);
        if (!__any_sync(0xffffffffu, redo)) break;
        if (redo) { entry = e; ch = run_chain<false>(sv, entry, seg_hi, out, 0, 0); }
      }
      // exclusive prefix of the counts
      int inc = ch.count;
#pragma unroll
      for (int s2 = 1; s2 < 32; s2 <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, s2); if (lane >= s2) inc += t; }
      total = __shfl_sync(0xffffffffu, inc, 31);
      const int obase = inc - ch.count;
      const lane_chain fin = run_chain<true>(sv, entry, seg_hi, out, obase, cap);
      bad = __any_sync(0xffffffffu, fin.bad && obase + fin.count < cap) ? 1 : 0;
    }
  }
  __syncwarp();
  for (int i = min(total, cap) + lane; i < cap; i += 32) out[i] = 0;
  return bad;
}

// The same scheme with all GD_CTA threads of a CTA on one stream (four times as many, four times shorter segments): what
// the chunked host pipelines use, where a call only carries ~1000 frames and a warp per frame leaves the GPU waiting on
// 1000 long serial chains.  Neighbour hand-over and the prefix sum go through shared memory.  Returns the malformed flag
// (uniform over the CTA).  sh: 2 * GD_CTA + 8 ints of shared scratch.
constexpr int GD_CTA = 128;
__device__ int decode_stream_cta(const uint8_t* __restrict__ d, uint64_t len, int32_t* __restrict__ out, int cap, int* sh) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  uint32_t* ends = (uint32_t*)sh;
  int* wsum = sh + GD_CTA;
  int total = 0, bad = 0;
  if (len > 1 && cap > 0) {                                    // uniform over the CTA
    stream_view sv;
    sv.k = d[0];
    if (sv.k > 32 || len > (1ull << 28)) bad = 1;
    else {
      const uint8_t* p = d + 1;
      sv.nbits = (uint32_t)((len - 1) * 8);
      sv.skip = (int)((uintptr_t)p & 3) * 8;
      sv.w = (const uint32_t*)((uintptr_t)p & ~(uintptr_t)3);
      const uint32_t span = (uint32_t)sv.skip + sv.nbits;
      sv.nwords = (span + 31) >> 5;
      const int lastbits = (int)(span - 32u * (sv.nwords - 1));
      sv.last_mask = (lastbits >= 32) ? 0xFFFFFFFFu : ~(0xFFFFFFFFu >> lastbits);
      uint32_t S = (sv.nbits + GD_CTA - 1) / GD_CTA;
      S = (S + 31) & ~31u;
      if (S < 64) S = 64;
      const uint32_t seg_lo = min((uint32_t)tid * S, sv.nbits), seg_hi = min(seg_lo + S, sv.nbits);
      uint32_t entry = seg_lo;
      lane_chain ch = run_chain<false>(sv, entry, seg_hi, out, 0, 0);
      for (int round = 0; round < GD_CTA; round++) {
        __syncthreads();
        ends[tid] = ch.end;
        __syncthreads();
        uint32_t e = (tid == 0) ? 0u : ends[tid - 1];
        e = max(e, seg_lo);                                    // a chain never ends before the next segment starts
        const bool redo = (e != entry);
        if (!__syncthreads_or(redo)) break;
        if (redo) { entry = e; ch = run_chain<false>(sv, entry, seg_hi, out, 0, 0); }
      }
      // exclusive prefix of the counts: warp scan, then the warp totals
      int inc = ch.count;
#pragma unroll
      for (int s2 = 1; s2 < 32; s2 <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, s2); if (lane >= s2) inc += t; }
      __syncthreads();
      if (lane == 31) wsum[wid] = inc;
      __syncthreads();
      int before = 0;
      for (int w = 0; w < GD_CTA / 32; w++) { const int t = wsum[w]; if (w < wid) before += t; total += t; }
      const int obase = before + inc - ch.count;
      const lane_chain fin = run_chain<true>(sv, entry, seg_hi, out, obase, cap);
      bad = __syncthreads_or(fin.bad && obase + fin.count < cap) ? 1 : 0;
    }
  }
  __syncthreads();
  for (int i = min(total, cap) + tid; i < cap; i += GD_CTA) out[i] = 0;
  return bad;
}

// one CTA per frame
__global__ void __launch_bounds__(GD_CTA) p1_unpack_cta_kernel(const uint8_t* __restrict__ raw, const uint64_t* __restrict__ off,
                                                               int B, int nq, int ntq, int nlpc, int32_t* __restrict__ q,
                                                               int32_t* __restrict__ tq, int32_t* __restrict__ lpc,
                                                               int32_t* __restrict__ status) {
  __shared__ int sh[2 * GD_CTA + 8];
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    const uint8_t* r = raw + off[b];
    const uint64_t L = off[b + 1] - off[b];
    int st = 0;
    uint64_t pos = 0;
    uint64_t l_at = 0, l_len = 0, t_at = 0, t_len = 0, q_at = 0, q_len = 0;
    if (L < (lpc ? 6u : 4u)) st = 1;
    else {
      if (lpc) {                                               // profile2.py:66-67
        const uint64_t ll = ((uint64_t)r[0] << 8) | r[1];
        l_at = 2; l_len = min64(ll, L - 2);
        pos = 2 + l_len;
      }
      if (L - pos >= 4) {                                      // profile1.py:61-62
        const uint64_t tl = ((uint64_t)r[pos] << 24) | ((uint64_t)r[pos + 1] << 16) | ((uint64_t)r[pos + 2] << 8) | r[pos + 3];
        t_at = pos + 4; t_len = min64(tl, L - t_at);
        q_at = t_at + t_len; q_len = L - q_at;
      }
    }
    if (lpc) st |= decode_stream_cta(r + l_at, l_len, lpc + (size_t)b * nlpc, nlpc, sh) << 1;
    st |= decode_stream_cta(r + t_at, t_len, tq + (size_t)b * ntq, ntq, sh) << 1;
    st |= decode_stream_cta(r + q_at, q_len, q + (size_t)b * nq, nq, sh) << 1;
    if (status && threadIdx.x == 0) status[b] = st;
  }
}

// one warp per frame.  status[b]: bit 0 = payload too short for its prefixes (silent frame), bit 1 = malformed code
constexpr int GD_WARPS = 4;
__global__ void __launch_bounds__(GD_WARPS * 32) p1_unpack_kernel(const uint8_t* __restrict__ raw, const uint64_t* __restrict__ off,
                                                                  int B, int nq, int ntq, int nlpc, int32_t* __restrict__ q,
                                                                  int32_t* __restrict__ tq, int32_t* __restrict__ lpc,
                                                                  int32_t* __restrict__ status) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * GD_WARPS + (threadIdx.x >> 5);
  if (b >= B) return;
  const uint8_t* r = raw + off[b];
  const uint64_t L = off[b + 1] - off[b];
  int st = 0;
  uint64_t pos = 0;
  uint64_t l_at = 0, l_len = 0, t_at = 0, t_len = 0, q_at = 0, q_len = 0;
  if (L < (lpc ? 6u : 4u)) st = 1;
  else {
    if (lpc) {                                                 // profile2.py:66-67
      const uint64_t ll = ((uint64_t)r[0] << 8) | r[1];
      l_at = 2; l_len = min64(ll, L - 2);
      pos = 2 + l_len;
    }
    if (L - pos >= 4) {                                        // profile1.py:61-62
      const uint64_t tl = ((uint64_t)r[pos] << 24) | ((uint64_t)r[pos + 1] << 16) | ((uint64_t)r[pos + 2] << 8) | r[pos + 3];
      t_at = pos + 4; t_len = min64(tl, L - t_at);
      q_at = t_at + t_len; q_len = L - q_at;
    }
  }
  if (lpc) st |= decode_stream_warp(r + l_at, l_len, lpc + (size_t)b * nlpc, nlpc, lane) << 1;
  st |= decode_stream_warp(r + t_at, t_len, tq + (size_t)b * ntq, ntq, lane) << 1;
  st |= decode_stream_warp(r + q_at, q_len, q + (size_t)b * nq, nq, lane) << 1;
  if (status && lane == 0) status[b] = st;
}

}  // namespace

extern "C" size_t frad_p1_raw_bound(int N, int C, int with_lpc) {
  auto gol = [](size_t n) { return 1 + (n * 35 + 7) / 8; };
  size_t b = 4 + gol((size_t)FRAD_NBANDS * C) + gol((size_t)N * C);
  if (with_lpc) b += 2 + gol((size_t)FRAD_LPC_LEN * C);
  return (b + 15) & ~(size_t)15;
}

extern "C" int frad_p1_pack_dev(frad_ctx* ctx, const int32_t* q, const int32_t* tq, const int32_t* lpc, int B, int N, int C,
                                uint8_t* raw, size_t raw_stride, uint32_t* raw_len) {
  if (!ctx) return FRAD_E_ARG;
  if (B < 0 || N < 1 || C < 1 || !q || !tq || !raw || !raw_len || raw_stride < frad_p1_raw_bound(N, C, lpc != nullptr))
    return frad_fail(ctx, FRAD_E_ARG, "frad_p1_pack_dev: bad argument (raw_stride must be >= frad_p1_raw_bound)");
  if (B == 0) return FRAD_OK;
  FRAD_CUDA(ctx, cudaSetDevice(ctx->device));
  // a warp per frame (eight frames per CTA); FRAD_PACK_CTA=1 forces the CTA-per-frame form
  static const char* force = getenv("FRAD_PACK_CTA");
  if (force && force[0] == '1') {
    const int grid = B < ctx->sm_count * 8 ? B : ctx->sm_count * 8;
    p1_pack_kernel<256><<<grid, 256, 0, ctx->stream>>>(q, tq, lpc, B, N * C, FRAD_NBANDS * C, FRAD_LPC_LEN * C, raw, raw_stride, raw_len);
  } else {
    const int want = (B + GE_WARPS - 1) / GE_WARPS;
    const int grid = want < ctx->sm_count * 8 ? want : ctx->sm_count * 8;
    p1_pack_kernel<32><<<grid, 256, 0, ctx->stream>>>(q, tq, lpc, B, N * C, FRAD_NBANDS * C, FRAD_LPC_LEN * C, raw, raw_stride, raw_len);
  }
  FRAD_LAUNCH_CHECK(ctx);
  return FRAD_OK;
}

// internal form: appends at *running (device counter, may be NULL) and never writes beyond dst_cap
int frad_compact_append(frad_ctx* ctx, const uint8_t* src, size_t stride, const uint32_t* len, int B, uint8_t* dst, uint64_t dst_cap,
                        uint64_t* dst_off, uint64_t* running) {
  FRAD_CUDA(ctx, cudaSetDevice(ctx->device));
  scan_len_kernel<<<1, 1024, 0, ctx->stream>>>(len, B, dst_off, (unsigned long long*)running);
  FRAD_LAUNCH_CHECK(ctx);
  if (B == 0) return FRAD_OK;
  const int grid = B < ctx->sm_count * 16 ? B : ctx->sm_count * 16;
  compact_kernel<<<grid, 128, 0, ctx->stream>>>(src, stride, dst_off, B, dst, dst_cap);
  FRAD_LAUNCH_CHECK(ctx);
  return FRAD_OK;
}

extern "C" int frad_compact_dev(frad_ctx* ctx, const uint8_t* src, size_t stride, const uint32_t* len, int B, uint8_t* dst,
                                size_t dst_cap, uint64_t* dst_off) {
  if (!ctx) return FRAD_E_ARG;
  if (B < 0 || !src || !len || !dst || !dst_off) return frad_fail(ctx, FRAD_E_ARG, "frad_compact_dev: bad argument");
  return frad_compact_append(ctx, src, stride, len, B, dst, dst_cap, dst_off, nullptr);
}

extern "C" int frad_p1_unpack_dev(frad_ctx* ctx, const uint8_t* raw, const uint64_t* raw_off, int B, int N, int C, int with_lpc,
                                  int32_t* q, int32_t* tq, int32_t* lpc, int32_t* status) {
  if (!ctx) return FRAD_E_ARG;
  if (B < 0 || N < 1 || C < 1 || !raw || !raw_off || !q || !tq || (with_lpc && !lpc))
    return frad_fail(ctx, FRAD_E_ARG, "frad_p1_unpack_dev: bad argument");
  if (B == 0) return FRAD_OK;
  FRAD_CUDA(ctx, cudaSetDevice(ctx->device));
  // few frames (the chunks of the host pipelines): a CTA per frame, so that the serial chains are four times shorter;
  // many frames: a warp per frame (less synchronisation per value).  FRAD_UNPACK_CTA=0/1 forces one of them.
  static const char* force = getenv("FRAD_UNPACK_CTA");
  const bool cta = force ? (force[0] == '1') : (B <= 16 * ctx->sm_count);
  if (cta) {
    p1_unpack_cta_kernel<<<B, GD_CTA, 0, ctx->stream>>>(raw, raw_off, B, N * C, FRAD_NBANDS * C, FRAD_LPC_LEN * C, q, tq,
                                                        with_lpc ? lpc : nullptr, status);
  } else {
    p1_unpack_kernel<<<(B + GD_WARPS - 1) / GD_WARPS, GD_WARPS * 32, 0, ctx->stream>>>(raw, raw_off, B, N * C, FRAD_NBANDS * C,
                                                                                     FRAD_LPC_LEN * C, q, tq, with_lpc ? lpc : nullptr, status);
  }
  FRAD_LAUNCH_CHECK(ctx);
  return FRAD_OK;
}

// ---- decoder cross-fade (decoder.py:28-46) + output sample format -----------------------------------------------------
// frames: f64 [B][N][C] straight from the inverse transform.  Frame b contributes its first cut = N*(ratio-1)/ratio
// samples; the first L = N - cut of them are blended with the tail of frame b-1 (prev_tail for b == 0):
//   out = rn(rn(cur * fade[i]) + rn(tail * fade[L-1-i]))      -- NumPy's two products and one sum, no FMA contraction
// and then converted to the requested sample format (pcmformat.py:49-62).  One thread per output sample, coalesced.
#include "devmath.cuh"
namespace {
__global__ void __launch_bounds__(256) overlap_kernel(const double* __restrict__ frames, long long B, int N, int C, int cut, int L,
                                                      const double* __restrict__ fade, const double* __restrict__ prev_tail,
                                                      void* __restrict__ out, int out_dtype, double* __restrict__ tail_out) {
  const long long per = (long long)cut * C, total = B * per;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const long long b = e / per;
    const int r = (int)(e - b * per), i = r / C, c = r - i * C;
    double v = frames[(b * N + i) * C + c];
    if (i < L && (b > 0 || prev_tail)) {
      const double t = (b > 0) ? frames[((b - 1) * N + cut + i) * C + c] : prev_tail[(long long)i * C + c];
      v = __dadd_rn(__dmul_rn(v, fade[i]), __dmul_rn(t, fade[L - 1 - i]));
    }
    store_pcm<double>(out, out_dtype, e, v);
  }
  if (tail_out) {
    const long long tl = (long long)L * C;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < tl; e += stride)
      tail_out[e] = frames[((B - 1) * N + cut) * C + e];
  }
}
}  // namespace

extern "C" int frad_overlap_dev(frad_ctx* ctx, const double* frames, int B, int N, int C, int ratio, const double* fade,
                                const double* prev_tail, void* out, int out_dtype, double* tail_out) {
  if (!ctx) return FRAD_E_ARG;
  if (B < 0 || N < 1 || C < 1 || ratio < 2 || !frames || !fade || !out || !frad_dtype_size(out_dtype))
    return frad_fail(ctx, FRAD_E_ARG, "frad_overlap_dev: bad argument");
  if (B == 0) return FRAD_OK;
  FRAD_CUDA(ctx, cudaSetDevice(ctx->device));
  const int cut = (int)((long long)N * (ratio - 1) / ratio), L = N - cut;
  const long long total = (long long)B * cut * C;
  long long want = (total + 255) / 256;
  const long long cap = (long long)ctx->sm_count * 16;
  overlap_kernel<<<(int)(want < cap ? (want < 1 ? 1 : want) : cap), 256, 0, ctx->stream>>>(frames, B, N, C, cut, L, fade, prev_tail, out,
                                                                                      out_dtype, tail_out);
  FRAD_LAUNCH_CHECK(ctx);
  return FRAD_OK;
}
