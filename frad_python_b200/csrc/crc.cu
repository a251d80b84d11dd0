// Batched CRC-32/ISO-HDLC (zlib.crc32, tools/asfh.py:4,69) and CRC-16/ARC (common.py:4-10).
//
// A message is cut into segments (host-chosen, so that every warp of the persistent grid gets about the
// same number of them); one warp per segment.  The 4-byte-aligned body of a segment is right-aligned
// onto groups of 32 words: lane l owns word l of every group and runs a Horner recurrence whose step is
// "advance by one whole group" (128 bytes), so every load is a fully coalesced 128-byte warp request and
// each word costs four table look-ups (slicing-by-4 tables pre-multiplied by x^(8*124)).
// Those look-ups are data dependent; to keep them free of bank conflicts the advance table is replicated
// once per lane in shared memory (layout [byte position][byte value][lane], 128 KiB): lane l only ever
// touches bank l.  The 32 lane remainders are folded by a five-level shuffle tree (level j advances by
// 4*2^j bytes), the extra x^(8*124) factor is removed with an inverse-advance table, and the <= 3 unaligned
// head / tail bytes are handled bytewise.  Segment remainders are combined by a small second kernel
// (advance by the length of what follows with power-of-two advance tables).  All tables are built on the
// host from the polynomial alone.
#include "common.cuh"


namespace {

struct crc_tables {            // one CRC flavour; every [4][256] table maps (byte position k, byte value) of a state
  uint32_t byte_t[256];        // classic byte table
  uint32_t slice[4][256];      // advance by 4 bytes
  uint32_t adv[4][256];        // advance by 128 bytes
  uint32_t inv[4][256];        // inverse of "advance by 124 bytes"
  uint32_t lvl[5][4][256];     // advance by 4 * 2^j bytes, j = 0..4 (fold tree)
  uint32_t pw[32][4][256];     // advance by 2^j bytes, j = 0..31 (segment combine)
};
struct crc_tables_all { crc_tables t32, t16; };

constexpr int kCrcWarps = 32;                       // warps per CTA, one CTA per SM
constexpr size_t kCrcSmem = sizeof(uint32_t) * (4 * 256 * 32 + 5 * 4 * 256 + 4 * 256 + 256);

uint32_t feed_zero_bytes(const uint32_t* bt, uint32_t s, long long n) {
  for (long long i = 0; i < n; i++) s = (s >> 8) ^ bt[s & 0xFF];
  return s;
}
uint32_t host_tab4(const uint32_t (*t)[256], uint32_t x) {
  return t[0][x & 0xFF] ^ t[1][(x >> 8) & 0xFF] ^ t[2][(x >> 16) & 0xFF] ^ t[3][x >> 24];
}

void build_tables(uint32_t poly, int width, crc_tables* t) {
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ poly : c >> 1;
    t->byte_t[i] = c;
  }
  // States are kept as 32-bit values for both widths: the bits above the CRC width are message bytes that have been
  // XORed in but not consumed yet (a whole word is injected at once), and the byte step is linear on all 32 bits.
  for (int k = 0; k < 4; k++) for (uint32_t i = 0; i < 256; i++) t->pw[0][k][i] = feed_zero_bytes(t->byte_t, i << (8 * k), 1);
  for (int j = 1; j < 32; j++)
    for (int k = 0; k < 4; k++) for (uint32_t i = 0; i < 256; i++)
      t->pw[j][k][i] = host_tab4(t->pw[j - 1], t->pw[j - 1][k][i]);
  memcpy(t->slice, t->pw[2], sizeof(t->slice));
  memcpy(t->adv, t->pw[7], sizeof(t->adv));
  for (int j = 0; j < 5; j++) memcpy(t->lvl[j], t->pw[2 + j], sizeof(t->slice));
  // matrix of A = "feed 124 zero bytes" on width-bit states, then its inverse over GF(2)
  uint64_t rows[32];   // row r: bit c set if output bit r depends on input bit c; augmented with identity in the high half
  uint32_t col[32];
  for (int c = 0; c < width; c++) col[c] = feed_zero_bytes(t->byte_t, 1u << c, 124);
  for (int r = 0; r < width; r++) {
    uint64_t v = 0;
    for (int c = 0; c < width; c++) if ((col[c] >> r) & 1) v |= 1ull << c;
    rows[r] = v | (1ull << (32 + r));
  }
  for (int c = 0; c < width; c++) {
    int piv = -1;
    for (int r = c; r < width; r++) if ((rows[r] >> c) & 1) { piv = r; break; }
    uint64_t tmp = rows[c]; rows[c] = rows[piv]; rows[piv] = tmp;    // A is invertible (x is a unit mod P)
    for (int r = 0; r < width; r++) if (r != c && ((rows[r] >> c) & 1)) rows[r] ^= rows[c];
  }
  // inverse matrix: output bit c = parity(inv_row[c] & input)  where inv_row[c] = rows[c] >> 32
  for (int k = 0; k < 4; k++) for (uint32_t i = 0; i < 256; i++) {
    uint32_t in = i << (8 * k), out = 0;
    for (int c = 0; c < width; c++) if (__builtin_parityll((rows[c] >> 32) & in)) out |= 1u << c;
    t->inv[k][i] = (width == 32 || k < 2) ? out : 0;
  }
}

__device__ __forceinline__ uint32_t tab4(const uint32_t* t, uint32_t x) {       // t: [4][256]
  return t[x & 0xFF] ^ t[256 + ((x >> 8) & 0xFF)] ^ t[512 + ((x >> 16) & 0xFF)] ^ t[768 + (x >> 24)];
}
// lane-private copy of a [4][256] table: element (k, i) of lane l at byte (k * 256 + i) * 128 + 4 l; `rep` is the
// shared-memory byte address of the lane's element (0, 0).  One byte extract, one shift-add and one load per look-up.
template <int K>
__device__ __forceinline__ uint32_t look_rep(uint32_t rep, uint32_t x) {
  const uint32_t f = __byte_perm(x, 0, 0x4440 + K);
  uint32_t v;
  asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(rep + (f << 7)), "n"(K * 256 * 128));
  return v;
}
__device__ __forceinline__ uint32_t tab4_rep(uint32_t rep, uint32_t x) {
  return look_rep<0>(rep, x) ^ look_rep<1>(rep, x) ^ look_rep<2>(rep, x) ^ look_rep<3>(rep, x);
}

struct seg_geom { uint32_t head, words, tail; int nseg; };
__device__ __forceinline__ seg_geom geometry(const uint8_t* p, uint32_t L, uint32_t seg_words) {
  seg_geom g;
  g.head = (uint32_t)((4 - ((uintptr_t)p & 3)) & 3);
  if (g.head > L) g.head = L;
  g.words = (L - g.head) >> 2;
  g.tail = (L - g.head) & 3;
  g.nseg = (int)((g.words + seg_words - 1) / seg_words);
  if (g.nseg < 1) g.nseg = 1;
  return g;
}

template <bool IS32>
__global__ void __launch_bounds__(kCrcWarps * 32, 1)
crc_kernel(const uint8_t* __restrict__ buf, const uint64_t* __restrict__ off, const uint32_t* __restrict__ len, size_t stride,
           uint32_t ulen, int B, int S, uint32_t seg_words, const crc_tables* __restrict__ gt, uint32_t* __restrict__ out,
           uint32_t* __restrict__ partial) {
  extern __shared__ __align__(16) uint32_t crc_smem[];
  uint32_t* adv_rep = crc_smem;                       // [4][256][32]
  uint32_t* lvl = adv_rep + 4 * 256 * 32;             // [5][4][256]
  uint32_t* inv = lvl + 5 * 4 * 256;                  // [4][256]
  uint32_t* byte_t = inv + 4 * 256;                   // [256]
  {
    const uint32_t* a = &gt->adv[0][0];
    for (int i = threadIdx.x; i < 4 * 256 * 32; i += blockDim.x) adv_rep[i] = __ldg(a + (i >> 5));
    const uint32_t* l = &gt->lvl[0][0][0];
    for (int i = threadIdx.x; i < 5 * 4 * 256; i += blockDim.x) lvl[i] = __ldg(l + i);
    const uint32_t* v = &gt->inv[0][0];
    for (int i = threadIdx.x; i < 4 * 256; i += blockDim.x) inv[i] = __ldg(v + i);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) byte_t[i] = __ldg(&gt->byte_t[i]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t rep = (uint32_t)__cvta_generic_to_shared(adv_rep + lane);
  const uint32_t init = IS32 ? 0xFFFFFFFFu : 0u;
  const long long units = (long long)B * S;
  for (long long u = (long long)blockIdx.x * kCrcWarps + (threadIdx.x >> 5); u < units; u += (long long)gridDim.x * kCrcWarps) {
    const int m = (int)(u / S), i = (int)(u - (long long)m * S);
    const uint8_t* p = buf + (off ? off[m] : (uint64_t)m * stride);
    const uint32_t L = len ? len[m] : ulen;
    const seg_geom g = geometry(p, L, seg_words);
    if (i >= g.nseg) continue;
    uint32_t s = 0;
    if (i == 0) {
      s = init;
      for (uint32_t b = 0; b < g.head; b++) s = (s >> 8) ^ byte_t[(s ^ p[b]) & 0xFF];
    }
    const uint32_t w0 = (uint32_t)i * seg_words;
    const uint32_t nw = (g.words > w0) ? min(seg_words, g.words - w0) : 0u;
    if (nw > 0) {
      const int pad = (int)((32u - (nw & 31u)) & 31u);
      const int G = (int)((nw + pad) >> 5);
      // word j of the segment (j = 32 g + lane - pad; j < 0 is virtual zero padding in front)
      const uint32_t* wp = (const uint32_t*)(p + g.head) + w0 + lane - pad;
      uint32_t x = (lane >= pad) ? __ldg(wp) : 0u;
      if (lane == pad) x ^= s;                     // inject the running state into the first real word
      uint32_t acc = tab4_rep(rep, x);
      int gi = 1;
      for (; gi + 8 <= G; gi += 8) {
        // one lane per 128-byte line pulls the groups 24..31 ahead into L2 (the loads below then see L2 latency)
        if (lane < 8 && gi + 24 + lane < G) asm volatile("prefetch.global.L2 [%0];" ::"l"(wp - lane + (size_t)(gi + 24 + lane) * 32));
        uint32_t xs[8];
#pragma unroll
        for (int q = 0; q < 8; q++) xs[q] = __ldg(wp + (size_t)(gi + q) * 32);
#pragma unroll
        for (int q = 0; q < 8; q++) acc = tab4_rep(rep, acc ^ xs[q]);
      }
      for (; gi < G; gi++) acc = tab4_rep(rep, acc ^ __ldg(wp + (size_t)gi * 32));
      // fold the lanes: r = sum_l A_{4(31-l)}(acc_l), then undo the extra 124-byte advance
      // (only the lanes that still carry a partial result look up: fewer lanes, fewer bank conflicts)
#pragma unroll
      for (int j = 0; j < 5; j++) {
        const uint32_t nxt = __shfl_down_sync(0xffffffffu, acc, 1 << j);
        if ((lane & ((2 << j) - 1)) == 0) acc = tab4(lvl + j * 1024, acc) ^ nxt;
      }
      s = tab4(inv, __shfl_sync(0xffffffffu, acc, 0));
    }
    if (i == g.nseg - 1) {
      const uint8_t* t = p + g.head + (size_t)g.words * 4;
      for (uint32_t b = 0; b < g.tail; b++) s = (s >> 8) ^ byte_t[(s ^ t[b]) & 0xFF];
    }
    if (lane == 0) {
      if (S == 1) out[m] = IS32 ? ~s : (s & 0xFFFFu);
      else partial[(size_t)m * S + i] = s;
    }
  }
}

// one thread per message: s = R_0; s = A^{bytes of segment i}(s) ^ R_i
template <bool IS32>
__global__ void crc_combine_kernel(const uint8_t* __restrict__ buf, const uint64_t* __restrict__ off, const uint32_t* __restrict__ len,
                                   size_t stride, uint32_t ulen, int B, int S, uint32_t seg_words,
                                   const crc_tables* __restrict__ gt, const uint32_t* __restrict__ partial, uint32_t* __restrict__ out) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= B) return;
  const uint8_t* p = buf + (off ? off[m] : (uint64_t)m * stride);
  const uint32_t L = len ? len[m] : ulen;
  const seg_geom g = geometry(p, L, seg_words);
  uint32_t s = partial[(size_t)m * S];
  for (int i = 1; i < g.nseg; i++) {
    const uint32_t w0 = (uint32_t)i * seg_words;
    uint32_t nbytes = 4u * min(seg_words, g.words - w0) + ((i == g.nseg - 1) ? g.tail : 0u);
    for (int j = 0; nbytes; j++, nbytes >>= 1) {
      if (nbytes & 1u) {
        const uint32_t* t = &gt->pw[j][0][0];
        s = __ldg(t + (s & 0xFF)) ^ __ldg(t + 256 + ((s >> 8) & 0xFF)) ^ __ldg(t + 512 + ((s >> 16) & 0xFF)) ^ __ldg(t + 768 + (s >> 24));
      }
    }
    s ^= partial[(size_t)m * S + i];
  }
  out[m] = IS32 ? ~s : (s & 0xFFFFu);
}

int ensure_tables(frad_ctx* ctx) {
  if (ctx->crc_tables) return FRAD_OK;
  crc_tables_all* h = new crc_tables_all();
  build_tables(0xEDB88320u, 32, &h->t32);
  build_tables(0xA001u, 16, &h->t16);
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, sizeof(crc_tables_all));
  if (e == cudaSuccess) e = frad_upload(d, h, sizeof(crc_tables_all));
  delete h;
  if (e != cudaSuccess) { cudaGetLastError(); if (d) cudaFree(d); return frad_fail(ctx, FRAD_E_NOMEM, "crc table upload failed"); }
  ctx->crc_tables = d;
  return FRAD_OK;
}

template <bool IS32>
int crc_dev(frad_ctx* ctx, const uint8_t* buf, const uint64_t* off, const uint32_t* len, size_t stride, uint32_t ulen,
            int B, uint32_t* out) {
  if (!ctx) return FRAD_E_ARG;
  if (B < 0 || !out || (!buf && B > 0)) return frad_fail(ctx, FRAD_E_ARG, "frad_crc*_dev: bad argument");
  if (B == 0) return FRAD_OK;
  FRAD_CUDA(ctx, cudaSetDevice(ctx->device));
  int rc = ensure_tables(ctx);
  if (rc) return rc;
  const crc_tables_all* all = (const crc_tables_all*)ctx->crc_tables;
  const crc_tables* t = IS32 ? &all->t32 : &all->t16;
  // segmentation: aim at >= 8 segments per resident warp, segments of at least 1 KiB (multiples of 32 words)
  const uint32_t max_len = ulen;          // with len[]: its upper bound; 0 = unknown: one segment per message
  const long long warps = (long long)ctx->sm_count * kCrcWarps;
  int S = 1;
  uint32_t seg_words = 0x40000000u;
  if (max_len >= 2048) {
    const uint32_t wmax = max_len / 4 + 1;
    long long want = (8 * warps + B - 1) / B;
    const long long most = wmax / 256;
    if (want > most) want = most;
    if (want > 1) {
      seg_words = (uint32_t)(((wmax + want - 1) / want + 31) / 32 * 32);
      S = (int)((wmax + seg_words - 1) / seg_words);
    }
  }
  uint32_t* partial = nullptr;
  if (S > 1) {
    partial = (uint32_t*)frad_scratch(ctx, 3, sizeof(uint32_t) * (size_t)B * S);
    if (!partial) return FRAD_E_NOMEM;
  }
  const long long units = (long long)B * S;
  long long want_ctas = (units + kCrcWarps - 1) / kCrcWarps;
  const int grid = (int)(want_ctas < ctx->sm_count ? want_ctas : ctx->sm_count);
  auto k = crc_kernel<IS32>;
  FRAD_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCrcSmem));
  k<<<grid, kCrcWarps * 32, kCrcSmem, ctx->stream>>>(buf, off, len, stride, ulen, B, S, seg_words, t, out, partial);
  FRAD_LAUNCH_CHECK(ctx);
  if (S > 1) {
    crc_combine_kernel<IS32><<<(B + 127) / 128, 128, 0, ctx->stream>>>(buf, off, len, stride, ulen, B, S, seg_words, t, partial, out);
    FRAD_LAUNCH_CHECK(ctx);
  }
  return FRAD_OK;
}

}  // namespace

extern "C" int frad_crc32_dev(frad_ctx* ctx, const uint8_t* buf, const uint64_t* off, const uint32_t* len,
                              size_t stride, uint32_t ulen, int B, uint32_t* out) {
  return crc_dev<true>(ctx, buf, off, len, stride, ulen, B, out);
}
extern "C" int frad_crc16_dev(frad_ctx* ctx, const uint8_t* buf, const uint64_t* off, const uint32_t* len,
                              size_t stride, uint32_t ulen, int B, uint32_t* out) {
  return crc_dev<false>(ctx, buf, off, len, stride, ulen, B, out);
}
