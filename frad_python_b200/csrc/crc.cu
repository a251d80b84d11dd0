// Batched CRC-32/ISO-HDLC (zlib.crc32, tools/asfh.py:4,69) and CRC-16/ARC (common.py:4-10).
//
// One warp per message.  The 4-byte-aligned body is cut into groups of 32 words; lane l owns word l
// of every group and runs a Horner recurrence whose step is "advance by one whole group" (128 bytes),
// so every load is a fully coalesced 128-byte warp request and each word costs four table lookups
// (slicing-by-4 tables pre-multiplied by x^(8*124)).  The 32 lane remainders are folded by lane 0,
// the x^(8*124) factor is removed with an inverse-advance table, and the unaligned head / short tail
// are handled bytewise.  All tables are built on the host from the polynomial alone.
#include "common.cuh"

namespace {

struct crc_tables {            // one CRC flavour
  uint32_t byte_t[256];        // classic byte table
  uint32_t slice[4][256];      // slicing-by-4: state after (byte at position k, then the rest of the word)
  uint32_t adv[4][256];        // slice tables followed by 124 zero bytes
  uint32_t inv[4][256];        // inverse of "feed 124 zero bytes", by byte of the state
};
struct crc_tables_all { crc_tables t32, t16; };

uint32_t feed_zero_bytes(const uint32_t* bt, uint32_t s, int n) {
  for (int i = 0; i < n; i++) s = (s >> 8) ^ bt[s & 0xFF];
  return s;
}

void build_tables(uint32_t poly, int width, crc_tables* t) {
  for (uint32_t i = 0; i < 256; i++) {
    uint32_t c = i;
    for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ poly : c >> 1;
    t->byte_t[i] = c;
  }
  // word x = b0 | b1<<8 | b2<<16 | b3<<24 fed from state 0: b0 is followed by 3 more bytes, b3 by none
  for (uint32_t i = 0; i < 256; i++) {
    t->slice[3][i] = t->byte_t[i];                         // byte b3 (x >> 24)
    t->slice[2][i] = feed_zero_bytes(t->byte_t, t->byte_t[i], 1);
    t->slice[1][i] = feed_zero_bytes(t->byte_t, t->byte_t[i], 2);
    t->slice[0][i] = feed_zero_bytes(t->byte_t, t->byte_t[i], 3);
    for (int k = 0; k < 4; k++) t->adv[k][i] = feed_zero_bytes(t->byte_t, t->slice[k][i], 124);
  }
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

__device__ __forceinline__ uint32_t tab4(const uint32_t (*t)[256], uint32_t x) {
  return t[0][x & 0xFF] ^ t[1][(x >> 8) & 0xFF] ^ t[2][(x >> 16) & 0xFF] ^ t[3][x >> 24];
}

template <bool IS32>
__global__ void __launch_bounds__(256) crc_kernel(const uint8_t* __restrict__ buf, const uint64_t* __restrict__ off,
                                                  const uint32_t* __restrict__ len, size_t stride, uint32_t ulen, int B,
                                                  const crc_tables* __restrict__ gt, uint32_t* __restrict__ out) {
  __shared__ crc_tables st;
  {
    const uint32_t* src = (const uint32_t*)gt;
    uint32_t* dst = (uint32_t*)&st;
    for (int i = threadIdx.x; i < (int)(sizeof(crc_tables) / 4); i += blockDim.x) dst[i] = src[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const uint32_t init = IS32 ? 0xFFFFFFFFu : 0u;
  for (long long m = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5); m < B;
       m += (long long)gridDim.x * warps_per_block) {
    const uint8_t* p = buf + (off ? off[m] : (uint64_t)m * stride);
    const uint32_t L = len ? len[m] : ulen;
    uint32_t head = (uint32_t)((4 - ((uintptr_t)p & 3)) & 3);
    if (head > L) head = L;
    const uint32_t groups = (L - head) / 128;
    uint32_t s = init;
    if (lane == 0) for (uint32_t i = 0; i < head; i++) s = (s >> 8) ^ st.byte_t[(s ^ p[i]) & 0xFF];
    if (groups > 0) {
      s = __shfl_sync(0xffffffffu, s, 0);
      const uint32_t* w = (const uint32_t*)(p + head) + lane;
      uint32_t acc = 0;
      uint32_t x = w[0];
      if (lane == 0) x ^= s;                       // inject the running state into the first word
      acc = tab4(st.adv, x);
      for (uint32_t g = 1; g < groups; g++) acc = tab4(st.adv, acc ^ w[(size_t)g * 32]);
      // fold the lanes: r = sum_l A_{4(31-l)}(acc_l), then undo the extra 124-zero-byte advance
      uint32_t r = 0;
      for (int l = 0; l < 32; l++) {
        const uint32_t a = __shfl_sync(0xffffffffu, acc, l);
        r = tab4(st.slice, r) ^ a;
      }
      s = tab4(st.inv, r);
    }
    if (lane == 0) {
      for (uint32_t i = head + groups * 128; i < L; i++) s = (s >> 8) ^ st.byte_t[(s ^ p[i]) & 0xFF];
      out[m] = IS32 ? ~s : (s & 0xFFFFu);
    }
  }
}

int ensure_tables(frad_ctx* ctx) {
  if (ctx->crc_tables) return FRAD_OK;
  crc_tables_all* h = new crc_tables_all();
  build_tables(0xEDB88320u, 32, &h->t32);
  build_tables(0xA001u, 16, &h->t16);
  void* d = nullptr;
  cudaError_t e = cudaMalloc(&d, sizeof(crc_tables_all));
  if (e == cudaSuccess) e = cudaMemcpy(d, h, sizeof(crc_tables_all), cudaMemcpyHostToDevice);
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
  const int threads = 256, wpb = threads / 32;
  long long want = ((long long)B + wpb - 1) / wpb;
  long long cap = (long long)ctx->sm_count * 8;
  int grid = (int)(want < cap ? want : cap);
  crc_kernel<IS32><<<grid, threads, 0, ctx->stream>>>(buf, off, len, stride, ulen, B, t, out);
  FRAD_LAUNCH_CHECK(ctx);
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
