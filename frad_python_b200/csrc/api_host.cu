// Host-pointer entry points: copy in, run the *_dev path, copy out, synchronise.
// These are what a ctypes / cgo / JNI binding of the reference's per-frame seams would call.
#include "common.cuh"

namespace {
struct arena { frad_ctx* ctx; char* base; size_t used, cap; bool bad; };
// carve 256-byte aligned pieces out of scratch slot 2
void* take(arena& a, size_t bytes) {
  size_t o = (a.used + 255) & ~(size_t)255;
  if (o + bytes > a.cap) { a.bad = true; return nullptr; }
  a.used = o + bytes;
  return a.base + o;
}
size_t rounded(size_t b) { return ((b + 255) & ~(size_t)255) + 256; }
}  // namespace

#define ARENA(ctx, a, total)                                              \
  arena a; a.ctx = (ctx); a.used = 0; a.cap = (total); a.bad = false;     \
  a.base = (char*)frad_scratch((ctx), 2, (total));                        \
  if (!a.base) return FRAD_E_NOMEM;

extern "C" int frad_p0_encode(frad_ctx* ctx, const void* pcm, int pcm_dtype, int64_t hop, int B, int N, int C,
                              int bits, int little_endian, int prec,
                              uint8_t* payload, size_t payload_stride, double* peak, uint32_t* crc32) {
  if (!ctx) return FRAD_E_ARG;
  const int es = frad_dtype_size(pcm_dtype);
  if (!es || B < 0 || N < 1 || C < 1 || hop < 0) return frad_fail(ctx, FRAD_E_ARG, "frad_p0_encode: bad argument");
  if (B == 0) return FRAD_OK;
  const size_t in_bytes = ((size_t)(B - 1) * (size_t)hop + (size_t)N) * C * es;
  const size_t out_bytes = (size_t)B * payload_stride;
  ARENA(ctx, a, rounded(in_bytes) + rounded(out_bytes) + rounded(8 * (size_t)B) + rounded(4 * (size_t)B));
  void* d_in = take(a, in_bytes);
  uint8_t* d_out = (uint8_t*)take(a, out_bytes);
  double* d_peak = (double*)take(a, 8 * (size_t)B);
  uint32_t* d_crc = (uint32_t*)take(a, 4 * (size_t)B);
  if (a.bad) return frad_fail(ctx, FRAD_E_NOMEM, "internal arena overflow");
  int rc;
  if ((rc = frad_h2d(ctx, d_in, pcm, in_bytes))) return rc;
  if ((rc = frad_p0_encode_dev(ctx, d_in, pcm_dtype, hop, B, N, C, bits, little_endian, prec, d_out, payload_stride, d_peak))) return rc;
  if (crc32) {
    const size_t used = frad_p0_payload_bytes(N, C, bits);
    if ((rc = frad_crc32_dev(ctx, d_out, nullptr, nullptr, payload_stride, (uint32_t)used, B, d_crc))) return rc;
    if ((rc = frad_d2h(ctx, crc32, d_crc, 4 * (size_t)B))) return rc;
  }
  if ((rc = frad_d2h(ctx, payload, d_out, out_bytes))) return rc;
  if (peak && (rc = frad_d2h(ctx, peak, d_peak, 8 * (size_t)B))) return rc;
  return frad_sync(ctx);
}

extern "C" int frad_p0_decode(frad_ctx* ctx, const uint8_t* payload, size_t payload_stride, int B, int N, int C,
                              int bits, int little_endian, int prec, void* pcm, int pcm_dtype) {
  if (!ctx) return FRAD_E_ARG;
  const int es = frad_dtype_size(pcm_dtype);
  if (!es || B < 0 || N < 1 || C < 1) return frad_fail(ctx, FRAD_E_ARG, "frad_p0_decode: bad argument");
  if (B == 0) return FRAD_OK;
  const size_t in_bytes = (size_t)B * payload_stride;
  const size_t out_bytes = (size_t)B * N * C * es;
  ARENA(ctx, a, rounded(in_bytes) + rounded(out_bytes));
  uint8_t* d_in = (uint8_t*)take(a, in_bytes);
  void* d_out = take(a, out_bytes);
  if (a.bad) return frad_fail(ctx, FRAD_E_NOMEM, "internal arena overflow");
  int rc;
  if ((rc = frad_h2d(ctx, d_in, payload, in_bytes))) return rc;
  if ((rc = frad_p0_decode_dev(ctx, d_in, payload_stride, B, N, C, bits, little_endian, prec, d_out, pcm_dtype))) return rc;
  if ((rc = frad_d2h(ctx, pcm, d_out, out_bytes))) return rc;
  return frad_sync(ctx);
}

extern "C" int frad_p1_encode(frad_ctx* ctx, const void* pcm, int pcm_dtype, int64_t hop, int B, int N, int C,
                              int bits, int srate, double loss_level, int prec,
                              int32_t* q, int32_t* tq, int32_t* lpc) {
  if (!ctx) return FRAD_E_ARG;
  const int es = frad_dtype_size(pcm_dtype);
  if (!es || B < 0 || N < 1 || C < 1 || hop < 0) return frad_fail(ctx, FRAD_E_ARG, "frad_p1_encode: bad argument");
  if (B == 0) return FRAD_OK;
  const size_t in_bytes = ((size_t)(B - 1) * (size_t)hop + (size_t)N) * C * es;
  const size_t q_bytes = (size_t)B * N * C * 4, tq_bytes = (size_t)B * FRAD_NBANDS * C * 4, l_bytes = (size_t)B * FRAD_LPC_LEN * C * 4;
  ARENA(ctx, a, rounded(in_bytes) + rounded(q_bytes) + rounded(tq_bytes) + rounded(l_bytes));
  void* d_in = take(a, in_bytes);
  int32_t* d_q = (int32_t*)take(a, q_bytes);
  int32_t* d_tq = (int32_t*)take(a, tq_bytes);
  int32_t* d_l = lpc ? (int32_t*)take(a, l_bytes) : nullptr;
  if (a.bad) return frad_fail(ctx, FRAD_E_NOMEM, "internal arena overflow");
  int rc;
  if ((rc = frad_h2d(ctx, d_in, pcm, in_bytes))) return rc;
  if ((rc = frad_p1_encode_dev(ctx, d_in, pcm_dtype, hop, B, N, C, bits, srate, loss_level, prec, d_q, d_tq, d_l))) return rc;
  if ((rc = frad_d2h(ctx, q, d_q, q_bytes))) return rc;
  if ((rc = frad_d2h(ctx, tq, d_tq, tq_bytes))) return rc;
  if (lpc && (rc = frad_d2h(ctx, lpc, d_l, l_bytes))) return rc;
  return frad_sync(ctx);
}

extern "C" int frad_p1_decode(frad_ctx* ctx, const int32_t* q, const int32_t* tq, const int32_t* lpc,
                              int B, int N, int C, int bits, int srate, int prec, void* pcm, int pcm_dtype) {
  if (!ctx) return FRAD_E_ARG;
  const int es = frad_dtype_size(pcm_dtype);
  if (!es || B < 0 || N < 1 || C < 1) return frad_fail(ctx, FRAD_E_ARG, "frad_p1_decode: bad argument");
  if (B == 0) return FRAD_OK;
  const size_t q_bytes = (size_t)B * N * C * 4, tq_bytes = (size_t)B * FRAD_NBANDS * C * 4, l_bytes = (size_t)B * FRAD_LPC_LEN * C * 4;
  const size_t out_bytes = (size_t)B * N * C * es;
  ARENA(ctx, a, rounded(out_bytes) + rounded(q_bytes) + rounded(tq_bytes) + rounded(l_bytes));
  int32_t* d_q = (int32_t*)take(a, q_bytes);
  int32_t* d_tq = (int32_t*)take(a, tq_bytes);
  int32_t* d_l = lpc ? (int32_t*)take(a, l_bytes) : nullptr;
  void* d_out = take(a, out_bytes);
  if (a.bad) return frad_fail(ctx, FRAD_E_NOMEM, "internal arena overflow");
  int rc;
  if ((rc = frad_h2d(ctx, d_q, q, q_bytes))) return rc;
  if ((rc = frad_h2d(ctx, d_tq, tq, tq_bytes))) return rc;
  if (lpc && (rc = frad_h2d(ctx, d_l, lpc, l_bytes))) return rc;
  if ((rc = frad_p1_decode_dev(ctx, d_q, d_tq, d_l, B, N, C, bits, srate, prec, d_out, pcm_dtype))) return rc;
  if ((rc = frad_d2h(ctx, pcm, d_out, out_bytes))) return rc;
  return frad_sync(ctx);
}

// ---- CRC / RS over host buffers ----------------------------------------------------------------
static int stage_index(frad_ctx* ctx, arena& a, const uint64_t* off, const uint32_t* len, int B,
                       const uint64_t** d_off, const uint32_t** d_len) {
  *d_off = nullptr; *d_len = nullptr;
  if (a.bad) return frad_fail(ctx, FRAD_E_NOMEM, "internal arena overflow");
  int rc;
  if (off) {
    void* p = take(a, 8 * (size_t)B);
    if ((rc = frad_h2d(ctx, p, off, 8 * (size_t)B))) return rc;
    *d_off = (const uint64_t*)p;
  }
  if (len) {
    void* p = take(a, 4 * (size_t)B);
    if ((rc = frad_h2d(ctx, p, len, 4 * (size_t)B))) return rc;
    *d_len = (const uint32_t*)p;
  }
  return FRAD_OK;
}

template <bool IS32>
static int crc_host(frad_ctx* ctx, const uint8_t* buf, const uint64_t* off, const uint32_t* len,
                    size_t stride, uint32_t ulen, int B, size_t total, uint32_t* out) {
  if (!ctx) return FRAD_E_ARG;
  if (B < 0 || !out) return frad_fail(ctx, FRAD_E_ARG, "frad_crc: bad argument");
  if (B == 0) return FRAD_OK;
  ARENA(ctx, a, rounded(total) + rounded(12 * (size_t)B) * 2 + rounded(4 * (size_t)B));
  uint8_t* d_buf = (uint8_t*)take(a, total ? total : 1);
  uint32_t* d_out = (uint32_t*)take(a, 4 * (size_t)B);
  const uint64_t* d_off; const uint32_t* d_len;
  if (a.bad) return frad_fail(ctx, FRAD_E_NOMEM, "internal arena overflow");
  int rc;
  if ((rc = frad_h2d(ctx, d_buf, buf, total))) return rc;
  if ((rc = stage_index(ctx, a, off, len, B, &d_off, &d_len))) return rc;
  rc = IS32 ? frad_crc32_dev(ctx, d_buf, d_off, d_len, stride, ulen, B, d_out)
            : frad_crc16_dev(ctx, d_buf, d_off, d_len, stride, ulen, B, d_out);
  if (rc) return rc;
  if ((rc = frad_d2h(ctx, out, d_out, 4 * (size_t)B))) return rc;
  return frad_sync(ctx);
}

extern "C" int frad_crc32(frad_ctx* ctx, const uint8_t* buf, const uint64_t* off, const uint32_t* len,
                          size_t stride, uint32_t ulen, int B, size_t total_bytes, uint32_t* out) {
  return crc_host<true>(ctx, buf, off, len, stride, ulen, B, total_bytes, out);
}
extern "C" int frad_crc16(frad_ctx* ctx, const uint8_t* buf, const uint64_t* off, const uint32_t* len,
                          size_t stride, uint32_t ulen, int B, size_t total_bytes, uint32_t* out) {
  return crc_host<false>(ctx, buf, off, len, stride, ulen, B, total_bytes, out);
}

extern "C" int frad_rs_encode(frad_ctx* ctx, const uint8_t* buf, const uint64_t* off, const uint32_t* len,
                              size_t stride, uint32_t ulen, int B, size_t total_bytes, int dsize, int csize,
                              uint8_t* out, const uint64_t* out_off, size_t out_stride, size_t out_total_bytes) {
  if (!ctx) return FRAD_E_ARG;
  if (B < 0 || !out) return frad_fail(ctx, FRAD_E_ARG, "frad_rs_encode: bad argument");
  if (B == 0) return FRAD_OK;
  ARENA(ctx, a, rounded(total_bytes) + rounded(out_total_bytes) + 3 * rounded(8 * (size_t)B));
  uint8_t* d_buf = (uint8_t*)take(a, total_bytes ? total_bytes : 1);
  uint8_t* d_out = (uint8_t*)take(a, out_total_bytes ? out_total_bytes : 1);
  const uint64_t *d_off, *d_ooff = nullptr; const uint32_t* d_len;
  if (a.bad) return frad_fail(ctx, FRAD_E_NOMEM, "internal arena overflow");
  int rc;
  if ((rc = frad_h2d(ctx, d_buf, buf, total_bytes))) return rc;
  if ((rc = stage_index(ctx, a, off, len, B, &d_off, &d_len))) return rc;
  if (out_off) {
    void* p = take(a, 8 * (size_t)B);
    if ((rc = frad_h2d(ctx, p, out_off, 8 * (size_t)B))) return rc;
    d_ooff = (const uint64_t*)p;
  }
  if ((rc = frad_rs_encode_dev(ctx, d_buf, d_off, d_len, stride, ulen, B, dsize, csize, d_out, d_ooff, out_stride))) return rc;
  if ((rc = frad_d2h(ctx, out, d_out, out_total_bytes))) return rc;
  return frad_sync(ctx);
}

extern "C" int frad_rs_decode(frad_ctx* ctx, const uint8_t* buf, const uint64_t* off, const uint32_t* len,
                              size_t stride, uint32_t ulen, int B, size_t total_bytes, int dsize, int csize, int repair,
                              uint8_t* out, const uint64_t* out_off, size_t out_stride, size_t out_total_bytes,
                              int32_t* status, int32_t* fixed) {
  if (!ctx) return FRAD_E_ARG;
  if (B < 0 || !out) return frad_fail(ctx, FRAD_E_ARG, "frad_rs_decode: bad argument");
  if (B == 0) return FRAD_OK;
  ARENA(ctx, a, rounded(total_bytes) + rounded(out_total_bytes) + 5 * rounded(8 * (size_t)B));
  uint8_t* d_buf = (uint8_t*)take(a, total_bytes ? total_bytes : 1);
  uint8_t* d_out = (uint8_t*)take(a, out_total_bytes ? out_total_bytes : 1);
  int32_t* d_status = (int32_t*)take(a, 4 * (size_t)B);
  int32_t* d_fixed = (int32_t*)take(a, 4 * (size_t)B);
  const uint64_t *d_off, *d_ooff = nullptr; const uint32_t* d_len;
  if (a.bad) return frad_fail(ctx, FRAD_E_NOMEM, "internal arena overflow");
  int rc;
  if ((rc = frad_h2d(ctx, d_buf, buf, total_bytes))) return rc;
  if ((rc = stage_index(ctx, a, off, len, B, &d_off, &d_len))) return rc;
  if (out_off) {
    void* p = take(a, 8 * (size_t)B);
    if ((rc = frad_h2d(ctx, p, out_off, 8 * (size_t)B))) return rc;
    d_ooff = (const uint64_t*)p;
  }
  if ((rc = frad_rs_decode_dev(ctx, d_buf, d_off, d_len, stride, ulen, B, dsize, csize, repair, d_out, d_ooff, out_stride, d_status, d_fixed))) return rc;
  if ((rc = frad_d2h(ctx, out, d_out, out_total_bytes))) return rc;
  if (status && (rc = frad_d2h(ctx, status, d_status, 4 * (size_t)B))) return rc;
  if (fixed && (rc = frad_d2h(ctx, fixed, d_fixed, 4 * (size_t)B))) return rc;
  return frad_sync(ctx);
}
