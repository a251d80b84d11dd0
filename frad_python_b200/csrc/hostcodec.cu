// Host-side Exp-Golomb-Rice bit coder of fourier/tools/p1tools.py:46-74 (native C++: the reference's
// pure-Python bit-string loop is 74-85 % of its profile-1 CPU time).  The bit format is the
// reference's: one byte k = ceil(log2(max|v|)) (0 when all values are 0), then per value
// u = 2v-1 (v > 0) else -2v, w = u + 2^k written as (bitlen(w) - k - 1) zeros followed by w in
// bitlen(w) bits, MSB first, zero-padded to a byte.  Deflate / inflate stay with zlib on the host.
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#include "../../include/fradb200.h"

namespace {

inline int bitlen64(uint64_t v) { return v ? 64 - __builtin_clzll(v) : 0; }

struct bit_writer {
  uint8_t* out; size_t cap, pos; uint64_t acc; int nacc; bool overflow;
  void put(uint64_t v, int nbits) {          // nbits <= 57
    acc = (acc << nbits) | v; nacc += nbits;
    while (nacc >= 8) {
      nacc -= 8;
      if (pos < cap) out[pos] = (uint8_t)(acc >> nacc); else overflow = true;
      pos++;
    }
  }
  void zeros(int n) { while (n > 0) { int c = n > 32 ? 32 : n; put(0, c); n -= c; } }
  void flush() { if (nacc > 0) { int pad = 8 - nacc; put(0, pad); } }
};

int64_t encode_one(const int32_t* v, size_t n, uint8_t* out, size_t cap) {
  if (n == 0) { if (cap >= 1) out[0] = 0; return 1; }
  uint64_t mx = 0;
  for (size_t i = 0; i < n; i++) { int64_t x = v[i]; uint64_t a = (uint64_t)(x < 0 ? -x : x); if (a > mx) mx = a; }
  const int k = (mx <= 1) ? 0 : bitlen64(mx - 1);
  bit_writer bw{out, cap, 0, 0, 0, false};
  bw.put((uint64_t)k, 8);
  for (size_t i = 0; i < n; i++) {
    const int64_t x = v[i];
    const uint64_t u = (x > 0) ? (uint64_t)(2 * x - 1) : (uint64_t)(-2 * x);
    const uint64_t w = u + (1ull << k);
    const int nb = bitlen64(w);
    bw.zeros(nb - (k + 1));
    bw.put(w, nb);
  }
  bw.flush();
  return bw.overflow ? -(int64_t)bw.pos : (int64_t)bw.pos;
}

// returns the number of decoded values (may exceed cap: surplus values are counted, not stored)
int64_t decode_one(const uint8_t* d, size_t len, int32_t* out, size_t cap) {
  if (len == 0) return 0;
  const int k = d[0];
  const uint8_t* p = d + 1;
  const size_t nbits = (len - 1) * 8;
  size_t pos = 0;
  int64_t count = 0;
  auto bit = [&](size_t i) -> int { return (p[i >> 3] >> (7 - (i & 7))) & 1; };
  while (pos < nbits) {
    size_t q = pos;
    // skip zero bits (bytewise when aligned)
    while (q < nbits) {
      if ((q & 7) == 0 && p[q >> 3] == 0) { q += 8; continue; }
      if (bit(q)) break;
      q++;
    }
    if (q >= nbits) break;                       // no set bit left: padding
    const size_t m = q - pos;
    size_t width = 2 * m + (size_t)k + 1;
    if (pos + width > nbits) width = nbits - pos;   // truncated tail: Python slicing semantics
    uint64_t w = 0;
    const size_t start = pos + m;                // leading zeros contribute nothing
    for (size_t i = start; i < pos + width; i++) w = (w << 1) | (uint64_t)bit(i);
    const int64_t nv = (int64_t)w - (int64_t)(k < 63 ? (1ull << k) : 0);
    const int64_t val = (nv & 1) ? ((nv + 1) >> 1) : -(nv >> 1);
    if ((size_t)count < cap) out[count] = (int32_t)val;
    count++;
    pos += width;
  }
  return count;
}

template <typename F>
void parallel_for(int n, int threads, F f) {
  if (threads <= 1 || n < 64) { for (int i = 0; i < n; i++) f(i); return; }
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++)
    pool.emplace_back([=]() { for (int i = t; i < n; i += threads) f(i); });
  for (auto& th : pool) th.join();
}

}  // namespace

extern "C" int64_t frad_golomb_encode_host(const int32_t* vals, size_t n, uint8_t* out, size_t cap) {
  return encode_one(vals, n, out, cap);
}
extern "C" int64_t frad_golomb_decode_host(const uint8_t* data, size_t len, int32_t* out, size_t cap) {
  return decode_one(data, len, out, cap);
}

// B equal-length integer streams -> B byte strings at out + i*out_stride; out_len[i] receives the length
// (negative: needed size when it did not fit).  threads <= 0: hardware concurrency.
extern "C" int frad_golomb_encode_batch_host(const int32_t* vals, int B, size_t n_per, uint8_t* out, size_t out_stride,
                                             int64_t* out_len, int threads) {
  if (B < 0 || (B > 0 && (!vals || !out || !out_len))) return FRAD_E_ARG;
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  parallel_for(B, threads, [=](int i) { out_len[i] = encode_one(vals + (size_t)i * n_per, n_per, out + (size_t)i * out_stride, out_stride); });
  return FRAD_OK;
}

// B byte strings (off/len) -> B integer rows of `cap` values (zero-filled tails); count[i] = decoded values
extern "C" int frad_golomb_decode_batch_host(const uint8_t* data, const uint64_t* off, const uint32_t* len, int B,
                                             int32_t* out, size_t cap, int64_t* count, int threads) {
  if (B < 0 || (B > 0 && (!data || !off || !len || !out || !count))) return FRAD_E_ARG;
  if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
  parallel_for(B, threads, [=](int i) {
    int32_t* row = out + (size_t)i * cap;
    const int64_t c = decode_one(data + off[i], len[i], row, cap);
    for (size_t j = (size_t)(c < 0 ? 0 : c); j < cap; j++) row[j] = 0;
    count[i] = c;
  });
  return FRAD_OK;
}
