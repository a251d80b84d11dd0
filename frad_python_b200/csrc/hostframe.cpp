// Host-side batch framing around the device calls (no device work, no arithmetic of the hot path).
//
// The reference frames, deflates and parses one frame at a time in Python (libfrad/encoder.py:96-104, decoder.py:51-104,
// fourier/profile1.py:50,59, tools/asfh.py:51-134), which caps a batched codec at about 1000 x real time however fast the
// GPU is.  These entry points do the same byte work for a whole batch in one call:
//   frad_deflate_frames / frad_inflate_frames   raw deflate of every frame payload on a thread pool -- the general-purpose
//                                               byte compressor stays on the host (north_star); same library (zlib) and
//                                               parameters as zlib.compress(raw, wbits=-15), so the bytes are identical
//   frad_asfh_write_frames                      frame headers + payloads of a run of frames -> one byte string (asfh.py:51-73)
//   frad_asfh_scan                              header walk over a buffer: the complete frames at its front that share one
//                                               parameter set (asfh.py:98-134, decoder.py:82-91)
#include <zlib.h>

#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/fradb200.h"

namespace {

int pick_threads(int threads, int B) {
  if (threads <= 0) {
    unsigned hc = std::thread::hardware_concurrency();
    threads = hc ? (int)hc : 4;
    if (threads > 32) threads = 32;
  }
  if (threads > B) threads = B;
  return threads < 1 ? 1 : threads;
}

template <class F>
void parallel_frames(int B, int threads, F body) {
  threads = pick_threads(threads, (B + 15) / 16);
  std::atomic<int> next(0);
  auto work = [&](int tix) {
    for (;;) {
      const int i0 = next.fetch_add(16);
      if (i0 >= B) break;
      const int i1 = i0 + 16 < B ? i0 + 16 : B;
      body(tix, i0, i1);
    }
  };
  if (threads == 1) { work(0); return; }
  std::vector<std::thread> pool;
  for (int t = 1; t < threads; t++) pool.emplace_back(work, t);
  work(0);
  for (auto& th : pool) th.join();
}

const uint8_t kSign[4] = {0xff, 0xd0, 0xd2, 0x98};
const int kSrates[12] = {96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000};

int samples_of_index(int idx) { const int base[4] = {128, 160, 192, 224}; return base[idx & 3] << (idx >> 2); }
int srate_index(uint32_t srate) {                                  // smallest legal rate >= srate (profiles.py get_valid_srate)
  int best = -1;
  for (int i = 0; i < 12; i++) if ((uint32_t)kSrates[i] >= srate && (best < 0 || kSrates[i] < kSrates[best])) best = i;
  return best < 0 ? 0 : best;
}
int samples_index(uint32_t n) {                                    // smallest legal size >= n
  int best = -1;
  for (int i = 0; i < 32; i++) { const int s = samples_of_index(i); if ((uint32_t)s >= n && (best < 0 || s < samples_of_index(best))) best = i; }
  return best < 0 ? 31 : best;
}
void be32(uint8_t* p, uint32_t v) { p[0] = v >> 24; p[1] = v >> 16; p[2] = v >> 8; p[3] = v; }
uint32_t rd32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

}  // namespace

extern "C" size_t frad_deflate_bound(size_t raw_bytes) {
  // deflateBound for raw streams: the stored-block worst case plus slack
  return raw_bytes + (raw_bytes >> 12) + (raw_bytes >> 14) + (raw_bytes >> 25) + 13 + 8;
}

// Frame i = raw[raw_off[i] .. raw_off[i+1]) -> raw deflate stream at out + out_off[i]; out_off[B] = total bytes.
// slot_stride: scratch layout -- frame i is first written at out + i * slot_stride (>= frad_deflate_bound of the longest
// frame), then the frames are moved together.  Returns FRAD_E_ARG when a frame does not fit its slot.
extern "C" int frad_deflate_frames(const uint8_t* raw, const uint64_t* raw_off, int B, uint8_t* out, size_t slot_stride,
                                   uint64_t* out_off, int threads) {
  if (B < 0 || !raw_off || !out_off || (B > 0 && (!raw || !out))) return FRAD_E_ARG;
  std::vector<uint32_t> len((size_t)B, 0);
  std::atomic<int> bad(0);
  parallel_frames(B, threads, [&](int, int i0, int i1) {
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    if (deflateInit2(&zs, Z_DEFAULT_COMPRESSION, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { bad = 1; return; }
    for (int i = i0; i < i1; i++) {
      if (i > i0) deflateReset(&zs);
      zs.next_in = const_cast<Bytef*>(raw + raw_off[i]);
      zs.avail_in = (uInt)(raw_off[i + 1] - raw_off[i]);
      zs.next_out = out + (size_t)i * slot_stride;
      zs.avail_out = (uInt)slot_stride;
      if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { bad = 1; break; }
      len[i] = (uint32_t)(slot_stride - zs.avail_out);
    }
    deflateEnd(&zs);
  });
  if (bad) return FRAD_E_ARG;
  uint64_t pos = 0;
  for (int i = 0; i < B; i++) {                                   // move together (frame 0 already sits at 0)
    out_off[i] = pos;
    if (pos != (uint64_t)i * slot_stride) std::memmove(out + pos, out + (size_t)i * slot_stride, len[i]);
    pos += len[i];
  }
  out_off[B] = pos;
  return FRAD_OK;
}

// Frame i = pay[pay_off[i] .. pay_off[i] + pay_len[i]) (raw deflate) -> inflated bytes, frames moved together in out;
// out_off[B+1].  A frame that fails to inflate contributes zero bytes (profile1.py:59-60: it decodes to silence); output
// beyond slot_stride bytes per frame is dropped (no legal frame is longer than frad_p1_raw_bound).
extern "C" int frad_inflate_frames(const uint8_t* pay, const uint64_t* pay_off, const uint32_t* pay_len, int B, uint8_t* out,
                                   size_t slot_stride, uint64_t* out_off, int threads) {
  if (B < 0 || !pay_off || !pay_len || !out_off || (B > 0 && (!pay || !out))) return FRAD_E_ARG;
  std::vector<uint32_t> len((size_t)B, 0);
  parallel_frames(B, threads, [&](int, int i0, int i1) {
    z_stream zs;
    std::memset(&zs, 0, sizeof(zs));
    if (inflateInit2(&zs, -15) != Z_OK) return;
    for (int i = i0; i < i1; i++) {
      if (i > i0) inflateReset(&zs);
      zs.next_in = const_cast<Bytef*>(pay + pay_off[i]);
      zs.avail_in = (uInt)pay_len[i];
      zs.next_out = out + (size_t)i * slot_stride;
      zs.avail_out = (uInt)slot_stride;
      const int rc = inflate(&zs, Z_FINISH);
      // zlib.decompress raises unless the stream ends cleanly; Z_BUF_ERROR with a full slot is our own truncation
      const bool truncated = (rc == Z_BUF_ERROR || rc == Z_OK) && zs.avail_out == 0;
      len[i] = (rc == Z_STREAM_END || truncated) ? (uint32_t)(slot_stride - zs.avail_out) : 0u;
    }
    inflateEnd(&zs);
  });
  uint64_t pos = 0;
  for (int i = 0; i < B; i++) {
    out_off[i] = pos;
    if (len[i] && pos != (uint64_t)i * slot_stride) std::memmove(out + pos, out + (size_t)i * slot_stride, len[i]);
    pos += len[i];
  }
  out_off[B] = pos;
  return FRAD_OK;
}

// tools/asfh.py:51-73 for a run of frames that share everything but payload, frame size, depth index and checksum.
// depth_idx / fsize may be NULL (then depth0 / fsize0 for every frame); crc NULL = no checksum field value (compact
// profiles without ECC carry none).  flush_each: a force-flush marker (asfh.py:75-87) after every frame.
// Returns the number of bytes written, or -(bytes needed) when out_cap is too small.
extern "C" int64_t frad_asfh_write_frames(int B, int profile, int ecc, int little_endian, const uint8_t* depth_idx, int depth0,
                                          int channels, uint32_t srate, const uint32_t* fsize, uint32_t fsize0, int overlap_ratio,
                                          int ecc_dsize, int ecc_csize, const uint8_t* pay, const uint64_t* pay_off,
                                          const uint32_t* pay_len, const uint32_t* crc, int flush_each, uint8_t* out, size_t out_cap) {
  if (B < 0 || (B > 0 && (!pay || !pay_off || !pay_len))) return 0;
  const bool compact = (profile == 1 || profile == 2);
  const size_t hdr = compact ? (ecc ? 16 : 12) : 32;
  size_t need = 0;
  for (int i = 0; i < B; i++) need += hdr + pay_len[i] + ((flush_each && compact) ? 12 : 0);
  if (need > out_cap || !out) return -(int64_t)need;
  uint8_t* p = out;
  for (int i = 0; i < B; i++) {
    const int depth = depth_idx ? depth_idx[i] : depth0;
    const uint32_t fs = fsize ? fsize[i] : fsize0;
    const uint8_t pfb = (uint8_t)((profile << 5) | ((ecc ? 1 : 0) << 4) | ((little_endian ? 1 : 0) << 3) | (depth & 7));
    std::memcpy(p, kSign, 4);
    be32(p + 4, pay_len[i]);
    p[8] = pfb;
    uint16_t css = 0;
    if (compact) {
      css = (uint16_t)(((channels - 1) << 10) | (srate_index(srate) << 6) | (samples_index(fs) << 1));
      p[9] = css >> 8; p[10] = css & 0xff;
      p[11] = (uint8_t)(overlap_ratio - 1 > 0 ? overlap_ratio - 1 : 0);
      if (ecc) {
        p[12] = (uint8_t)ecc_dsize; p[13] = (uint8_t)ecc_csize;
        const uint32_t c = crc ? crc[i] : 0;
        p[14] = (c >> 8) & 0xff; p[15] = c & 0xff;
      }
    } else {
      p[9] = (uint8_t)(channels - 1); p[10] = (uint8_t)ecc_dsize; p[11] = (uint8_t)ecc_csize;
      be32(p + 12, srate);
      std::memset(p + 16, 0, 8);
      be32(p + 24, fs);
      be32(p + 28, crc ? crc[i] : 0);
    }
    p += hdr;
    std::memcpy(p, pay + pay_off[i], pay_len[i]);
    p += pay_len[i];
    if (flush_each && compact) {
      std::memcpy(p, kSign, 4);
      std::memset(p + 4, 0, 4);
      p[8] = pfb;
      const uint16_t cf = (uint16_t)((((channels > 1 ? channels : 1) - 1) << 10) | (srate_index(srate) << 6) | (samples_index(fs) << 1) | 1);
      p[9] = cf >> 8; p[10] = cf & 0xff; p[11] = 0;
      p += 12;
    }
  }
  return (int64_t)(p - out);
}

// Walks buf[pos .. len): as long as a complete frame with a header of the usual kind starts exactly at the cursor and has
// the same parameters as the first one found, records it and moves on.  Stops (without consuming) at anything else: bytes
// that are not a frame signature, an incomplete header or payload, a force-flush marker, the 64-bit length escape, a change
// of any parameter.  The caller's per-frame state machine takes over from *next_pos.
// params[12] = profile, ecc, little_endian, depth index, channels, srate, fsize, overlap_ratio, ecc_dsize, ecc_csize,
//              header bytes, 0.   pay_pos / pay_len / crc: one entry per frame found (at most max_frames).
extern "C" int frad_asfh_scan(const uint8_t* buf, size_t len, size_t pos, int max_frames, int64_t* params, uint64_t* pay_pos,
                              uint32_t* pay_len, uint32_t* crc, size_t* next_pos) {
  int n = 0;
  int64_t cur[12] = {0};
  while (n < max_frames) {
    if (len - pos < 12 || std::memcmp(buf + pos, kSign, 4) != 0) break;
    const uint8_t* h = buf + pos;
    const uint32_t frm = rd32(h + 4);
    if (frm == 0xFFFFFFFFu) break;
    const int profile = h[8] >> 5, ecc = (h[8] >> 4) & 1, le = (h[8] >> 3) & 1, depth = h[8] & 7;
    int64_t p[12] = {profile, ecc, le, depth, 0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t c = 0;
    size_t hdr;
    if (profile == 1 || profile == 2) {
      const uint16_t css = (uint16_t)((h[9] << 8) | h[10]);
      if (css & 1) break;                                            // force flush
      p[4] = (css >> 10) + 1;
      if (((css >> 6) & 15) >= 12) break;                            // not a legal rate index: leave it to the caller
      p[5] = kSrates[(css >> 6) & 15];
      p[6] = samples_of_index((css >> 1) & 31);
      p[7] = h[11] ? h[11] + 1 : 0;
      hdr = ecc ? 16 : 12;
      if (len - pos < hdr) break;
      if (ecc) { p[8] = h[12]; p[9] = h[13]; c = ((uint32_t)h[14] << 8) | h[15]; }
    } else {
      hdr = 32;
      if (len - pos < hdr) break;
      p[4] = h[9] + 1; p[8] = h[10]; p[9] = h[11];
      p[5] = rd32(h + 12); p[6] = rd32(h + 24);
      c = rd32(h + 28);
    }
    p[10] = (int64_t)hdr;
    if (len - pos - hdr < frm) break;                                // payload incomplete
    if (n == 0) std::memcpy(cur, p, sizeof(cur));
    else if (std::memcmp(cur, p, sizeof(cur)) != 0) break;
    pay_pos[n] = pos + hdr; pay_len[n] = frm; crc[n] = c;
    n++;
    pos += hdr + frm;
  }
  if (n) std::memcpy(params, cur, sizeof(cur));
  *next_pos = pos;
  return n;
}
