/* fradb200 -- C ABI of the B200-native FrAD hot path (libfradb200.so).
 *
 * Drop-in boundary for the numeric seams of H4n-uL/FrAD_Python (paths relative to the
 * reference's src/libfrad/):
 *   fourier/profile0.py:14,46   analogue / digital   -> frad_p0_encode*, frad_p0_decode*
 *   fourier/profile4.py:14,46   analogue / digital   -> frad_p4_encode*, frad_p4_decode*
 *   fourier/profile1.py:15,54   analogue / digital   -> frad_p1_encode*, frad_p1_decode*  (numeric part;
 *   fourier/profile2.py:15,58   analogue / digital   -> frad_p2_encode*, frad_p2_decode*   bit strings via frad_p1_pack/unpack_dev,
 *                                                                                          deflate stays on the host)
 *   tools/ecc.py:6,14           encode / decode      -> frad_rs_encode*, frad_rs_decode*
 *   tools/asfh.py:4,69 (zlib.crc32), common.py:6 (crc16_ansi) -> frad_crc32*, frad_crc16*
 *
 * Conventions: every function returns 0 on success or a negative FRAD_E_* code; the text is
 * available from frad_last_error().  No exceptions or exit() cross the ABI.  The caller owns all
 * buffers.  A context is bound to one GPU and one CUDA stream and is NOT thread-safe; distinct
 * contexts are independent (one per GPU / host thread).  There is no CPU fallback: without a
 * usable CUDA device frad_ctx_create fails.
 *
 * "*_dev" entry points take DEVICE pointers and only enqueue work on the context's stream
 * (results valid after frad_sync).  The entry points without the suffix take HOST pointers
 * (ideally from frad_host_alloc, i.e. pinned), copy in, run, copy out and synchronise.
 *
 * Layouts (all C-contiguous):
 *   pcm        [..][C]   interleaved samples; frame b starts at sample b*hop (hop == N without overlap)
 *   p0 payload [B][stride] bytes, value index k*C + c, N*C*bits/8 bytes used per frame (12 bit: ceil(1.5*N*C))
 *   q          [B][N*C]  int32, index k*C + c          (profile1.py:34-36)
 *   tq         [B][27*C] int32, index band*C + c       (profile1.py:38-40)
 *   lpc        [B][13*C] int32, index i*C + c          (profile2.py:46)
 */
#ifndef FRADB200_H
#define FRADB200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct frad_ctx frad_ctx;

enum {
  FRAD_OK = 0,
  FRAD_E_CUDA = -1,      /* CUDA runtime error (see frad_last_error) */
  FRAD_E_ARG = -2,       /* invalid argument */
  FRAD_E_NOMEM = -3,     /* device / pinned allocation failed */
  FRAD_E_UNSUPPORTED = -4
};

/* sample formats of PCM buffers (src/libfrad/backend/pcmformat.py:4-62); little-endian host order */
enum {
  FRAD_F32 = 0, FRAD_F64 = 1,
  FRAD_S8 = 2, FRAD_S16 = 3, FRAD_S32 = 4, FRAD_S64 = 5,
  FRAD_U8 = 6, FRAD_U16 = 7, FRAD_U32 = 8, FRAD_U64 = 9,
  FRAD_F16 = 10
};

/* compute precision of the transform: FRAD_PREC_REF follows the reference:
 * forward = f32 for f16/f32 PCM, f64 otherwise; inverse = f64. */
enum { FRAD_PREC_REF = 0, FRAD_PREC_F32 = 1, FRAD_PREC_F64 = 2 };

#define FRAD_NBANDS 27
#define FRAD_LPC_LEN 13

/* ---- context, memory, stream --------------------------------------------------------------- */
int  frad_ctx_create(int device, frad_ctx** out);
void frad_ctx_destroy(frad_ctx* ctx);
const char* frad_last_error(frad_ctx* ctx);      /* ctx may be NULL: last creation error */
int  frad_sync(frad_ctx* ctx);
void* frad_stream(frad_ctx* ctx);                /* the cudaStream_t work is enqueued on */
int  frad_device_count(void);
/* "domain:bus:device.function" of a GPU (>= 16 bytes), e.g. to look up its NUMA node under /sys/bus/pci/devices/ so that a
 * rank's pinned staging buffers are allocated next to the GPU they feed (frad_python_b200/shard.py: bind_host_near_gpu) */
int  frad_device_pci_bus_id(int device, char* out, int cap);
int  frad_sm_count(frad_ctx* ctx);
uint64_t frad_kernel_launches(frad_ctx* ctx);    /* kernels launched through this ctx so far */

/* test / diagnostic switches (default 0): results must not depend on them beyond the stated tolerances */
enum { FRAD_OPT_FORCE_GENERIC = 1,   /* bypass the warp-FFT kernels (N = 2048 and the other N = 64 R1 lengths): generic kernels only */
       FRAD_OPT_EXACT_QUANT = 2 };   /* profile-1 quantiser: always evaluate the f64 formula */
int  frad_set_option(frad_ctx* ctx, int option, int value);

/* Which transform kernels a frame length runs on (needs no context or device): 2 = the N = 2048 kernels (the reference's default
 * frame, src/tools/cli.py:26), 1 = the multi-sequence warp-FFT kernels (N = 64 R1 < 2048 with R1 a power of two or one of 3, 5, 6,
 * 7, 10, 12, 14, 20, 24, 28: fourteen of the compact sizes of fourier/profiles.py:14-23), 0 = the generic shared-memory kernels
 * (every other length the reference accepts: same results, lower throughput).  Channel counts that do not divide a tile of
 * class 1, profile 2 at class 1 and misaligned buffers also take the generic kernels. */
int  frad_frame_length_class(int N);

void* frad_host_alloc(frad_ctx* ctx, size_t bytes);   /* pinned host memory, NULL on failure */
void  frad_host_free(frad_ctx* ctx, void* p);
int   frad_host_register(frad_ctx* ctx, void* p, size_t bytes);   /* page-lock caller-owned memory (e.g. a shared mapping) */
int   frad_host_unregister(frad_ctx* ctx, void* p);
void* frad_dev_alloc(frad_ctx* ctx, size_t bytes);    /* device memory, NULL on failure */
void  frad_dev_free(frad_ctx* ctx, void* p);
int  frad_h2d(frad_ctx* ctx, void* dst_dev, const void* src_host, size_t bytes);   /* async on the stream */
int  frad_d2h(frad_ctx* ctx, void* dst_host, const void* src_dev, size_t bytes);   /* async on the stream */
int  frad_memset_dev(frad_ctx* ctx, void* dst_dev, int value, size_t bytes);

/* CUDA-event timing on the context's stream: record returns a slot id in [0, 64) */
int  frad_event_record(frad_ctx* ctx, int slot);
int  frad_event_elapsed_ms(frad_ctx* ctx, int slot_start, int slot_stop, float* ms);

/* ---- profile 0: DCT archive (profile0.py) -------------------------------------------------- */
/* peak[b] receives max|X| of frame b (as double) so the caller can apply the depth bump of
 * profile0.py:24-26; payload bytes are already produced for the requested depth. */
size_t frad_p0_payload_bytes(int N, int C, int bits);
int frad_p0_encode_dev(frad_ctx* ctx, const void* pcm, int pcm_dtype, int64_t hop, int B, int N, int C,
                       int bits, int little_endian, int prec,
                       uint8_t* payload, size_t payload_stride, double* peak);
int frad_p0_decode_dev(frad_ctx* ctx, const uint8_t* payload, size_t payload_stride, int B, int N, int C,
                       int bits, int little_endian, int prec,
                       void* pcm, int pcm_dtype);     /* pcm_dtype: FRAD_F64 (reference) or FRAD_F32 */
int frad_p0_encode(frad_ctx* ctx, const void* pcm, int pcm_dtype, int64_t hop, int B, int N, int C,
                   int bits, int little_endian, int prec,
                   uint8_t* payload, size_t payload_stride, double* peak, uint32_t* crc32 /* may be NULL */);
int frad_p0_decode(frad_ctx* ctx, const uint8_t* payload, size_t payload_stride, int B, int N, int C,
                   int bits, int little_endian, int prec, void* pcm, int pcm_dtype);

/* ---- profile 4: PCM archive (fourier/profile4.py:14-63; the reference CLI's default profile, src/tools/cli.py:28) ---- */
/* No transform: the samples themselves go through profile 0's value packing (cast to f16 / f32 / f64, 12 / 24 / 48-bit
 * truncation) in their interleaved order; peak[b] = max|sample| of frame b for the depth bump of profile4.py:23-25.
 * Payload size and layout are those of frad_p0_payload_bytes. */
int frad_p4_encode_dev(frad_ctx* ctx, const void* pcm, int pcm_dtype, int64_t hop, int B, int N, int C,
                       int bits, int little_endian, uint8_t* payload, size_t payload_stride, double* peak);
int frad_p4_decode_dev(frad_ctx* ctx, const uint8_t* payload, size_t payload_stride, int B, int N, int C,
                       int bits, int little_endian, void* pcm, int pcm_dtype);
int frad_p4_encode(frad_ctx* ctx, const void* pcm, int pcm_dtype, int64_t hop, int B, int N, int C,
                   int bits, int little_endian, uint8_t* payload, size_t payload_stride, double* peak,
                   uint32_t* crc32 /* may be NULL */);
int frad_p4_decode(frad_ctx* ctx, const uint8_t* payload, size_t payload_stride, int B, int N, int C,
                   int bits, int little_endian, void* pcm, int pcm_dtype);

/* ---- profile 1 / 2: psychoacoustic quantisation (profile1.py, profile2.py, p1tools, p2tools) - */
/* N must be one of the 32 compact frame sizes; srate one of the 12 compact rates; the caller pads
 * short frames with zeros (profile1.py:19).  lpc == NULL selects profile 1, otherwise profile 2. */
int frad_p1_encode_dev(frad_ctx* ctx, const void* pcm, int pcm_dtype, int64_t hop, int B, int N, int C,
                       int bits, int srate, double loss_level, int prec,
                       int32_t* q, int32_t* tq, int32_t* lpc);
int frad_p1_decode_dev(frad_ctx* ctx, const int32_t* q, const int32_t* tq, const int32_t* lpc,
                       int B, int N, int C, int bits, int srate, int prec, void* pcm, int pcm_dtype);
int frad_p1_encode(frad_ctx* ctx, const void* pcm, int pcm_dtype, int64_t hop, int B, int N, int C,
                   int bits, int srate, double loss_level, int prec,
                   int32_t* q, int32_t* tq, int32_t* lpc);
int frad_p1_decode(frad_ctx* ctx, const int32_t* q, const int32_t* tq, const int32_t* lpc,
                   int B, int N, int C, int bits, int srate, int prec, void* pcm, int pcm_dtype);

/* ---- profile 1 / 2 frame payloads before deflate, packed on the device ---------------------------------- */
/* The byte string the reference hands to zlib (profile1.py:43-46: >I len || gol(tq) || gol(q); profile2.py:48-51:
 * >H len || gol(lpc) || >I len || gol(tq) || gol(q)) with gol() = Exp-Golomb-Rice (p1tools.py:46-60).
 * frad_p1_raw_bound: worst-case bytes of one frame (a multiple of 16); raw_stride must be at least that.
 * pack: q/tq[/lpc] -> raw[b*raw_stride ..], raw_len[b].  lpc == NULL selects the profile-1 layout.
 * compact: B strided byte strings -> one contiguous buffer; dst_off[B+1] receives the offsets (dst_off[B] = total);
 *          strings that would end beyond dst_cap are not written (check dst_off[B] <= dst_cap).
 * unpack: raw_off[B+1] delimits the (inflated) payloads inside raw; missing values are zero (profile1.py:67-68),
 *         status[b] (may be NULL): bit 0 = shorter than its length prefixes (decodes to silence),
 *         bit 1 = a code no 32-bit value can produce (decoding of that stream stopped there). */
size_t frad_p1_raw_bound(int N, int C, int with_lpc);
int frad_p1_pack_dev(frad_ctx* ctx, const int32_t* q, const int32_t* tq, const int32_t* lpc, int B, int N, int C,
                     uint8_t* raw, size_t raw_stride, uint32_t* raw_len);
int frad_compact_dev(frad_ctx* ctx, const uint8_t* src, size_t stride, const uint32_t* len, int B,
                     uint8_t* dst, size_t dst_cap, uint64_t* dst_off);
int frad_p1_unpack_dev(frad_ctx* ctx, const uint8_t* raw, const uint64_t* raw_off, int B, int N, int C, int with_lpc,
                       int32_t* q, int32_t* tq, int32_t* lpc, int32_t* status);
/* Host pointers, pipelined: PCM -> payloads (raw_off[B+1]; FRAD_E_NOMEM with raw_off[B] = bytes needed when raw_cap is
 * too small) and payloads -> PCM.  profile is 1 or 2. */
int frad_p1_encode_raw(frad_ctx* ctx, const void* pcm, int pcm_dtype, int64_t hop, int B, int N, int C,
                       int bits, int srate, double loss_level, int prec, int profile,
                       uint8_t* raw, size_t raw_cap, uint64_t* raw_off);
int frad_p1_decode_raw(frad_ctx* ctx, const uint8_t* raw, const uint64_t* raw_off, int B, int N, int C,
                       int bits, int srate, int prec, int profile, void* pcm, int pcm_dtype, int32_t* status);

/* ---- decoder cross-fade (decoder.py:28-46) fused with the output sample format (pcmformat.py:49-62) ------------- */
/* frames: f64 [B][N][C] (device).  Each frame yields its first cut = N*(ratio-1)/ratio samples; the first L = N - cut of
 * them are blended with the previous frame's tail: out = cur*fade[i] + tail*fade[L-1-i] (fade: L window values, device).
 * prev_tail: f64 [L][C] from the previous call or NULL (then frame 0 is not faded); out: [B*cut][C] samples of out_dtype;
 * tail_out (may be NULL): f64 [L][C], the tail of frame B-1 for the next call. */
int frad_overlap_dev(frad_ctx* ctx, const double* frames, int B, int N, int C, int ratio, const double* fade,
                     const double* prev_tail, void* out, int out_dtype, double* tail_out);
/* Host pointers, pipelined: payloads (as for frad_p1_decode_raw) -> cross-faded PCM in pcm_dtype, [B*cut][C]. */
int frad_p1_decode_stream(frad_ctx* ctx, const uint8_t* raw, const uint64_t* raw_off, int B, int N, int C,
                          int bits, int srate, int prec, int profile, int ratio, const double* fade,
                          const double* prev_tail, void* pcm, int pcm_dtype, double* tail_out, int32_t* status);

/* ---- CRC (zlib.crc32 / crc16_ansi) ---------------------------------------------------------- */
/* Message i is buf[off[i] .. off[i]+len[i]); off == NULL means i*stride, len == NULL means every message has ulen bytes.
 * With len != NULL, ulen is an upper bound of len[] (the *_dev entry points cannot read device memory on the host):
 * 0 = unknown, which is correct but selects the one-message-per-warp / one-codeword-per-thread kernels. */
int frad_crc32_dev(frad_ctx* ctx, const uint8_t* buf, const uint64_t* off, const uint32_t* len,
                   size_t stride, uint32_t ulen, int B, uint32_t* out);
int frad_crc16_dev(frad_ctx* ctx, const uint8_t* buf, const uint64_t* off, const uint32_t* len,
                   size_t stride, uint32_t ulen, int B, uint32_t* out);
int frad_crc32(frad_ctx* ctx, const uint8_t* buf, const uint64_t* off, const uint32_t* len,
               size_t stride, uint32_t ulen, int B, size_t total_bytes, uint32_t* out);
int frad_crc16(frad_ctx* ctx, const uint8_t* buf, const uint64_t* off, const uint32_t* len,
               size_t stride, uint32_t ulen, int B, size_t total_bytes, uint32_t* out);

/* ---- Reed-Solomon ECC (tools/ecc.py; GF(2^8) 0x11D, alpha = 2, fcr = 0) ----------------------- */
/* Message i (len[i] bytes at off[i]) is split into dsize-byte chunks; each becomes chunk || csize
 * parity bytes at out + out_off[i].  Encoded length = len + ceil(len/dsize)*csize. */
size_t frad_rs_encoded_len(size_t len, int dsize, int csize);
size_t frad_rs_decoded_len(size_t len, int dsize, int csize);
int frad_rs_encode_dev(frad_ctx* ctx, const uint8_t* buf, const uint64_t* off, const uint32_t* len,
                       size_t stride, uint32_t ulen, int B, int dsize, int csize,
                       uint8_t* out, const uint64_t* out_off, size_t out_stride);
/* repair != 0: correct up to csize/2 byte errors per codeword, otherwise zero-fill the chunk
 * (ecc.py:19-22); repair == 0: strip parity only (ecc.py:23).
 * status (may be NULL) [B]: number of codewords zero-filled in message i (0 = all recovered);
 * fixed (may be NULL) [B]: total corrected byte errors in message i. */
int frad_rs_decode_dev(frad_ctx* ctx, const uint8_t* buf, const uint64_t* off, const uint32_t* len,
                       size_t stride, uint32_t ulen, int B, int dsize, int csize, int repair,
                       uint8_t* out, const uint64_t* out_off, size_t out_stride,
                       int32_t* status, int32_t* fixed);
int frad_rs_encode(frad_ctx* ctx, const uint8_t* buf, const uint64_t* off, const uint32_t* len,
                   size_t stride, uint32_t ulen, int B, size_t total_bytes, int dsize, int csize,
                   uint8_t* out, const uint64_t* out_off, size_t out_stride, size_t out_total_bytes);
int frad_rs_decode(frad_ctx* ctx, const uint8_t* buf, const uint64_t* off, const uint32_t* len,
                   size_t stride, uint32_t ulen, int B, size_t total_bytes, int dsize, int csize, int repair,
                   uint8_t* out, const uint64_t* out_off, size_t out_stride, size_t out_total_bytes,
                   int32_t* status, int32_t* fixed);

/* ---- host-side batch framing (no device work; csrc/hostframe.cpp) ------------------------------------------------------
 * What the reference does per frame in Python around the seams above, for a whole batch in one call: raw deflate / inflate of
 * the profile 1 / 2 payloads (fourier/profile1.py:50,59 -- zlib, same parameters as zlib.compress(raw, wbits=-15), on a
 * thread pool; the byte compressor stays on the host), the frame headers (tools/asfh.py:51-73) and the header walk of the
 * decoder (tools/asfh.py:98-134, decoder.py:82-91).  threads <= 0: one per hardware thread (at most 32).
 * deflate: frame i = raw[raw_off[i] .. raw_off[i+1]) -> out + out_off[i]; out must hold B * slot_stride bytes with
 *          slot_stride >= frad_deflate_bound(longest frame); out_off[B] = total.
 * inflate: frame i = pay[pay_off[i] .. +pay_len[i]) -> out + out_off[i] (same scratch rule; slot_stride = the longest legal
 *          payload, frad_p1_raw_bound); a frame that does not inflate contributes zero bytes (it decodes to silence).
 * write:   headers + payloads of B frames sharing one parameter set -> out; returns bytes written or -(bytes needed).
 * scan:    the complete frames at buf[pos..] that share one parameter set; stops before anything else (see hostframe.cpp). */
size_t frad_deflate_bound(size_t raw_bytes);
int frad_deflate_frames(const uint8_t* raw, const uint64_t* raw_off, int B, uint8_t* out, size_t slot_stride,
                        uint64_t* out_off, int threads);
int frad_inflate_frames(const uint8_t* pay, const uint64_t* pay_off, const uint32_t* pay_len, int B, uint8_t* out,
                        size_t slot_stride, uint64_t* out_off, int threads);
int64_t frad_asfh_write_frames(int B, int profile, int ecc, int little_endian, const uint8_t* depth_idx, int depth0,
                               int channels, uint32_t srate, const uint32_t* fsize, uint32_t fsize0, int overlap_ratio,
                               int ecc_dsize, int ecc_csize, const uint8_t* pay, const uint64_t* pay_off,
                               const uint32_t* pay_len, const uint32_t* crc, int flush_each, uint8_t* out, size_t out_cap);
int frad_asfh_scan(const uint8_t* buf, size_t len, size_t pos, int max_frames, int64_t* params, uint64_t* pay_pos,
                   uint32_t* pay_len, uint32_t* crc, size_t* next_pos);

#ifdef __cplusplus
}
#endif
#endif /* FRADB200_H */
