// Host-pointer entry points: copy in, run the *_dev path, copy out, synchronise.
// These are what a ctypes / cgo / JNI binding of the reference's per-frame seams would call.
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <sched.h>
#include <vector>
#include "common.cuh"

static uint32_t host_max_len(const uint32_t* len, uint32_t ulen, int B) {
  if (!len) return ulen;
  uint32_t m = 0;
  for (int i = 0; i < B; i++) if (len[i] > m) m = len[i];
  return m ? m : 1;
}

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

// ---- pipelined transform entry points ---------------------------------------------------------------
// The batch is cut into chunks of frames; H2D copies run on ctx->s_in, kernels on ctx->stream, D2H copies on
// ctx->s_out, chained by events, so PCIe traffic in both directions overlaps the kernels (pinned host memory
// makes the copies truly asynchronous; pageable memory still works, just without overlap).
namespace {
static int env_int(const char* name, int dflt, int lo) {
  const char* v = getenv(name);
  if (!v) return dflt;
  const int x = atoi(v);
  return x < lo ? lo : x;
}
static const int kChunkFrames = env_int("FRAD_CHUNK", 1024, 1);
// Host waits of the pacing below.  FRAD_WAIT=yield (default): poll the event and give the core away between polls -- as
// quick as a spinning wait when cores are free, and other ranks' threads get the core when they are not (one rank per GPU
// with two pipelines each is 16 waiting threads on a 16-core box).  spin: cudaEventSynchronize on a spinning event;
// block: cudaEventBlockingSync (measured on B200, one rank: 26 % slower end to end -- a wake-up per chunk and stream).
static int wait_mode() {
  const char* v = getenv("FRAD_WAIT");
  if (!v) return 0;
  return (v[0] == 's') ? 1 : (v[0] == 'b') ? 2 : 0;
}
static const int kWait = wait_mode();
static inline cudaError_t wait_event(cudaEvent_t e) {
  if (kWait != 0) return cudaEventSynchronize(e);
  for (;;) {
    const cudaError_t r = cudaEventQuery(e);
    if (r != cudaErrorNotReady) return r;
    sched_yield();
  }
}

// FRAD_TRACE=1: every pipelined call prints one line per chunk to stderr -- host submit times and device completion times
// of its input copy, kernels and output copy, in ms since the call started (CLOCK_MONOTONIC of the start included, so
// the lines of two contexts can be laid side by side).  Diagnostics only: the events then carry time stamps.
static const int kTrace = env_int("FRAD_TRACE", 0, 0);
static inline double mono_ms() {
  timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  return 1e3 * (double)ts.tv_sec + 1e-6 * (double)ts.tv_nsec;
}
struct pipeline {
  frad_ctx* ctx; int nchunks; int chunk;
  cudaEvent_t* in_done; cudaEvent_t* k_done; cudaEvent_t* aux; cudaEvent_t* out_done;     // one event of each kind per chunk
  cudaEvent_t fin, fence;
  double t0; std::vector<double> h_in, h_run, h_out;                                       // trace: host submit times
};

int pipeline_begin(frad_ctx* ctx, int B, pipeline* p) {
  p->ctx = ctx;
  p->chunk = kChunkFrames;
  p->nchunks = (B + p->chunk - 1) / p->chunk;
  while ((int)ctx->pipe_events.size() < 4 * p->nchunks + 2) {
    cudaEvent_t e;
    FRAD_CUDA(ctx, cudaEventCreateWithFlags(&e, (kTrace ? 0 : cudaEventDisableTiming) | (kWait == 2 ? cudaEventBlockingSync : 0)));
    ctx->pipe_events.push_back(e);
  }
  p->in_done = ctx->pipe_events.data();
  p->k_done = ctx->pipe_events.data() + p->nchunks;
  p->aux = ctx->pipe_events.data() + 2 * p->nchunks + 2;
  p->out_done = ctx->pipe_events.data() + 3 * p->nchunks + 2;
  p->fin = ctx->pipe_events[2 * p->nchunks + 1];
  // the copy streams must not run ahead of work already queued on the main stream (buffer reuse)
  cudaEvent_t fence = ctx->pipe_events[2 * p->nchunks];
  p->fence = fence;
  if (kTrace) {
    p->t0 = mono_ms();
    p->h_in.assign(p->nchunks, 0.0); p->h_run.assign(p->nchunks, 0.0); p->h_out.assign(p->nchunks, 0.0);
  }
  FRAD_CUDA(ctx, cudaEventRecord(fence, ctx->stream));
  FRAD_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, fence, 0));
  FRAD_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, fence, 0));
  return FRAD_OK;
}

// Nothing may still be in flight towards the caller's buffers or the shared arena when a call returns -- also when it
// returns an error from the middle of the pipeline.
static void pipeline_drain(frad_ctx* ctx) {
  cudaStreamSynchronize(ctx->s_in);
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->s_out);
}

// Input chunk i covers the samples [i chunk hop, (i + 1) chunk hop) (the last one runs to the end of the input).  The frames
// of chunk i read up to sample (f1 - 1) hop + N: returns the last input chunk that must have arrived before they run.
static int last_input_chunk(int i, int chunk, int nchunks, int B, int64_t hop, int N) {
  const int f1 = ((i + 1) * chunk < B) ? (i + 1) * chunk : B;
  if (hop <= 0) return nchunks - 1;                       // every frame reads the same samples: they travel with the last chunk
  const int64_t end = (int64_t)(f1 - 1) * hop + N;        // one past the last sample read
  int64_t j = (end - 1) / ((int64_t)chunk * hop);
  if (j > nchunks - 1) j = nchunks - 1;
  return (int)(j < i ? i : j);
}

// h2d(i, f0, f1, stream), run(f0, f1) [enqueues on ctx->stream], d2h(f0, f1, stream)
// need(i): index of the last input chunk the kernels of chunk i read (i itself when frames do not overlap)
// fin(stream): optional last step on the output stream before the final synchronisation
struct no_fin { int operator()(cudaStream_t) const { return FRAD_OK; } };
struct need_self { int operator()(int i) const { return i; } };
template <class NEED, class H2D, class RUN, class D2H, class FIN = no_fin>
int pipeline_run_inner(pipeline& p, int B, NEED need, H2D h2d, RUN run, D2H d2h, FIN fin) {
  frad_ctx* ctx = p.ctx;
  int rc;
  // The host paces the copy queues: a chunk's input copy is submitted when the previous one has completed, its output
  // copy when its kernels have finished and the previous output copy has left, so no stream ever holds more than one
  // ready copy.  Measured on B200 with a decoder and an encoder running side by side on two contexts (one D2H-heavy, the
  // other H2D-heavy): with the whole batch queued up front, or with three chunks of look-ahead, the second context's
  // copies are only served after the first call has finished (9.9 ms for the pair = the sum of the two calls); paced
  // like this the two calls interleave chunk by chunk (8.3 ms).  A call running alone pays about 3 % for the pacing.
  static const int kDepth = env_int("FRAD_COPY_DEPTH", 1, 1);   // ready copies per stream
  int submitted_in = 0;
  auto queue_in_upto = [&](int last) -> int {
    for (; submitted_in <= last && submitted_in < p.nchunks; submitted_in++) {
      const int i = submitted_in;
      if (i >= kDepth) FRAD_CUDA(ctx, wait_event(p.in_done[i - kDepth]));
      const int f0 = i * p.chunk, f1 = (f0 + p.chunk < B) ? f0 + p.chunk : B;
      if (kTrace) p.h_in[i] = mono_ms() - p.t0;
      int r = h2d(i, f0, f1, ctx->s_in);
      if (r) return r;
      FRAD_CUDA(ctx, cudaEventRecord(p.in_done[i], ctx->s_in));
    }
    return FRAD_OK;
  };
  auto queue_out = [&](int i) -> int {                     // chunk i's kernels are done: its results can leave
    const int f0 = i * p.chunk, f1 = (f0 + p.chunk < B) ? f0 + p.chunk : B;
    FRAD_CUDA(ctx, wait_event(p.k_done[i]));
    if (i >= kDepth) FRAD_CUDA(ctx, wait_event(p.out_done[i - kDepth]));
    if (kTrace) p.h_out[i] = mono_ms() - p.t0;
    int r = d2h(f0, f1, ctx->s_out);
    if (r) return r;
    FRAD_CUDA(ctx, cudaEventRecord(p.out_done[i], ctx->s_out));
    return FRAD_OK;
  };
  for (int i = 0; i < p.nchunks; i++) {
    const int f0 = i * p.chunk, f1 = (f0 + p.chunk < B) ? f0 + p.chunk : B;
    const int last = need(i);
    if ((rc = queue_in_upto(last))) return rc;
    FRAD_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, p.in_done[last], 0));    // s_in is in order: covers every earlier chunk
    if (kTrace) p.h_run[i] = mono_ms() - p.t0;
    if ((rc = run(f0, f1))) return rc;
    FRAD_CUDA(ctx, cudaEventRecord(p.k_done[i], ctx->stream));
    if (i + 1 < p.nchunks && (rc = queue_in_upto(need(i + 1)))) return rc;   // next input while these kernels run
    if (i > 0 && (rc = queue_out(i - 1))) return rc;
  }
  if ((rc = queue_out(p.nchunks - 1))) return rc;
  if ((rc = fin(ctx->s_out))) return rc;
  FRAD_CUDA(ctx, cudaEventRecord(p.fin, ctx->s_out));      // the kernels of every chunk precede its output copy
  FRAD_CUDA(ctx, wait_event(p.fin));
  if (kTrace) {
    const double t1 = mono_ms() - p.t0;
    for (int i = 0; i < p.nchunks; i++) {
      float a = 0, k = 0, o = 0;
      cudaEventElapsedTime(&a, p.fence, p.in_done[i]);
      cudaEventElapsedTime(&k, p.fence, p.k_done[i]);
      cudaEventElapsedTime(&o, p.fence, p.out_done[i]);
      fprintf(stderr, "{\"trace\": \"%p\", \"t0\": %.3f, \"chunk\": %d, \"h_in\": %.3f, \"in_done\": %.3f, \"h_run\": %.3f, \"k_done\": %.3f, "
                      "\"h_out\": %.3f, \"out_done\": %.3f, \"end\": %.3f}\n", (void*)ctx, p.t0, i, p.h_in[i], a, p.h_run[i], k, p.h_out[i], o, t1);
    }
  }
  return FRAD_OK;
}
template <class NEED, class H2D, class RUN, class D2H, class FIN = no_fin>
int pipeline_run(pipeline& p, int B, NEED need, H2D h2d, RUN run, D2H d2h, FIN fin = FIN()) {
  const int rc = pipeline_run_inner(p, B, need, h2d, run, d2h, fin);
  if (rc) pipeline_drain(p.ctx);
  return rc;
}
// frames of N samples every hop samples
struct need_hop {
  const pipeline* p; int B; int64_t hop; int N;
  int operator()(int i) const { return last_input_chunk(i, p->chunk, p->nchunks, B, hop, N); }
};

inline int copy_async(frad_ctx* ctx, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s) {
  if (bytes) FRAD_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, kind, s));
  return FRAD_OK;
}
}  // namespace

#define H2D_ cudaMemcpyHostToDevice
#define D2H_ cudaMemcpyDeviceToHost

static int lossless_encode(frad_ctx* ctx, int profile, const void* pcm, int pcm_dtype, int64_t hop, int B, int N, int C,
                           int bits, int little_endian, int prec,
                           uint8_t* payload, size_t payload_stride, double* peak, uint32_t* crc32) {
  if (!ctx) return FRAD_E_ARG;
  const int es = frad_dtype_size(pcm_dtype);
  if (!es || B < 0 || N < 1 || C < 1 || hop < 0) return frad_fail(ctx, FRAD_E_ARG, "frad_p0_encode: bad argument");
  if (B == 0) return FRAD_OK;
  FRAD_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t samp = (size_t)C * es;                                  // bytes per multichannel sample
  const size_t in_bytes = ((size_t)(B - 1) * (size_t)hop + (size_t)N) * samp;
  const size_t out_bytes = (size_t)B * payload_stride;
  ARENA(ctx, a, rounded(in_bytes) + rounded(out_bytes) + rounded(8 * (size_t)B) + rounded(4 * (size_t)B));
  char* d_in = (char*)take(a, in_bytes);
  uint8_t* d_out = (uint8_t*)take(a, out_bytes);
  double* d_peak = (double*)take(a, 8 * (size_t)B);
  uint32_t* d_crc = (uint32_t*)take(a, 4 * (size_t)B);
  if (a.bad) return frad_fail(ctx, FRAD_E_NOMEM, "internal arena overflow");
  const size_t used = frad_p0_payload_bytes(N, C, bits);
  pipeline p;
  int rc = pipeline_begin(ctx, B, &p);
  if (rc) return rc;
  return pipeline_run(p, B, need_hop{&p, B, hop, N},
    [&](int, int f0, int f1, cudaStream_t s) {
      const size_t lo = (size_t)f0 * hop * samp;
      const size_t hi = (f1 == B) ? in_bytes : (size_t)f1 * hop * samp;
      return copy_async(ctx, d_in + lo, (const char*)pcm + lo, hi - lo, H2D_, s);
    },
    [&](int f0, int f1) {
      int r = (profile == 4)
                  ? frad_p4_encode_dev(ctx, d_in + (size_t)f0 * hop * samp, pcm_dtype, hop, f1 - f0, N, C, bits, little_endian,
                                       d_out + (size_t)f0 * payload_stride, payload_stride, d_peak + f0)
                  : frad_p0_encode_dev(ctx, d_in + (size_t)f0 * hop * samp, pcm_dtype, hop, f1 - f0, N, C, bits, little_endian, prec,
                                       d_out + (size_t)f0 * payload_stride, payload_stride, d_peak + f0);
      if (!r && crc32) r = frad_crc32_dev(ctx, d_out + (size_t)f0 * payload_stride, nullptr, nullptr, payload_stride,
                                          (uint32_t)used, f1 - f0, d_crc + f0);
      return r;
    },
    [&](int f0, int f1, cudaStream_t s) {
      int r = copy_async(ctx, payload + (size_t)f0 * payload_stride, d_out + (size_t)f0 * payload_stride,
                         (size_t)(f1 - f0) * payload_stride, D2H_, s);
      if (!r && peak) r = copy_async(ctx, peak + f0, d_peak + f0, 8 * (size_t)(f1 - f0), D2H_, s);
      if (!r && crc32) r = copy_async(ctx, crc32 + f0, d_crc + f0, 4 * (size_t)(f1 - f0), D2H_, s);
      return r;
    });
}

extern "C" int frad_p0_encode(frad_ctx* ctx, const void* pcm, int pcm_dtype, int64_t hop, int B, int N, int C,
                              int bits, int little_endian, int prec,
                              uint8_t* payload, size_t payload_stride, double* peak, uint32_t* crc32) {
  return lossless_encode(ctx, 0, pcm, pcm_dtype, hop, B, N, C, bits, little_endian, prec, payload, payload_stride, peak, crc32);
}
extern "C" int frad_p4_encode(frad_ctx* ctx, const void* pcm, int pcm_dtype, int64_t hop, int B, int N, int C,
                              int bits, int little_endian, uint8_t* payload, size_t payload_stride, double* peak, uint32_t* crc32) {
  return lossless_encode(ctx, 4, pcm, pcm_dtype, hop, B, N, C, bits, little_endian, FRAD_PREC_REF, payload, payload_stride, peak, crc32);
}

static int lossless_decode(frad_ctx* ctx, int profile, const uint8_t* payload, size_t payload_stride, int B, int N, int C,
                           int bits, int little_endian, int prec, void* pcm, int pcm_dtype) {
  if (!ctx) return FRAD_E_ARG;
  const int es = frad_dtype_size(pcm_dtype);
  if (!es || B < 0 || N < 1 || C < 1) return frad_fail(ctx, FRAD_E_ARG, "frad_p0_decode: bad argument");
  if (B == 0) return FRAD_OK;
  FRAD_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t in_bytes = (size_t)B * payload_stride;
  const size_t fbytes = (size_t)N * C * es;
  ARENA(ctx, a, rounded(in_bytes) + rounded((size_t)B * fbytes));
  uint8_t* d_in = (uint8_t*)take(a, in_bytes);
  char* d_out = (char*)take(a, (size_t)B * fbytes);
  if (a.bad) return frad_fail(ctx, FRAD_E_NOMEM, "internal arena overflow");
  pipeline p;
  int rc = pipeline_begin(ctx, B, &p);
  if (rc) return rc;
  return pipeline_run(p, B, need_self(),
    [&](int, int f0, int f1, cudaStream_t s) {
      return copy_async(ctx, d_in + (size_t)f0 * payload_stride, payload + (size_t)f0 * payload_stride,
                        (size_t)(f1 - f0) * payload_stride, H2D_, s);
    },
    [&](int f0, int f1) {
      return (profile == 4)
                 ? frad_p4_decode_dev(ctx, d_in + (size_t)f0 * payload_stride, payload_stride, f1 - f0, N, C, bits, little_endian,
                                      d_out + (size_t)f0 * fbytes, pcm_dtype)
                 : frad_p0_decode_dev(ctx, d_in + (size_t)f0 * payload_stride, payload_stride, f1 - f0, N, C, bits, little_endian,
                                      prec, d_out + (size_t)f0 * fbytes, pcm_dtype);
    },
    [&](int f0, int f1, cudaStream_t s) {
      return copy_async(ctx, (char*)pcm + (size_t)f0 * fbytes, d_out + (size_t)f0 * fbytes, (size_t)(f1 - f0) * fbytes, D2H_, s);
    });
}
extern "C" int frad_p0_decode(frad_ctx* ctx, const uint8_t* payload, size_t payload_stride, int B, int N, int C,
                              int bits, int little_endian, int prec, void* pcm, int pcm_dtype) {
  return lossless_decode(ctx, 0, payload, payload_stride, B, N, C, bits, little_endian, prec, pcm, pcm_dtype);
}
extern "C" int frad_p4_decode(frad_ctx* ctx, const uint8_t* payload, size_t payload_stride, int B, int N, int C,
                              int bits, int little_endian, void* pcm, int pcm_dtype) {
  return lossless_decode(ctx, 4, payload, payload_stride, B, N, C, bits, little_endian, FRAD_PREC_REF, pcm, pcm_dtype);
}

extern "C" int frad_p1_encode(frad_ctx* ctx, const void* pcm, int pcm_dtype, int64_t hop, int B, int N, int C,
                              int bits, int srate, double loss_level, int prec,
                              int32_t* q, int32_t* tq, int32_t* lpc) {
  if (!ctx) return FRAD_E_ARG;
  const int es = frad_dtype_size(pcm_dtype);
  if (!es || B < 0 || N < 1 || C < 1 || hop < 0) return frad_fail(ctx, FRAD_E_ARG, "frad_p1_encode: bad argument");
  if (B == 0) return FRAD_OK;
  FRAD_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t samp = (size_t)C * es;
  const size_t in_bytes = ((size_t)(B - 1) * (size_t)hop + (size_t)N) * samp;
  const size_t qf = (size_t)N * C, tf = (size_t)FRAD_NBANDS * C, lf = (size_t)FRAD_LPC_LEN * C;   // ints per frame
  ARENA(ctx, a, rounded(in_bytes) + rounded(4 * qf * B) + rounded(4 * tf * B) + rounded(4 * lf * B));
  char* d_in = (char*)take(a, in_bytes);
  int32_t* d_q = (int32_t*)take(a, 4 * qf * B);
  int32_t* d_tq = (int32_t*)take(a, 4 * tf * B);
  int32_t* d_l = lpc ? (int32_t*)take(a, 4 * lf * B) : nullptr;
  if (a.bad) return frad_fail(ctx, FRAD_E_NOMEM, "internal arena overflow");
  pipeline p;
  int rc = pipeline_begin(ctx, B, &p);
  if (rc) return rc;
  return pipeline_run(p, B, need_hop{&p, B, hop, N},
    [&](int, int f0, int f1, cudaStream_t s) {
      const size_t lo = (size_t)f0 * hop * samp;
      const size_t hi = (f1 == B) ? in_bytes : (size_t)f1 * hop * samp;
      return copy_async(ctx, d_in + lo, (const char*)pcm + lo, hi - lo, H2D_, s);
    },
    [&](int f0, int f1) {
      return frad_p1_encode_dev(ctx, d_in + (size_t)f0 * hop * samp, pcm_dtype, hop, f1 - f0, N, C, bits, srate, loss_level,
                                prec, d_q + f0 * qf, d_tq + f0 * tf, d_l ? d_l + f0 * lf : nullptr);
    },
    [&](int f0, int f1, cudaStream_t s) {
      const size_t nfr = (size_t)(f1 - f0);
      int r = copy_async(ctx, q + f0 * qf, d_q + f0 * qf, 4 * qf * nfr, D2H_, s);
      if (!r) r = copy_async(ctx, tq + f0 * tf, d_tq + f0 * tf, 4 * tf * nfr, D2H_, s);
      if (!r && lpc) r = copy_async(ctx, lpc + f0 * lf, d_l + f0 * lf, 4 * lf * nfr, D2H_, s);
      return r;
    });
}

extern "C" int frad_p1_decode(frad_ctx* ctx, const int32_t* q, const int32_t* tq, const int32_t* lpc,
                              int B, int N, int C, int bits, int srate, int prec, void* pcm, int pcm_dtype) {
  if (!ctx) return FRAD_E_ARG;
  const int es = frad_dtype_size(pcm_dtype);
  if (!es || B < 0 || N < 1 || C < 1) return frad_fail(ctx, FRAD_E_ARG, "frad_p1_decode: bad argument");
  if (B == 0) return FRAD_OK;
  FRAD_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t qf = (size_t)N * C, tf = (size_t)FRAD_NBANDS * C, lf = (size_t)FRAD_LPC_LEN * C;
  const size_t fbytes = qf * es;
  ARENA(ctx, a, rounded(fbytes * B) + rounded(4 * qf * B) + rounded(4 * tf * B) + rounded(4 * lf * B));
  int32_t* d_q = (int32_t*)take(a, 4 * qf * B);
  int32_t* d_tq = (int32_t*)take(a, 4 * tf * B);
  int32_t* d_l = lpc ? (int32_t*)take(a, 4 * lf * B) : nullptr;
  char* d_out = (char*)take(a, fbytes * B);
  if (a.bad) return frad_fail(ctx, FRAD_E_NOMEM, "internal arena overflow");
  pipeline p;
  int rc = pipeline_begin(ctx, B, &p);
  if (rc) return rc;
  return pipeline_run(p, B, need_self(),
    [&](int, int f0, int f1, cudaStream_t s) {
      const size_t nfr = (size_t)(f1 - f0);
      int r = copy_async(ctx, d_q + f0 * qf, q + f0 * qf, 4 * qf * nfr, H2D_, s);
      if (!r) r = copy_async(ctx, d_tq + f0 * tf, tq + f0 * tf, 4 * tf * nfr, H2D_, s);
      if (!r && lpc) r = copy_async(ctx, d_l + f0 * lf, lpc + f0 * lf, 4 * lf * nfr, H2D_, s);
      return r;
    },
    [&](int f0, int f1) {
      return frad_p1_decode_dev(ctx, d_q + f0 * qf, d_tq + f0 * tf, d_l ? d_l + f0 * lf : nullptr, f1 - f0, N, C, bits, srate,
                                prec, d_out + (size_t)f0 * fbytes, pcm_dtype);
    },
    [&](int f0, int f1, cudaStream_t s) {
      return copy_async(ctx, (char*)pcm + (size_t)f0 * fbytes, d_out + (size_t)f0 * fbytes, (size_t)(f1 - f0) * fbytes, D2H_, s);
    });
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
  const uint32_t mx = host_max_len(len, ulen, B);        // lets the kernel cut long messages into balanced segments
  rc = IS32 ? frad_crc32_dev(ctx, d_buf, d_off, d_len, stride, mx, B, d_out)
            : frad_crc16_dev(ctx, d_buf, d_off, d_len, stride, mx, B, d_out);
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
  rc = frad_rs_encode_dev(ctx, d_buf, d_off, d_len, stride, host_max_len(len, ulen, B), B, dsize, csize, d_out, d_ooff, out_stride);
  if (rc) return rc;
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
  rc = frad_rs_decode_dev(ctx, d_buf, d_off, d_len, stride, host_max_len(len, ulen, B), B, dsize, csize, repair, d_out, d_ooff,
                          out_stride, d_status, d_fixed);
  if (rc) return rc;
  if ((rc = frad_d2h(ctx, out, d_out, out_total_bytes))) return rc;
  if (status && (rc = frad_d2h(ctx, status, d_status, 4 * (size_t)B))) return rc;
  if (fixed && (rc = frad_d2h(ctx, fixed, d_fixed, 4 * (size_t)B))) return rc;
  return frad_sync(ctx);
}

// ---- profile 1 / 2 with the bit packing on the device: PCM <-> frame payloads before deflate -------------------
// Same pipeline as above, but what crosses PCIe on the coded side is the Exp-Golomb-Rice byte string of every frame
// (profile1.py:43-46 / profile2.py:48-51) instead of the int32 streams: about five times fewer bytes.
extern "C" int frad_overlap_dev(frad_ctx* ctx, const double* frames, int B, int N, int C, int ratio, const double* fade,
                                const double* prev_tail, void* out, int out_dtype, double* tail_out);
int frad_compact_append(frad_ctx* ctx, const uint8_t* src, size_t stride, const uint32_t* len, int B, uint8_t* dst, uint64_t dst_cap,
                        uint64_t* dst_off, uint64_t* running);   // golomb.cu

extern "C" int frad_p1_encode_raw(frad_ctx* ctx, const void* pcm, int pcm_dtype, int64_t hop, int B, int N, int C,
                                  int bits, int srate, double loss_level, int prec, int profile,
                                  uint8_t* raw, size_t raw_cap, uint64_t* raw_off) {
  if (!ctx) return FRAD_E_ARG;
  const int es = frad_dtype_size(pcm_dtype);
  if (!es || B < 0 || N < 1 || C < 1 || hop < 0 || !raw_off || (profile != 1 && profile != 2) || (!raw && raw_cap))
    return frad_fail(ctx, FRAD_E_ARG, "frad_p1_encode_raw: bad argument");
  raw_off[0] = 0;
  if (B == 0) return FRAD_OK;
  FRAD_CUDA(ctx, cudaSetDevice(ctx->device));
  const bool p2 = profile == 2;
  const size_t samp = (size_t)C * es;
  const size_t in_bytes = ((size_t)(B - 1) * (size_t)hop + (size_t)N) * samp;
  const size_t qf = (size_t)N * C, tf = (size_t)FRAD_NBANDS * C, lf = (size_t)FRAD_LPC_LEN * C;
  const size_t stride = frad_p1_raw_bound(N, C, p2);
  const size_t cf = (size_t)(B < kChunkFrames ? B : kChunkFrames);     // frames per chunk: the int / strided scratch is per chunk
  ARENA(ctx, a, rounded(in_bytes) + rounded(4 * qf * cf) + rounded(4 * tf * cf) + rounded(4 * lf * cf) + rounded(stride * cf) +
                    rounded(4 * cf) + rounded(raw_cap) + rounded(8 * ((size_t)B + 1)) + rounded(8));
  char* d_in = (char*)take(a, in_bytes);
  int32_t* d_q = (int32_t*)take(a, 4 * qf * cf);
  int32_t* d_tq = (int32_t*)take(a, 4 * tf * cf);
  int32_t* d_l = p2 ? (int32_t*)take(a, 4 * lf * cf) : nullptr;
  uint8_t* d_strided = (uint8_t*)take(a, stride * cf);
  uint32_t* d_len = (uint32_t*)take(a, 4 * cf);
  uint8_t* d_raw = (uint8_t*)take(a, raw_cap ? raw_cap : 1);
  uint64_t* d_off = (uint64_t*)take(a, 8 * ((size_t)B + 1));
  uint64_t* d_total = (uint64_t*)take(a, 8);
  if (a.bad) return frad_fail(ctx, FRAD_E_NOMEM, "internal arena overflow");
  FRAD_CUDA(ctx, cudaMemsetAsync(d_total, 0, 8, ctx->stream));
  pipeline p;
  int rc = pipeline_begin(ctx, B, &p);
  if (rc) return rc;
  bool overflow = false;
  auto fetch_bytes = [&](int i, cudaStream_t s) -> int {       // payload bytes of chunk i, once its offsets are on the host
    const int f0 = i * p.chunk, f1 = (f0 + p.chunk < B) ? f0 + p.chunk : B;
    FRAD_CUDA(ctx, wait_event(p.aux[i]));
    if (overflow || raw_off[f1] > raw_cap) { overflow = true; return (int)FRAD_OK; }
    return copy_async(ctx, raw + raw_off[f0], d_raw + raw_off[f0], (size_t)(raw_off[f1] - raw_off[f0]), D2H_, s);
  };
  rc = pipeline_run(p, B, need_hop{&p, B, hop, N},
    [&](int, int f0, int f1, cudaStream_t s) {
      const size_t lo = (size_t)f0 * hop * samp;
      const size_t hi = (f1 == B) ? in_bytes : (size_t)f1 * hop * samp;
      return copy_async(ctx, d_in + lo, (const char*)pcm + lo, hi - lo, H2D_, s);
    },
    [&](int f0, int f1) {
      const int nb = f1 - f0;
      int r = frad_p1_encode_dev(ctx, d_in + (size_t)f0 * hop * samp, pcm_dtype, hop, nb, N, C, bits, srate, loss_level, prec,
                                 d_q, d_tq, d_l);
      if (!r) r = frad_p1_pack_dev(ctx, d_q, d_tq, d_l, nb, N, C, d_strided, stride, d_len);
      if (!r) r = frad_compact_append(ctx, d_strided, stride, d_len, nb, d_raw, raw_cap, d_off + f0, d_total);
      return r;
    },
    [&](int f0, int f1, cudaStream_t s) -> int {
      // The byte range of a chunk is only known on the device: its offsets come back first, the bytes follow one chunk
      // later -- by then the offsets have normally arrived, so the host does not stall the queue of kernels and copies
      // waiting for one small transfer (which may sit behind another context's large copies on the D2H engine).
      const int i = f0 / p.chunk;
      int r = copy_async(ctx, raw_off + f0, d_off + f0, 8 * (size_t)(f1 - f0 + 1), D2H_, s);
      if (r) return r;
      FRAD_CUDA(ctx, cudaEventRecord(p.aux[i], s));
      if (i > 0) r = fetch_bytes(i - 1, s);
      return r;
    },
    [&](cudaStream_t s) -> int { return fetch_bytes(p.nchunks - 1, s); });
  if (rc) return rc;
  if (overflow) return frad_fail(ctx, FRAD_E_NOMEM, "frad_p1_encode_raw: raw_cap too small (raw_off[B] holds the size needed)");
  return FRAD_OK;
}

extern "C" int frad_p1_decode_raw(frad_ctx* ctx, const uint8_t* raw, const uint64_t* raw_off, int B, int N, int C,
                                  int bits, int srate, int prec, int profile, void* pcm, int pcm_dtype, int32_t* status) {
  if (!ctx) return FRAD_E_ARG;
  const int es = frad_dtype_size(pcm_dtype);
  if (!es || B < 0 || N < 1 || C < 1 || !raw_off || (profile != 1 && profile != 2))
    return frad_fail(ctx, FRAD_E_ARG, "frad_p1_decode_raw: bad argument");
  if (B == 0) return FRAD_OK;
  FRAD_CUDA(ctx, cudaSetDevice(ctx->device));
  const bool p2 = profile == 2;
  const size_t qf = (size_t)N * C, tf = (size_t)FRAD_NBANDS * C, lf = (size_t)FRAD_LPC_LEN * C;
  const size_t fbytes = qf * es;
  const size_t total = (size_t)raw_off[B];
  const size_t cf = (size_t)(B < kChunkFrames ? B : kChunkFrames);
  ARENA(ctx, a, rounded(total + 16) + rounded(8 * ((size_t)B + 1)) + rounded(4 * qf * cf) + rounded(4 * tf * cf) + rounded(4 * lf * cf) +
                    rounded(fbytes * B) + rounded(4 * (size_t)B));
  uint8_t* d_raw = (uint8_t*)take(a, total + 16);
  uint64_t* d_off = (uint64_t*)take(a, 8 * ((size_t)B + 1));
  int32_t* d_q = (int32_t*)take(a, 4 * qf * cf);
  int32_t* d_tq = (int32_t*)take(a, 4 * tf * cf);
  int32_t* d_l = p2 ? (int32_t*)take(a, 4 * lf * cf) : nullptr;
  char* d_out = (char*)take(a, fbytes * B);
  int32_t* d_status = (int32_t*)take(a, 4 * (size_t)B);
  if (a.bad) return frad_fail(ctx, FRAD_E_NOMEM, "internal arena overflow");
  pipeline p;
  int rc = pipeline_begin(ctx, B, &p);
  if (rc) return rc;
  rc = pipeline_run(p, B, need_self(),
    [&](int i, int f0, int f1, cudaStream_t s) {
      int r = FRAD_OK;
      if (i == 0) r = copy_async(ctx, d_off, raw_off, 8 * ((size_t)B + 1), H2D_, s);
      if (!r) r = copy_async(ctx, d_raw + raw_off[f0], raw + raw_off[f0], (size_t)(raw_off[f1] - raw_off[f0]), H2D_, s);
      return r;
    },
    [&](int f0, int f1) {
      const int nb = f1 - f0;
      int r = frad_p1_unpack_dev(ctx, d_raw, d_off + f0, nb, N, C, p2 ? 1 : 0, d_q, d_tq, d_l, d_status + f0);
      if (!r) r = frad_p1_decode_dev(ctx, d_q, d_tq, d_l, nb, N, C, bits, srate, prec, d_out + (size_t)f0 * fbytes, pcm_dtype);
      return r;
    },
    [&](int f0, int f1, cudaStream_t s) {
      int r = copy_async(ctx, (char*)pcm + (size_t)f0 * fbytes, d_out + (size_t)f0 * fbytes, (size_t)(f1 - f0) * fbytes, D2H_, s);
      if (!r && status) r = copy_async(ctx, status + f0, d_status + f0, 4 * (size_t)(f1 - f0), D2H_, s);
      return r;
    });
  return rc;
}

// Payloads -> PCM as the reference's Decoder emits it for a run of equal-sized lossy frames: inverse transform, cross-fade
// with the previous frame (decoder.py:28-46) and conversion to the output sample format (pcmformat.py:49-62), all on the
// device; per frame only cut = N*(ratio-1)/ratio samples come back.  fade: L = N - cut window values (host);
// prev_tail / tail_out: f64 [L][C] carried between calls (prev_tail NULL: first frame is not faded).
extern "C" int frad_p1_decode_stream(frad_ctx* ctx, const uint8_t* raw, const uint64_t* raw_off, int B, int N, int C,
                                     int bits, int srate, int prec, int profile, int ratio, const double* fade,
                                     const double* prev_tail, void* pcm, int pcm_dtype, double* tail_out, int32_t* status) {
  if (!ctx) return FRAD_E_ARG;
  const int es = frad_dtype_size(pcm_dtype);
  if (!es || B < 0 || N < 1 || C < 1 || !raw_off || (profile != 1 && profile != 2) || ratio < 2 || !fade)
    return frad_fail(ctx, FRAD_E_ARG, "frad_p1_decode_stream: bad argument");
  if (B == 0) return FRAD_OK;
  FRAD_CUDA(ctx, cudaSetDevice(ctx->device));
  const bool p2 = profile == 2;
  const int cut = (int)((long long)N * (ratio - 1) / ratio), L = N - cut;
  const size_t qf = (size_t)N * C, tf = (size_t)FRAD_NBANDS * C, lf = (size_t)FRAD_LPC_LEN * C;
  const size_t fbytes = qf * 8, obytes = (size_t)cut * C * es, tbytes = (size_t)L * C * 8;
  const size_t total = (size_t)raw_off[B];
  const size_t cf = (size_t)(B < kChunkFrames ? B : kChunkFrames);
  ARENA(ctx, a, rounded(total + 16) + rounded(8 * ((size_t)B + 1)) + rounded(4 * qf * cf) + rounded(4 * tf * cf) + rounded(4 * lf * cf) +
                    rounded(fbytes * B) + rounded(obytes * B) + rounded(4 * (size_t)B) + rounded(8 * (size_t)L) + 2 * rounded(tbytes));
  uint8_t* d_raw = (uint8_t*)take(a, total + 16);
  uint64_t* d_off = (uint64_t*)take(a, 8 * ((size_t)B + 1));
  int32_t* d_q = (int32_t*)take(a, 4 * qf * cf);
  int32_t* d_tq = (int32_t*)take(a, 4 * tf * cf);
  int32_t* d_l = p2 ? (int32_t*)take(a, 4 * lf * cf) : nullptr;
  double* d_frames = (double*)take(a, fbytes * B);
  char* d_out = (char*)take(a, obytes * B);
  int32_t* d_status = (int32_t*)take(a, 4 * (size_t)B);
  double* d_fade = (double*)take(a, 8 * (size_t)L);
  double* d_prev = (double*)take(a, tbytes);
  double* d_tail = (double*)take(a, tbytes);
  if (a.bad) return frad_fail(ctx, FRAD_E_NOMEM, "internal arena overflow");
  pipeline p;
  int rc = pipeline_begin(ctx, B, &p);
  if (rc) return rc;
  rc = pipeline_run(p, B, need_self(),
    [&](int i, int f0, int f1, cudaStream_t s) {
      int r = FRAD_OK;
      if (i == 0) {
        r = copy_async(ctx, d_off, raw_off, 8 * ((size_t)B + 1), H2D_, s);
        if (!r) r = copy_async(ctx, d_fade, fade, 8 * (size_t)L, H2D_, s);
        if (!r && prev_tail) r = copy_async(ctx, d_prev, prev_tail, tbytes, H2D_, s);
      }
      if (!r) r = copy_async(ctx, d_raw + raw_off[f0], raw + raw_off[f0], (size_t)(raw_off[f1] - raw_off[f0]), H2D_, s);
      return r;
    },
    [&](int f0, int f1) {
      const int nb = f1 - f0;
      int r = frad_p1_unpack_dev(ctx, d_raw, d_off + f0, nb, N, C, p2 ? 1 : 0, d_q, d_tq, d_l, d_status + f0);
      if (!r) r = frad_p1_decode_dev(ctx, d_q, d_tq, d_l, nb, N, C, bits, srate, prec, d_frames + (size_t)f0 * qf, FRAD_F64);
      // the chunk's first frame fades against the last frame of the previous chunk, which sits right before it in d_frames
      const double* tail_in = (f0 == 0) ? (prev_tail ? d_prev : nullptr) : d_frames + ((size_t)(f0 - 1) * N + cut) * C;
      if (!r) r = frad_overlap_dev(ctx, d_frames + (size_t)f0 * qf, nb, N, C, ratio, d_fade, tail_in, d_out + (size_t)f0 * obytes,
                                   pcm_dtype, (f1 == B) ? d_tail : nullptr);
      return r;
    },
    [&](int f0, int f1, cudaStream_t s) {
      int r = copy_async(ctx, (char*)pcm + (size_t)f0 * obytes, d_out + (size_t)f0 * obytes, (size_t)(f1 - f0) * obytes, D2H_, s);
      if (!r && status) r = copy_async(ctx, status + f0, d_status + f0, 4 * (size_t)(f1 - f0), D2H_, s);
      if (!r && tail_out && f1 == B) r = copy_async(ctx, tail_out, d_tail, tbytes, D2H_, s);
      return r;
    });
  return rc;
}
