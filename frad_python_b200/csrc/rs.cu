// Batched Reed-Solomon over GF(2^8) for tools/ecc.py:6-25 (reedsolo.RSCodec semantics: primitive
// polynomial 0x11D, alpha = 2, fcr = 0, systematic, shortened last chunk).
//
// Encoding is polynomial division by g(x), run as a byte-wide LFSR: one thread per codeword keeps the
// csize parity bytes packed in 32-bit registers; each data byte costs one shared-memory row fetch of
// the table row[f] = f * g (csize bytes), a funnel-shift of the registers and an XOR.  Decoding
// re-runs the same LFSR over the received data: remainder == received parity  <=>  all syndromes are
// zero, so clean codewords cost exactly one encode.  Only dirty codewords run Berlekamp-Massey /
// Chien / Forney (per thread, from the csize-byte remainder: S_j = rem(alpha^j)).
#include "common.cuh"


namespace {

constexpr int kMaxParityWords = 64;     // csize <= 254

struct rs_table {                        // device layout: header then rows
  int csize, nw;                         // parity bytes, 32-bit words per row
  uint8_t gf_exp[512];
  uint8_t gf_log[256];
  // followed by uint32_t rows[256][nw]            (one LFSR step: symbol f entering an empty register)
  // then     uint32_t slice[3][256][nw]        (the same symbol followed by 1, 2, 3 zero-input steps)
};

void host_gf(uint8_t* ex, uint8_t* lg) {
  int x = 1;
  for (int i = 0; i < 255; i++) { ex[i] = (uint8_t)x; lg[x] = (uint8_t)i; x <<= 1; if (x & 0x100) x ^= 0x11D; }
  for (int i = 255; i < 512; i++) ex[i] = ex[i - 255];
  lg[0] = 0;
}

const rs_table* get_table(frad_ctx* ctx, int csize) {
  auto it = ctx->rs_tables.find((uint32_t)csize);
  if (it != ctx->rs_tables.end()) return (const rs_table*)it->second;
  const int nw = (csize + 3) / 4;
  const size_t bytes = sizeof(rs_table) + (size_t)4 * 256 * nw * 4;
  std::vector<uint8_t> h(bytes, 0);
  rs_table* t = (rs_table*)h.data();
  t->csize = csize; t->nw = nw;
  host_gf(t->gf_exp, t->gf_log);
  auto mul = [&](int a, int b) -> int { return (a && b) ? t->gf_exp[t->gf_log[a] + t->gf_log[b]] : 0; };
  // g(x) = prod_{i<csize} (x - alpha^i), highest degree first, g[0] = 1
  std::vector<int> g(1, 1);
  for (int i = 0; i < csize; i++) {
    std::vector<int> n(g.size() + 1, 0);
    for (size_t j = 0; j < g.size(); j++) { n[j] ^= g[j]; n[j + 1] ^= mul(g[j], t->gf_exp[i]); }
    g = n;
  }
  uint8_t* rows = h.data() + sizeof(rs_table);
  for (int f = 0; f < 256; f++) for (int j = 0; j < csize; j++) rows[(size_t)f * nw * 4 + j] = (uint8_t)mul(f, g[j + 1]);
  // slicing tables: advance the register by further zero-input steps (par <- shift(par) ^ rows[par[0]])
  const size_t tsz = (size_t)256 * nw * 4;
  for (int k = 1; k <= 3; k++) {
    const uint8_t* prev = rows + (k - 1) * tsz;
    uint8_t* cur = rows + k * tsz;
    for (int f = 0; f < 256; f++) {
      const uint8_t* pr = prev + (size_t)f * nw * 4;
      uint8_t* cr = cur + (size_t)f * nw * 4;
      const int fb = csize > 0 ? pr[0] : 0;
      for (int j = 0; j < csize; j++) cr[j] = (uint8_t)((j + 1 < csize ? pr[j + 1] : 0) ^ rows[(size_t)fb * nw * 4 + j]);
    }
  }
  void* d = nullptr;
  if (cudaMalloc(&d, bytes) != cudaSuccess || frad_upload(d, h.data(), bytes) != cudaSuccess) {
    cudaGetLastError();
    frad_fail(ctx, FRAD_E_NOMEM, "rs table upload failed");
    return nullptr;
  }
  ctx->rs_tables[(uint32_t)csize] = d;
  return (const rs_table*)d;
}

struct rs_smem_view {
  const uint8_t* gf_exp; const uint8_t* gf_log; const uint32_t* rows; int csize, nw;
};

__device__ __forceinline__ int gmul(const rs_smem_view& s, int a, int b) {
  return (a && b) ? s.gf_exp[s.gf_log[a] + s.gf_log[b]] : 0;
}
__device__ __forceinline__ int gdiv(const rs_smem_view& s, int a, int b) {   // b != 0
  return a ? s.gf_exp[s.gf_log[a] + 255 - s.gf_log[b]] : 0;
}

// LFSR remainder of `n` data bytes; par[] little-endian packed, byte j <-> coefficient x^(csize-1-j)
template <int NW>
__device__ __forceinline__ void lfsr(const rs_smem_view& s, const uint8_t* __restrict__ d, int n, uint32_t* par) {
#pragma unroll
  for (int w = 0; w < NW; w++) par[w] = 0;
  for (int i = 0; i < n; i++) {
    const uint32_t f = (par[0] ^ d[i]) & 0xFF;
    const uint32_t* row = s.rows + f * NW;
#pragma unroll
    for (int w = 0; w < NW - 1; w++) par[w] = __funnelshift_r(par[w], par[w + 1], 8) ^ row[w];
    par[NW - 1] = (par[NW - 1] >> 8) ^ row[NW - 1];
  }
}
__device__ void lfsr_any(const rs_smem_view& s, const uint8_t* __restrict__ d, int n, uint32_t* par) {
  const int NW = s.nw;
  for (int w = 0; w < NW; w++) par[w] = 0;
  for (int i = 0; i < n; i++) {
    const uint32_t f = (par[0] ^ d[i]) & 0xFF;
    const uint32_t* row = s.rows + f * NW;
    for (int w = 0; w < NW - 1; w++) par[w] = ((par[w] >> 8) | (par[w + 1] << 24)) ^ row[w];
    par[NW - 1] = (par[NW - 1] >> 8) ^ row[NW - 1];
  }
}
__device__ __forceinline__ void run_lfsr(const rs_smem_view& s, const uint8_t* d, int n, uint32_t* par) {
  switch (s.nw) {
    case 1: lfsr<1>(s, d, n, par); break;
    case 2: lfsr<2>(s, d, n, par); break;
    case 3: lfsr<3>(s, d, n, par); break;
    case 4: lfsr<4>(s, d, n, par); break;
    case 5: lfsr<5>(s, d, n, par); break;
    case 6: lfsr<6>(s, d, n, par); break;
    case 7: lfsr<7>(s, d, n, par); break;
    case 8: lfsr<8>(s, d, n, par); break;
    default: lfsr_any(s, d, n, par); break;
  }
}

__device__ __forceinline__ rs_smem_view load_tables(const rs_table* gt, unsigned char* smem) {
  const int nw = gt->nw;
  const size_t bytes = sizeof(rs_table) + (size_t)256 * nw * 4;
  const uint32_t* src = (const uint32_t*)gt;
  uint32_t* dst = (uint32_t*)smem;
  for (size_t i = threadIdx.x; i < bytes / 4; i += blockDim.x) dst[i] = src[i];
  __syncthreads();
  const rs_table* st = (const rs_table*)smem;
  rs_smem_view v;
  v.gf_exp = st->gf_exp; v.gf_log = st->gf_log; v.rows = (const uint32_t*)(smem + sizeof(rs_table));
  v.csize = st->csize; v.nw = nw;
  return v;
}

__global__ void __launch_bounds__(128) rs_encode_kernel(const uint8_t* __restrict__ buf, const uint64_t* __restrict__ off,
                                                        const uint32_t* __restrict__ len, size_t stride, uint32_t ulen, int B,
                                                        int dsize, const rs_table* __restrict__ gt,
                                                        uint8_t* __restrict__ out, const uint64_t* __restrict__ out_off, size_t out_stride) {
  extern __shared__ __align__(16) unsigned char smem[];
  const rs_smem_view s = load_tables(gt, smem);
  const int csize = s.csize, block = dsize + csize;
  for (int m = blockIdx.x; m < B; m += gridDim.x) {
    const uint8_t* src = buf + (off ? off[m] : (uint64_t)m * stride);
    uint8_t* dst = out + (out_off ? out_off[m] : (uint64_t)m * out_stride);
    const uint32_t L = len ? len[m] : ulen;
    const int ncw = (int)((L + dsize - 1) / dsize);
    for (int k = threadIdx.x; k < ncw; k += blockDim.x) {
      const uint8_t* d = src + (size_t)k * dsize;
      const int n = (int)(((size_t)(k + 1) * dsize <= L) ? dsize : (L - (size_t)k * dsize));
      uint32_t par[kMaxParityWords];
      run_lfsr(s, d, n, par);
      uint8_t* o = dst + (size_t)k * block;
      for (int i = 0; i < n; i++) o[i] = d[i];
      for (int j = 0; j < csize; j++) o[n + j] = (uint8_t)(par[j >> 2] >> (8 * (j & 3)));
    }
  }
}

// full decode of one dirty codeword.  word: n received bytes (data || parity) in global memory,
// rem: csize remainder bytes (computed parity xor received parity), out: n - csize corrected data bytes.
// Returns number of corrected bytes, or -1 when undecodable.
__device__ int rs_repair_word(const rs_smem_view& s, const uint8_t* word, int n, const uint8_t* rem, uint8_t* out) {
  const int nsym = s.csize, dlen = n - nsym;
  uint8_t S[256], lam[256], prev[256], tmp[256], omega[256];
  // S_j = rem(alpha^j): rem byte i carries x^(nsym-1-i)
  for (int j = 0; j < nsym; j++) {
    int acc = 0;
    const int root = s.gf_exp[j % 255];
    for (int i = 0; i < nsym; i++) acc = gmul(s, acc, root) ^ rem[i];
    S[j] = (uint8_t)acc;
  }
  // Berlekamp-Massey, polynomials lowest degree first
  for (int i = 0; i <= nsym; i++) { lam[i] = 0; prev[i] = 0; }
  lam[0] = 1; prev[0] = 1;
  int L = 0, m = 1, b = 1, lam_len = 1, prev_len = 1;
  for (int it = 0; it < nsym; it++) {
    int d = S[it];
    for (int i = 1; i <= L && i < lam_len; i++) d ^= gmul(s, lam[i], S[it - i]);
    if (d == 0) { m++; continue; }
    const int coef = gdiv(s, d, b);
    int new_len = lam_len;
    if (prev_len + m > new_len) new_len = prev_len + m;
    if (new_len > nsym + 1) new_len = nsym + 1;          // higher terms can never survive the 2L <= nsym test
    if (2 * L <= it) {
      for (int i = 0; i < lam_len; i++) tmp[i] = lam[i];
      const int old_len = lam_len;
      for (int i = lam_len; i < new_len; i++) lam[i] = 0;
      for (int i = 0; i < prev_len && i + m < new_len; i++) lam[i + m] ^= (uint8_t)gmul(s, coef, prev[i]);
      lam_len = new_len;
      for (int i = 0; i < old_len; i++) prev[i] = tmp[i];
      prev_len = old_len;
      L = it + 1 - L; b = d; m = 1;
    } else {
      for (int i = lam_len; i < new_len; i++) lam[i] = 0;
      for (int i = 0; i < prev_len && i + m < new_len; i++) lam[i + m] ^= (uint8_t)gmul(s, coef, prev[i]);
      lam_len = new_len;
      m++;
    }
  }
  int deg = lam_len - 1;
  while (deg > 0 && lam[deg] == 0) deg--;
  if (2 * L > nsym || deg != L) return -1;
  // Omega = S(x) * Lambda(x) mod x^nsym
  for (int i = 0; i < nsym; i++) {
    int acc = 0;
    const int jm = (i < deg) ? i : deg;
    for (int j = 0; j <= jm; j++) acc ^= gmul(s, lam[j], S[i - j]);
    omega[i] = (uint8_t)acc;
  }
  // Chien over the positions that exist + Forney
  for (int i = 0; i < dlen; i++) out[i] = word[i];
  int found = 0;
  for (int i = 0; i < n; i++) {
    const int p = n - 1 - i;
    const int xinv = s.gf_exp[(255 - (p % 255)) % 255];
    int ev = 0;
    for (int j = deg; j >= 0; j--) ev = gmul(s, ev, xinv) ^ lam[j];
    if (ev != 0) continue;
    found++;
    int den = 0;     // Lambda'(xinv): odd-degree terms only
    for (int j = deg; j >= 1; j--) { den = gmul(s, den, xinv); if (j & 1) den ^= lam[j]; }
    // Horner above evaluates sum_{j odd} lam[j] xinv^(j-1)
    if (den == 0) return -1;
    int num = 0;
    for (int j = nsym - 1; j >= 0; j--) num = gmul(s, num, xinv) ^ omega[j];
    num = gmul(s, num, s.gf_exp[p % 255]);
    const int e = gdiv(s, num, den);
    if (i < dlen) out[i] ^= (uint8_t)e;
  }
  // The locator generates all nsym syndromes with length L <= nsym/2; when it also splits into L
  // distinct roots at existing positions the Forney pattern reproduces every syndrome, so the
  // corrected word is a codeword and reedsolo's syndrome re-check cannot fail.
  if (found != deg) return -1;
  return deg;
}

__global__ void __launch_bounds__(128) rs_decode_kernel(const uint8_t* __restrict__ buf, const uint64_t* __restrict__ off,
                                                        const uint32_t* __restrict__ len, size_t stride, uint32_t ulen, int B,
                                                        int dsize, int repair, const rs_table* __restrict__ gt,
                                                        uint8_t* __restrict__ out, const uint64_t* __restrict__ out_off, size_t out_stride,
                                                        int32_t* __restrict__ status, int32_t* __restrict__ fixed) {
  extern __shared__ __align__(16) unsigned char smem[];
  const rs_smem_view s = load_tables(gt, smem);
  const int csize = s.csize, block = dsize + csize;
  for (int m = blockIdx.x; m < B; m += gridDim.x) {
    const uint8_t* src = buf + (off ? off[m] : (uint64_t)m * stride);
    uint8_t* dst = out + (out_off ? out_off[m] : (uint64_t)m * out_stride);
    const uint32_t L = len ? len[m] : ulen;
    const int ncw = (int)((L + block - 1) / block);
    int my_fail = 0, my_fixed = 0;
    for (int k = threadIdx.x; k < ncw; k += blockDim.x) {
      const uint8_t* w = src + (size_t)k * block;
      const int n = (int)(((size_t)(k + 1) * block <= L) ? block : (L - (size_t)k * block));
      const int dlen = (n > csize) ? n - csize : 0;
      uint8_t* o = dst + (size_t)k * dsize;
      if (!repair) { for (int i = 0; i < dlen; i++) o[i] = w[i]; continue; }
      if (n <= csize) { my_fail++; continue; }            // no data bytes: reedsolo raises, ecc.py:22 writes b''
      uint32_t par[kMaxParityWords];
      run_lfsr(s, w, dlen, par);
      uint8_t rem[256];
      int dirty = 0;
      for (int j = 0; j < csize; j++) {
        rem[j] = (uint8_t)(par[j >> 2] >> (8 * (j & 3))) ^ w[dlen + j];
        dirty |= rem[j];
      }
      if (!dirty) { for (int i = 0; i < dlen; i++) o[i] = w[i]; continue; }
      const int r = rs_repair_word(s, w, n, rem, o);
      if (r < 0) { for (int i = 0; i < dlen; i++) o[i] = 0; my_fail++; }
      else my_fixed += r;
    }
    if (status || fixed) {
      // block-wide totals for this message
      __shared__ int s_fail, s_fixed;
      if (threadIdx.x == 0) { s_fail = 0; s_fixed = 0; }
      __syncthreads();
      if (my_fail) atomicAdd(&s_fail, my_fail);
      if (my_fixed) atomicAdd(&s_fixed, my_fixed);
      __syncthreads();
      if (threadIdx.x == 0) { if (status) status[m] = s_fail; if (fixed) fixed[m] = s_fixed; }
      __syncthreads();
    }
  }
}


// ---- tile kernel: one warp = 32 consecutive codewords of one message -----------------------------------
// Coalesced staging of the tile into shared memory (odd row pitch: conflict-free per-lane rows), LFSR by
// SL symbols per step (slicing tables: one row fetch of csize bytes per data byte; with SL = 2 the register
// shift is a byte permute), coalesced write-back of data || parity (encode) or of the data (decode).
//
// The row fetches are data dependent, so a plain table costs a ~3-way bank conflict per fetch.  Here every
// table row is cut into planes of 16, 8 or 4 bytes and each plane is replicated 8, 16 or 32 times
// (128 bytes per table entry and plane): a lane reads copy (lane mod replication), so the 8 / 16 / 32 lanes
// that share one shared-memory wavefront of a 128- / 64- / 32-bit load always hit distinct banks.
// The kernel is then bound by shared-memory bandwidth alone: csize bytes per data byte at 128 B/clk/SM.
constexpr int kRsMaxWarps = 16;
constexpr int kPlaneWords = 256 * 32;                 // 32 KiB per plane and slicing table

template <int NW> struct rs_planes {
  static constexpr int N4 = NW / 4, N2 = (NW % 4) / 2, N1 = NW % 2, COUNT = N4 + N2 + N1;
  static constexpr int SL = (COUNT <= 2) ? 2 : 1;     // symbols per LFSR step (what fits next to the staging rows)
};

// XOR the NW-word row of symbol f (table t) into acc[]; base pointers already carry the lane's replica offset
template <int NW>
__device__ __forceinline__ void fetch_row(const uint32_t* const* lane_base, int t, uint32_t f, uint32_t* acc) {
  typedef rs_planes<NW> PL;
  const uint32_t eo = f << 5;                          // 128 bytes per entry
  int pl = t * PL::COUNT, w = 0;
#pragma unroll
  for (int i = 0; i < PL::N4; i++, pl++, w += 4) {
    const uint4 v = *(const uint4*)(lane_base[pl] + eo);
    acc[w] ^= v.x; acc[w + 1] ^= v.y; acc[w + 2] ^= v.z; acc[w + 3] ^= v.w;
  }
  if (PL::N2) { const uint2 v = *(const uint2*)(lane_base[pl] + eo); acc[w] ^= v.x; acc[w + 1] ^= v.y; pl++; w += 2; }
  if (PL::N1) { acc[w] ^= lane_base[pl][eo]; }
}

template <int NW, bool DECODE>
__global__ void __launch_bounds__(kRsMaxWarps * 32, 1) rs_tile_kernel(
    const uint8_t* __restrict__ buf, const uint64_t* __restrict__ off, const uint32_t* __restrict__ len, size_t stride,
    uint32_t ulen, int B, int dsize, int repair, const rs_table* __restrict__ gt, uint8_t* __restrict__ out,
    const uint64_t* __restrict__ out_off, size_t out_stride, int32_t* __restrict__ status, int32_t* __restrict__ fixed,
    int tiles_per_msg, int pitch_words) {
  typedef rs_planes<NW> PL;
  constexpr int SL = PL::SL;
  extern __shared__ __align__(16) unsigned char smem[];
  const int csize = NW * 4, block = dsize + csize;
  const int nwarps = blockDim.x >> 5;
  // shared memory: [SL][COUNT] planes | rs_table header (GF tables) | staging rows
  uint32_t* planes = (uint32_t*)smem;
  constexpr size_t plane_bytes = (size_t)SL * PL::COUNT * kPlaneWords * 4;
  rs_table* st = (rs_table*)(smem + plane_bytes);
  {
    const uint32_t* src = (const uint32_t*)gt;
    uint32_t* dstw = (uint32_t*)st;
    for (size_t i = threadIdx.x; i < sizeof(rs_table) / 4; i += blockDim.x) dstw[i] = src[i];
    const uint32_t* rows_g = (const uint32_t*)((const unsigned char*)gt + sizeof(rs_table));
    // table t = 0 is "symbol followed by SL - 1 zero-input steps" (first of the group), t = SL - 1 the plain rows
    for (int t = 0; t < SL; t++) {
      const uint32_t* tab = rows_g + (size_t)(SL - 1 - t) * 256 * NW;
      int pl = t * PL::COUNT, wo = 0;
      for (int k = 0; k < PL::COUNT; k++, pl++) {
        const int pw = (k < PL::N4) ? 4 : ((k < PL::N4 + PL::N2) ? 2 : 1);
        uint32_t* dstp = planes + (size_t)pl * kPlaneWords;
        for (int e = threadIdx.x; e < kPlaneWords; e += blockDim.x) {
          const int f = e >> 5, c = (e & 31) % pw;
          dstp[e] = __ldg(tab + f * NW + wo + c);
        }
        wo += pw;
      }
    }
  }
  __syncthreads();
  rs_smem_view sv;
  sv.gf_exp = st->gf_exp; sv.gf_log = st->gf_log; sv.rows = nullptr; sv.csize = csize; sv.nw = NW;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t* lane_base[SL * PL::COUNT];
#pragma unroll
  for (int t = 0; t < SL; t++) {
#pragma unroll
    for (int k = 0; k < PL::COUNT; k++) {
      const int pw = (k < PL::N4) ? 4 : ((k < PL::N4 + PL::N2) ? 2 : 1);
      lane_base[t * PL::COUNT + k] = planes + (size_t)(t * PL::COUNT + k) * kPlaneWords + (lane % (32 / pw)) * pw;
    }
  }
  const size_t stage_off = (plane_bytes + sizeof(rs_table) + 15) & ~(size_t)15;
  uint32_t* rows_w = (uint32_t*)(smem + stage_off) + (size_t)w * 32 * pitch_words;       // [32][pitch_words]
  const int rowb = DECODE ? block : dsize;             // bytes per input row
  const int outb = DECODE ? dsize : block;             // bytes per output row
  const long long total_tiles = (long long)B * tiles_per_msg;
  const long long wstep = (long long)gridDim.x * nwarps;
  for (long long t = (long long)blockIdx.x * nwarps + w; t < total_tiles; t += wstep) {
    const int m = (int)(t / tiles_per_msg), kt = (int)(t - (long long)m * tiles_per_msg);
    const uint8_t* src = buf + (off ? off[m] : (uint64_t)m * stride);
    uint8_t* dst = out + (out_off ? out_off[m] : (uint64_t)m * out_stride);
    const uint32_t L = len ? len[m] : ulen;
    const int ncw = (int)((L + rowb - 1) / rowb);
    const int k0 = kt * 32;
    if (k0 >= ncw) continue;
    const int nrows = min(32, ncw - k0);
    const uint8_t* tin = src + (size_t)k0 * rowb;
    const uint32_t tbytes_in = min((uint32_t)(nrows * rowb), L - (uint32_t)k0 * rowb);
    __syncwarp();
    // ---- stage in: row by row, lanes along the row (asynchronous 4-byte copies global -> shared, so the whole
    //      tile is in flight at once and nothing is staged through registers) ----
    if ((((uintptr_t)tin) & 3) == 0 && (rowb & 3) == 0) {
      const int rw = rowb >> 2;
      const int full = (int)(tbytes_in / (uint32_t)rowb);                 // complete rows; at most one short row follows
      const uint32_t* gsrc = (const uint32_t*)tin + lane;
      uint32_t sa = (uint32_t)__cvta_generic_to_shared(rows_w + lane);
      const uint32_t sstep = (uint32_t)pitch_words * 4;
      if (rw <= 32) {
        const bool act = lane < rw;
#pragma unroll 4
        for (int r = 0; r < full; r++, gsrc += rw, sa += sstep)
          if (act) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
      } else {
#pragma unroll 1
        for (int r = 0; r < full; r++, gsrc += rw, sa += sstep) {
#pragma unroll 1
          for (int c = 0; c + lane < rw; c += 32)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa + (uint32_t)c * 4), "l"(gsrc + c) : "memory");
        }
      }
      if (full < nrows) {                                                  // short last row of the message
        const uint32_t left = tbytes_in - (uint32_t)full * (uint32_t)rowb;
        const int wr = (int)(left >> 2);
#pragma unroll 1
        for (int c = 0; c + lane < wr; c += 32)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa + (uint32_t)c * 4), "l"(gsrc + c) : "memory");
        if (lane < (int)(left & 3)) ((uint8_t*)(rows_w + full * pitch_words))[wr * 4 + lane] = tin[(size_t)full * rowb + wr * 4 + lane];
      }
      {   // pull this warp's next tile towards L2 meanwhile
        const long long tn = t + wstep;
        if (tn < total_tiles && !off) {
          const int mn = (int)(tn / tiles_per_msg), ktn = (int)(tn - (long long)mn * tiles_per_msg);
          const char* nx = (const char*)buf + (uint64_t)mn * stride + (size_t)ktn * 32 * rowb;
          if (lane * 128 < 32 * rowb) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + lane * 128));
        }
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
    } else {
      int r = lane / rowb, c = lane - r * rowb;
      for (uint32_t bi = lane; bi < tbytes_in; bi += 32) {
        ((uint8_t*)(rows_w + r * pitch_words))[c] = tin[bi];
        c += 32;
        while (c >= rowb) { c -= rowb; r++; }
      }
    }
    __syncwarp();
    // ---- per-lane codeword ----
    int my_fail = 0, my_fixed = 0;
    int rowlen = 0, dlen = 0;
    if (lane < nrows) {
      const uint32_t start = (uint32_t)(k0 + lane) * rowb;
      rowlen = (int)min((uint32_t)rowb, L - start);
      dlen = DECODE ? max(rowlen - csize, 0) : rowlen;
      uint32_t* row = rows_w + lane * pitch_words;
      uint32_t par[NW];
#pragma unroll
      for (int i = 0; i < NW; i++) par[i] = 0;
      const int nq = (!DECODE || repair) ? (dlen >> 2) : 0;
#pragma unroll 2
      for (int j = 0; j < nq; j++) {
        const uint32_t dw = row[j];
        if constexpr (SL == 2) {
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const uint32_t x = par[0] ^ (h ? (dw >> 16) : dw);
#pragma unroll
            for (int i = 0; i < NW; i++) par[i] = __byte_perm(par[i], (i + 1 < NW) ? par[i + 1] : 0u, 0x5432);
            fetch_row<NW>(lane_base, 0, x & 0xFF, par);
            fetch_row<NW>(lane_base, 1, (x >> 8) & 0xFF, par);
          }
        } else {
#pragma unroll
          for (int h = 0; h < 4; h++) {
            const uint32_t f = (par[0] ^ (dw >> (8 * h))) & 0xFF;
#pragma unroll
            for (int i = 0; i < NW; i++) par[i] = (i + 1 < NW) ? __funnelshift_r(par[i], par[i + 1], 8) : (par[i] >> 8);
            fetch_row<NW>(lane_base, 0, f, par);
          }
        }
      }
      const uint8_t* rb = (const uint8_t*)row;
      for (int j = nq * 4; j < ((!DECODE || repair) ? dlen : 0); j++) {
        const uint32_t f = (par[0] ^ rb[j]) & 0xFF;
#pragma unroll
        for (int i = 0; i < NW; i++) par[i] = (i + 1 < NW) ? __funnelshift_r(par[i], par[i + 1], 8) : (par[i] >> 8);
        fetch_row<NW>(lane_base, SL - 1, f, par);
      }
      if (!DECODE) {
        uint8_t* pb = (uint8_t*)row + dlen;              // parity follows the data inside the staged row
        if ((dlen & 3) == 0) {
#pragma unroll
          for (int i = 0; i < NW; i++) row[(dlen >> 2) + i] = par[i];
        } else {
          for (int j = 0; j < csize; j++) pb[j] = (uint8_t)(par[j >> 2] >> (8 * (j & 3)));
        }
      } else if (repair) {
        if (rowlen <= csize) my_fail++;
        else {
          uint8_t rem[32];
          int dirty = 0;
          for (int j = 0; j < csize; j++) {
            rem[j] = (uint8_t)(par[j >> 2] >> (8 * (j & 3))) ^ rb[dlen + j];
            dirty |= rem[j];
          }
          if (dirty) {
            uint8_t fixedrow[256];
            const int rfix = rs_repair_word(sv, rb, rowlen, rem, fixedrow);
            uint8_t* wb = (uint8_t*)row;
            if (rfix < 0) { for (int i = 0; i < dlen; i++) wb[i] = 0; my_fail++; }
            else { for (int i = 0; i < dlen; i++) wb[i] = fixedrow[i]; my_fixed += rfix; }
          }
        }
      }
    }
    __syncwarp();
    // ---- write back: rows are contiguous in the output (all rows full except possibly the last) ----
    uint8_t* tout = dst + (size_t)k0 * outb;
    const int my_out = DECODE ? dlen : (dlen + csize);
    if ((((uintptr_t)tout) & 3) == 0 && (outb & 3) == 0) {
      const uint32_t* row = rows_w + lane;
      uint32_t* orow = (uint32_t*)tout + lane;
      const int ow = outb >> 2;
      const int last_len = __shfl_sync(0xffffffffu, my_out, nrows - 1);   // only the last row can be short
      const int full = (last_len == outb) ? nrows : nrows - 1;
      if (ow <= 32) {
        const bool act = lane < ow;
#pragma unroll 4
        for (int r = 0; r < full; r++, row += pitch_words, orow += ow)
          if (act) *orow = *row;
      } else {
#pragma unroll 1
        for (int r = 0; r < full; r++, row += pitch_words, orow += ow) {
#pragma unroll 1
          for (int j = 0; j + lane < ow; j += 32) orow[j] = row[j];
        }
      }
      if (full < nrows) {
#pragma unroll 1
        for (int j = 0; j + lane < (last_len >> 2); j += 32) orow[j] = row[j];
        if (lane < (last_len & 3)) ((uint8_t*)(orow - lane))[(last_len & ~3) + lane] = ((const uint8_t*)(row - lane))[(last_len & ~3) + lane];
      }
    } else {
      for (int r = 0; r < nrows; r++) {
        const int rl = __shfl_sync(0xffffffffu, my_out, r);
        const uint8_t* row = (const uint8_t*)(rows_w + r * pitch_words);
        uint8_t* orow = tout + (size_t)r * outb;
        for (int j = lane; j < rl; j += 32) orow[j] = row[j];
      }
    }
    if (DECODE && (status || fixed)) {
#pragma unroll
      for (int sft = 16; sft > 0; sft >>= 1) {
        my_fail += __shfl_xor_sync(0xffffffffu, my_fail, sft);
        my_fixed += __shfl_xor_sync(0xffffffffu, my_fixed, sft);
      }
      if (lane == 0) {
        if (status && my_fail) atomicAdd(&status[m], my_fail);
        if (fixed && my_fixed) atomicAdd(&fixed[m], my_fixed);
      }
    }
  }
}

template <int NW, bool DECODE>
int launch_tiles_nw(frad_ctx* ctx, const uint8_t* buf, const uint64_t* off, const uint32_t* len, size_t stride, uint32_t ulen, int B,
                    int dsize, int repair, const rs_table* t, uint8_t* out, const uint64_t* out_off, size_t out_stride,
                    int32_t* status, int32_t* fixed, uint32_t max_len) {
  typedef rs_planes<NW> PL;
  const int csize = NW * 4;
  const int rowb = DECODE ? dsize + csize : dsize;
  const int staged = dsize + csize;                                   // encode appends the parity inside the staged row
  const int pitch = ((staged + 3) / 4) | 1;
  const int tiles_per_msg = (int)(((max_len + rowb - 1) / rowb + 31) / 32);
  if (tiles_per_msg < 1) return FRAD_OK;
  const size_t fixed_bytes = (((size_t)PL::SL * PL::COUNT * kPlaneWords * 4 + sizeof(rs_table)) + 15) & ~(size_t)15;
  const size_t per_warp = (size_t)32 * pitch * 4;
  const long long total = (long long)B * tiles_per_msg;
  int warps = (int)((ctx->smem_optin - fixed_bytes) / per_warp);
  if (warps > kRsMaxWarps) warps = kRsMaxWarps;
  if (warps < 1) return frad_fail(ctx, FRAD_E_UNSUPPORTED, "rs tile kernel does not fit in shared memory");
  long long per_cta_even = (total + ctx->sm_count - 1) / ctx->sm_count;   // small batches: no idle warps around the tables
  if (per_cta_even < warps) warps = (int)(per_cta_even < 1 ? 1 : per_cta_even);
  const size_t smem = fixed_bytes + (size_t)warps * per_warp;
  long long want = (total + warps - 1) / warps;
  auto k = rs_tile_kernel<NW, DECODE>;
  FRAD_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)(want < ctx->sm_count ? want : ctx->sm_count);
  k<<<grid, warps * 32, smem, ctx->stream>>>(buf, off, len, stride, ulen, B, dsize, repair, t, out, out_off, out_stride, status, fixed,
                                            tiles_per_msg, pitch);
  FRAD_LAUNCH_CHECK(ctx);
  return FRAD_OK;
}

template <bool DECODE>
int launch_tiles(frad_ctx* ctx, const uint8_t* buf, const uint64_t* off, const uint32_t* len, size_t stride, uint32_t ulen, int B,
                 int dsize, int csize, int repair, const rs_table* t, uint8_t* out, const uint64_t* out_off, size_t out_stride,
                 int32_t* status, int32_t* fixed, uint32_t max_len) {
#define RS_PICK(NWV)                                                                                                         \
  case NWV:                                                                                                                  \
    return launch_tiles_nw<NWV, DECODE>(ctx, buf, off, len, stride, ulen, B, dsize, repair, t, out, out_off, out_stride, status, \
                                        fixed, max_len);
  switch (csize / 4) { RS_PICK(1) RS_PICK(2) RS_PICK(3) RS_PICK(4) RS_PICK(5) RS_PICK(6) RS_PICK(7) RS_PICK(8) default: return FRAD_E_UNSUPPORTED; }
#undef RS_PICK
}

bool tile_ok(int csize) { return csize >= 4 && csize <= 32 && (csize & 3) == 0; }
size_t table_bytes(int csize) { return sizeof(rs_table) + (size_t)256 * ((csize + 3) / 4) * 4; }

}  // namespace

extern "C" size_t frad_rs_encoded_len(size_t len, int dsize, int csize) {
  if (dsize <= 0) return len;
  return len + ((len + dsize - 1) / dsize) * (size_t)csize;
}
extern "C" size_t frad_rs_decoded_len(size_t len, int dsize, int csize) {
  const size_t block = (size_t)dsize + csize;
  if (block == 0) return 0;
  const size_t full = len / block, rem = len % block;
  return full * dsize + ((rem > (size_t)csize) ? rem - csize : 0);
}

extern "C" int frad_rs_encode_dev(frad_ctx* ctx, const uint8_t* buf, const uint64_t* off, const uint32_t* len,
                                  size_t stride, uint32_t ulen, int B, int dsize, int csize,
                                  uint8_t* out, const uint64_t* out_off, size_t out_stride) {
  if (!ctx) return FRAD_E_ARG;
  if (B < 0 || dsize < 1 || csize < 0 || dsize + csize > 255 || !out || (!buf && B > 0))
    return frad_fail(ctx, FRAD_E_ARG, "frad_rs_encode_dev: bad argument");
  if (B == 0) return FRAD_OK;
  FRAD_CUDA(ctx, cudaSetDevice(ctx->device));
  const rs_table* t = get_table(ctx, csize);
  if (!t) return FRAD_E_NOMEM;
  if (tile_ok(csize) && ulen) {                  // with len[]: ulen is its upper bound (0 = unknown: per-message kernels)
    const uint32_t mx = ulen;
    return launch_tiles<false>(ctx, buf, off, len, stride, ulen, B, dsize, csize, 1, t, out, out_off, out_stride, nullptr, nullptr, mx);
  }
  const size_t smem = table_bytes(csize);
  FRAD_CUDA(ctx, cudaFuncSetAttribute(rs_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = B < ctx->sm_count * 8 ? B : ctx->sm_count * 8;
  rs_encode_kernel<<<grid, 128, smem, ctx->stream>>>(buf, off, len, stride, ulen, B, dsize, t, out, out_off, out_stride);
  FRAD_LAUNCH_CHECK(ctx);
  return FRAD_OK;
}

extern "C" int frad_rs_decode_dev(frad_ctx* ctx, const uint8_t* buf, const uint64_t* off, const uint32_t* len,
                                  size_t stride, uint32_t ulen, int B, int dsize, int csize, int repair,
                                  uint8_t* out, const uint64_t* out_off, size_t out_stride,
                                  int32_t* status, int32_t* fixed) {
  if (!ctx) return FRAD_E_ARG;
  if (B < 0 || dsize < 1 || csize < 0 || dsize + csize > 255 || !out || (!buf && B > 0))
    return frad_fail(ctx, FRAD_E_ARG, "frad_rs_decode_dev: bad argument");
  if (B == 0) return FRAD_OK;
  FRAD_CUDA(ctx, cudaSetDevice(ctx->device));
  const rs_table* t = get_table(ctx, csize);
  if (!t) return FRAD_E_NOMEM;
  if (tile_ok(csize) && ulen) {                  // with len[]: ulen is its upper bound (0 = unknown: per-message kernels)
    const uint32_t mx = ulen;
    if (status) FRAD_CUDA(ctx, cudaMemsetAsync(status, 0, sizeof(int32_t) * (size_t)B, ctx->stream));
    if (fixed) FRAD_CUDA(ctx, cudaMemsetAsync(fixed, 0, sizeof(int32_t) * (size_t)B, ctx->stream));
    return launch_tiles<true>(ctx, buf, off, len, stride, ulen, B, dsize, csize, repair, t, out, out_off, out_stride, status, fixed, mx);
  }
  const size_t smem = table_bytes(csize);
  FRAD_CUDA(ctx, cudaFuncSetAttribute(rs_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = B < ctx->sm_count * 8 ? B : ctx->sm_count * 8;
  rs_decode_kernel<<<grid, 128, smem, ctx->stream>>>(buf, off, len, stride, ulen, B, dsize, repair, t, out, out_off, out_stride, status, fixed);
  FRAD_LAUNCH_CHECK(ctx);
  return FRAD_OK;
}
