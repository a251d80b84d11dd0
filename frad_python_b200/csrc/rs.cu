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

extern thread_local uint32_t g_rs_max_len;

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
  if (cudaMalloc(&d, bytes) != cudaSuccess || cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
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
// four symbols per step (slicing tables: 4 row fetches of csize bytes per 4 data bytes, the register shift
// becomes a register renaming), coalesced write-back of data || parity (encode) or of the data (decode).
size_t table_bytes_sliced(int csize);
constexpr int kRsWarps = 4;

template <int NW, bool DECODE>
__global__ void __launch_bounds__(kRsWarps * 32) rs_tile_kernel(
    const uint8_t* __restrict__ buf, const uint64_t* __restrict__ off, const uint32_t* __restrict__ len, size_t stride,
    uint32_t ulen, int B, int dsize, int repair, const rs_table* __restrict__ gt, uint8_t* __restrict__ out,
    const uint64_t* __restrict__ out_off, size_t out_stride, int32_t* __restrict__ status, int32_t* __restrict__ fixed,
    int tiles_per_msg, int pitch_words) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int csize = NW * 4, block = dsize + csize;
  const size_t tbytes = sizeof(rs_table) + (size_t)4 * 256 * NW * 4;
  {
    const uint32_t* src = (const uint32_t*)gt;
    uint32_t* dstw = (uint32_t*)smem;
    for (size_t i = threadIdx.x; i < tbytes / 4; i += blockDim.x) dstw[i] = src[i];
  }
  __syncthreads();
  const rs_table* st = (const rs_table*)smem;
  rs_smem_view sv;
  sv.gf_exp = st->gf_exp; sv.gf_log = st->gf_log; sv.rows = (const uint32_t*)(smem + sizeof(rs_table)); sv.csize = csize; sv.nw = NW;
  const uint32_t* T3 = sv.rows;                      // symbol is the last of the four
  const uint32_t* T2 = T3 + 256 * NW;
  const uint32_t* T1 = T2 + 256 * NW;
  const uint32_t* T0 = T1 + 256 * NW;                // symbol is the first of the four (3 more steps follow)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const size_t stage_off = (tbytes + 15) & ~(size_t)15;
  uint32_t* rows_w = (uint32_t*)(smem + stage_off) + (size_t)w * 32 * pitch_words;       // [32][pitch_words]
  const int rowb = DECODE ? block : dsize;             // bytes per input row
  const int outb = DECODE ? dsize : block;             // bytes per output row
  // tile enumeration: flat over (message, tile) for uniform batches, CTA-per-message otherwise
  const long long total_tiles = (long long)B * tiles_per_msg;
  const long long wstep = (long long)gridDim.x * kRsWarps;
  for (long long t = (long long)blockIdx.x * kRsWarps + w; t < total_tiles; t += wstep) {
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
    // ---- stage in ----
    if ((((uintptr_t)tin) & 3) == 0 && (rowb & 3) == 0) {
      const int rw = rowb >> 2, nwords = (int)(tbytes_in >> 2);
      int r = lane / rw, c = lane - r * rw;
      for (int wi = lane; wi < nwords; wi += 32) {
        rows_w[r * pitch_words + c] = __ldg((const uint32_t*)tin + wi);
        c += 32;
        while (c >= rw) { c -= rw; r++; }
      }
      for (uint32_t bi = (uint32_t)nwords * 4 + lane; bi < tbytes_in; bi += 32) {       // < 4 trailing bytes
        const int rr = (int)(bi / rowb), cc = (int)(bi - (uint32_t)rr * rowb);
        ((uint8_t*)(rows_w + rr * pitch_words))[cc] = tin[bi];
      }
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
      for (int j = 0; j < nq; j++) {
        const uint32_t x = par[0] ^ row[j];
        const uint32_t* a = T0 + (x & 0xFF) * NW;
        const uint32_t* b = T1 + ((x >> 8) & 0xFF) * NW;
        const uint32_t* c = T2 + ((x >> 16) & 0xFF) * NW;
        const uint32_t* d = T3 + (x >> 24) * NW;
        if constexpr ((NW & 1) == 0) {                   // 8-byte aligned rows: half as many shared-memory loads
          const uint2 *a2 = (const uint2*)a, *b2 = (const uint2*)b, *c2 = (const uint2*)c, *d2 = (const uint2*)d;
#pragma unroll
          for (int i = 0; i < NW / 2; i++) {
            const uint2 va = a2[i], vb = b2[i], vc = c2[i], vd = d2[i];
            par[2 * i] = par[2 * i + 1] ^ va.x ^ vb.x ^ vc.x ^ vd.x;
            par[2 * i + 1] = ((2 * i + 2 < NW) ? par[2 * i + 2] : 0u) ^ va.y ^ vb.y ^ vc.y ^ vd.y;
          }
        } else {
#pragma unroll
          for (int i = 0; i < NW; i++) par[i] = ((i + 1 < NW) ? par[i + 1] : 0u) ^ a[i] ^ b[i] ^ c[i] ^ d[i];
        }
      }
      const uint8_t* rb = (const uint8_t*)row;
      for (int j = nq * 4; j < ((!DECODE || repair) ? dlen : 0); j++) {
        const uint32_t f = (par[0] ^ rb[j]) & 0xFF;
        const uint32_t* rr = T3 + f * NW;
#pragma unroll
        for (int i = 0; i < NW; i++) par[i] = ((i + 1 < NW) ? __funnelshift_r(par[i], par[i + 1], 8) : (par[i] >> 8)) ^ rr[i];
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
    for (int r = 0; r < nrows; r++) {
      const int rl = __shfl_sync(0xffffffffu, DECODE ? dlen : (dlen + csize), r);
      const uint32_t* row = rows_w + r * pitch_words;
      uint8_t* orow = tout + (size_t)r * outb;
      if ((((uintptr_t)orow) & 3) == 0 && (rl & 3) == 0) {
        for (int j = lane; j < (rl >> 2); j += 32) ((uint32_t*)orow)[j] = row[j];
      } else {
        for (int j = lane; j < rl; j += 32) orow[j] = ((const uint8_t*)row)[j];
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

template <bool DECODE>
int launch_tiles(frad_ctx* ctx, const uint8_t* buf, const uint64_t* off, const uint32_t* len, size_t stride, uint32_t ulen, int B,
                 int dsize, int csize, int repair, const rs_table* t, uint8_t* out, const uint64_t* out_off, size_t out_stride,
                 int32_t* status, int32_t* fixed, uint32_t max_len) {
  const int rowb = DECODE ? dsize + csize : dsize;
  const int staged = dsize + csize;                                   // encode appends the parity inside the staged row
  const int pitch = ((staged + 3) / 4) | 1;
  const int tiles_per_msg = (int)(((max_len + rowb - 1) / rowb + 31) / 32);
  if (tiles_per_msg < 1) return FRAD_OK;
  const size_t tb = (table_bytes_sliced(csize) + 15) & ~(size_t)15;
  const size_t smem = tb + (size_t)kRsWarps * 32 * pitch * 4;
  const long long total = (long long)B * tiles_per_msg;
  long long want = (total + kRsWarps - 1) / kRsWarps;
  const void* k = nullptr;
#define RS_PICK(NWV) case NWV: k = (const void*)rs_tile_kernel<NWV, DECODE>; break;
  switch (csize / 4) { RS_PICK(1) RS_PICK(2) RS_PICK(3) RS_PICK(4) RS_PICK(5) RS_PICK(6) RS_PICK(7) RS_PICK(8) default: return FRAD_E_UNSUPPORTED; }
#undef RS_PICK
  FRAD_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  FRAD_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kRsWarps * 32, smem));
  if (per_sm < 1) per_sm = 1;
  const long long cap = (long long)per_sm * ctx->sm_count;
  const int grid = (int)(want < cap ? want : cap);
  void* args[] = {(void*)&buf, (void*)&off, (void*)&len, (void*)&stride, (void*)&ulen, (void*)&B, (void*)&dsize, (void*)&repair,
                  (void*)&t, (void*)&out, (void*)&out_off, (void*)&out_stride, (void*)&status, (void*)&fixed,
                  (void*)&tiles_per_msg, (void*)&pitch};
  FRAD_CUDA(ctx, cudaLaunchKernel(k, dim3(grid), dim3(kRsWarps * 32), args, smem, ctx->stream));
  FRAD_LAUNCH_CHECK(ctx);
  return FRAD_OK;
}

bool tile_ok(int csize) { return csize >= 4 && csize <= 32 && (csize & 3) == 0; }
size_t table_bytes(int csize) { return sizeof(rs_table) + (size_t)256 * ((csize + 3) / 4) * 4; }
size_t table_bytes_sliced(int csize) { return sizeof(rs_table) + (size_t)4 * 256 * ((csize + 3) / 4) * 4; }

}  // namespace

// upper bound of the message lengths behind a device-resident len[] array; set by the host-pointer entry
// points (which know the lengths) around their *_dev call, 0 = unknown (per-message kernels are used)
thread_local uint32_t g_rs_max_len = 0;
void frad_rs_set_max_len(uint32_t v) { g_rs_max_len = v; }

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
  if (tile_ok(csize) && (!len || g_rs_max_len)) {
    const uint32_t mx = len ? g_rs_max_len : ulen;
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
  if (tile_ok(csize) && (!len || g_rs_max_len)) {
    const uint32_t mx = len ? g_rs_max_len : ulen;
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
