// PNG encoder, per-thread logic (the caller's `out_image.save(path)`, test_wo_llava.py:97 / test.py:130, on the device).
//
// Everything here is written as PHASE functions of an explicit thread index: a kernel in png.cu is
//   phase_a(threadIdx.x, ...); __syncthreads(); phase_b(threadIdx.x, ...); ...
// and nothing else, with all block-shared state in a plain struct.  The same header compiles with g++ (no CUDA),
// where a test harness (tests/png_host_emul.cpp — test infrastructure, never linked into the product) runs the phases
// in a loop over blocks and threads: the indexing, the Huffman construction and the bit packing are checked on the
// host against zlib / PIL, and the GPU test then asserts the device output is byte-identical to that emulation.
//
// File layout produced (all on the device; the host only copies `result[0]` bytes out):
//   signature, IHDR, one IDAT chunk per deflate block, IEND.
// Stream: zlib header 78 01; the filtered scanlines (adaptive filter per row, libpng's minimum-sum-of-absolute-values
// rule) cut into 16 KB pieces; each piece is ONE dynamic-Huffman deflate block (RFC 1951 §3.2.7; literals, run matches
// at distance 1 through a single distance code), followed by an empty stored block (000 + pad + 00 00 FF FF) so every
// piece ends on a byte boundary and the pieces are independent — the unit of parallelism; then a final empty stored
// block and the Adler-32 of the filtered stream.  LZ77 is restricted to what needs no search and no cross-thread state:
// inside a thread's 64-byte run, six or more equal bytes become literal + one match at distance 1 (flat image regions
// filter to runs of zeros: 0.15 instead of 1 bit per byte).  A general match search buys a few percent on photographic
// residuals and is serial per stream; Huffman coding over per-16 KB statistics is where PNG's compression of such
// images comes from.
#pragma once
#include <stdint.h>
#include <string.h>

// In the product (nvcc) the phase functions exist as DEVICE code only — the library has no host version of the encoder and
// no CPU fallback; under g++ (the test harness) the same text is plain inline C++.  FD_HOST_HD marks the two size formulas the
// host side of the C-ABI needs as well.
#if defined(__CUDACC__)
#define FD_HD __device__ __forceinline__
#define FD_HOST_HD __host__ __device__ __forceinline__
#else
#define FD_HD inline
#define FD_HOST_HD inline
#endif
#if defined(__CUDA_ARCH__)
#define FD_PNG_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#define FD_PNG_ATOMIC_OR(p, v) atomicOr((p), (v))
#else
#define FD_PNG_ATOMIC_ADD(p, v) (*(p) += (v))
#define FD_PNG_ATOMIC_OR(p, v) (*(p) |= (v))
#endif

namespace fd {
namespace png {

// little-endian 32-bit load from a 4-byte aligned address
FD_HD uint32_t load_u32(const uint8_t* p) {
#if defined(__CUDA_ARCH__)
  return *reinterpret_cast<const uint32_t*>(p);
#else
  uint32_t v;
  memcpy(&v, p, 4);
  return v;
#endif
}

constexpr int kThreads = 256;                 // threads per block in every kernel below
constexpr int kChunk = 16384;                 // filtered bytes per deflate block
constexpr int kRun = kChunk / kThreads;       // consecutive bytes one thread encodes
constexpr int kRunPadWords = kRun / 4 + 1;    // words between two runs of a block staged in shared memory (17: the 32 lanes of a
                                              // warp, each walking its own run, start in 32 different banks)
constexpr int kSyms = 286;                    // literals 0..255, end-of-block, length codes 257..285
constexpr int kCodeStride = 288;              // words per block in the code table (286 padded)
constexpr int kHdrWords = 72;                 // bit-packed block header, <= 17 + 57 + 287*7 = 2083 bits
constexpr int kMinRun = 6;                    // equal bytes from which a run is coded as literal + match(distance 1)
constexpr int kMaxLitBits = 15, kMaxClBits = 7;
constexpr int kSigIhdrBytes = 8 + 25;         // signature + IHDR chunk
constexpr uint32_t kAdlerMod = 65521u;

// per-block metadata written by the statistics kernel
struct BlockMeta {
  uint32_t hdr_bits;        // bits of the dynamic-block header
  uint32_t payload_bytes;   // bytes of this block incl. the byte-aligning stored block (excl. zlib header / trailer)
};

// ------------------------------------------------------------------------------------------------ filters
FD_HD int paeth(int a, int b, int c) {
  const int p = a + b - c;
  const int pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
// filtered value of byte i of row y under filter f (PNG spec §9.2); prev == nullptr for the first row
FD_HD uint8_t filter_byte(const uint8_t* cur, const uint8_t* prev, int i, int bpp, int f) {
  const int x = cur[i];
  const int a = i >= bpp ? cur[i - bpp] : 0;
  const int b = prev ? prev[i] : 0;
  const int c = (prev && i >= bpp) ? prev[i - bpp] : 0;
  int pred = 0;
  if (f == 1) pred = a;
  else if (f == 2) pred = b;
  else if (f == 3) pred = (a + b) >> 1;
  else if (f == 4) pred = paeth(a, b, c);
  return (uint8_t)(x - pred);
}
FD_HD uint32_t abs_signed_byte(uint8_t v) { return v < 128 ? v : 256u - v; }

struct FilterShared {
  uint32_t part[5][kThreads];
  unsigned long long pa[kThreads], pb[kThreads];
  uint32_t sums[5];
  int choice;
};

// phase 1: per-thread sums of |filtered| for the five filters over bytes tid, tid + 256, ...
FD_HD void filter_phase_sums(const uint8_t* img, long long stride, int y, int rb, int bpp, int tid, FilterShared& sh) {
  const uint8_t* cur = img + (long long)y * stride;
  const uint8_t* prev = y > 0 ? cur - stride : nullptr;
  uint32_t s[5] = {0, 0, 0, 0, 0};
  for (int i = tid; i < rb; i += kThreads) {
#pragma unroll
    for (int f = 0; f < 5; ++f) s[f] += abs_signed_byte(filter_byte(cur, prev, i, bpp, f));
  }
#pragma unroll
  for (int f = 0; f < 5; ++f) sh.part[f][tid] = s[f];
}
// phase 2: threads 0..4 total one filter each
FD_HD void filter_phase_total(int tid, FilterShared& sh) {
  if (tid < 5) {
    uint32_t t = 0;
    for (int k = 0; k < kThreads; ++k) t += sh.part[tid][k];
    sh.sums[tid] = t;
  }
}
// phase 3: thread 0 picks the filter (lowest index among the minima, as libpng's strict "<" scan does)
FD_HD void filter_phase_choose(int tid, FilterShared& sh) {
  if (tid == 0) {
    int best = 0;
    for (int f = 1; f < 5; ++f)
      if (sh.sums[f] < sh.sums[best]) best = f;
    sh.choice = best;
  }
}
// phase 4: write the filtered row (filter byte + rb bytes) and this thread's Adler-32 partial sums:
//   A = sum d_i ; B = sum (n - i) d_i  over the row's n = rb + 1 bytes (i = position in the row)
FD_HD void filter_phase_write(const uint8_t* img, long long stride, int y, int rb, int bpp, uint8_t* filt, int tid, FilterShared& sh) {
  const uint8_t* cur = img + (long long)y * stride;
  const uint8_t* prev = y > 0 ? cur - stride : nullptr;
  const int f = sh.choice;
  const int n = rb + 1;
  uint8_t* row = filt + (long long)y * n;
  unsigned long long a = 0, b = 0;
  if (tid == 0) {
    row[0] = (uint8_t)f;
    a += (unsigned)f;
    b += (unsigned long long)n * (unsigned)f;
  }
  for (int i = tid; i < rb; i += kThreads) {
    const uint8_t v = filter_byte(cur, prev, i, bpp, f);
    row[1 + i] = v;
    a += v;
    b += (unsigned long long)(n - 1 - i) * v;
  }
  sh.pa[tid] = a;
  sh.pb[tid] = b;
}
// phase 5: thread 0 reduces the Adler partials of the row (mod 65521)
FD_HD void filter_phase_adler(int y, uint32_t* adler_rows, int tid, FilterShared& sh) {
  if (tid == 0) {
    unsigned long long a = 0, b = 0;
    for (int k = 0; k < kThreads; ++k) { a += sh.pa[k]; b += sh.pb[k] % kAdlerMod; }
    adler_rows[2 * y] = (uint32_t)(a % kAdlerMod);
    adler_rows[2 * y + 1] = (uint32_t)(b % kAdlerMod);
  }
}

// ------------------------------------------------------------------------------------------------ Huffman codes
struct CodeScratch {
  uint16_t order[2][288];        // radix-sort ping-pong of the used symbols
  uint32_t weight[2 * 288];
  uint16_t parent[2 * 288];
  uint8_t depth[2 * 288];
  uint32_t count[32];            // symbols per code length
  uint32_t next_code[17];
  uint8_t len[288];              // code lengths of the alphabet being built
  uint8_t lit_len[288];          // literal/length code lengths of the block
  uint32_t lit_next[17];         // first canonical code of every literal/length code length
  int dist_len;                  // 1 when the block has matches (one distance code of one bit), else 0
  uint8_t all_len[288];          // HLIT literal/length lengths + 1 distance length
  uint8_t tok_sym[288], tok_extra[288];
  uint32_t cl_freq[19];
  uint8_t cl_len[19];
  uint32_t cl_code[19];
};

FD_HD uint32_t reverse_bits(uint32_t v, int n) {          // the low n bits of v in reverse order, 1 <= n <= 32
  v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
  v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
  v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
  v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
  v = (v >> 16) | (v << 16);
  return v >> (32 - n);
}

// Length-limited Huffman code lengths for freq[0..nsym): optimal tree (two-queue construction over radix-sorted
// leaves), then, if a leaf is deeper than maxbits, the counts per length are repaired (one deepest code removed, one
// shorter code split, until the Kraft sum is 1) and lengths are re-dealt in frequency order.  Writes s.len[0..nsym).
FD_HD void huffman_lengths(const uint32_t* freq, int nsym, int maxbits, CodeScratch& s) {
  int n = 0;
  uint32_t fmax = 0;
  for (int i = 0; i < nsym; ++i) {
    s.len[i] = 0;
    if (freq[i]) { s.order[0][n++] = (uint16_t)i; if (freq[i] > fmax) fmax = freq[i]; }
  }
  if (n == 0) return;
  if (n == 1) {                       // a complete code needs two leaves: pair the symbol with a neighbour
    const int a = s.order[0][0];
    s.len[a] = 1;
    s.len[a == 0 ? 1 : a - 1] = 1;
    return;
  }
  // stable LSD radix sort by frequency, 8 bits per pass
  int cur = 0;
  for (int shift = 0; shift < 32 && (fmax >> shift) != 0; shift += 8) {
    uint32_t bucket[256];
    for (int k = 0; k < 256; ++k) bucket[k] = 0;
    for (int k = 0; k < n; ++k) ++bucket[(freq[s.order[cur][k]] >> shift) & 255u];
    uint32_t run = 0;
    for (int k = 0; k < 256; ++k) { const uint32_t c = bucket[k]; bucket[k] = run; run += c; }
    for (int k = 0; k < n; ++k) {
      const uint16_t sym = s.order[cur][k];
      s.order[cur ^ 1][bucket[(freq[sym] >> shift) & 255u]++] = sym;
    }
    cur ^= 1;
  }
  const uint16_t* sorted = s.order[cur];          // ascending frequency
  for (int k = 0; k < n; ++k) s.weight[k] = freq[sorted[k]];
  // two queues: leaves [leaf, n) and internal nodes [inode, next), both in ascending weight
  int leaf = 0, inode = n;
  for (int next = n; next < 2 * n - 1; ++next) {
    int pick[2];
    for (int j = 0; j < 2; ++j) {
      if (leaf < n && (inode >= next || s.weight[leaf] <= s.weight[inode])) pick[j] = leaf++;
      else pick[j] = inode++;
    }
    s.weight[next] = s.weight[pick[0]] + s.weight[pick[1]];
    s.parent[pick[0]] = (uint16_t)next;
    s.parent[pick[1]] = (uint16_t)next;
  }
  const int root = 2 * n - 2;
  s.depth[root] = 0;
  for (int k = root - 1; k >= 0; --k) {
    const int d = s.depth[s.parent[k]] + 1;
    s.depth[k] = (uint8_t)(d > 31 ? 31 : d);
  }
  for (int d = 0; d < 32; ++d) s.count[d] = 0;
  for (int k = 0; k < n; ++k) ++s.count[s.depth[k]];
  // enforce the limit
  for (int d = maxbits + 1; d < 32; ++d) { s.count[maxbits] += s.count[d]; s.count[d] = 0; }
  uint32_t total = 0;
  for (int d = maxbits; d >= 1; --d) total += s.count[d] << (maxbits - d);
  while (total > (1u << maxbits)) {
    --s.count[maxbits];
    for (int d = maxbits - 1; d >= 1; --d)
      if (s.count[d]) { --s.count[d]; s.count[d + 1] += 2; break; }
    --total;
  }
  // deal the lengths: rarest symbols get the longest codes
  int k = 0;
  for (int d = maxbits; d >= 1; --d)
    for (uint32_t c = 0; c < s.count[d]; ++c) s.len[sorted[k++]] = (uint8_t)d;
}

// canonical code values (RFC 1951 §3.2.2) from s.len[0..nsym); out[i] = (len << 16) | bit-reversed code
FD_HD void canonical_codes(int nsym, int maxbits, CodeScratch& s, uint32_t* out) {
  for (int d = 0; d <= maxbits; ++d) s.count[d] = 0;
  for (int i = 0; i < nsym; ++i) ++s.count[s.len[i]];
  s.count[0] = 0;
  uint32_t code = 0;
  s.next_code[0] = 0;
  for (int d = 1; d <= maxbits; ++d) { code = (code + s.count[d - 1]) << 1; s.next_code[d] = code; }
  for (int i = 0; i < nsym; ++i) {
    const int l = s.len[i];
    out[i] = l ? ((uint32_t)l << 16) | reverse_bits(s.next_code[l]++, l) : 0u;
  }
}

// single-owner bit writer (plain stores into a zeroed word array) for the block header
struct LocalBits {
  uint32_t* words;
  uint32_t nbits;
};
FD_HD void local_put(LocalBits& w, uint32_t v, int n) {
  const uint32_t pos = w.nbits;
  w.words[pos >> 5] |= v << (pos & 31);
  if ((pos & 31) + n > 32) w.words[(pos >> 5) + 1] |= v >> (32 - (pos & 31));
  w.nbits += n;
}

// ------------------------------------------------------------------------------------------------ tokens
// Length code of a match length 3..258 (RFC 1951 §3.2.5): symbol, number of extra bits, extra value.
FD_HD void length_code(int len, int& sym, int& ebits, int& eval) {
  if (len <= 10) { sym = 254 + len; ebits = 0; eval = 0; return; }
  if (len == 258) { sym = 285; ebits = 0; eval = 0; return; }
  int e = 1;
  while (len - 3 >= (8 << e)) ++e;             // len - 3 in [4 << e, 8 << e): four symbols of 2^e lengths each
  const int off = len - 3 - (4 << e);
  sym = 261 + 4 * e + (off >> e);
  ebits = e;
  eval = off & ((1 << e) - 1);
}
FD_HD int length_extra_bits(int sym) {         // extra bits of length symbol 257..285
  return (sym < 265 || sym == 285) ? 0 : (sym - 261) >> 2;
}
// The token that starts at byte i of a thread's run p[0..m): a literal, plus — when p[i] begins kMinRun or more equal
// bytes — one match (distance 1) over the rest of them.  Returns the bytes consumed; match_len = 0 for a bare literal.
// Tokens never look outside the run, so threads tokenise independently and the histogram, the bit count and the emit pass
// see the same tokens.
FD_HD int next_token(const uint8_t* p, int i, int m, int& match_len) {
  const uint8_t v = p[i];
  int r = 1;
  while (i + r < m && p[i + r] == v) ++r;
  if (r >= kMinRun) { match_len = r - 1; return r; }
  match_len = 0;
  return 1;
}

// The code of one deflate block from its token histogram hist[0..286) (literals, end-of-block, length symbols), in three
// steps so that the per-symbol part runs on all threads:
//   1 (one thread)   literal/length code lengths, first code of every length
//   2 (every thread) canonical code of a symbol (RFC 1951 §3.2.2: codes of one length are consecutive in symbol order, so a
//                    symbol's code is the first code of its length + the number of earlier symbols of that length) and the
//                    bits its occurrences take
//   3 (one thread)   the bit-packed dynamic-block header and the block's size
FD_HD void build_lit_lengths(uint32_t* hist, CodeScratch& s) {
  hist[256] = 1;                                           // one end-of-block per block
  uint32_t matches = 0;
  for (int i = 257; i < kSyms; ++i) matches += hist[i];
  s.dist_len = matches ? 1 : 0;
  huffman_lengths(hist, kSyms, kMaxLitBits, s);
  for (int d = 0; d <= kMaxLitBits; ++d) s.count[d] = 0;
  for (int i = 0; i < kSyms; ++i) { s.lit_len[i] = s.len[i]; ++s.count[s.len[i]]; }
  s.count[0] = 0;
  uint32_t code = 0;
  s.lit_next[0] = 0;
  for (int d = 1; d <= kMaxLitBits; ++d) { code = (code + s.count[d - 1]) << 1; s.lit_next[d] = code; }
}
// returns the bits all occurrences of `sym` take in the block (code + extra bits + distance code; end-of-block excluded)
FD_HD unsigned long long build_lit_code(const uint32_t* hist, const CodeScratch& s, uint32_t* codes, int sym) {
  if (sym >= kSyms) { codes[sym] = 0; return 0; }
  const int l = s.lit_len[sym];
  if (l == 0) { codes[sym] = 0; return 0; }
  uint32_t rank = 0;
  for (int j = 0; j < sym; ++j) rank += s.lit_len[j] == l;
  codes[sym] = ((uint32_t)l << 16) | reverse_bits(s.lit_next[l] + rank, l);
  if (sym < 256) return (unsigned long long)hist[sym] * l;
  if (sym == 256) return 0;
  return (unsigned long long)hist[sym] * (l + length_extra_bits(sym) + s.dist_len);
}
FD_HD void build_block_header(CodeScratch& s, unsigned long long data_bits, uint32_t* hdr, BlockMeta* meta) {
  const int eob_len = s.lit_len[256];
  int hlit = kSyms;
  while (hlit > 257 && s.lit_len[hlit - 1] == 0) --hlit;
  for (int i = 0; i < hlit; ++i) s.all_len[i] = s.lit_len[i];
  // one distance code (distance 1), one bit long when the block has matches, absent (zero bits) when it is all literals
  s.all_len[hlit] = (uint8_t)s.dist_len;
  // run-length tokens over the code lengths (§3.2.7: 16 = repeat previous 3-6, 17 = zeros 3-10, 18 = zeros 11-138)
  const int total_lens = hlit + 1;
  int ntok = 0;
  for (int c = 0; c < 19; ++c) s.cl_freq[c] = 0;
  for (int i = 0; i < total_lens;) {
    const int v = s.all_len[i];
    int run = 1;
    while (i + run < total_lens && s.all_len[i + run] == v) ++run;
    i += run;
    if (v == 0) {
      while (run >= 11) { const int t = run > 138 ? 138 : run; s.tok_sym[ntok] = 18; s.tok_extra[ntok++] = (uint8_t)(t - 11); run -= t; }
      if (run >= 3) { s.tok_sym[ntok] = 17; s.tok_extra[ntok++] = (uint8_t)(run - 3); run = 0; }
      while (run-- > 0) { s.tok_sym[ntok] = 0; s.tok_extra[ntok++] = 0; }
    } else {
      s.tok_sym[ntok] = (uint8_t)v; s.tok_extra[ntok++] = 0;
      --run;
      while (run >= 3) { const int t = run > 6 ? 6 : run; s.tok_sym[ntok] = 16; s.tok_extra[ntok++] = (uint8_t)(t - 3); run -= t; }
      while (run-- > 0) { s.tok_sym[ntok] = (uint8_t)v; s.tok_extra[ntok++] = 0; }
    }
  }
  for (int t = 0; t < ntok; ++t) ++s.cl_freq[s.tok_sym[t]];
  huffman_lengths(s.cl_freq, 19, kMaxClBits, s);
  for (int c = 0; c < 19; ++c) s.cl_len[c] = s.len[c];
  canonical_codes(19, kMaxClBits, s, s.cl_code);
  const int cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  int hclen = 19;
  while (hclen > 4 && s.cl_len[cl_order[hclen - 1]] == 0) --hclen;
  for (int w = 0; w < kHdrWords; ++w) hdr[w] = 0;
  LocalBits bits{hdr, 0};
  local_put(bits, 0, 1);                 // BFINAL = 0 (the stream is closed by a separate empty stored block)
  local_put(bits, 2, 2);                 // BTYPE = 10 dynamic Huffman
  local_put(bits, (uint32_t)(hlit - 257), 5);
  local_put(bits, 0, 5);                 // HDIST = 1 - 1
  local_put(bits, (uint32_t)(hclen - 4), 4);
  for (int k = 0; k < hclen; ++k) local_put(bits, s.cl_len[cl_order[k]], 3);
  for (int t = 0; t < ntok; ++t) {
    const int sym = s.tok_sym[t];
    local_put(bits, s.cl_code[sym] & 0xffffu, (int)(s.cl_code[sym] >> 16));
    if (sym == 16) local_put(bits, s.tok_extra[t], 2);
    else if (sym == 17) local_put(bits, s.tok_extra[t], 3);
    else if (sym == 18) local_put(bits, s.tok_extra[t], 7);
  }
  meta->hdr_bits = bits.nbits;
  // header + tokens + end-of-block + stored-block header (3 bits), padded to a byte, + LEN / NLEN of the empty block
  const unsigned long long bits_total = bits.nbits + data_bits + (unsigned)eob_len + 3u;
  meta->payload_bytes = (uint32_t)((bits_total + 7) / 8 + 4);
}

// ------------------------------------------------------------------------------------------------ a block in shared memory
// The 16 KB of a block, copied with coalesced word loads; run r (the 64 bytes thread r tokenises) starts at word r * 17.
struct StagedBlock {
  uint32_t words[kThreads * kRunPadWords];
};
FD_HD void stage_phase_load(const uint8_t* filt, long long n_total, long long block, int tid, StagedBlock& st) {
  const long long base = block * kChunk;
  for (int wi = tid; wi < kChunk / 4; wi += kThreads) {
    const long long off = base + 4ll * wi;
    uint32_t v = 0;
    if (off + 4 <= n_total) v = load_u32(filt + off);
    else
      for (int b = 0; b < 4; ++b)
        if (off + b < n_total) v |= (uint32_t)filt[off + b] << (8 * b);
    st.words[(wi / (kRun / 4)) * kRunPadWords + wi % (kRun / 4)] = v;
  }
}
// thread tid's run: pointer into the staged block and its length (kRun, fewer or none at the end of the stream)
FD_HD const uint8_t* staged_run(const StagedBlock& st, long long n_total, long long block, int tid, int& m) {
  const long long left = n_total - (block * kChunk + (long long)tid * kRun);
  m = left >= kRun ? kRun : (left > 0 ? (int)left : 0);
  return reinterpret_cast<const uint8_t*>(st.words + tid * kRunPadWords);
}

struct StatsShared {
  uint32_t hist[288];
  StagedBlock st;
  CodeScratch scratch;
  unsigned long long part_bits[kThreads];
};
FD_HD void stats_phase_stage(const uint8_t* filt, long long n_total, long long block, int tid, StatsShared& sh) {
  for (int i = tid; i < 288; i += kThreads) sh.hist[i] = 0;
  stage_phase_load(filt, n_total, block, tid, sh.st);
}
FD_HD void stats_phase_hist(long long n_total, long long block, int tid, StatsShared& sh) {
  int m;
  const uint8_t* p = staged_run(sh.st, n_total, block, tid, m);
  for (int i = 0; i < m;) {
    int match_len;
    const int adv = next_token(p, i, m, match_len);
    FD_PNG_ATOMIC_ADD(&sh.hist[p[i]], 1u);
    if (match_len) {
      int sym, ebits, eval;
      length_code(match_len, sym, ebits, eval);
      FD_PNG_ATOMIC_ADD(&sh.hist[sym], 1u);
    }
    i += adv;
  }
}
FD_HD void stats_phase_lengths(int tid, StatsShared& sh) {
  if (tid == 0) build_lit_lengths(sh.hist, sh.scratch);
}
FD_HD void stats_phase_codes(long long block, uint32_t* codes, int tid, StatsShared& sh) {
  unsigned long long bits = build_lit_code(sh.hist, sh.scratch, codes + block * kCodeStride, tid);
  if (tid + kThreads < kCodeStride) bits += build_lit_code(sh.hist, sh.scratch, codes + block * kCodeStride, tid + kThreads);
  sh.part_bits[tid] = bits;
}
FD_HD void stats_phase_header(long long block, uint32_t* hdrs, BlockMeta* meta, int tid, StatsShared& sh) {
  if (tid == 0) {
    unsigned long long data_bits = 0;
    for (int k = 0; k < kThreads; ++k) data_bits += sh.part_bits[k];
    build_block_header(sh.scratch, data_bits, hdrs + block * kHdrWords, meta + block);
  }
}

// ------------------------------------------------------------------------------------------------ CRC-32 / Adler-32
FD_HD uint32_t crc_table_entry(uint32_t n) {
  uint32_t c = n;
  for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
  return c;
}
FD_HD uint32_t crc_update_slow(uint32_t crc, const uint8_t* p, long long n) {     // crc is the running (inverted) register
  for (long long i = 0; i < n; ++i) crc = crc_table_entry((crc ^ p[i]) & 255u) ^ (crc >> 8);
  return crc;
}
FD_HD void put_be32(uint8_t* p, uint32_t v) {
  p[0] = (uint8_t)(v >> 24); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v;
}

// ------------------------------------------------------------------------------------------------ layout (one CTA)
// File offsets: IDAT chunk k starts at idat_off[k] (its 4-byte length field); its payload at idat_off[k] + 8.
// The first IDAT carries the 2-byte zlib header in front of block 0, the last one the final stored block + Adler-32.
// result[0] = file bytes, result[1] = 0 ok / 1 capacity exceeded (then nothing else is written).
struct LayoutShared {
  long long chunk_bytes[kThreads];       // bytes of the IDAT chunks of a thread's slice of blocks, then their exclusive scan
  uint32_t a[kThreads], b[kThreads];     // Adler-32 sums of a thread's slice of rows (as one segment)
  unsigned long long seg[kThreads];      // bytes in that segment
  int ok;
};
FD_HD long long idat_chunk_bytes(const BlockMeta* meta, long long k, long long nblocks) {
  return 12 + (long long)meta[k].payload_bytes + (k == 0 ? 2 : 0) + (k == nblocks - 1 ? 5 + 4 : 0);
}
// phase 1: every thread totals a contiguous slice of blocks and folds a contiguous slice of rows into one Adler segment.
//   Appending a segment of n bytes with sums (a, b) to a prefix with sums (A, B):  B += n * A + b ;  A += a   (mod 65521)
FD_HD void layout_phase_partial(int h, int w, int channels, long long nblocks, const BlockMeta* meta, const uint32_t* adler_rows, int tid,
                                LayoutShared& sh) {
  const long long per = (nblocks + kThreads - 1) / kThreads;
  long long bytes = 0;
  for (long long k = tid * per; k < (tid + 1) * per && k < nblocks; ++k) bytes += idat_chunk_bytes(meta, k, nblocks);
  sh.chunk_bytes[tid] = bytes;
  const unsigned long long n = (unsigned long long)w * channels + 1;
  const int rows_per = (h + kThreads - 1) / kThreads;
  unsigned long long A = 0, B = 0, len = 0;
  for (int y = tid * rows_per; y < (tid + 1) * rows_per && y < h; ++y) {
    B = (B + (n % kAdlerMod) * A + adler_rows[2 * y + 1]) % kAdlerMod;
    A = (A + adler_rows[2 * y]) % kAdlerMod;
    len += n;
  }
  sh.a[tid] = (uint32_t)A; sh.b[tid] = (uint32_t)B; sh.seg[tid] = len;
}
// phase 2: thread 0 scans the 256 partials, checks the capacity and writes the fixed parts of the file
FD_HD void layout_phase_scan(int h, int w, int channels, long long nblocks, uint8_t* out, long long capacity, long long* idat_off,
                             long long* result, int tid, LayoutShared& sh) {
  if (tid != 0) return;
  long long pos = kSigIhdrBytes;
  for (int t = 0; t < kThreads; ++t) { const long long c = sh.chunk_bytes[t]; sh.chunk_bytes[t] = pos; pos += c; }
  const long long total = pos + 12;      // IEND
  result[0] = total;
  result[1] = total > capacity ? 1 : 0;
  sh.ok = total <= capacity;
  if (!sh.ok) return;
  idat_off[nblocks] = pos;
  const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
  for (int i = 0; i < 8; ++i) out[i] = sig[i];
  uint8_t* ih = out + 8;
  put_be32(ih, 13);
  ih[4] = 'I'; ih[5] = 'H'; ih[6] = 'D'; ih[7] = 'R';
  put_be32(ih + 8, (uint32_t)w);
  put_be32(ih + 12, (uint32_t)h);
  ih[16] = 8;                                                            // bit depth
  ih[17] = channels == 1 ? 0 : channels == 2 ? 4 : channels == 3 ? 2 : 6;  // colour type
  ih[18] = 0; ih[19] = 0; ih[20] = 0;                                    // deflate, adaptive filtering, no interlace
  put_be32(ih + 21, ~crc_update_slow(0xffffffffu, ih + 4, 17));
  out[kSigIhdrBytes + 8] = 0x78;                                         // CMF: deflate, 32 KB window
  out[kSigIhdrBytes + 9] = 0x01;                                         // FLG: check bits, no preset dictionary
  unsigned long long A = 1, B = 0;                                       // Adler-32 of the whole filtered stream
  for (int t = 0; t < kThreads; ++t) {
    B = (B + (sh.seg[t] % kAdlerMod) * A + sh.b[t]) % kAdlerMod;
    A = (A + sh.a[t]) % kAdlerMod;
  }
  uint8_t* tail = out + pos - 4 - 9;                                     // in front of the last IDAT's CRC
  tail[0] = 0x01; tail[1] = 0x00; tail[2] = 0x00; tail[3] = 0xff; tail[4] = 0xff;   // BFINAL = 1, stored, LEN = 0
  put_be32(tail + 5, (uint32_t)((B << 16) | A));
  uint8_t* e = out + pos;
  put_be32(e, 0);
  e[4] = 'I'; e[5] = 'E'; e[6] = 'N'; e[7] = 'D';
  put_be32(e + 8, ~crc_update_slow(0xffffffffu, e + 4, 4));
}
// phase 3: every thread writes the offsets and the length / type fields of its slice of IDAT chunks
FD_HD void layout_phase_chunks(long long nblocks, const BlockMeta* meta, uint8_t* out, long long* idat_off, int tid, LayoutShared& sh) {
  if (!sh.ok) return;
  const long long per = (nblocks + kThreads - 1) / kThreads;
  long long pos = sh.chunk_bytes[tid];
  for (long long k = tid * per; k < (tid + 1) * per && k < nblocks; ++k) {
    const long long bytes = idat_chunk_bytes(meta, k, nblocks);
    idat_off[k] = pos;
    uint8_t* c = out + pos;
    put_be32(c, (uint32_t)(bytes - 12));
    c[4] = 'I'; c[5] = 'D'; c[6] = 'A'; c[7] = 'T';
    pos += bytes;
  }
}
FD_HD long long block_payload_offset(const long long* idat_off, long long block) {
  return idat_off[block] + 8 + (block == 0 ? 2 : 0);
}

// ------------------------------------------------------------------------------------------------ bit packing
// Threads of a block (and of neighbouring blocks, at the shared boundary words) OR their bits into the zeroed output.
struct BitWriter {
  uint32_t* words;
  unsigned long long widx;
  unsigned long long acc;
  int nbits;
};
FD_HD void bw_init(BitWriter& w, uint32_t* words, unsigned long long bitpos) {
  w.words = words; w.widx = bitpos >> 5; w.nbits = (int)(bitpos & 31); w.acc = 0;
}
FD_HD void bw_put(BitWriter& w, uint32_t v, int n) {          // n <= 32, v < 2^n
  w.acc |= (unsigned long long)v << w.nbits;
  w.nbits += n;
  if (w.nbits >= 32) {
    const uint32_t lo = (uint32_t)w.acc;
    if (lo) FD_PNG_ATOMIC_OR(&w.words[w.widx], lo);
    ++w.widx;
    w.acc >>= 32;
    w.nbits -= 32;
  }
}
FD_HD void bw_flush(BitWriter& w) {
  const uint32_t lo = (uint32_t)w.acc;
  if (w.nbits > 0 && lo) FD_PNG_ATOMIC_OR(&w.words[w.widx], lo);
  w.acc = 0; w.nbits = 0;
}

constexpr int kScanGroup = 16;                 // threads per group of the two-level scan of the bit counts
struct EncodeShared {
  uint32_t codes[kCodeStride];
  uint32_t cnt[kThreads];
  uint32_t group[kThreads / kScanGroup];
  StagedBlock st;
};
FD_HD void encode_phase_load(const uint8_t* filt, long long n_total, const uint32_t* codes, long long block, int tid, EncodeShared& sh) {
  for (int i = tid; i < kCodeStride; i += kThreads) sh.codes[i] = codes[block * kCodeStride + i];
  stage_phase_load(filt, n_total, block, tid, sh.st);
}
// bits of the tokens of this thread's run
FD_HD void encode_phase_count(long long n_total, long long block, int tid, EncodeShared& sh) {
  int m;
  const uint8_t* p = staged_run(sh.st, n_total, block, tid, m);
  uint32_t bits = 0;
  for (int i = 0; i < m;) {
    int match_len;
    const int adv = next_token(p, i, m, match_len);
    bits += sh.codes[p[i]] >> 16;
    if (match_len) {
      int sym, ebits, eval;
      length_code(match_len, sym, ebits, eval);
      bits += (sh.codes[sym] >> 16) + ebits + 1;       // + the one-bit distance code
    }
    i += adv;
  }
  sh.cnt[tid] = bits;
}
// totals of groups of 16 threads (first level of the scan)
FD_HD void encode_phase_groups(int tid, EncodeShared& sh) {
  if (tid < kThreads / kScanGroup) {
    uint32_t t = 0;
    for (int k = 0; k < kScanGroup; ++k) t += sh.cnt[tid * kScanGroup + k];
    sh.group[tid] = t;
  }
}
// thread 0 also writes the block header, the last thread the end-of-block code and the byte-aligning stored block
FD_HD void encode_phase_emit(long long n_total, long long block, const uint32_t* hdrs, const BlockMeta* meta,
                             const long long* idat_off, uint32_t* out_words, int tid, EncodeShared& sh) {
  const unsigned long long base_bit = 8ull * (unsigned long long)block_payload_offset(idat_off, block);
  const uint32_t hdr_bits = meta[block].hdr_bits;
  unsigned long long before = 0;
  for (int g = 0; g < tid / kScanGroup; ++g) before += sh.group[g];
  for (int k = (tid / kScanGroup) * kScanGroup; k < tid; ++k) before += sh.cnt[k];
  BitWriter bw;
  if (tid == 0) {
    bw_init(bw, out_words, base_bit);
    const uint32_t* hdr = hdrs + block * kHdrWords;
    uint32_t left = hdr_bits;
    for (int wd = 0; left > 0; ++wd) {
      const int n = left >= 32 ? 32 : (int)left;
      bw_put(bw, n == 32 ? hdr[wd] : (hdr[wd] & ((1u << n) - 1u)), n);
      left -= n;
    }
    bw_flush(bw);
  }
  bw_init(bw, out_words, base_bit + hdr_bits + before);
  int m;
  const uint8_t* p = staged_run(sh.st, n_total, block, tid, m);
  for (int i = 0; i < m;) {
    int match_len;
    const int adv = next_token(p, i, m, match_len);
    const uint32_t c = sh.codes[p[i]];
    bw_put(bw, c & 0xffffu, (int)(c >> 16));
    if (match_len) {
      int sym, ebits, eval;
      length_code(match_len, sym, ebits, eval);
      const uint32_t lc = sh.codes[sym];
      bw_put(bw, lc & 0xffffu, (int)(lc >> 16));
      bw_put(bw, (uint32_t)eval, ebits + 1);           // extra bits, then the distance code: a single 0 bit
    }
    i += adv;
  }
  if (tid == kThreads - 1) {
    const uint32_t eob = sh.codes[256];
    bw_put(bw, eob & 0xffffu, (int)(eob >> 16));
    bw_put(bw, 0, 3);                                    // BFINAL = 0, BTYPE = 00 (stored)
    const unsigned long long used = hdr_bits + before + sh.cnt[tid] + (eob >> 16) + 3;
    const int pad = (int)((8 - (used & 7)) & 7);
    if (pad) bw_put(bw, 0, pad);
    bw_put(bw, 0xffff0000u, 32);                         // LEN = 0, NLEN = 0xffff
  }
  bw_flush(bw);
}

// ------------------------------------------------------------------------------------------------ IDAT CRCs
// One CTA per IDAT chunk.  CRC-32 is linear over GF(2): for a message cut into pieces m_0 .. m_k,
//   crc(m_0 .. m_k) = XOR_i  x^(8 * bytes after piece i) * crc(m_i)   (mod P; the identity zlib's crc32_combine uses),
// so every thread takes the CRC of its own piece with the byte table, multiplies it by the power of x that stands for the
// bytes behind it (square-and-multiply over x^(8 * 2^k), 32-step carry-less multiplications), and one thread XORs the parts.
// Polynomials in the reflected representation of the CRC register: bit 31 = x^0, bit 0 = x^31.
FD_HD uint32_t gf2_mulmod(uint32_t a, uint32_t b) {
  uint32_t p = 0;
  for (int i = 0; i < 32; ++i) {
    if (a & (0x80000000u >> i)) p ^= b;
    b = (b & 1u) ? (b >> 1) ^ 0xedb88320u : b >> 1;       // b *= x
  }
  return p;
}
constexpr int kCrcPowers = 20;                 // an IDAT chunk (at most 9/8 of 16 KB + header) is far below 2^20 bytes
struct CrcShared {
  uint32_t table[256];
  uint32_t pow8[kCrcPowers];     // x^(8 * 2^k) mod P
  uint32_t part[kThreads];
};
FD_HD void crc_phase_tables(int tid, CrcShared& sh) {
  for (int i = tid; i < 256; i += kThreads) sh.table[i] = crc_table_entry((uint32_t)i);
  if (tid < kCrcPowers) {
    uint32_t v = 0x40000000u;                              // x^1
    for (int k = 0; k < tid + 3; ++k) v = gf2_mulmod(v, v);
    sh.pow8[tid] = v;
  }
}
// CRC range of a chunk: its type + payload, n bytes from out + idat_off[block] + 4
FD_HD void crc_phase_pieces(const uint8_t* out, const long long* idat_off, long long block, int tid, CrcShared& sh) {
  const uint8_t* p = out + idat_off[block] + 4;
  const long long n = idat_off[block + 1] - idat_off[block] - 8;
  const long long piece = (n + kThreads - 1) / kThreads;
  const long long begin = tid * piece;
  long long end = begin + piece;
  if (end > n) end = n;
  uint32_t part = 0;
  if (begin < end) {
    uint32_t crc = 0xffffffffu;
    for (long long i = begin; i < end; ++i) crc = sh.table[(crc ^ p[i]) & 255u] ^ (crc >> 8);
    crc = ~crc;
    unsigned long long behind = (unsigned long long)(n - end);
    uint32_t shift = 0x80000000u;                          // x^0
    for (int k = 0; behind; ++k, behind >>= 1)
      if (behind & 1) shift = gf2_mulmod(shift, sh.pow8[k]);
    part = gf2_mulmod(shift, crc);
  }
  sh.part[tid] = part;
}
FD_HD void crc_phase_store(uint8_t* out, const long long* idat_off, long long block, int tid, const CrcShared& sh) {
  if (tid == 0) {
    uint32_t crc = 0;
    for (int k = 0; k < kThreads; ++k) crc ^= sh.part[k];
    put_be32(out + idat_off[block + 1] - 4, crc);
  }
}

FD_HOST_HD long long num_blocks(long long n_total) { return (n_total + kChunk - 1) / kChunk; }
// generous bound on the file size: a block's literals cost at most 9 bits per byte under its own optimal code
FD_HOST_HD long long file_capacity(long long n_total) {
  return kSigIhdrBytes + 12 + n_total + n_total / 8 + num_blocks(n_total) * (12 + kHdrWords * 4 + 16) + 64;
}

}  // namespace png
}  // namespace fd
