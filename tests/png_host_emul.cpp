// TEST INFRASTRUCTURE (never linked into the product): runs the per-thread phase functions of the device PNG encoder
// (faithdiff_dev_b200/csrc/png_core.cuh) on the host, block by block and thread by thread, in exactly the order the kernels
// of csrc/png.cu call them (a __syncthreads() there = the end of a `for tid` loop here).  Built with g++ by
// tests/test_cpu_png.py; its output is checked against zlib / PIL on the CPU and is the byte-exact expectation of
// the GPU test.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../faithdiff_dev_b200/csrc/png_core.cuh"

using namespace fd::png;

// Order in which the threads of a block run inside one phase: 0 ascending, 1 descending, 2 a fresh pseudo-random permutation
// per phase.  On the device the order is arbitrary, so a file that depends on it would mean two threads of one phase touch the
// same shared word (a missing __syncthreads()); the tests require identical bytes for all three.
static int g_order_mode = 0;
static unsigned g_order_state = 1;
extern "C" void emul_set_thread_order(int mode, unsigned seed) { g_order_mode = mode; g_order_state = seed ? seed : 1; }
static const int* thread_order(int n) {
  static int perm[kThreads];
  for (int i = 0; i < n; ++i) perm[i] = g_order_mode == 1 ? n - 1 - i : i;
  if (g_order_mode == 2)
    for (int i = n - 1; i > 0; --i) {
      g_order_state = g_order_state * 1664525u + 1013904223u;
      const int j = (int)((g_order_state >> 8) % (unsigned)(i + 1));
      const int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
    }
  return perm;
}
#define FOR_THREADS(t, n) for (const int *ord_ = thread_order(n), *it_ = ord_; it_ < ord_ + (n); ++it_) for (int t = *it_, once_ = 1; once_; once_ = 0)

extern "C" long long emul_png_file_capacity(int h, int w, int channels) {
  const long long cap = file_capacity((long long)h * ((long long)w * channels + 1));
  return (cap + 3) & ~3ll;
}

// returns 0 ok, 1 capacity exceeded; *file_bytes = size of the PNG in `out`; `filtered_out` (optional) gets the filtered stream
extern "C" int emul_png_encode(const unsigned char* image, int h, int w, int channels, long long stride, unsigned char* out,
                               long long capacity, long long* file_bytes, unsigned char* filtered_out) {
  const int rb = w * channels;
  const long long n_total = (long long)h * (rb + 1);
  const long long nb = num_blocks(n_total);
  std::vector<uint8_t> filt(n_total);
  std::vector<uint32_t> adler_rows(2 * (size_t)h), codes(nb * kCodeStride), hdrs(nb * kHdrWords);
  std::vector<BlockMeta> meta(nb);
  std::vector<long long> idat_off(nb + 1);
  long long result[2] = {0, 0};
  memset(out, 0, capacity);
  {
    FilterShared* sh = new FilterShared;
    for (int y = 0; y < h; ++y) {
      FOR_THREADS(t, kThreads) filter_phase_sums(image, stride, y, rb, channels, t, *sh);
      FOR_THREADS(t, kThreads) filter_phase_total(t, *sh);
      FOR_THREADS(t, kThreads) filter_phase_choose(t, *sh);
      FOR_THREADS(t, kThreads) filter_phase_write(image, stride, y, rb, channels, filt.data(), t, *sh);
      FOR_THREADS(t, kThreads) filter_phase_adler(y, adler_rows.data(), t, *sh);
    }
    delete sh;
  }
  {
    StatsShared* sh = new StatsShared;
    for (long long b = 0; b < nb; ++b) {
      FOR_THREADS(t, kThreads) stats_phase_stage(filt.data(), n_total, b, t, *sh);
      FOR_THREADS(t, kThreads) stats_phase_hist(n_total, b, t, *sh);
      FOR_THREADS(t, kThreads) stats_phase_lengths(t, *sh);
      FOR_THREADS(t, kThreads) stats_phase_codes(b, codes.data(), t, *sh);
      FOR_THREADS(t, kThreads) stats_phase_header(b, hdrs.data(), meta.data(), t, *sh);
    }
    delete sh;
  }
  {
    LayoutShared* sh = new LayoutShared;
    FOR_THREADS(t, kThreads) layout_phase_partial(h, w, channels, nb, meta.data(), adler_rows.data(), t, *sh);
    FOR_THREADS(t, kThreads) layout_phase_scan(h, w, channels, nb, out, capacity, idat_off.data(), result, t, *sh);
    FOR_THREADS(t, kThreads) layout_phase_chunks(nb, meta.data(), out, idat_off.data(), t, *sh);
    delete sh;
  }
  *file_bytes = result[0];
  if (result[1]) return 1;
  {
    EncodeShared* sh = new EncodeShared;
    for (long long b = 0; b < nb; ++b) {
      FOR_THREADS(t, kThreads) encode_phase_load(filt.data(), n_total, codes.data(), b, t, *sh);
      FOR_THREADS(t, kThreads) encode_phase_count(n_total, b, t, *sh);
      FOR_THREADS(t, kThreads) encode_phase_groups(t, *sh);
      FOR_THREADS(t, kThreads)
        encode_phase_emit(n_total, b, hdrs.data(), meta.data(), idat_off.data(), reinterpret_cast<uint32_t*>(out), t, *sh);
    }
    delete sh;
  }
  {
    CrcShared* sh = new CrcShared;
    for (long long b = 0; b < nb; ++b) {
      FOR_THREADS(t, kThreads) crc_phase_tables(t, *sh);
      FOR_THREADS(t, kThreads) crc_phase_pieces(out, idat_off.data(), b, t, *sh);
      FOR_THREADS(t, kThreads) crc_phase_store(out, idat_off.data(), b, t, *sh);
    }
    delete sh;
  }
  if (filtered_out) memcpy(filtered_out, filt.data(), n_total);
  return 0;
}

// Huffman construction alone, for the length-limit test: lens[nsym] out
extern "C" void emul_huffman_lengths(const unsigned* freq, int nsym, int maxbits, unsigned char* lens) {
  CodeScratch* s = new CodeScratch;
  huffman_lengths(freq, nsym, maxbits, *s);
  for (int i = 0; i < nsym; ++i) lens[i] = s->len[i];
  delete s;
}
