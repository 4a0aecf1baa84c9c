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
      for (int t = 0; t < kThreads; ++t) filter_phase_sums(image, stride, y, rb, channels, t, *sh);
      for (int t = 0; t < kThreads; ++t) filter_phase_total(t, *sh);
      for (int t = 0; t < kThreads; ++t) filter_phase_choose(t, *sh);
      for (int t = 0; t < kThreads; ++t) filter_phase_write(image, stride, y, rb, channels, filt.data(), t, *sh);
      for (int t = 0; t < kThreads; ++t) filter_phase_adler(y, adler_rows.data(), t, *sh);
    }
    delete sh;
  }
  {
    StatsShared* sh = new StatsShared;
    for (long long b = 0; b < nb; ++b) {
      for (int t = 0; t < kThreads; ++t) stats_phase_stage(filt.data(), n_total, b, t, *sh);
      for (int t = 0; t < kThreads; ++t) stats_phase_hist(n_total, b, t, *sh);
      for (int t = 0; t < kThreads; ++t) stats_phase_lengths(t, *sh);
      for (int t = 0; t < kThreads; ++t) stats_phase_codes(b, codes.data(), t, *sh);
      for (int t = 0; t < kThreads; ++t) stats_phase_header(b, hdrs.data(), meta.data(), t, *sh);
    }
    delete sh;
  }
  {
    LayoutShared* sh = new LayoutShared;
    for (int t = 0; t < kThreads; ++t) layout_phase_partial(h, w, channels, nb, meta.data(), adler_rows.data(), t, *sh);
    for (int t = 0; t < kThreads; ++t) layout_phase_scan(h, w, channels, nb, out, capacity, idat_off.data(), result, t, *sh);
    for (int t = 0; t < kThreads; ++t) layout_phase_chunks(nb, meta.data(), out, idat_off.data(), t, *sh);
    delete sh;
  }
  *file_bytes = result[0];
  if (result[1]) return 1;
  {
    EncodeShared* sh = new EncodeShared;
    for (long long b = 0; b < nb; ++b) {
      for (int t = 0; t < kThreads; ++t) encode_phase_load(filt.data(), n_total, codes.data(), b, t, *sh);
      for (int t = 0; t < kThreads; ++t) encode_phase_count(n_total, b, t, *sh);
      for (int t = 0; t < kThreads; ++t) encode_phase_groups(t, *sh);
      for (int t = 0; t < kThreads; ++t)
        encode_phase_emit(n_total, b, hdrs.data(), meta.data(), idat_off.data(), reinterpret_cast<uint32_t*>(out), t, *sh);
    }
    delete sh;
  }
  {
    CrcShared* sh = new CrcShared;
    for (long long b = 0; b < nb; ++b) {
      for (int t = 0; t < kThreads; ++t) crc_phase_tables(t, *sh);
      for (int t = 0; t < kThreads; ++t) crc_phase_pieces(out, idat_off.data(), b, t, *sh);
      for (int t = 0; t < kThreads; ++t) crc_phase_store(out, idat_off.data(), b, t, *sh);
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
