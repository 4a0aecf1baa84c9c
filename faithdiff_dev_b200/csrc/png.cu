// PNG encoder on the device: the caller's `out_image.save(path)` (test_wo_llava.py:97, test.py:130) without shipping the
// raw 24.9 MB 4K image over PCIe and through the host's zlib (1-2 s per 4K image with PIL: more than a whole restoration
// on 8 GPUs).  Five launches on one stream; the per-thread logic and the file format are in png_core.cuh.
//   filter  : one CTA per scanline   — adaptive filter choice, filtered row, per-row Adler-32 partials
//   stats   : one CTA per 16 KB block — block staged in shared memory, token histogram, the block's Huffman code + header
//   layout  : one CTA                — prefix sum of the block sizes, Adler-32 combine, PNG framing, zlib header / trailer
//   encode  : one CTA per block      — every thread packs 64 bytes; bit offsets from a two-level scan of the bit counts
//   crc     : one CTA per IDAT chunk — per-thread piece CRCs combined through GF(2) multiplications by powers of x
#include "common.cuh"
#include "png_core.cuh"
#include "../../include/faithdiff_b200.h"

namespace fd {
extern int64_t g_launches;
using namespace png;

__global__ void __launch_bounds__(kThreads) png_filter_kernel(const uint8_t* __restrict__ img, long long stride, int rb, int bpp,
                                                             uint8_t* __restrict__ filt, uint32_t* __restrict__ adler_rows) {
  __shared__ FilterShared sh;
  const int y = blockIdx.x, tid = threadIdx.x;
  filter_phase_sums(img, stride, y, rb, bpp, tid, sh);
  __syncthreads();
  filter_phase_total(tid, sh);
  __syncthreads();
  filter_phase_choose(tid, sh);
  __syncthreads();
  filter_phase_write(img, stride, y, rb, bpp, filt, tid, sh);
  __syncthreads();
  filter_phase_adler(y, adler_rows, tid, sh);
}

__global__ void __launch_bounds__(kThreads) png_stats_kernel(const uint8_t* __restrict__ filt, long long n_total, uint32_t* __restrict__ codes,
                                                            uint32_t* __restrict__ hdrs, BlockMeta* __restrict__ meta) {
  __shared__ StatsShared sh;
  const long long block = blockIdx.x;
  const int tid = threadIdx.x;
  stats_phase_stage(filt, n_total, block, tid, sh);
  __syncthreads();
  stats_phase_hist(n_total, block, tid, sh);
  __syncthreads();
  stats_phase_lengths(tid, sh);
  __syncthreads();
  stats_phase_codes(block, codes, tid, sh);
  __syncthreads();
  stats_phase_header(block, hdrs, meta, tid, sh);
}

__global__ void __launch_bounds__(kThreads) png_layout_kernel(int h, int w, int channels, long long nblocks, const BlockMeta* __restrict__ meta,
                                                             const uint32_t* __restrict__ adler_rows, uint8_t* __restrict__ out, long long capacity,
                                                             long long* __restrict__ idat_off, long long* __restrict__ result) {
  __shared__ LayoutShared sh;
  const int tid = threadIdx.x;
  layout_phase_partial(h, w, channels, nblocks, meta, adler_rows, tid, sh);
  __syncthreads();
  layout_phase_scan(h, w, channels, nblocks, out, capacity, idat_off, result, tid, sh);
  __syncthreads();
  layout_phase_chunks(nblocks, meta, out, idat_off, tid, sh);
}

__global__ void __launch_bounds__(kThreads) png_encode_kernel(const uint8_t* __restrict__ filt, long long n_total, const uint32_t* __restrict__ codes,
                                                             const uint32_t* __restrict__ hdrs, const BlockMeta* __restrict__ meta,
                                                             const long long* __restrict__ idat_off, const long long* __restrict__ result,
                                                             uint32_t* __restrict__ out_words) {
  __shared__ EncodeShared sh;
  if (result[1] != 0) return;            // capacity exceeded: the layout kernel wrote nothing, neither do we
  const long long block = blockIdx.x;
  const int tid = threadIdx.x;
  encode_phase_load(filt, n_total, codes, block, tid, sh);
  __syncthreads();
  encode_phase_count(n_total, block, tid, sh);
  __syncthreads();
  encode_phase_groups(tid, sh);
  __syncthreads();
  encode_phase_emit(n_total, block, hdrs, meta, idat_off, out_words, tid, sh);
}

__global__ void __launch_bounds__(kThreads) png_crc_kernel(uint8_t* __restrict__ out, const long long* __restrict__ idat_off,
                                                          const long long* __restrict__ result) {
  __shared__ CrcShared sh;
  if (result[1] != 0) return;
  const long long block = blockIdx.x;
  const int tid = threadIdx.x;
  crc_phase_tables(tid, sh);
  __syncthreads();
  crc_phase_pieces(out, idat_off, block, tid, sh);
  __syncthreads();
  crc_phase_store(out, idat_off, block, tid, sh);
}

// workspace layout (32-bit words): adler_rows [2h] | codes [nb * kCodeStride] | hdrs [nb * kHdrWords] | meta [nb * 2] | pad | idat_off [nb + 1] (64-bit)
static long long work_words(long long h, long long nb) {
  long long wds = 2 * h + nb * kCodeStride + nb * kHdrWords + nb * 2;
  wds += wds & 1;
  return wds + 2 * (nb + 1);
}
}  // namespace fd

extern "C" int64_t fd_png_filtered_bytes(int32_t h, int32_t w, int32_t channels) {
  return (int64_t)h * ((int64_t)w * channels + 1);
}
extern "C" int64_t fd_png_work_words(int32_t h, int32_t w, int32_t channels) {
  return fd::work_words(h, fd::png::num_blocks(fd_png_filtered_bytes(h, w, channels)));
}
extern "C" int64_t fd_png_file_capacity(int32_t h, int32_t w, int32_t channels) {
  const int64_t cap = fd::png::file_capacity(fd_png_filtered_bytes(h, w, channels));
  return (cap + 3) & ~(int64_t)3;
}

extern "C" int fd_png_encode(const uint8_t* image, int32_t h, int32_t w, int32_t channels, int64_t row_stride_bytes, uint8_t* filtered,
                             uint32_t* work, uint8_t* out, int64_t out_capacity, int64_t* result, void* stream_) {
  using namespace fd;
  using namespace fd::png;
  FD_REQUIRE(image && filtered && work && out && result, "fd_png_encode: null pointer");
  FD_REQUIRE(h > 0 && w > 0 && channels >= 1 && channels <= 4, "fd_png_encode: bad shape %d x %d x %d", h, w, channels);
  FD_REQUIRE(row_stride_bytes >= (int64_t)w * channels, "fd_png_encode: row stride %lld < row bytes", (long long)row_stride_bytes);
  FD_REQUIRE(((uintptr_t)out % 4) == 0 && ((uintptr_t)filtered % 4) == 0 && ((uintptr_t)work % 8) == 0 && ((uintptr_t)result % 8) == 0 && out_capacity % 4 == 0,
             "fd_png_encode: out / filtered must be 4-byte aligned, the capacity a multiple of 4, work / result 8-byte aligned");
  const long long n_total = fd_png_filtered_bytes(h, w, channels);
  FD_REQUIRE(n_total < (1ll << 31), "fd_png_encode: image too large (%lld filtered bytes)", n_total);
  const long long nb = num_blocks(n_total);
  cudaStream_t st = (cudaStream_t)stream_;
  uint32_t* adler_rows = work;
  uint32_t* codes = adler_rows + 2ll * h;
  uint32_t* hdrs = codes + nb * kCodeStride;
  BlockMeta* meta = reinterpret_cast<BlockMeta*>(hdrs + nb * kHdrWords);
  long long* idat_off = reinterpret_cast<long long*>(work + (work_words(h, nb) - 2 * (nb + 1)));
  static_assert(sizeof(BlockMeta) == 8 && sizeof(long long) == sizeof(int64_t), "workspace layout");
  FD_CHECK_CUDA(cudaMemsetAsync(out, 0, (size_t)out_capacity, st));
  png_filter_kernel<<<(unsigned)h, kThreads, 0, st>>>(image, (long long)row_stride_bytes, w * channels, channels, filtered, adler_rows);
  png_stats_kernel<<<(unsigned)nb, kThreads, 0, st>>>(filtered, n_total, codes, hdrs, meta);
  png_layout_kernel<<<1, kThreads, 0, st>>>(h, w, channels, nb, meta, adler_rows, out, (long long)out_capacity, idat_off,
                                      reinterpret_cast<long long*>(result));
  png_encode_kernel<<<(unsigned)nb, kThreads, 0, st>>>(filtered, n_total, codes, hdrs, meta, idat_off,
                                                       reinterpret_cast<const long long*>(result), reinterpret_cast<uint32_t*>(out));
  png_crc_kernel<<<(unsigned)nb, kThreads, 0, st>>>(out, idat_off, reinterpret_cast<const long long*>(result));
  g_launches += 5;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}
