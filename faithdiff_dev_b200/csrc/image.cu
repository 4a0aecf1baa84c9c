// Image pre/post-processing and the colour fix (utils/color_fix.py) on the GPU.
//   preprocess : VaeImageProcessor.preprocess   (pipeline_FaithDiff_tlc.py:1246)  u8 HWC -> [-1,1]
//   postprocess: VaeImageProcessor.postprocess  (pipeline_FaithDiff_tlc.py:1451)  -> u8 HWC
//   AdaIN      : utils/color_fix.py:14-27, :44-71  (unbiased variance + 1e-5)
//   wavelet    : utils/color_fix.py:29-42, :73-118 (5 levels of replicate-padded dilated [1 2 1]^2/16 blur)
#include "common.cuh"
#include "../../include/faithdiff_b200.h"

namespace fd {
extern int64_t g_launches;

__global__ void preprocess_kernel(const uint8_t* __restrict__ rgb, __nv_bfloat16* __restrict__ out_nhwc, float* __restrict__ out_nchw,
                                  long long hw) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw * 3; i += (long long)gridDim.x * blockDim.x) {
    float v = __fdiv_rn((float)rgb[i], 255.f);
    v = __fsub_rn(__fmul_rn(2.f, v), 1.f);
    if (out_nhwc) out_nhwc[i] = __float2bfloat16_rn(v);
    if (out_nchw) out_nchw[(i % 3) * hw + i / 3] = v;
  }
}

// x: NHWC bf16 [h,w,3] (mode 1) or NCHW fp32 [3,h,w] (mode 0) -> u8 HWC
__global__ void postprocess_kernel(const void* __restrict__ x, int nhwc_bf16, uint8_t* __restrict__ rgb, long long hw) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw * 3; i += (long long)gridDim.x * blockDim.x) {
    float t;
    if (nhwc_bf16) {
      // bf16 tensor arithmetic: each op rounds to bf16
      t = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[i]);
      t = __bfloat162float(__float2bfloat16_rn(__fmul_rn(t, 0.5f)));
      t = __bfloat162float(__float2bfloat16_rn(__fadd_rn(t, 0.5f)));
    } else {
      t = reinterpret_cast<const float*>(x)[(i % 3) * hw + i / 3];
      t = __fadd_rn(__fmul_rn(t, 0.5f), 0.5f);
    }
    t = fminf(fmaxf(t, 0.f), 1.f);
    rgb[i] = (uint8_t)rintf(__fmul_rn(t, 255.f));
  }
}

// ------------------------------------------------------------------ AdaIN
// stats[img][ch][0..1] = (sum, sumsq) in double
__global__ void adain_stats_kernel(const uint8_t* __restrict__ tgt, const uint8_t* __restrict__ src, long long hw, double* __restrict__ stats) {
  __shared__ double red[12][8];
  double acc[12];
#pragma unroll
  for (int j = 0; j < 12; ++j) acc[j] = 0.0;
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < hw; p += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const float a = __fdiv_rn((float)tgt[p * 3 + ch], 255.f);
      const float b = __fdiv_rn((float)src[p * 3 + ch], 255.f);
      acc[ch * 2] += a; acc[ch * 2 + 1] += (double)a * a;
      acc[6 + ch * 2] += b; acc[6 + ch * 2 + 1] += (double)b * b;
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 12; ++j) {
    double v = acc[j];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[j][wid] = v;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    double v = 0.0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) v += red[threadIdx.x][i];
    atomicAdd(stats + threadIdx.x, v);
  }
}
__global__ void adain_apply_kernel(const uint8_t* __restrict__ tgt, uint8_t* __restrict__ out, long long hw, const double* __restrict__ stats) {
  __shared__ float cm[3], cs[3], sm_[3], ss[3];
  if (threadIdx.x < 6) {
    const int img = threadIdx.x / 3, ch = threadIdx.x % 3;
    const double s = stats[img * 6 + ch * 2], q = stats[img * 6 + ch * 2 + 1];
    const double n = (double)hw;
    const double mean = s / n;
    const double var = (q - s * mean) / (n - 1.0) + 1e-5;
    if (img == 0) { cm[ch] = (float)mean; cs[ch] = (float)sqrt(var); } else { sm_[ch] = (float)mean; ss[ch] = (float)sqrt(var); }
  }
  __syncthreads();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw * 3; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i % 3);
    const float x = __fdiv_rn((float)tgt[i], 255.f);
    float r = __fdiv_rn(__fsub_rn(x, cm[ch]), cs[ch]);
    r = __fadd_rn(__fmul_rn(r, ss[ch]), sm_[ch]);
    r = fminf(fmaxf(r, 0.f), 1.f);
    out[i] = (uint8_t)(__fmul_rn(r, 255.f));   // ToPILImage: mul(255).byte() truncates
  }
}

// ------------------------------------------------------------------ wavelet
// planes are fp32 CHW per image: buf[img][ch][h][w]
__global__ void wavelet_init_kernel(const uint8_t* __restrict__ tgt, const uint8_t* __restrict__ src, float* __restrict__ cur,
                                    float* __restrict__ high, long long hw) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw * 3; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / 3;
    const int ch = (int)(i % 3);
    cur[ch * hw + p] = __fdiv_rn((float)tgt[i], 255.f);
    cur[(3 + ch) * hw + p] = __fdiv_rn((float)src[i], 255.f);
    high[ch * hw + p] = 0.f;
  }
}
// one level for both images (blockIdx.z = img*3 + ch): low = blur(cur); content image also accumulates high += cur - low
__global__ void wavelet_level_kernel(const float* __restrict__ cur, float* __restrict__ nxt, float* __restrict__ high, int h, int w, int radius) {
  const int plane = blockIdx.z;
  const long long hw = (long long)h * w;
  const float* in = cur + plane * hw;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= w || y >= h) return;
  const int ym = max(y - radius, 0), yp = min(y + radius, h - 1);
  const int xm = max(x - radius, 0), xp = min(x + radius, w - 1);
  // same accumulation order as a direct 3x3 correlation (row-major taps)
  float acc = 0.f;
  acc = __fmaf_rn(in[(long long)ym * w + xm], 0.0625f, acc);
  acc = __fmaf_rn(in[(long long)ym * w + x], 0.125f, acc);
  acc = __fmaf_rn(in[(long long)ym * w + xp], 0.0625f, acc);
  acc = __fmaf_rn(in[(long long)y * w + xm], 0.125f, acc);
  acc = __fmaf_rn(in[(long long)y * w + x], 0.25f, acc);
  acc = __fmaf_rn(in[(long long)y * w + xp], 0.125f, acc);
  acc = __fmaf_rn(in[(long long)yp * w + xm], 0.0625f, acc);
  acc = __fmaf_rn(in[(long long)yp * w + x], 0.125f, acc);
  acc = __fmaf_rn(in[(long long)yp * w + xp], 0.0625f, acc);
  const long long o = (long long)y * w + x;
  nxt[plane * hw + o] = acc;
  if (plane < 3) high[plane * hw + o] = __fadd_rn(high[plane * hw + o], __fsub_rn(in[o], acc));
}
__global__ void wavelet_final_kernel(const float* __restrict__ high, const float* __restrict__ style_low, uint8_t* __restrict__ out, long long hw) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hw * 3; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / 3;
    const int ch = (int)(i % 3);
    float r = __fadd_rn(high[ch * hw + p], style_low[ch * hw + p]);
    r = fminf(fmaxf(r, 0.f), 1.f);
    out[i] = (uint8_t)(__fmul_rn(r, 255.f));
  }
}
// ------------------------------------------------------------------ Pillow-compatible resample (LANCZOS upscale of the caller)
// PIL's ImagingResample for 8-bit images (what `image.resize((w, h), Image.LANCZOS)` at test_wo_llava.py:81 runs): a horizontal
// then a vertical pass, taps normalised and rounded to 22-bit fixed point on the host (image_process.resample_taps), int32
// accumulation started at 1 << 21, arithmetic shift, clamp to u8, and a u8 intermediate image between the passes.  Integer work:
// the result is bit-identical to PIL.  bounds[2*i] = first source index, bounds[2*i+1] = tap count of output index i.
constexpr int kResampleBits = 32 - 8 - 2;
__device__ __forceinline__ uint8_t resample_clip8(int v) {
  v >>= kResampleBits;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

template <int CH>
__global__ void resample_h_kernel(const uint8_t* __restrict__ in, int in_w, long long rows, uint8_t* __restrict__ out, int out_w,
                                  const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  const long long total = rows * out_w;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int xx = (int)(i % out_w);
    const long long y = i / out_w;
    const int xmin = __ldg(bounds + 2 * xx), n = __ldg(bounds + 2 * xx + 1);
    const int* k = kk + (long long)xx * ksize;
    const uint8_t* src = in + (y * in_w + xmin) * CH;
    int acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = 1 << (kResampleBits - 1);
    for (int x = 0; x < n; ++x) {
      const int w = __ldg(k + x);
#pragma unroll
      for (int c = 0; c < CH; ++c) acc[c] += (int)src[x * CH + c] * w;
    }
    uint8_t* dst = out + i * CH;
#pragma unroll
    for (int c = 0; c < CH; ++c) dst[c] = resample_clip8(acc[c]);
  }
}

// the vertical pass does not see channels: every byte column of a row is filtered alike; VEC bytes per thread
template <int VEC>
__global__ void resample_v_kernel(const uint8_t* __restrict__ in, long long row_bytes, uint8_t* __restrict__ out, int out_h,
                                  const int* __restrict__ bounds, const int* __restrict__ kk, int ksize) {
  const long long per_row = row_bytes / VEC;
  const long long total = per_row * out_h;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int yy = (int)(i / per_row);
    const long long col = (i - (long long)yy * per_row) * VEC;
    const int ymin = __ldg(bounds + 2 * yy), n = __ldg(bounds + 2 * yy + 1);
    const int* k = kk + (long long)yy * ksize;
    int acc[VEC];
#pragma unroll
    for (int c = 0; c < VEC; ++c) acc[c] = 1 << (kResampleBits - 1);
    for (int y = 0; y < n; ++y) {
      const int w = __ldg(k + y);
      const uint8_t* src = in + (long long)(ymin + y) * row_bytes + col;
      if (VEC == 4) {
        const uint32_t u = __ldg(reinterpret_cast<const uint32_t*>(src));
        acc[0] += (int)(u & 0xffu) * w; acc[1 % VEC] += (int)((u >> 8) & 0xffu) * w;
        acc[2 % VEC] += (int)((u >> 16) & 0xffu) * w; acc[3 % VEC] += (int)(u >> 24) * w;
      } else {
#pragma unroll
        for (int c = 0; c < VEC; ++c) acc[c] += (int)src[c] * w;
      }
    }
    uint8_t* dst = out + (long long)yy * row_bytes + col;
    if (VEC == 4) {
      *reinterpret_cast<uint32_t*>(dst) = (uint32_t)resample_clip8(acc[0]) | ((uint32_t)resample_clip8(acc[1 % VEC]) << 8) |
                                          ((uint32_t)resample_clip8(acc[2 % VEC]) << 16) | ((uint32_t)resample_clip8(acc[3 % VEC]) << 24);
    } else {
#pragma unroll
      for (int c = 0; c < VEC; ++c) dst[c] = resample_clip8(acc[c]);
    }
  }
}
}  // namespace fd

using namespace fd;

static unsigned ew_grid(long long total) {
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  return (unsigned)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

extern "C" int fd_image_preprocess(const uint8_t* rgb, void* out_nhwc_bf16, void* out_nchw_f32, int32_t h, int32_t w, void* stream_) {
  FD_REQUIRE(rgb && (out_nhwc_bf16 || out_nchw_f32), "fd_image_preprocess: null pointer");
  const long long hw = (long long)h * w;
  preprocess_kernel<<<ew_grid(hw * 3), 256, 0, (cudaStream_t)stream_>>>(rgb, (__nv_bfloat16*)out_nhwc_bf16, (float*)out_nchw_f32, hw);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fd_image_postprocess(const void* x, int32_t x_is_nhwc_bf16, uint8_t* rgb, int32_t h, int32_t w, void* stream_) {
  FD_REQUIRE(x && rgb, "fd_image_postprocess: null pointer");
  const long long hw = (long long)h * w;
  postprocess_kernel<<<ew_grid(hw * 3), 256, 0, (cudaStream_t)stream_>>>(x, x_is_nhwc_bf16, rgb, hw);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int64_t fd_colorfix_work_floats(int32_t h, int32_t w) { return (int64_t)15 * h * w + 64; }

extern "C" int fd_adain_color_fix(const uint8_t* target, const uint8_t* source, uint8_t* out, int32_t h, int32_t w, float* work,
                                  void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(target && source && out && work && ((uintptr_t)work % 8) == 0, "fd_adain_color_fix: bad arguments");
  const long long hw = (long long)h * w;
  double* stats = reinterpret_cast<double*>(work);
  FD_CHECK_CUDA(cudaMemsetAsync(stats, 0, 12 * sizeof(double), stream));
  adain_stats_kernel<<<ew_grid(hw) / 2 + 1, 256, 0, stream>>>(target, source, hw, stats);
  adain_apply_kernel<<<ew_grid(hw * 3), 256, 0, stream>>>(target, out, hw, stats);
  g_launches += 2;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fd_wavelet_color_fix(const uint8_t* target, const uint8_t* source, uint8_t* out, int32_t h, int32_t w, float* work,
                                    void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(target && source && out && work, "fd_wavelet_color_fix: bad arguments");
  const long long hw = (long long)h * w;
  float* a = work;              // 6 planes
  float* b = work + 6 * hw;     // 6 planes
  float* high = work + 12 * hw; // 3 planes
  wavelet_init_kernel<<<ew_grid(hw * 3), 256, 0, stream>>>(target, source, a, high, hw);
  dim3 block(32, 8), grid((w + 31) / 32, (h + 7) / 8, 6);
  for (int lvl = 0; lvl < 5; ++lvl) {
    wavelet_level_kernel<<<grid, block, 0, stream>>>(a, b, high, h, w, 1 << lvl);
    float* t = a; a = b; b = t;
  }
  wavelet_final_kernel<<<ew_grid(hw * 3), 256, 0, stream>>>(high, a + 3 * hw, out, hw);
  g_launches += 7;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fd_resample_u8_horizontal(const uint8_t* in, int32_t in_w, int64_t rows, uint8_t* out, int32_t out_w, int32_t channels,
                                         const int32_t* bounds, const int32_t* taps, int32_t ksize, void* stream_) {
  FD_REQUIRE(in && out && bounds && taps, "fd_resample_u8_horizontal: null pointer");
  FD_REQUIRE(channels >= 1 && channels <= 4 && in_w > 0 && out_w > 0 && rows >= 0 && ksize > 0, "fd_resample_u8_horizontal: bad shape");
  if (rows == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream_;
  const unsigned grid = ew_grid(rows * out_w);
#define FD_RH(C) resample_h_kernel<C><<<grid, 256, 0, st>>>(in, in_w, (long long)rows, out, out_w, bounds, taps, ksize)
  if (channels == 1) FD_RH(1); else if (channels == 2) FD_RH(2); else if (channels == 3) FD_RH(3); else FD_RH(4);
#undef FD_RH
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fd_resample_u8_vertical(const uint8_t* in, int64_t row_bytes, uint8_t* out, int32_t out_h, const int32_t* bounds,
                                       const int32_t* taps, int32_t ksize, void* stream_) {
  FD_REQUIRE(in && out && bounds && taps, "fd_resample_u8_vertical: null pointer");
  FD_REQUIRE(row_bytes > 0 && out_h >= 0 && ksize > 0, "fd_resample_u8_vertical: bad shape");
  if (out_h == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream_;
  const bool vec = row_bytes % 4 == 0 && ((uintptr_t)in % 4 == 0) && ((uintptr_t)out % 4 == 0);
  if (vec) resample_v_kernel<4><<<ew_grid(row_bytes / 4 * out_h), 256, 0, st>>>(in, (long long)row_bytes, out, out_h, bounds, taps, ksize);
  else resample_v_kernel<1><<<ew_grid(row_bytes * out_h), 256, 0, st>>>(in, (long long)row_bytes, out, out_h, bounds, taps, ksize);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}
