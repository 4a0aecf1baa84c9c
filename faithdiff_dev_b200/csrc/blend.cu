// Tiled-latent kernels: tile gather (LocalAttention.grids), Gaussian overlap blend (grids_inverse), the fused
// per-step tail (CFG combine + Euler step + blend + re-tile + scale_model_input) and the VAE seam cross-fade.
// Reference: FaithDiff/pipelines/pipeline_FaithDiff_tlc.py:160-252, :1326, :1363-1373, :1409;
//            FaithDiff/models/unet_2d_condition_vae_extension.py:150-162.
// Bit-exactness: every arithmetic step reproduces the reference's dtype promotion and rounding order with explicit
// round-to-nearest intrinsics (no FMA contraction): low-precision tile × fp64 weight, accumulated into an fp32 canvas
// through an fp64 add, in row-major tile order; fp32 divide; round-to-nearest-even cast.
#include "common.cuh"
#include "../../include/faithdiff_b200.h"

namespace fd {
extern int64_t g_launches;

__device__ __forceinline__ float load_lowp(const void* p, int dtype, long long i) {
  if (dtype == 0) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  if (dtype == 1) return __half2float(reinterpret_cast<const __half*>(p)[i]);
  return reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void store_lowp(void* p, int dtype, long long i, float v) {
  if (dtype == 0) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
  else if (dtype == 1) reinterpret_cast<__half*>(p)[i] = __float2half_rn(v);
  else reinterpret_cast<float*>(p)[i] = v;
}
// round an fp32 value through the storage dtype (what a torch elementwise op on a bf16/fp16 tensor does)
__device__ __forceinline__ float round_lowp(float v, int dtype) {
  if (dtype == 0) return __bfloat162float(__float2bfloat16_rn(v));
  if (dtype == 1) return __half2float(__float2half_rn(v));
  return v;
}

__global__ void tile_gather_kernel(const uint8_t* __restrict__ canvas, uint8_t* __restrict__ tiles,
                                   const int* __restrict__ org, int ntiles, int c, int h, int w, int k1, int k2, int eb) {
  const long long total = (long long)ntiles * c * k1 * k2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int lx = (int)(i % k2);
    long long t = i / k2;
    const int ly = (int)(t % k1); t /= k1;
    const int ch = (int)(t % c);
    const int tl = (int)(t / c);
    const long long src = ((long long)ch * h + org[2 * tl] + ly) * w + org[2 * tl + 1] + lx;
    if (eb == 2) reinterpret_cast<uint16_t*>(tiles)[i] = reinterpret_cast<const uint16_t*>(canvas)[src];
    else reinterpret_cast<uint32_t*>(tiles)[i] = reinterpret_cast<const uint32_t*>(canvas)[src];
  }
}

// mode 0: plain blend of given tiles.  mode 1: fused CFG + Euler update per tile before blending.
struct BlendParams {
  const void* tiles;      // mode 0: [T,C,k1,k2]; mode 1: current latents x [T,C,k1,k2]
  const void* eps;        // mode 1: UNet output [T,2,C,k1,k2]
  int dtype;              // dtype of eps (mode 1) / of the tiles (mode 0): the model dtype, whose rounding the CFG arithmetic follows
  int x_dtype;            // dtype of the latent state x (tiles in, next_tiles out): the model dtype (reference cast semantics) or fp32
  const double* weights;  // [k1,k2] fp64 (channel-invariant)
  const int* org;
  int ntiles, c, h, w, k1, k2;
  float guidance, sigma, sigma_next, inv_sigma, inv_next_div;
  float* canvas_f32;
  void* canvas_lowp;
  int lowp_dtype;
  void* next_tiles;
  void* next_scaled;
  float* canvas_num;      // accumulate-only output (multi-GPU), else nullptr
  int accumulate_only;
};

__device__ __forceinline__ float euler_step(const BlendParams& p, int t, int ch, int ly, int lx) {
  const long long plane = (long long)p.k1 * p.k2;
  const long long off = (long long)ly * p.k2 + lx;
  const float eu = load_lowp(p.eps, p.dtype, ((long long)(t * 2 + 0) * p.c + ch) * plane + off);
  const float et = load_lowp(p.eps, p.dtype, ((long long)(t * 2 + 1) * p.c + ch) * plane + off);
  // noise_pred = eu + g * (et - eu), each op rounded to the tensor dtype (pipeline…:1365).  With the latent state kept in fp32
  // (x_dtype = 2) the guidance arithmetic stays in fp32 as well: its three roundings act on 5x-amplified differences and were the
  // largest remaining bf16 term of the per-step latent error; it also makes this tail and the N-rank unit form the same arithmetic.
  const int rd = p.x_dtype == 2 ? 2 : p.dtype;
  float d = round_lowp(__fsub_rn(et, eu), rd);
  d = round_lowp(__fmul_rn(d, p.guidance), rd);
  const float e = round_lowp(__fadd_rn(eu, d), rd);
  // EulerDiscreteScheduler.step (gamma = 0): fp32 sample; sigma*eps rounded to the model dtype
  const float x = load_lowp(p.tiles, p.x_dtype, ((long long)t * p.c + ch) * plane + off);
  const float se = round_lowp(__fmul_rn(e, p.sigma), rd);
  const float pred0 = __fsub_rn(x, se);
  // torch's CUDA true-divide by a host scalar multiplies by the fp32 reciprocal (BinaryDivTrueKernel.cu); the
  // reference runs on CUDA, so that is the arithmetic reproduced here
  const float deriv = __fmul_rn(__fsub_rn(x, pred0), p.inv_sigma);
  const float dt = __fsub_rn(p.sigma_next, p.sigma);
  return round_lowp(__fadd_rn(x, __fmul_rn(deriv, dt)), p.x_dtype);
}

__global__ void tile_blend_kernel(const BlendParams p) {
  const long long total = (long long)p.c * p.h * p.w;
  const long long plane = (long long)p.k1 * p.k2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % p.w);
    const int y = (int)((i / p.w) % p.h);
    const int ch = (int)(i / ((long long)p.w * p.h));
    float pred = p.accumulate_only ? p.canvas_num[i] : 0.f, cnt = 0.f;
    for (int t = 0; t < p.ntiles; ++t) {
      const int ly = y - p.org[2 * t], lx = x - p.org[2 * t + 1];
      if (ly < 0 || ly >= p.k1 || lx < 0 || lx >= p.k2) continue;
      const double wgt = p.weights[(long long)ly * p.k2 + lx];
      float v;
      if (p.eps) v = euler_step(p, t, ch, ly, lx);
      else v = load_lowp(p.tiles, p.x_dtype, ((long long)t * p.c + ch) * plane + (long long)ly * p.k2 + lx);
      pred = __double2float_rn(__dadd_rn((double)pred, __dmul_rn((double)v, wgt)));
      cnt = __double2float_rn(__dadd_rn((double)cnt, wgt));
    }
    if (p.accumulate_only) { p.canvas_num[i] = pred; continue; }
    const float r = __fdiv_rn(pred, cnt);
    if (p.canvas_f32) p.canvas_f32[i] = r;
    if (p.canvas_lowp) store_lowp(p.canvas_lowp, p.lowp_dtype, i, r);
    if (p.next_tiles) {
      const float rl = round_lowp(r, p.x_dtype);
      const float rs = round_lowp(__fmul_rn(rl, p.inv_next_div), p.dtype);
      for (int t = 0; t < p.ntiles; ++t) {
        const int ly = y - p.org[2 * t], lx = x - p.org[2 * t + 1];
        if (ly < 0 || ly >= p.k1 || lx < 0 || lx >= p.k2) continue;
        const long long o = ((long long)t * p.c + ch) * plane + (long long)ly * p.k2 + lx;
        store_lowp(p.next_tiles, p.x_dtype, o, rl);
        if (p.next_scaled) store_lowp(p.next_scaled, p.dtype, o, rs);
      }
    }
  }
}

// (tile, CFG-branch) units — the finer multi-GPU work split.  CFG combine and the Euler update are linear in eps:
//   x' = x + dt * ((1-g) * eps_uncond + g * eps_text)
// so every unit deposits  w * coef_u * eps_u  (coef = (1-g)*dt or g*dt) and the branch-0 unit of a tile also deposits w * x.
// Summed over all ranks (all-reduce) this is the blend numerator; fp32 combination (no intermediate bf16 roundings).
__global__ void unit_accumulate_kernel(const void* __restrict__ x_tiles, int x_dtype, const void* __restrict__ eps_units, int dtype,
                                       const double* __restrict__ weights, const int* __restrict__ org,
                                       const int* __restrict__ unit_tile, const float* __restrict__ unit_coef,
                                       const int* __restrict__ unit_addx, int nunits, int c, int h, int w, int k1, int k2,
                                       float* __restrict__ canvas_num, int overwrite) {
  const long long total = (long long)c * h * w;
  const long long plane = (long long)k1 * k2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int y = (int)((i / w) % h);
    const int ch = (int)(i / ((long long)w * h));
    double acc = overwrite ? 0.0 : (double)canvas_num[i];
    for (int u = 0; u < nunits; ++u) {
      const int t = unit_tile[u];
      const int ly = y - org[2 * t], lx = x - org[2 * t + 1];
      if (ly < 0 || ly >= k1 || lx < 0 || lx >= k2) continue;
      const long long off = (long long)ly * k2 + lx;
      float v = __fmul_rn(unit_coef[u], load_lowp(eps_units, dtype, ((long long)u * c + ch) * plane + off));
      if (unit_addx[u]) v = __fadd_rn(v, load_lowp(x_tiles, x_dtype, ((long long)t * c + ch) * plane + off));
      acc = __dadd_rn(acc, __dmul_rn((double)v, weights[off]));
    }
    canvas_num[i] = __double2float_rn(acc);
  }
}

__global__ void canvas_finalize_kernel(const float* __restrict__ num, const float* __restrict__ count, float* __restrict__ out,
                                       void* lowp, int lowp_dtype, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float r = __fdiv_rn(num[i], count[i]);
    if (out) out[i] = r;
    if (lowp) store_lowp(lowp, lowp_dtype, i, r);
  }
}

// multi-GPU step tail after the exchange, one launch: divide by the count map, cast, re-tile (LocalAttention.grids of the next
// step) and scale_model_input — the gather-form tail of tile_blend_kernel on an already summed numerator
__global__ void canvas_finalize_retile_kernel(const float* __restrict__ num, const float* __restrict__ count, float* __restrict__ out,
                                              void* lowp, int dtype, int x_dtype, const int* __restrict__ org, int ntiles, int c, int h,
                                              int w, int k1, int k2, float inv_next_div, void* next_tiles, void* next_scaled) {
  const long long total = (long long)c * h * w;
  const long long plane = (long long)k1 * k2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const int y = (int)((i / w) % h);
    const int ch = (int)(i / ((long long)w * h));
    const float r = __fdiv_rn(num[i], count[i]);
    if (out) out[i] = r;
    if (lowp) store_lowp(lowp, dtype, i, r);
    const float rl = round_lowp(r, x_dtype);
    const float rs = round_lowp(__fmul_rn(rl, inv_next_div), dtype);
    for (int t = 0; t < ntiles; ++t) {
      const int ly = y - org[2 * t], lx = x - org[2 * t + 1];
      if (ly < 0 || ly >= k1 || lx < 0 || lx >= k2) continue;
      const long long o = ((long long)t * c + ch) * plane + (long long)ly * k2 + lx;
      store_lowp(next_tiles, x_dtype, o, rl);
      store_lowp(next_scaled, dtype, o, rs);
    }
  }
}

// scale_model_input: out = round_lowp(x * (1/div))  (EulerDiscreteScheduler.scale_model_input, pipeline…:1336; torch CUDA
// divides by a host scalar through its fp32 reciprocal)
__global__ void scale_div_kernel(const void* __restrict__ x, int x_dtype, void* __restrict__ out, int dtype, float inv_div, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    store_lowp(out, dtype, i, __fmul_rn(load_lowp(x, x_dtype, i), inv_div));
}

// b[..., y, :] = a[..., -e+y, :] * (1 - y/e) + b[..., y, :] * (y/e)   (vertical) — torch rounding per op
__global__ void seam_blend_kernel(const void* __restrict__ a, void* __restrict__ b, int dtype, int c, int ah, int aw, int bh,
                                  int bw, int extent, int vertical) {
  const int len = vertical ? (bw < aw ? bw : aw) : (bh < ah ? bh : ah);
  const long long total = (long long)c * extent * len;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int l = (int)(i % len);
    const int s = (int)((i / len) % extent);
    const int ch = (int)(i / ((long long)len * extent));
    long long ia, ib;
    if (vertical) { ia = ((long long)ch * ah + (ah - extent + s)) * aw + l; ib = ((long long)ch * bh + s) * bw + l; }
    else { ia = ((long long)ch * ah + l) * aw + (aw - extent + s); ib = ((long long)ch * bh + l) * bw + s; }
    const float wa = (float)(1.0 - (double)s / (double)extent), wb = (float)((double)s / (double)extent);
    const float t1 = round_lowp(__fmul_rn(load_lowp(a, dtype, ia), wa), dtype);
    const float t2 = round_lowp(__fmul_rn(load_lowp(b, dtype, ib), wb), dtype);
    store_lowp(b, dtype, ib, __fadd_rn(t1, t2));
  }
}
}  // namespace fd

using namespace fd;

static unsigned blend_grid(long long total) {
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  return (unsigned)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

extern "C" int fd_tile_gather(const void* canvas, void* tiles, const int32_t* org, int32_t ntiles, int32_t c, int32_t h, int32_t w,
                              int32_t k1, int32_t k2, int32_t elem_bytes, void* stream_) {
  FD_REQUIRE(canvas && tiles && org && (elem_bytes == 2 || elem_bytes == 4), "fd_tile_gather: bad arguments");
  const long long total = (long long)ntiles * c * k1 * k2;
  tile_gather_kernel<<<blend_grid(total), 256, 0, (cudaStream_t)stream_>>>((const uint8_t*)canvas, (uint8_t*)tiles, org, ntiles, c, h,
                                                                         w, k1, k2, elem_bytes);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

static int launch_blend(const BlendParams& p, void* stream_) {
  const long long total = (long long)p.c * p.h * p.w;
  tile_blend_kernel<<<blend_grid(total), 256, 0, (cudaStream_t)stream_>>>(p);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fd_tile_blend(const void* tiles, int32_t tile_dtype, const double* weights, const int32_t* org, int32_t ntiles,
                             int32_t c, int32_t h, int32_t w, int32_t k1, int32_t k2, float* canvas_f32, void* canvas_lowp,
                             int32_t lowp_dtype, void* stream_) {
  FD_REQUIRE(tiles && weights && org && (canvas_f32 || canvas_lowp), "fd_tile_blend: null pointer");
  BlendParams p = {};
  p.tiles = tiles; p.dtype = tile_dtype; p.x_dtype = tile_dtype; p.weights = weights; p.org = org; p.ntiles = ntiles; p.c = c; p.h = h;
  p.w = w; p.k1 = k1; p.k2 = k2; p.canvas_f32 = canvas_f32; p.canvas_lowp = canvas_lowp; p.lowp_dtype = lowp_dtype;
  return launch_blend(p, stream_);
}

extern "C" int fd_cfg_euler_blend(const void* eps, const void* x_tiles, int32_t dtype, int32_t x_dtype, const double* weights, const int32_t* org,
                                  int32_t ntiles, int32_t c, int32_t h, int32_t w, int32_t k1, int32_t k2, float guidance,
                                  float sigma, float sigma_next, float next_scale_div, float* canvas_f32, void* canvas_lowp,
                                  void* next_tiles, void* next_scaled, void* stream_) {
  FD_REQUIRE(eps && x_tiles && weights && org, "fd_cfg_euler_blend: null pointer");
  FD_REQUIRE(x_dtype == dtype || x_dtype == 2, "fd_cfg_euler_blend: the latent state is kept in the model dtype or in fp32");
  BlendParams p = {};
  p.tiles = x_tiles; p.eps = eps; p.dtype = dtype; p.x_dtype = x_dtype; p.weights = weights; p.org = org; p.ntiles = ntiles; p.c = c; p.h = h; p.w = w;
  p.k1 = k1; p.k2 = k2; p.guidance = guidance; p.sigma = sigma; p.sigma_next = sigma_next; p.inv_sigma = 1.0f / sigma;
  p.inv_next_div = 1.0f / next_scale_div;
  p.canvas_f32 = canvas_f32; p.canvas_lowp = canvas_lowp; p.lowp_dtype = dtype; p.next_tiles = next_tiles; p.next_scaled = next_scaled;
  return launch_blend(p, stream_);
}

extern "C" int fd_tile_accumulate(const void* tiles, const void* eps, int32_t dtype, const double* weights, const int32_t* org,
                                  int32_t ntiles, int32_t c, int32_t h, int32_t w, int32_t k1, int32_t k2, float guidance,
                                  float sigma, float sigma_next, float* canvas_num, void* stream_) {
  FD_REQUIRE(tiles && weights && org && canvas_num, "fd_tile_accumulate: null pointer");
  BlendParams p = {};
  p.tiles = tiles; p.eps = eps; p.dtype = dtype; p.x_dtype = dtype; p.weights = weights; p.org = org; p.ntiles = ntiles; p.c = c; p.h = h; p.w = w;
  p.k1 = k1; p.k2 = k2; p.guidance = guidance; p.sigma = sigma; p.sigma_next = sigma_next; p.inv_sigma = 1.0f / sigma;
  p.canvas_num = canvas_num;
  p.accumulate_only = 1;
  return launch_blend(p, stream_);
}

extern "C" int fd_canvas_finalize(const float* canvas_num, const float* count, float* canvas_f32, void* canvas_lowp,
                                  int32_t lowp_dtype, int64_t n, void* stream_) {
  FD_REQUIRE(canvas_num && count, "fd_canvas_finalize: null pointer");
  canvas_finalize_kernel<<<blend_grid(n), 256, 0, (cudaStream_t)stream_>>>(canvas_num, count, canvas_f32, canvas_lowp, lowp_dtype, n);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fd_seam_blend(const void* a, void* b, int32_t dtype, int32_t c, int32_t ah, int32_t aw, int32_t bh, int32_t bw,
                             int32_t extent, int32_t vertical, void* stream_) {
  FD_REQUIRE(a && b, "fd_seam_blend: null pointer");
  int e = extent;
  if (vertical) { if (e > ah) e = ah; if (e > bh) e = bh; } else { if (e > aw) e = aw; if (e > bw) e = bw; }
  if (e <= 0) return 0;
  const int len = vertical ? (bw < aw ? bw : aw) : (bh < ah ? bh : ah);
  seam_blend_kernel<<<blend_grid((long long)c * e * len), 256, 0, (cudaStream_t)stream_>>>(a, b, dtype, c, ah, aw, bh, bw, e, vertical);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fd_scale_div(const void* x, int32_t x_dtype, void* out, int32_t dtype, float div, int64_t n, void* stream_) {
  FD_REQUIRE(x && out && dtype >= 0 && dtype <= 2 && x_dtype >= 0 && x_dtype <= 2, "fd_scale_div: bad arguments");
  scale_div_kernel<<<blend_grid(n), 256, 0, (cudaStream_t)stream_>>>(x, x_dtype, out, dtype, 1.0f / div, n);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fd_unit_accumulate(const void* x_tiles, int32_t x_dtype, const void* eps_units, int32_t dtype, const double* weights,
                                  const int32_t* origins_ij, const int32_t* unit_tile, const float* unit_coef,
                                  const int32_t* unit_addx, int32_t nunits, int32_t c, int32_t h, int32_t w, int32_t k1,
                                  int32_t k2, float* canvas_num, void* stream_) {
  FD_REQUIRE(x_tiles && eps_units && weights && origins_ij && unit_tile && unit_coef && unit_addx && canvas_num,
             "fd_unit_accumulate: null pointer");
  unit_accumulate_kernel<<<blend_grid((long long)c * h * w), 256, 0, (cudaStream_t)stream_>>>(
      x_tiles, x_dtype, eps_units, dtype, weights, origins_ij, unit_tile, unit_coef, unit_addx, nunits, c, h, w, k1, k2, canvas_num, 0);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fd_unit_numerator(const void* x_tiles, int32_t x_dtype, const void* eps_units, int32_t dtype, const double* weights,
                                 const int32_t* origins_ij, const int32_t* unit_tile, const float* unit_coef,
                                 const int32_t* unit_addx, int32_t nunits, int32_t c, int32_t h, int32_t w, int32_t k1,
                                 int32_t k2, float* canvas_num, void* stream_) {
  FD_REQUIRE(x_tiles && eps_units && weights && origins_ij && unit_tile && unit_coef && unit_addx && canvas_num,
             "fd_unit_numerator: null pointer");
  unit_accumulate_kernel<<<blend_grid((long long)c * h * w), 256, 0, (cudaStream_t)stream_>>>(
      x_tiles, x_dtype, eps_units, dtype, weights, origins_ij, unit_tile, unit_coef, unit_addx, nunits, c, h, w, k1, k2, canvas_num, 1);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fd_canvas_finalize_retile(const float* canvas_num, const float* count, float* canvas_f32, void* canvas_lowp,
                                         int32_t dtype, int32_t x_dtype, const int32_t* origins_ij, int32_t ntiles, int32_t c, int32_t h,
                                         int32_t w, int32_t k1, int32_t k2, float next_scale_div, void* next_tiles, void* next_scaled,
                                         void* stream_) {
  FD_REQUIRE(canvas_num && count && origins_ij && next_tiles && next_scaled && dtype >= 0 && dtype <= 2 && (x_dtype == dtype || x_dtype == 2),
             "fd_canvas_finalize_retile: bad arguments");
  canvas_finalize_retile_kernel<<<blend_grid((long long)c * h * w), 256, 0, (cudaStream_t)stream_>>>(
      canvas_num, count, canvas_f32, canvas_lowp, dtype, x_dtype, origins_ij, ntiles, c, h, w, k1, k2, 1.0f / next_scale_div, next_tiles,
      next_scaled);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}
