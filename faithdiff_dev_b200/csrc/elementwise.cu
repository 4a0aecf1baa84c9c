// Small / bandwidth-bound helper kernels: layout changes at the NCHW module boundary, nearest upsample, direct 3x3
// convs with a tiny channel count on one side, tiny-M linears (timestep-embedding MLPs), sinusoidal embedding,
// the FaithDiff batch-axis attention, row softmax.
#include "common.cuh"
#include "../../include/faithdiff_b200.h"

namespace fd {
extern int64_t g_launches;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }

// ------------------------------------------------------------------ NCHW <-> NHWC (32x32 smem transpose tiles)
template <typename TI>
__global__ void nchw_to_nhwc_kernel(const TI* __restrict__ x, __nv_bfloat16* __restrict__ out, int c, long long hw, int ldc) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int cc = c0 + i;
    const long long p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (cc < c && p < hw) ? to_f<TI>(x[((long long)b * c + cc) * hw + p]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long p = p0 + i;
    const int cc = c0 + threadIdx.x;
    if (cc < ldc && p < hw) out[((long long)b * hw + p) * ldc + cc] = __float2bfloat16_rn(tile[threadIdx.x][i]);   // channels [c, ldc) = 0
  }
}
template <typename TO>
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, TO* __restrict__ out, int c, long long hw, long long ldc) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const long long p = p0 + i;
    const int cc = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (cc < c && p < hw) ? __bfloat162float(x[((long long)b * hw + p) * ldc + cc]) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int cc = c0 + i;
    const long long p = p0 + threadIdx.x;
    if (cc < c && p < hw) out[((long long)b * c + cc) * hw + p] = from_f<TO>(tile[threadIdx.x][i]);
  }
}

// ------------------------------------------------------------------ out = alpha*a + beta*b on strided bf16 row views
__global__ void axpby_kernel(const __nv_bfloat16* __restrict__ a, long long lda, const __nv_bfloat16* __restrict__ b, long long ldb,
                             __nv_bfloat16* __restrict__ out, long long ldo, int v, long long total, float alpha, float beta) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / v;
    const int c = (int)(i - r * v) * 8;
    const uint4 ua = __ldg(reinterpret_cast<const uint4*>(a + r * lda + c));
    const uint4 ub = __ldg(reinterpret_cast<const uint4*>(b + r * ldb + c));
    const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w}, wb[4] = {ub.x, ub.y, ub.z, ub.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = unpack_bf16(wa[j]), fb = unpack_bf16(wb[j]);
      o[j] = pack_bf16(fmaf(alpha, fa.x, beta * fb.x), fmaf(alpha, fa.y, beta * fb.y));
    }
    *reinterpret_cast<uint4*>(out + r * ldo + c) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------ nearest 2x upsample, NHWC, 16-byte vectors
__global__ void upsample_nearest_kernel(const uint4* __restrict__ x, uint4* __restrict__ out, int h, int w, int oh, int ow,
                                        int v, long long total) {
  const float sy = (float)h / (float)oh, sx = (float)w / (float)ow;   // torch 'nearest': src = min(floor(dst*scale), in-1)
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % v);
    long long t = i / v;
    const int ox = (int)(t % ow); t /= ow;
    const int oy = (int)(t % oh);
    const long long b = t / oh;
    const int iy = min((int)floorf(oy * sy), h - 1), ix = min((int)floorf(ox * sx), w - 1);
    out[i] = __ldg(x + ((b * h + iy) * w + ix) * v + cv);
  }
}

// ------------------------------------------------------------------ direct conv3x3, tiny cin (<= 8): thread per (pixel, 8 couts)
// in_mode: 0 NHWC bf16, 1 NCHW bf16, 2 NCHW fp32, 3 NCHW fp16
__device__ __forceinline__ float load_in(const void* x, int in_mode, long long b, int c, int y, int xx, int cin, int h, int w) {
  if (in_mode == 0) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[((b * h + y) * w + xx) * cin + c]);
  const long long idx = ((b * cin + c) * h + y) * w + xx;
  if (in_mode == 1) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[idx]);
  if (in_mode == 2) return reinterpret_cast<const float*>(x)[idx];
  return __half2float(reinterpret_cast<const __half*>(x)[idx]);
}
// weights staged TRANSPOSED in smem ([9*CIN][cout] fp32) so that the 8-cout slices read by consecutive threads are
// consecutive 32-byte chunks (conflict-free); inputs live in registers (CIN is a compile-time constant).
template <int CIN>
__global__ void conv3x3_smallcin_kernel(const void* __restrict__ x, int in_mode, const __nv_bfloat16* __restrict__ w,
                                        const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, long long out_ld,
                                        int batch, int h, int wd, int cout) {
  extern __shared__ float sw[];   // [9*CIN][cout]
  constexpr int KK = 9 * CIN;
  for (int i = threadIdx.x; i < cout * KK; i += blockDim.x) {
    const int oc = i / KK, k = i - oc * KK;
    sw[k * cout + oc] = __bfloat162float(w[i]);
  }
  __syncthreads();
  const int ng = cout >> 3;
  const long long total = (long long)batch * h * wd * ng;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % ng);
    long long t = i / ng;
    const int xx = (int)(t % wd); t /= wd;
    const int y = (int)(t % h);
    const long long b = t / h;
    float in[KK];
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = y + tap / 3 - 1, xq = xx + tap % 3 - 1;
      const bool ok = yy >= 0 && yy < h && xq >= 0 && xq < wd;
#pragma unroll
      for (int c = 0; c < CIN; ++c) in[tap * CIN + c] = ok ? load_in(x, in_mode, b, c, yy, xq, CIN, h, wd) : 0.f;
    }
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = bias ? bias[g * 8 + j] : 0.f;
#pragma unroll
    for (int k = 0; k < KK; ++k) {
      const float4 w0 = *reinterpret_cast<const float4*>(sw + k * cout + g * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(sw + k * cout + g * 8 + 4);
      o[0] += in[k] * w0.x; o[1] += in[k] * w0.y; o[2] += in[k] * w0.z; o[3] += in[k] * w0.w;
      o[4] += in[k] * w1.x; o[5] += in[k] * w1.y; o[6] += in[k] * w1.z; o[7] += in[k] * w1.w;
    }
    uint4 u;
    u.x = pack_bf16(o[0], o[1]); u.y = pack_bf16(o[2], o[3]); u.z = pack_bf16(o[4], o[5]); u.w = pack_bf16(o[6], o[7]);
    *reinterpret_cast<uint4*>(out + (((b * h + y) * wd + xx) * out_ld + g * 8)) = u;
  }
}

// tiny cout (<= 8): warp per pixel, lanes split the channel vectors, warp-shuffle reduction.
// out_mode: 0 NHWC bf16, 1 NCHW bf16, 2 NCHW fp32, 3 NCHW fp16
__global__ void conv3x3_smallcout_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                         const float* __restrict__ bias, void* __restrict__ out, int out_mode, int batch,
                                         int h, int wd, int cin, int cout) {
  extern __shared__ __nv_bfloat16 swb[];   // [cout][9*cin]
  const int kk = 9 * cin;
  for (int i = threadIdx.x; i < cout * kk / 8; i += blockDim.x)
    reinterpret_cast<uint4*>(swb)[i] = __ldg(reinterpret_cast<const uint4*>(w) + i);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int V = cin >> 3;
  const long long total = (long long)batch * h * wd;
  const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long pix = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5; pix < total; pix += warps) {
    const int xx = (int)(pix % wd);
    const int y = (int)((pix / wd) % h);
    const long long b = pix / ((long long)wd * h);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = y + tap / 3 - 1, xq = xx + tap % 3 - 1;
      if (yy < 0 || yy >= h || xq < 0 || xq >= wd) continue;
      const __nv_bfloat16* px = x + ((b * h + yy) * wd + xq) * cin;
      for (int v = lane; v < V; v += 32) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(px) + v);
        float f[8];
        float2 t;
        t = unpack_bf16(u.x); f[0] = t.x; f[1] = t.y; t = unpack_bf16(u.y); f[2] = t.x; f[3] = t.y;
        t = unpack_bf16(u.z); f[4] = t.x; f[5] = t.y; t = unpack_bf16(u.w); f[6] = t.x; f[7] = t.y;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (j < cout) {
            const uint4 wu = *reinterpret_cast<const uint4*>(swb + j * kk + tap * cin + v * 8);
            float2 a;
            a = unpack_bf16(wu.x); acc[j] += f[0] * a.x + f[1] * a.y;
            a = unpack_bf16(wu.y); acc[j] += f[2] * a.x + f[3] * a.y;
            a = unpack_bf16(wu.z); acc[j] += f[4] * a.x + f[5] * a.y;
            a = unpack_bf16(wu.w); acc[j] += f[6] * a.x + f[7] * a.y;
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = warp_sum(acc[j]);
    if (lane < cout) {
      float r = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) if (j == lane) r = acc[j];
      r += bias ? bias[lane] : 0.f;
      const long long hwp = (long long)y * wd + xx;
      if (out_mode == 0) reinterpret_cast<__nv_bfloat16*>(out)[pix * cout + lane] = __float2bfloat16_rn(r);
      else {
        const long long idx = (b * cout + lane) * ((long long)h * wd) + hwp;
        if (out_mode == 1) reinterpret_cast<__nv_bfloat16*>(out)[idx] = __float2bfloat16_rn(r);
        else if (out_mode == 2) reinterpret_cast<float*>(out)[idx] = r;
        else reinterpret_cast<__half*>(out)[idx] = __float2half_rn(r);
      }
    }
  }
}

// ------------------------------------------------------------------ 1x1 conv, <= 8 channels each side, NCHW
__global__ void conv1x1_small_kernel(const void* __restrict__ x, int in_dtype, const float* __restrict__ w,
                                     const float* __restrict__ bias, void* __restrict__ out, int out_dtype, int cin, int cout,
                                     long long hw, float in_scale, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / hw, p = i % hw;
    float in[8];
    for (int c = 0; c < cin; ++c) {
      const long long idx = (b * cin + c) * hw + p;
      float v = in_dtype == 0 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[idx])
              : in_dtype == 1 ? __half2float(reinterpret_cast<const __half*>(x)[idx]) : reinterpret_cast<const float*>(x)[idx];
      in[c] = v * in_scale;
    }
    for (int o = 0; o < cout; ++o) {
      float acc = bias ? bias[o] : 0.f;
      for (int c = 0; c < cin; ++c) acc += in[c] * w[o * cin + c];
      const long long idx = (b * cout + o) * hw + p;
      if (out_dtype == 0) reinterpret_cast<__nv_bfloat16*>(out)[idx] = __float2bfloat16_rn(acc);
      else if (out_dtype == 1) reinterpret_cast<__half*>(out)[idx] = __float2half_rn(acc);
      else reinterpret_cast<float*>(out)[idx] = acc;
    }
  }
}

// ------------------------------------------------------------------ tiny-M linear (M <= 8): warp per output column
__global__ void linear_smallm_kernel(const float* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                     const float* __restrict__ bias, const float* __restrict__ addend,
                                     float* __restrict__ out, int m, int n, int k,
                                     int act_in_silu, int act_out_silu) {
  const int lane = threadIdx.x & 31;
  const int col = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (col >= n) return;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  const __nv_bfloat16* wr = w + (long long)col * k;
  for (int kk = lane * 8; kk < k; kk += 256) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(wr + kk));
    float f[8];
    float2 t;
    t = unpack_bf16(u.x); f[0] = t.x; f[1] = t.y; t = unpack_bf16(u.y); f[2] = t.x; f[3] = t.y;
    t = unpack_bf16(u.z); f[4] = t.x; f[5] = t.y; t = unpack_bf16(u.w); f[6] = t.x; f[7] = t.y;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < m) {
        const float* xr = x + (long long)i * k + kk;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float xv = xr[j];
          if (act_in_silu) xv = xv / (1.f + expf(-xv));
          acc[i] += xv * f[j];
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = warp_sum(acc[i]);
  if (lane == 0) {
    for (int i = 0; i < m; ++i) {
      float r = acc[i] + (bias ? bias[col] : 0.f);
      if (act_out_silu) r = r / (1.f + expf(-r));
      if (addend) r += addend[(long long)i * n + col];
      out[(long long)i * n + col] = r;
    }
  }
}

// diffusers Timesteps(flip_sin_to_cos=True, downscale_freq_shift=0): [cos(v f_j), sin(v f_j)], f_j = exp(-ln(1e4) j / half)
__global__ void timestep_embedding_kernel(const float* __restrict__ values, float* __restrict__ out, int n, int dim,
                                          int inner, long long outer_ld) {
  const int half = dim >> 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * half) return;
  const int r = i / half, j = i % half;
  const float expo = -logf(10000.f) * (float)j;
  const float f = expf(expo / (float)half);
  const float a = values[r] * f;
  float* o = out + (long long)(r / inner) * outer_ld + (long long)(r % inner) * dim;
  o[j] = cosf(a);
  o[half + j] = sinf(a);
}

// ------------------------------------------------------------------ FaithDiff batch-axis attention
// qkv [G, L, HW, 3E] bf16 (q | k | v), out [G, L, HW, E]; sequence axis = L (the reference's batch), L <= 4.
__global__ void batch_axis_attention_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out,
                                            int L, long long hw, int E, int heads, long long total) {
  const int d = E / heads, dv = d >> 3;
  const float scale = rsqrtf((float)d);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int hd = (int)(i % heads);
    long long t = i / heads;
    const long long p = t % hw;
    const long long g = t / hw;
    const __nv_bfloat16* base = qkv + ((g * L) * hw + p) * 3 * E + hd * d;
    const long long lstride = hw * 3 * (long long)E;
    float s[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) s[a][b] = 0.f;
    for (int v = 0; v < dv; ++v) {
      float q[4][8], k[4][8];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        if (a < L) {
          const uint4 uq = __ldg(reinterpret_cast<const uint4*>(base + a * lstride) + v);
          const uint4 uk = __ldg(reinterpret_cast<const uint4*>(base + a * lstride + E) + v);
          float2 f;
          f = unpack_bf16(uq.x); q[a][0] = f.x; q[a][1] = f.y; f = unpack_bf16(uq.y); q[a][2] = f.x; q[a][3] = f.y;
          f = unpack_bf16(uq.z); q[a][4] = f.x; q[a][5] = f.y; f = unpack_bf16(uq.w); q[a][6] = f.x; q[a][7] = f.y;
          f = unpack_bf16(uk.x); k[a][0] = f.x; k[a][1] = f.y; f = unpack_bf16(uk.y); k[a][2] = f.x; k[a][3] = f.y;
          f = unpack_bf16(uk.z); k[a][4] = f.x; k[a][5] = f.y; f = unpack_bf16(uk.w); k[a][6] = f.x; k[a][7] = f.y;
        }
      }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if (a < L && b < L) {
#pragma unroll
            for (int j = 0; j < 8; ++j) s[a][b] += q[a][j] * k[b][j];
          }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (a < L) {
        float mx = -INFINITY;
#pragma unroll
        for (int b = 0; b < 4; ++b) if (b < L) { s[a][b] *= scale; mx = fmaxf(mx, s[a][b]); }
        float sum = 0.f;
#pragma unroll
        for (int b = 0; b < 4; ++b) if (b < L) { s[a][b] = expf(s[a][b] - mx); sum += s[a][b]; }
        const float inv = 1.f / sum;
#pragma unroll
        for (int b = 0; b < 4; ++b) if (b < L) s[a][b] *= inv;
      }
    }
    for (int v = 0; v < dv; ++v) {
      float vv[4][8];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        if (b < L) {
          const uint4 u = __ldg(reinterpret_cast<const uint4*>(base + b * lstride + 2 * E) + v);
          float2 f;
          f = unpack_bf16(u.x); vv[b][0] = f.x; vv[b][1] = f.y; f = unpack_bf16(u.y); vv[b][2] = f.x; vv[b][3] = f.y;
          f = unpack_bf16(u.z); vv[b][4] = f.x; vv[b][5] = f.y; f = unpack_bf16(u.w); vv[b][6] = f.x; vv[b][7] = f.y;
        }
      }
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        if (a < L) {
          float o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            o[j] = 0.f;
#pragma unroll
            for (int b = 0; b < 4; ++b) if (b < L) o[j] += s[a][b] * vv[b][j];
          }
          uint4 u;
          u.x = pack_bf16(o[0], o[1]); u.y = pack_bf16(o[2], o[3]); u.z = pack_bf16(o[4], o[5]); u.w = pack_bf16(o[6], o[7]);
          *(reinterpret_cast<uint4*>(out + ((g * L + a) * hw + p) * E + hd * d) + v) = u;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ row softmax fp32 -> bf16 (block per row)
__global__ void softmax_rows_kernel(const float* __restrict__ s, long long ld_s, __nv_bfloat16* __restrict__ p,
                                    long long ld_p, long long cols, float alpha) {
  __shared__ float red[32];
  const long long row = blockIdx.x;
  const float* sr = s + row * ld_s;
  float mx = -INFINITY;
  for (long long c = threadIdx.x; c < cols; c += blockDim.x) mx = fmaxf(mx, sr[c] * alpha);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = -INFINITY;
  for (int i = 0; i < (blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (long long c = threadIdx.x; c < cols; c += blockDim.x) sum += __expf(sr[c] * alpha - mx);
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = 0.f;
  for (int i = 0; i < (blockDim.x >> 5); ++i) sum += red[i];
  const float inv = 1.f / sum;
  __nv_bfloat16* pr = p + row * ld_p;
  for (long long c = threadIdx.x; c < cols; c += blockDim.x) pr[c] = __float2bfloat16_rn(__expf(sr[c] * alpha - mx) * inv);
}
}  // namespace fd

using namespace fd;

static inline unsigned grid_for(long long total, int threads, int per_sm = 16) {
  long long blocks = (total + threads - 1) / threads;
  const long long cap = (long long)num_sms() * per_sm;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

extern "C" int fd_nchw_to_nhwc_padded(const void* x, int32_t in_dtype, void* out, int32_t batch, int32_t c, int32_t ldc, int64_t hw,
                                      void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(x && out && in_dtype >= 0 && in_dtype <= 2 && ldc >= c, "fd_nchw_to_nhwc: bad arguments");
  dim3 grid((unsigned)((hw + 31) / 32), (ldc + 31) / 32, batch), block(32, 8);
  if (in_dtype == 0) nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, c, hw, ldc);
  else if (in_dtype == 1) nchw_to_nhwc_kernel<__half><<<grid, block, 0, stream>>>((const __half*)x, (__nv_bfloat16*)out, c, hw, ldc);
  else nchw_to_nhwc_kernel<float><<<grid, block, 0, stream>>>((const float*)x, (__nv_bfloat16*)out, c, hw, ldc);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fd_nchw_to_nhwc(const void* x, int32_t in_dtype, void* out, int32_t batch, int32_t c, int64_t hw, void* stream_) {
  return fd_nchw_to_nhwc_padded(x, in_dtype, out, batch, c, c, hw, stream_);
}

extern "C" int fd_nhwc_to_nchw_strided(const void* x, int64_t ldc, void* out, int32_t out_dtype, int32_t batch, int32_t c, int64_t hw,
                                       void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(x && out && out_dtype >= 0 && out_dtype <= 2 && ldc >= c, "fd_nhwc_to_nchw: bad arguments");
  dim3 grid((unsigned)((hw + 31) / 32), (c + 31) / 32, batch), block(32, 8);
  if (out_dtype == 0) nhwc_to_nchw_kernel<__nv_bfloat16><<<grid, block, 0, stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, c, hw, ldc);
  else if (out_dtype == 1) nhwc_to_nchw_kernel<__half><<<grid, block, 0, stream>>>((const __nv_bfloat16*)x, (__half*)out, c, hw, ldc);
  else nhwc_to_nchw_kernel<float><<<grid, block, 0, stream>>>((const __nv_bfloat16*)x, (float*)out, c, hw, ldc);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fd_nhwc_to_nchw(const void* x, void* out, int32_t out_dtype, int32_t batch, int32_t c, int64_t hw, void* stream_) {
  return fd_nhwc_to_nchw_strided(x, c, out, out_dtype, batch, c, hw, stream_);
}

extern "C" int fd_upsample_nearest_nhwc(const void* x, void* out, int32_t batch, int32_t h, int32_t w, int32_t oh, int32_t ow,
                                        int32_t c, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(x && out && c % 8 == 0, "fd_upsample_nearest_nhwc: channels must be a multiple of 8");
  const long long total = (long long)batch * oh * ow * (c / 8);
  upsample_nearest_kernel<<<grid_for(total, 256), 256, 0, stream>>>((const uint4*)x, (uint4*)out, h, w, oh, ow, c / 8, total);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fd_axpby_bf16(const void* a, int64_t lda, const void* b, int64_t ldb, void* out, int64_t ldo, int64_t rows,
                             int32_t cols, float alpha, float beta, void* stream_) {
  FD_REQUIRE(a && b && out && rows >= 0 && cols > 0, "fd_axpby_bf16: bad arguments");
  FD_REQUIRE(cols % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldo % 8 == 0 && ((uintptr_t)a % 16) == 0 && ((uintptr_t)b % 16) == 0 &&
                 ((uintptr_t)out % 16) == 0, "fd_axpby_bf16: cols / pitches must be multiples of 8 and pointers 16-byte aligned");
  if (rows == 0) return 0;
  const long long total = rows * (cols / 8);
  axpby_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream_>>>((const __nv_bfloat16*)a, lda, (const __nv_bfloat16*)b, ldb,
                                                                       (__nv_bfloat16*)out, ldo, cols / 8, total, alpha, beta);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fd_conv3x3_small(const void* x, int32_t in_mode, const void* w, const float* bias, void* out, int32_t out_mode,
                                int64_t out_ld, int32_t batch, int32_t h, int32_t wd, int32_t cin, int32_t cout, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(x && w && out, "fd_conv3x3_small: null pointer");
  if (cin <= 8) {
    FD_REQUIRE(cout % 8 == 0 && out_mode == 0, "fd_conv3x3_small: small-cin path needs cout%%8==0 and NHWC bf16 output");
    if (out_ld <= 0) out_ld = cout;
    FD_REQUIRE(out_ld % 8 == 0 && ((uintptr_t)out % 16) == 0, "fd_conv3x3_small: output pitch/alignment");
    const size_t smem = (size_t)cout * 9 * cin * sizeof(float);
    FD_REQUIRE(smem <= 96 * 1024, "fd_conv3x3_small: weights too large for smem");
    FD_REQUIRE(cin == 3 || cin == 4 || cin == 8, "fd_conv3x3_small: small-cin path supports cin 3, 4 or 8 (got %d)", cin);
    const long long total = (long long)batch * h * wd * (cout / 8);
    const unsigned grid = grid_for(total, 256, 8);
#define FD_LAUNCH_SMALLCIN(C)                                                                                              \
    {                                                                                                                      \
      static bool attr = false;                                                                                            \
      if (!attr) { FD_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_smallcin_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); attr = true; } \
      conv3x3_smallcin_kernel<C><<<grid, 256, smem, stream>>>(x, in_mode, (const __nv_bfloat16*)w, bias, (__nv_bfloat16*)out, out_ld, batch, h, wd, cout); \
    }
    if (cin == 3) FD_LAUNCH_SMALLCIN(3) else if (cin == 4) FD_LAUNCH_SMALLCIN(4) else FD_LAUNCH_SMALLCIN(8)
#undef FD_LAUNCH_SMALLCIN
  } else {
    FD_REQUIRE(cout <= 8 && cin % 8 == 0 && in_mode == 0, "fd_conv3x3_small: small-cout path needs cout<=8, cin%%8==0, NHWC bf16 input");
    const size_t smem = (size_t)cout * 9 * cin * sizeof(__nv_bfloat16);
    FD_REQUIRE(smem <= 96 * 1024, "fd_conv3x3_small: weights too large for smem");
    static bool attr = false;
    if (!attr) { FD_CHECK_CUDA(cudaFuncSetAttribute(conv3x3_smallcout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024)); attr = true; }
    const long long total = (long long)batch * h * wd * 32;
    conv3x3_smallcout_kernel<<<grid_for(total, 256, 8), 256, smem, stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)w, bias,
                                                                           out, out_mode, batch, h, wd, cin, cout);
  }
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fd_linear_smallm(const float* x, const void* w, const float* bias, const float* addend, float* out, int32_t m,
                                int32_t n, int32_t k, int32_t act_in_silu, int32_t act_out_silu, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(x && w && out && m >= 1 && m <= 8 && k % 8 == 0, "fd_linear_smallm: need 1<=m<=8 and k%%8==0 (m=%d k=%d)", m, k);
  const long long threads = (long long)n * 32;
  linear_smallm_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(x, (const __nv_bfloat16*)w, bias, addend, out, m, n, k,
                                                                             act_in_silu, act_out_silu);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fd_timestep_embedding(const float* values, float* out, int32_t n, int32_t dim, int32_t inner, int64_t outer_ld,
                                     void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(values && out && dim % 2 == 0, "fd_timestep_embedding: bad arguments");
  const int total = n * (dim / 2);
  if (inner <= 0) inner = 1;
  timestep_embedding_kernel<<<(total + 127) / 128, 128, 0, stream>>>(values, out, n, dim, inner, outer_ld);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fd_batch_axis_attention(const void* qkv, void* out, int32_t groups, int32_t seq, int64_t hw, int32_t embed,
                                       int32_t heads, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(qkv && out && seq >= 1 && seq <= 4 && embed % heads == 0 && (embed / heads) % 8 == 0,
             "fd_batch_axis_attention: need seq<=4 and head_dim%%8==0 (seq=%d embed=%d heads=%d)", seq, embed, heads);
  const long long total = (long long)groups * hw * heads;
  batch_axis_attention_kernel<<<grid_for(total, 128, 16), 128, 0, stream>>>((const __nv_bfloat16*)qkv, (__nv_bfloat16*)out, seq, hw,
                                                                          embed, heads, total);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fd_softmax_rows(const float* s, int64_t ld_s, void* p, int64_t ld_p, int64_t rows, int64_t cols, float alpha,
                               void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(s && p && rows > 0 && cols > 0, "fd_softmax_rows: bad arguments");
  softmax_rows_kernel<<<(unsigned)rows, 256, 0, stream>>>(s, ld_s, (__nv_bfloat16*)p, ld_p, cols, alpha);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fd_conv1x1_small(const void* x, int32_t in_dtype, const float* w, const float* bias, void* out, int32_t out_dtype,
                                int32_t batch, int32_t cin, int32_t cout, int64_t hw, float in_scale, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(x && w && out && cin >= 1 && cin <= 8 && cout >= 1 && cout <= 8, "fd_conv1x1_small: need 1..8 channels per side");
  const long long total = (long long)batch * hw;
  conv1x1_small_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x, in_dtype, w, bias, out, out_dtype, cin, cout, hw, in_scale, total);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}
