// Persistent, warp-specialised tcgen05 implicit-GEMM for sm_100a.
//   warp 0 lane 0 : TMA producer   (A tile 128x64 bf16 + B tile block_n x 64 bf16 per stage, 128B swizzle)
//   warp 1 lane 0 : MMA issuer     (tcgen05.mma cta_group::1 kind::f16, M=128, N=block_n, K=16; fp32 accum in TMEM)
//   warps 2..5    : epilogue       (tcgen05.ld -> bias / timestep row-bias / activation / GEGLU / residual -> global)
// TMEM holds two accumulator stages (2 x 256 columns) so the epilogue of tile i overlaps the mainloop of tile i+1.
// A operand is either a row-major matrix (2-D tensor map) or an NHWC image read through 4-D tensor maps with the
// 3x3 filter taps expressed as shifted TMA boxes (out-of-bounds = zero fill = the conv's zero padding);
// stride-2 convs use four phase-subsampled tensor maps of the same image (no data rearrangement).
#include "common.cuh"
#include "../../include/faithdiff_b200.h"

namespace fd {
extern int64_t g_launches;

struct GemmKParams {
  CUtensorMap amap[4];
  CUtensorMap bmap;
  int mode;            // 0 gemm, 1 conv
  int M, N;
  int m_tiles, n_tiles, block_n, stages;
  int taps, kb_per_tap, kb_src0, ctot, two_src;
  int tap_map[9], tap_dy[9], tap_dx[9];
  int out_h, out_w, bw, bh, tiles_x, tiles_per_img;
  // epilogue
  const float* bias;
  const float* rowbias;
  long long rows_per_batch, ld_rowbias;
  const __nv_bfloat16* residual;
  long long ld_res;
  void* out;
  long long ld_out;
  float alpha;
  int act, bias_along_m, out_fp32;
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kATileBytes = kBlockM * kBlockK * 2;   // 16 KB
constexpr int kAccCols = 256;                        // TMEM columns per accumulator stage
constexpr int kMaxStages = 8;
constexpr int kThreads = 192;

struct __align__(8) PipeBarriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreads, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmKParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b_tile_bytes = p.block_n * kBlockK * 2;
  const int stage_bytes = kATileBytes + b_tile_bytes;
  PipeBarriers* bars = reinterpret_cast<PipeBarriers*>(smem + (size_t)p.stages * stage_bytes);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.bmap);
    tma_prefetch_desc(&p.amap[0]);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], 1);
      mbar_init(&bars->tmem_empty[a], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  const int num_tiles = p.m_tiles * p.n_tiles;
  const int num_kb = p.taps * p.kb_per_tap;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------------ TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile % p.m_tiles;
        const int n_blk = tile / p.m_tiles;
        int img = 0, y0 = 0, x0 = 0;
        if (p.mode != 0) {
          img = m_blk / p.tiles_per_img;
          const int r = m_blk - img * p.tiles_per_img;
          y0 = (r / p.tiles_x) * p.bh;
          x0 = (r % p.tiles_x) * p.bw;
        }
        for (int tap = 0; tap < p.taps; ++tap) {
          for (int kb = 0; kb < p.kb_per_tap; ++kb) {
            mbar_wait(&bars->empty[stage], phase ^ 1);
            uint8_t* sa = smem + (size_t)stage * stage_bytes;
            uint8_t* sb = sa + kATileBytes;
            mbar_arrive_expect_tx(&bars->full[stage], (uint32_t)stage_bytes);
            int src, c0;
            if (p.two_src) {
              src = kb >= p.kb_src0;
              c0 = (src ? kb - p.kb_src0 : kb) * kBlockK;
            } else {
              src = p.tap_map[tap];
              c0 = kb * kBlockK;
            }
            if (p.mode == 0)
              tma_load_2d(&p.amap[src], &bars->full[stage], sa, c0, m_blk * kBlockM);
            else
              tma_load_4d(&p.amap[src], &bars->full[stage], sa, c0, x0 + p.tap_dx[tap], y0 + p.tap_dy[tap], img);
            tma_load_2d(&p.bmap, &bars->full[stage], sb, tap * p.ctot + kb * kBlockK, n_blk * p.block_n);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------------------------ MMA issuer (single thread)
      const uint32_t idesc = umma_idesc(kBlockM, p.block_n, /*bf16*/ 1, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&bars->tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * kAccCols;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t sb = sa + kATileBytes;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            const uint64_t adesc = umma_desc_sw128(sa + k * 32, 16, 1024);
            const uint64_t bdesc = umma_desc_sw128(sb + k * 32, 16, 1024);
            umma_f16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0);
          }
          tc_commit(&bars->empty[stage]);          // frees this smem stage once the MMAs above retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        tc_commit(&bars->tmem_full[acc]);          // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // -------------------------------------------------------------------- epilogue warps (2..5)
    const int quad = warp & 3;                      // TMEM lane quadrant this warp may read
    const int row_in_tile = quad * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile % p.m_tiles;
      const int n_blk = tile / p.m_tiles;
      long long row;
      bool valid;
      int sample;
      if (p.mode == 0) {
        row = (long long)m_blk * kBlockM + row_in_tile;
        valid = row < p.M;
        sample = p.rowbias ? (int)(row / p.rows_per_batch) : 0;
      } else {
        const int img = m_blk / p.tiles_per_img;
        const int r = m_blk - img * p.tiles_per_img;
        const int y = (r / p.tiles_x) * p.bh + row_in_tile / p.bw;
        const int x = (r % p.tiles_x) * p.bw + row_in_tile % p.bw;
        valid = (y < p.out_h) && (x < p.out_w);
        row = ((long long)img * p.out_h + y) * p.out_w + x;
        sample = img;
      }
      mbar_wait(&bars->tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * kAccCols;
      const float bias_m = (p.bias && p.bias_along_m && valid) ? p.bias[row] : 0.f;
      for (int c = 0; c < p.block_n; c += 32) {
        uint32_t v[32];
        tmem_ld32(t_row + c, v);
        tmem_ld_wait();
        const int n0 = n_blk * p.block_n + c;
        if (!valid || n0 >= p.N) continue;
        if (p.act == 2) {
          // GEGLU: columns interleaved (h, g) -> 16 outputs
          const int o0 = n0 >> 1;
          float o[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int n = n0 + 2 * j;
            float h = __uint_as_float(v[2 * j]) * p.alpha;
            float g = __uint_as_float(v[2 * j + 1]) * p.alpha;
            if (n < p.N) {
              if (p.bias) { h += __ldg(p.bias + n); g += __ldg(p.bias + n + 1); }
            }
            o[j] = h * gelu_erf_f(g);
          }
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ld_out + o0;
#pragma unroll
          for (int g8 = 0; g8 < 2; ++g8) {
            if (n0 + g8 * 16 >= p.N) break;
            uint4 st;
            st.x = pack_bf16(o[g8 * 8 + 0], o[g8 * 8 + 1]);
            st.y = pack_bf16(o[g8 * 8 + 2], o[g8 * 8 + 3]);
            st.z = pack_bf16(o[g8 * 8 + 4], o[g8 * 8 + 5]);
            st.w = pack_bf16(o[g8 * 8 + 6], o[g8 * 8 + 7]);
            *reinterpret_cast<uint4*>(op + g8 * 8) = st;
          }
          continue;
        }
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const int n = n0 + g8 * 8;
          if (n >= p.N) break;
          float x[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[g8 * 8 + j]) * p.alpha;
          if (p.bias) {
            if (p.bias_along_m) {
#pragma unroll
              for (int j = 0; j < 8; ++j) x[j] += bias_m;
            } else {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
              x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
              x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
            }
          }
          if (p.rowbias) {
            const float* rb = p.rowbias + (long long)sample * p.ld_rowbias + n;
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(rb));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(rb + 4));
            x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
            x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
          }
          if (p.act == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = silu_f(x[j]);
          } else if (p.act == 3) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = quick_gelu_f(x[j]);
          }
          if (p.residual) {
            const uint4 r = __ldg(reinterpret_cast<const uint4*>(p.residual + row * p.ld_res + n));
            float2 f;
            f = unpack_bf16(r.x); x[0] += f.x; x[1] += f.y;
            f = unpack_bf16(r.y); x[2] += f.x; x[3] += f.y;
            f = unpack_bf16(r.z); x[4] += f.x; x[5] += f.y;
            f = unpack_bf16(r.w); x[6] += f.x; x[7] += f.y;
          }
          if (p.out_fp32) {
            float* op = reinterpret_cast<float*>(p.out) + row * p.ld_out + n;
            *reinterpret_cast<float4*>(op) = make_float4(x[0], x[1], x[2], x[3]);
            *reinterpret_cast<float4*>(op + 4) = make_float4(x[4], x[5], x[6], x[7]);
          } else {
            uint4 st;
            st.x = pack_bf16(x[0], x[1]); st.y = pack_bf16(x[2], x[3]);
            st.z = pack_bf16(x[4], x[5]); st.w = pack_bf16(x[6], x[7]);
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ld_out + n) = st;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------- host side
static int pick_block_n(long long m_tiles, long long N, int sms) {
  int best = 32;
  double best_cost = 1e30;
  for (int bn = 256; bn >= 32; bn -= 32) {
    const long long tiles = m_tiles * ((N + bn - 1) / bn);
    const long long waves = (tiles + sms - 1) / sms;
    const double per_tile = fmax(bn / 2.0, 32.0 + bn / 4.0) + 24.0;   // MMA-bound vs smem-bound + fixed overhead
    const double cost = waves * per_tile;
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

static void pick_patch(int H, int W, int* bw, int* bh) {
  double best = 1e30;
  for (int w = 128; w >= 8; w >>= 1) {
    const int h = 128 / w;
    const double waste = (double)((W + w - 1) / w * w) * ((H + h - 1) / h * h) / ((double)W * H);
    if (waste < best - 1e-9) { best = waste; *bw = w; *bh = h; }
  }
}

}  // namespace fd

using namespace fd;

extern "C" int fd_gemm(const fd_gemm_desc* d, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d && d->a[0] && d->b && d->out, "fd_gemm: null operand");
  FD_REQUIRE(d->mode >= 0 && d->mode <= 2, "fd_gemm: bad mode %d", d->mode);
  GemmKParams p;
  memset(&p, 0, sizeof(p));
  const int nsrc = d->a[1] ? 2 : 1;
  const long long c0 = d->a_cols[0], c1 = nsrc == 2 ? d->a_cols[1] : 0;
  const long long ctot = c0 + c1;
  FD_REQUIRE(ctot % 8 == 0 && c0 % 8 == 0, "fd_gemm: K/channels per source must be multiples of 8 (got %lld,%lld)", c0, c1);
  FD_REQUIRE(nsrc == 1 || c0 % kBlockK == 0, "fd_gemm: first of two sources must have a multiple of 64 columns");
  FD_REQUIRE(d->mode != 2 || nsrc == 1, "fd_gemm: stride-2 conv takes one source");
  FD_REQUIRE(d->N % 8 == 0, "fd_gemm: N must be a multiple of 8 (got %lld)", (long long)d->N);
  FD_REQUIRE(d->act != 2 || (d->N % 16 == 0 && !d->out_fp32 && !d->residual && !d->rowbias),
             "fd_gemm: GEGLU needs N%%16==0, bf16 out, no residual/rowbias");
  for (int s = 0; s < nsrc; ++s)
    FD_REQUIRE(((uintptr_t)d->a[s] % 16) == 0 && d->a_ld[s] % 8 == 0, "fd_gemm: A source %d must be 16B aligned with pitch%%8==0", s);
  FD_REQUIRE(((uintptr_t)d->b % 16) == 0 && ((uintptr_t)d->out % 16) == 0 && d->ld_out % 8 == 0, "fd_gemm: B/out alignment");
  FD_REQUIRE(!d->residual || (((uintptr_t)d->residual % 16) == 0 && d->ld_res % 8 == 0), "fd_gemm: residual alignment");

  p.mode = d->mode == 0 ? 0 : 1;
  p.N = (int)d->N;
  p.taps = d->mode == 0 ? 1 : 9;
  p.ctot = (int)ctot;
  p.kb_per_tap = (int)((c0 + kBlockK - 1) / kBlockK + (c1 + kBlockK - 1) / kBlockK);
  p.kb_src0 = (int)((c0 + kBlockK - 1) / kBlockK);
  p.two_src = nsrc == 2;
  const long long ktot = (long long)p.taps * ctot;

  long long M;
  if (d->mode == 0) {
    M = d->M;
    p.m_tiles = (int)((M + kBlockM - 1) / kBlockM);
  } else {
    const int H = d->in_h, W = d->in_w;
    FD_REQUIRE(d->batch > 0 && H > 0 && W > 0 && (d->pad_lo == 0 || d->pad_lo == 1), "fd_gemm: bad conv geometry");
    if (d->mode == 1) {
      FD_REQUIRE(d->pad_lo == 1, "fd_gemm: stride-1 conv supports pad 1 only");
      p.out_h = H; p.out_w = W;
    } else {
      p.out_h = d->pad_lo ? (H - 1) / 2 + 1 : (H - 2) / 2 + 1;
      p.out_w = d->pad_lo ? (W - 1) / 2 + 1 : (W - 2) / 2 + 1;
    }
    pick_patch(p.out_h, p.out_w, &p.bw, &p.bh);
    p.tiles_x = (p.out_w + p.bw - 1) / p.bw;
    p.tiles_per_img = p.tiles_x * ((p.out_h + p.bh - 1) / p.bh);
    p.m_tiles = p.tiles_per_img * d->batch;
    M = (long long)d->batch * p.out_h * p.out_w;
    for (int t = 0; t < 9; ++t) {
      const int r = t / 3, s = t % 3;
      if (d->mode == 1) {
        p.tap_map[t] = 0; p.tap_dy[t] = r - 1; p.tap_dx[t] = s - 1;
      } else {
        const int ty = r - d->pad_lo, tx = s - d->pad_lo;
        const int py = ((ty % 2) + 2) % 2, px = ((tx % 2) + 2) % 2;
        p.tap_map[t] = py * 2 + px; p.tap_dy[t] = (ty - py) / 2; p.tap_dx[t] = (tx - px) / 2;
      }
    }
  }
  p.M = (int)M;
  FD_REQUIRE(M > 0 && M < (1ll << 31), "fd_gemm: bad M");
  const int sms = num_sms();
  p.block_n = d->block_n ? d->block_n : pick_block_n(p.m_tiles, d->N, sms);
  FD_REQUIRE(p.block_n % 32 == 0 && p.block_n >= 32 && p.block_n <= 256, "fd_gemm: bad block_n %d", p.block_n);
  p.n_tiles = (int)((d->N + p.block_n - 1) / p.block_n);
  const int stage_bytes = kATileBytes + p.block_n * kBlockK * 2;
  const int smem_budget = 227 * 1024 - (int)sizeof(PipeBarriers) - 64;
  p.stages = smem_budget / stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  const size_t smem_bytes = (size_t)p.stages * stage_bytes + sizeof(PipeBarriers);

  // tensor maps
  if (d->mode == 0) {
    for (int s = 0; s < nsrc; ++s) {
      uint64_t dims[2] = {(uint64_t)d->a_cols[s], (uint64_t)M};
      uint64_t str[1] = {(uint64_t)d->a_ld[s] * 2};
      uint32_t box[2] = {kBlockK, kBlockM};
      if (encode_tensor_map(&p.amap[s], d->a[s], 2, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                            CU_TENSOR_MAP_SWIZZLE_128B)) return 3;
    }
  } else if (d->mode == 1) {
    for (int s = 0; s < nsrc; ++s) {
      const uint64_t ld = (uint64_t)d->a_ld[s];
      uint64_t dims[4] = {(uint64_t)d->a_cols[s], (uint64_t)d->in_w, (uint64_t)d->in_h, (uint64_t)d->batch};
      uint64_t str[3] = {ld * 2, ld * 2 * d->in_w, ld * 2 * d->in_w * d->in_h};
      uint32_t box[4] = {kBlockK, (uint32_t)p.bw, (uint32_t)p.bh, 1};
      if (encode_tensor_map(&p.amap[s], d->a[s], 4, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                            CU_TENSOR_MAP_SWIZZLE_128B)) return 3;
    }
  } else {
    const uint64_t ld = (uint64_t)d->a_ld[0];
    for (int py = 0; py < 2; ++py)
      for (int px = 0; px < 2; ++px) {
        const int hh = (d->in_h - py + 1) / 2, ww = (d->in_w - px + 1) / 2;
        if (hh <= 0 || ww <= 0) { p.amap[py * 2 + px] = p.amap[0]; continue; }
        const uint8_t* base = (const uint8_t*)d->a[0] + ((uint64_t)py * d->in_w + px) * ld * 2;
        uint64_t dims[4] = {(uint64_t)c0, (uint64_t)ww, (uint64_t)hh, (uint64_t)d->batch};
        uint64_t str[3] = {ld * 4, ld * 4 * d->in_w, ld * 2 * d->in_w * d->in_h};
        uint32_t box[4] = {kBlockK, (uint32_t)p.bw, (uint32_t)p.bh, 1};
        if (encode_tensor_map(&p.amap[py * 2 + px], base, 4, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                              CU_TENSOR_MAP_SWIZZLE_128B)) return 3;
      }
  }
  {
    uint64_t dims[2] = {(uint64_t)ktot, (uint64_t)d->N};
    uint64_t str[1] = {(uint64_t)ktot * 2};
    uint32_t box[2] = {kBlockK, (uint32_t)p.block_n};
    if (encode_tensor_map(&p.bmap, d->b, 2, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                          CU_TENSOR_MAP_SWIZZLE_128B)) return 3;
  }
  p.bias = d->bias; p.rowbias = d->rowbias; p.rows_per_batch = d->rows_per_batch > 0 ? d->rows_per_batch : M;
  p.ld_rowbias = d->ld_rowbias > 0 ? d->ld_rowbias : d->N;
  FD_REQUIRE(!d->rowbias || (((uintptr_t)d->rowbias % 16) == 0 && p.ld_rowbias % 4 == 0), "fd_gemm: rowbias alignment");
  p.residual = (const __nv_bfloat16*)d->residual; p.ld_res = d->ld_res;
  p.out = d->out; p.ld_out = d->ld_out; p.alpha = d->alpha == 0.f ? 1.f : d->alpha;
  p.act = d->act; p.bias_along_m = d->bias_along_m; p.out_fp32 = d->out_fp32;

  static bool attr_set = false;
  if (!attr_set) {
    FD_CHECK_CUDA(cudaFuncSetAttribute(gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < sms ? tiles : sms;
  gemm_tcgen05_kernel<<<grid, kThreads, smem_bytes, stream>>>(p);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}
