// Persistent, warp-specialised tcgen05 implicit-GEMM for sm_100a.
//   warp 0 lane 0 : TMA producer   (A tile 128x64 bf16 + B tile block_n x 64 bf16 per stage, 128B swizzle)
//   warp 1 lane 0 : MMA issuer     (tcgen05.mma cta_group::1 kind::f16, M=128, N=block_n, K=16; fp32 accum in TMEM)
//   warps 2..5    : epilogue       (tcgen05.ld -> bias / timestep row-bias / activation / GEGLU / residual -> global)
// TMEM holds two accumulator stages (2 x 256 columns) so the epilogue of tile i overlaps the mainloop of tile i+1.
// A operand is either a row-major matrix (2-D tensor map) or an NHWC image read through 4-D tensor maps with the
// 3x3 filter taps expressed as shifted TMA boxes (out-of-bounds = zero fill = the conv's zero padding);
// stride-2 convs use four phase-subsampled tensor maps of the same image (no data rearrangement).
// Thread-block clusters (CM x CN CTAs, up to 2x2) share operand tiles through TMA multicast: the CN CTAs of one tile
// row each fetch 1/CN of the A tile and multicast it to the row, the CM CTAs of one tile column each fetch 1/CM of the B
// tile — per-SM L2->SMEM traffic per k-block drops from 16K+N*128 B to 16K/CN + N*128/CM B, which is what bounds a
// 1-CTA tile (64 B/clk/SM from L2 feeds only ~55% of the tensor pipe at 128x160).
#include "common.cuh"
#include "../../include/faithdiff_b200.h"

namespace fd {
extern int64_t g_launches;

struct GemmKParams {
  CUtensorMap amap[4];
  CUtensorMap bmap;
  CUtensorMap rmap;    // residual, box = one epilogue chunk (32 rows x 32 columns, SWIZZLE_64B)
  CUtensorMap omap;    // output,   box = one epilogue chunk (32 rows x 32 columns; GEGLU: x 16 columns, SWIZZLE_32B)
  int mode;            // 0 gemm, 1 conv
  int M, N;
  int m_tiles, n_tiles, block_n, stages;
  int taps, kb_per_tap, kb_src0, ctot, two_src;
  int tap_map[9], tap_dy[9], tap_dx[9];
  int out_h, out_w, bw, bh, tiles_x, tiles_per_img, batch;
  int cm, cn, m_super, n_super;      // cluster shape and super-tile grid
  int group_m;                       // tile rasterisation: concurrent CTAs cover group_m x (148/group_m) tiles
  int pair;                          // 1: CTA pair (cta_group::2, UMMA M=256); cm = 2, cn = 1, B tile split across the pair
  int a_part_dy, a_part_dx;          // conv: pixel offset of A part j inside the tile patch (split along y or x)
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
  int tma_epi;         // 1: the epilogue moves residual and output through shared memory with TMA (bf16 output, 16 B aligned)
  int slot_bytes;      // bytes of one epilogue staging slot (a 32-row chunk of the output)
  // LayerNorm folding (fd_gemm_ln, opt-in).  kLn = 2 (producer): per-row (sum, sum of squares) of the bf16-rounded output, one
  // float2 slot per (n tile, chunk half).  kLn = 1 (consumer, A = the raw residual stream, B = W*gamma): the epilogue computes
  // rstd_r * acc - rstd_r*mean_r * colsum[n] + bias'[n] from the producer's slots, i.e. LayerNorm(x) W^T + b without the
  // LayerNorm pass.
  const float2* ln_stats_in;
  int ln_in_ld, ln_in_n;
  const float* ln_colsum;
  float ln_inv_c, ln_eps;
  float2* ln_stats_out;
  int ln_out_ld;
  // GroupNorm statistics from the producer (kLn = 3): per (32-row epilogue box, column pair) sum / sum of squares of the
  // bf16-rounded outputs, float2 [4 * m_tiles, gn_ld]; the consumer GroupNorm becomes a single streaming pass (norm.cu)
  float2* gn_stats_out;
  int gn_ld;
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kATileBytes = kBlockM * kBlockK * 2;   // 16 KB
constexpr int kAccCols = 256;                        // TMEM columns per accumulator stage
constexpr int kMaxStages = 8;
constexpr int kThreads = 320;                        // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quadrant)

struct __align__(16) PipeBarriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint64_t res_full[8][2];     // per epilogue warp and staging slot: residual chunk landed
  uint32_t tmem_base;
  uint32_t pad_[3];
  float bias_s[8][128];        // per epilogue warp: the bias of the (up to four) 32-column chunks it owns in the current tile
};

// Grouped rasterisation of the linear (super-)tile index: consecutive indices walk down a band of `group_m` tile rows before
// moving to the next tile column, so the ~148 concurrently running CTAs cover a roughly square patch of the output: every
// A and B tile is then requested by ~12 CTAs at a time instead of one operand being hammered by all of them (L2 hot spot).
__device__ __forceinline__ void raster_tile(int st, int m_super, int n_super, int group_m, int& mp, int& np) {
  const int per_group = group_m * n_super;
  const int g = st / per_group;
  const int first_m = g * group_m;
  const int rows = min(m_super - first_m, group_m);
  const int r = st - g * per_group;
  mp = first_m + r % rows;
  np = r / rows;
}

// kPair = true instantiates the cta_group::2 (CTA-pair, UMMA M=256) variant; a kernel containing cta_group::2
// instructions can only be launched with an even cluster size, hence two instantiations.
template <bool kPair, int kEpi, int kLn = 0>
__global__ void __launch_bounds__(kThreads, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmKParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b_tile_bytes = (kPair ? p.block_n / 2 : p.block_n) * kBlockK * 2;   // pair: each CTA holds half of B
  const int stage_bytes = kATileBytes + b_tile_bytes;
  uint8_t* slots = smem + (size_t)p.stages * stage_bytes;          // 8 warps x 2 slots (1024-aligned: stage_bytes % 2048 == 0)
  PipeBarriers* bars = reinterpret_cast<PipeBarriers*>(slots + ((kEpi != 0 || p.tma_epi) ? 16 * p.slot_bytes : 0));

  const int csize = p.cm * p.cn;
  uint32_t crank = 0;
  if (csize > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const int ci = crank % p.cm, cj = crank / p.cm;          // position of this CTA inside the cluster (m, n); pair: ci = rank
  uint16_t amask = 0, bmask = 0;                              // CTAs sharing my A tile (same ci) / my B tile (same cj)
  for (int j = 0; j < p.cn; ++j) amask |= (uint16_t)(1u << (ci + j * p.cm));
  for (int i = 0; i < p.cm; ++i) bmask |= (uint16_t)(1u << (i + cj * p.cm));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.bmap);
    tma_prefetch_desc(&p.amap[0]);
    if (kEpi != 0 || p.tma_epi) { tma_prefetch_desc(&p.omap); if (p.residual) tma_prefetch_desc(&p.rmap); }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], kPair ? 1 : p.cm + p.cn - 1);   // every CTA that multicasts into this stage must see it freed
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->tmem_full[a], 1);
      mbar_init(&bars->tmem_empty[a], kPair ? 16 : 8);    // 8 epilogue warps (pair: leader collects both CTAs')
    }
    for (int w = 0; w < 8; ++w) { mbar_init(&bars->res_full[w][0], 1); mbar_init(&bars->res_full[w][1], 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (kPair) { tmem_alloc2(&bars->tmem_base, 512); tmem_relinquish2(); }
    else { tmem_alloc(&bars->tmem_base, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if (csize > 1) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  pdl_launch_dependents();     // the next kernel may begin its own prologue on SMs this grid no longer needs
  pdl_wait();                  // ... and this one touches global memory only after its predecessor has fully completed

  const int num_super = p.m_super * p.n_super;
  const int num_clusters = gridDim.x / csize;
  const int cluster_id = blockIdx.x / csize;
  const int num_kb = p.taps * p.kb_per_tap;
  const int a_part_bytes = kATileBytes / p.cn, b_part_bytes = kPair ? b_tile_bytes : b_tile_bytes / p.cm;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------------------------ TMA producer
      int stage = 0;
      uint32_t phase = 0;
      for (int st = cluster_id; st < num_super; st += num_clusters) {
        int mp_, np_;
        raster_tile(st, p.m_super, p.n_super, p.group_m, mp_, np_);
        const int m_blk = mp_ * p.cm + ci;
        const int n_blk = np_ * p.cn + cj;
        int img = 0, y0 = 0, x0 = 0;
        if (p.mode != 0) {
          img = m_blk / p.tiles_per_img;
          const int r = m_blk - img * p.tiles_per_img;
          y0 = (r / p.tiles_x) * p.bh + cj * p.a_part_dy;     // my 1/CN part of the A patch
          x0 = (r % p.tiles_x) * p.bw + cj * p.a_part_dx;
        }
        for (int tap = 0; tap < p.taps; ++tap) {
          for (int kb = 0; kb < p.kb_per_tap; ++kb) {
            mbar_wait(&bars->empty[stage], phase ^ 1);
            uint8_t* sa = smem + (size_t)stage * stage_bytes;
            uint8_t* sb = sa + kATileBytes;
            if constexpr (kPair) {
              // both CTAs' bytes complete on the leader's barrier; only the leader arms it
              const uint32_t lead_bar = mapa_u32(smem_u32(&bars->full[stage]), 0);
              if (crank == 0) mbar_arrive_expect_tx(&bars->full[stage], 2u * (uint32_t)stage_bytes);
              const int c0p = kb * kBlockK;
              const int kcolp = tap * p.ctot + kb * kBlockK;
              if (p.mode == 0)
                tma_load_2d_cg2(&p.amap[p.two_src ? (kb >= p.kb_src0) : 0], lead_bar, sa,
                                p.two_src && kb >= p.kb_src0 ? (kb - p.kb_src0) * kBlockK : c0p, m_blk * kBlockM);
              else
                tma_load_4d_cg2(&p.amap[p.two_src ? (kb >= p.kb_src0) : p.tap_map[tap]], lead_bar, sa,
                                p.two_src && kb >= p.kb_src0 ? (kb - p.kb_src0) * kBlockK : c0p, x0 + p.tap_dx[tap],
                                y0 + p.tap_dy[tap], img);
              tma_load_2d_cg2(&p.bmap, lead_bar, sb, kcolp, n_blk * p.block_n + (int)crank * (p.block_n / 2));
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
              continue;
            }
            mbar_arrive_expect_tx(&bars->full[stage], (uint32_t)stage_bytes);
            int src, c0;
            if (p.two_src) {
              src = kb >= p.kb_src0;
              c0 = (src ? kb - p.kb_src0 : kb) * kBlockK;
            } else {
              src = p.tap_map[tap];
              c0 = kb * kBlockK;
            }
            if (p.cn == 1) {
              if (p.mode == 0)
                tma_load_2d(&p.amap[src], &bars->full[stage], sa, c0, m_blk * kBlockM);
              else
                tma_load_4d(&p.amap[src], &bars->full[stage], sa, c0, x0 + p.tap_dx[tap], y0 + p.tap_dy[tap], img);
            } else {
              uint8_t* dst = sa + cj * a_part_bytes;
              if (p.mode == 0)
                tma_load_2d_mc(&p.amap[src], &bars->full[stage], dst, c0, m_blk * kBlockM + cj * (kBlockM / p.cn), amask);
              else
                tma_load_4d_mc(&p.amap[src], &bars->full[stage], dst, c0, x0 + p.tap_dx[tap], y0 + p.tap_dy[tap], img, amask);
            }
            const int kcol = tap * p.ctot + kb * kBlockK;
            if (p.cm == 1)
              tma_load_2d(&p.bmap, &bars->full[stage], sb, kcol, n_blk * p.block_n);
            else
              tma_load_2d_mc(&p.bmap, &bars->full[stage], sb + ci * b_part_bytes, kcol,
                             n_blk * p.block_n + ci * (p.block_n / p.cm), bmask);
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && !(kPair && crank != 0)) {
      // ------------------------------------------------------------------ MMA issuer (single thread; pair: leader CTA only)
      const uint32_t idesc = umma_idesc(kPair ? 2 * kBlockM : kBlockM, p.block_n, /*bf16*/ 1, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      const uint16_t cmask = amask | bmask;
      for (int st = cluster_id; st < num_super; st += num_clusters) {
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
            if constexpr (kPair) umma_f16_cg2(d_tmem, adesc, bdesc, idesc, (kb | k) != 0);
            else umma_f16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0);
          }
          // frees this smem stage (in every CTA that multicasts into it / in both CTAs of a pair) once the MMAs above retire
          if constexpr (kPair) tc_commit_cg2_mc(&bars->empty[stage], 3);
          else { if (csize == 1) tc_commit(&bars->empty[stage]); else tc_commit_mc(&bars->empty[stage], cmask); }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if constexpr (kPair) tc_commit_cg2_mc(&bars->tmem_full[acc], 3);   // accumulator complete -> both CTAs' epilogues
        else tc_commit(&bars->tmem_full[acc]);
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // -------------------------------------------------------------------- epilogue warps (2..9)
    const int quad = warp & 3;                      // TMEM lane quadrant this warp may read
    const int chalf = (warp - 2) >> 2;              // the two warps of a quadrant alternate 32-column chunks
    const int row_in_tile = quad * 32 + lane;
    const int nchunks = p.block_n >> 5;
    // kEpi != 0 fixes the flag combination at compile time (the generic epilogue is ~2400 SASS instructions per chunk with
    // a taken branch around every unused option; the profile showed the epilogue warps stalled on instruction fetch)
    const int act = kEpi == 1 ? 0 : (kEpi == 2 ? 2 : p.act);
    const bool tma_epi = kEpi != 0 || p.tma_epi;
    const bool bias_m_mode = kEpi == 0 && p.bias_along_m;
    const bool has_res = kEpi != 2 && p.residual != nullptr;
    const bool col_bias = p.bias && !bias_m_mode;
    float* bias_w = bars->bias_s[warp - 2];
    float* csum_w = reinterpret_cast<float*>(bars + 1) + (warp - 2) * 128;   // kLn == 1: colsum strip (the host reserves these 4 KB)
    // TMA epilogue state: two staging slots per warp, used alternately by the running chunk sequence number `qseq`
    uint8_t* slot0 = slots + (size_t)(warp - 2) * 2 * p.slot_bytes;
    uint64_t* rbar = bars->res_full[warp - 2];
    uint32_t qseq = 0;
    const int wx = p.mode ? (quad * 32) % p.bw : 0, wy = p.mode ? (quad * 32) / p.bw : 0;   // conv: this warp's 32 rows inside the tile's pixel patch
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int st = cluster_id; st < num_super; st += num_clusters) {
      int mp_, np_;
      raster_tile(st, p.m_super, p.n_super, p.group_m, mp_, np_);
      const int m_blk = mp_ * p.cm + ci;
      const int n_blk = np_ * p.cn + cj;
      long long row;
      bool valid;
      int sample;
      int bx = 0, by = 0, bimg = 0;                 // conv: TMA box origin of this warp's rows
      if (p.mode == 0) {
        row = (long long)m_blk * kBlockM + row_in_tile;
        valid = row < p.M;
        sample = (p.rowbias && valid) ? (int)(row / p.rows_per_batch) : 0;
      } else {
        const int img = m_blk / p.tiles_per_img;
        const int r = m_blk - img * p.tiles_per_img;
        const int ty = (r / p.tiles_x) * p.bh, tx = (r % p.tiles_x) * p.bw;
        const int y = ty + row_in_tile / p.bw;
        const int x = tx + row_in_tile % p.bw;
        valid = (img < p.batch) && (y < p.out_h) && (x < p.out_w);
        row = ((long long)img * p.out_h + y) * p.out_w + x;
        sample = valid ? img : 0;
        bx = tx + wx; by = ty + wy; bimg = img;
      }
      const int ncol0 = n_blk * p.block_n;
      const int box_row0 = m_blk * kBlockM + quad * 32;   // gemm mode: first of this warp's 32 rows

      // ---- residual prefetch (two chunks ahead; the first two are requested before the accumulator is even ready)
      //  TMA path:    one 32x32 box per chunk into the staging slot the result will later be stored from
      //  direct path: each thread reads its own row (16 B per load) into registers
      auto issue_res_load = [&](int c, uint32_t q) {           // one lane
        uint64_t* bar = &rbar[q & 1];
        uint8_t* dst = slot0 + (q & 1) * p.slot_bytes;
        mbar_arrive_expect_tx(bar, 2048);
        if (p.mode == 0) tma_load_2d(&p.rmap, bar, dst, ncol0 + c * 32, box_row0);
        else tma_load_4d(&p.rmap, bar, dst, ncol0 + c * 32, bx, by, bimg);
      };
      const __nv_bfloat16* res_row = (!tma_epi && has_res && valid) ? p.residual + row * p.ld_res : nullptr;
      uint4 r0[4], r1[4];
      auto load_res = [&](int c, uint4 (&dst)[4]) {
#pragma unroll
        for (int g8 = 0; g8 < 4; ++g8) {
          const int n = ncol0 + c * 32 + g8 * 8;
          dst[g8] = (res_row && n < p.N) ? __ldg(reinterpret_cast<const uint4*>(res_row + n)) : make_uint4(0u, 0u, 0u, 0u);
        }
      };
      if (has_res) {
        if (tma_epi) {
          if (lane == 0) {
            bulk_wait_read<0>();                    // my earlier stores have finished reading both slots
            if (chalf < nchunks && ncol0 + chalf * 32 < p.N) issue_res_load(chalf, qseq);
            if (chalf + 2 < nchunks && ncol0 + (chalf + 2) * 32 < p.N) issue_res_load(chalf + 2, qseq + 1);
          }
        } else {
          if (chalf < nchunks) load_res(chalf, r0);
          if (chalf + 2 < nchunks) load_res(chalf + 2, r1);
        }
      }
      // Column bias: one coalesced read per chunk into this warp's shared-memory strip while the accumulator is still being
      // computed; the per-chunk code then takes it with broadcast LDS.128.  (Read with LDG inside the chunk loop, the eight
      // dependent L1/L2 round trips per chunk cost 6-10 us per GEMM.)
      if (col_bias) {
        __syncwarp();
        int k = 0;
        for (int c = chalf; c < nchunks; c += 2, ++k) {
          const int n = ncol0 + c * 32 + lane;
          bias_w[k * 32 + lane] = n < p.N ? __ldg(p.bias + n) : 0.f;
        }
        __syncwarp();
      }
      float ln_rstd = 0.f, ln_nrm = 0.f;            // kLn == 1: rstd and -rstd*mean of this thread's row
      float ln_s1 = 0.f, ln_s2 = 0.f;               // kLn == 2: sum / sum of squares of this thread's row over its chunks of the tile
      if constexpr (kLn == 1) {
        __syncwarp();
        int k = 0;
        for (int c = chalf; c < nchunks; c += 2, ++k) {
          const int n = ncol0 + c * 32 + lane;
          csum_w[k * 32 + lane] = n < p.N ? __ldg(p.ln_colsum + n) : 0.f;
        }
        __syncwarp();
        if (valid) {
          const float2* sp = p.ln_stats_in + row * p.ln_in_ld;
          float s1 = 0.f, s2 = 0.f;
          for (int k2 = 0; k2 < p.ln_in_n; ++k2) {
            const float2 t = __ldg(sp + k2);
            s1 += t.x;
            s2 += t.y;
          }
          const float mean = s1 * p.ln_inv_c;
          const float var = fmaxf(fmaf(s2, p.ln_inv_c, -mean * mean), 0.f);
          ln_rstd = rsqrtf(var + p.ln_eps);
          ln_nrm = -ln_rstd * mean;
        }
      }
      mbar_wait(&bars->tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * kAccCols;
      const float bias_m = (bias_m_mode && p.bias && valid) ? p.bias[row] : 0.f;
      const float* rb_row = (kEpi != 2 && p.rowbias) ? p.rowbias + (long long)sample * p.ld_rowbias : nullptr;

      // alpha, bias, per-sample row bias and activation on 8 consecutive columns n..n+7 (group g8 of the chunk)
      auto affine_act8 = [&](float (&x)[8], const uint32_t* v8, const float* bs8, int n) {
        if constexpr (kLn == 1) {
          const float* cs8 = csum_w + (bs8 - bias_w);
          const float4 c0 = *reinterpret_cast<const float4*>(cs8);
          const float4 c1 = *reinterpret_cast<const float4*>(cs8 + 4);
          const float cs[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = fmaf(__uint_as_float(v8[j]), ln_rstd, ln_nrm * cs[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v8[j]) * p.alpha;
        }
        if (p.bias) {
          if (bias_m_mode) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] += bias_m;
          } else {
            const float4 b0 = *reinterpret_cast<const float4*>(bs8);
            const float4 b1 = *reinterpret_cast<const float4*>(bs8 + 4);
            x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
            x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
          }
        }
        if (rb_row) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(rb_row + n));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(rb_row + n + 4));
          x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
          x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
        }
        if (act == 1) {
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = silu_f(x[j]);
        } else if (act == 3) {
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = quick_gelu_f(x[j]);
        } else if (act == 4) {
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = gelu_erf_f(x[j]);
        } else if (act == 5) {                       // LeakyReLU(0.2): BSRNet dense blocks / upsampling convs
#pragma unroll
          for (int j = 0; j < 8; ++j) x[j] = fmaxf(x[j], 0.2f * x[j]);
        }
      };
      // GEGLU: the chunk's columns are interleaved (h, g) pairs -> 16 outputs h * gelu(g)
      auto geglu16 = [&](float (&o)[16], const uint32_t (&v)[32], const float* bs) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float h, g;
          if constexpr (kLn == 1) {
            const float2 cs = *reinterpret_cast<const float2*>(csum_w + (bs - bias_w) + 2 * j);
            h = fmaf(__uint_as_float(v[2 * j]), ln_rstd, ln_nrm * cs.x);
            g = fmaf(__uint_as_float(v[2 * j + 1]), ln_rstd, ln_nrm * cs.y);
          } else {
            h = __uint_as_float(v[2 * j]) * p.alpha;
            g = __uint_as_float(v[2 * j + 1]) * p.alpha;
          }
          if (col_bias) {
            const float2 b = *reinterpret_cast<const float2*>(bs + 2 * j);
            h += b.x; g += b.y;
          }
          o[j] = h * gelu_erf_f(g);
        }
      };
      auto release_acc = [&]() {                    // all of this warp's TMEM reads of the tile are done
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kPair && crank != 0) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->tmem_empty[acc]), 0));
          else mbar_arrive(&bars->tmem_empty[acc]);
        }
      };

      if (tma_epi) {
        // ---------------------------------------------------------------- TMA epilogue
        // Every thread owns one output row.  Written straight to global memory that is 32 different 128-byte lines per
        // warp instruction, and the LSU's one-line-per-cycle pace made the epilogue (not the MMA) the bound of every
        // GEMM here.  Instead the row's 64-byte piece goes to a swizzled (conflict-free) staging slot and one TMA store
        // per chunk writes the 32x32 box as full sectors; the residual comes in the same way, into the same slot.
        const uint32_t sw64 = (uint32_t)((lane >> 1) & 3), sw32 = (uint32_t)((lane >> 2) & 1);
        for (int c = chalf; c < nchunks; c += 2) {
          const int n0 = ncol0 + c * 32;
          if (n0 >= p.N) break;
          uint32_t v[32];
          tmem_ld32(t_row + c * 32, v);
          tmem_ld_wait();
          if (c + 2 >= nchunks || n0 + 64 >= p.N) release_acc();
          uint8_t* slot = slot0 + (qseq & 1) * p.slot_bytes;
          if (has_res) {
            mbar_wait(&rbar[qseq & 1], (qseq >> 1) & 1);
          } else {
            if (lane == 0) bulk_wait_read<1>();     // the store issued from this slot two chunks ago has read it
            __syncwarp();
          }
          const float* bs = bias_w + ((c - chalf) >> 1) * 32;
          if (kEpi != 3 && act == 2) {
            float o[16];
            geglu16(o, v, bs);
            uint8_t* rowp = slot + lane * 32;
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              uint4 stv;
              stv.x = pack_bf16(o[g * 8 + 0], o[g * 8 + 1]); stv.y = pack_bf16(o[g * 8 + 2], o[g * 8 + 3]);
              stv.z = pack_bf16(o[g * 8 + 4], o[g * 8 + 5]); stv.w = pack_bf16(o[g * 8 + 6], o[g * 8 + 7]);
              *reinterpret_cast<uint4*>(rowp + ((g ^ sw32) << 4)) = stv;
            }
          } else {
            uint8_t* rowp = slot + lane * 64;
#pragma unroll
            for (int g8 = 0; g8 < 4; ++g8) {
              float x[8];
              affine_act8(x, v + g8 * 8, bs + g8 * 8, n0 + g8 * 8);
              uint4* sp = reinterpret_cast<uint4*>(rowp + ((g8 ^ sw64) << 4));
              if (has_res) {
                const uint4 rv = *sp;
                float2 f;
                f = unpack_bf16(rv.x); x[0] += f.x; x[1] += f.y;
                f = unpack_bf16(rv.y); x[2] += f.x; x[3] += f.y;
                f = unpack_bf16(rv.z); x[4] += f.x; x[5] += f.y;
                f = unpack_bf16(rv.w); x[6] += f.x; x[7] += f.y;
              }
              uint4 stv;
              stv.x = pack_bf16(x[0], x[1]); stv.y = pack_bf16(x[2], x[3]);
              stv.z = pack_bf16(x[4], x[5]); stv.w = pack_bf16(x[6], x[7]);
              *sp = stv;
              if constexpr (kLn == 2) {             // statistics of what the consumer will read: the ROUNDED values
                float2 f;
                f = unpack_bf16(stv.x); ln_s1 += f.x + f.y; ln_s2 = fmaf(f.x, f.x, fmaf(f.y, f.y, ln_s2));
                f = unpack_bf16(stv.y); ln_s1 += f.x + f.y; ln_s2 = fmaf(f.x, f.x, fmaf(f.y, f.y, ln_s2));
                f = unpack_bf16(stv.z); ln_s1 += f.x + f.y; ln_s2 = fmaf(f.x, f.x, fmaf(f.y, f.y, ln_s2));
                f = unpack_bf16(stv.w); ln_s1 += f.x + f.y; ln_s2 = fmaf(f.x, f.x, fmaf(f.y, f.y, ln_s2));
              }
            }
          }
          fence_proxy_async();
          __syncwarp();
          if constexpr (kLn == 3) {
            // column-pair sums over this warp's 32 rows, read back from the staging slot (what the consumer will read: the
            // ROUNDED values): lane = (pair, row parity); rows 2k + parity are 64 B apart, so a warp-wide load is conflict-free
            const unsigned vmask = __ballot_sync(0xffffffffu, valid);
            const int pr = lane & 15, hf = lane >> 4;
            const uint8_t* colp = slot + (pr & 3) * 4;
            const int g8c = pr >> 2;
            uint64_t s2 = 0ull, q2 = 0ull;
#pragma unroll
            for (int k2 = 0; k2 < 16; ++k2) {
              const int rr = 2 * k2 + hf;
              uint32_t u = *reinterpret_cast<const uint32_t*>(colp + rr * 64 + ((g8c ^ ((rr >> 1) & 3)) << 4));
              if (!((vmask >> rr) & 1u)) u = 0u;
              const uint64_t f = pack_f2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
              s2 = add_f2(s2, f);
              q2 = fma_f2(f, f, q2);
            }
            float sa, sb, qa, qb;
            unpack_f2(s2, sa, sb);
            unpack_f2(q2, qa, qb);
            float ssum = sa + sb, qsum = qa + qb;
            ssum += __shfl_xor_sync(0xffffffffu, ssum, 16);
            qsum += __shfl_xor_sync(0xffffffffu, qsum, 16);
            if (hf == 0 && n0 + 2 * pr < p.N)
              p.gn_stats_out[(long long)(m_blk * 4 + quad) * p.gn_ld + (n0 >> 1) + pr] = make_float2(ssum, qsum);
            __syncwarp();                           // lane 0 may re-arm the slot for a residual load below
          }
          if (lane == 0) {
            const int oc = (kEpi != 3 && act == 2) ? (n0 >> 1) : n0;
            if (p.mode == 0) tma_store_2d(&p.omap, slot, oc, box_row0);
            else tma_store_4d(&p.omap, slot, oc, bx, by, bimg);
            bulk_commit();
            if (has_res && c + 4 < nchunks && n0 + 128 < p.N) {
              bulk_wait_read<0>();                  // the slot is reused for the residual two chunks on
              issue_res_load(c + 4, qseq + 2);
            }
          }
          ++qseq;
        }
        if (chalf >= nchunks || ncol0 + chalf * 32 >= p.N) release_acc();   // this warp owned no chunk of the tile
        if constexpr (kLn == 2) {
          if (valid) p.ln_stats_out[row * p.ln_out_ld + n_blk * 2 + chalf] = make_float2(ln_s1, ln_s2);   // zeros if no chunk
        }
      } else if constexpr (kEpi == 0) {
        // ---------------------------------------------------------------- direct epilogue (fp32 or unaligned output)
        auto process = [&](int c, const uint4 (&res)[4]) {
          uint32_t v[32];
          tmem_ld32(t_row + c * 32, v);
          tmem_ld_wait();
          const int n0 = ncol0 + c * 32;
          if (!valid || n0 >= p.N) return;
          const float* bs = bias_w + ((c - chalf) >> 1) * 32;
          if (act == 2) {
            float o[16];
            geglu16(o, v, bs);
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ld_out + (n0 >> 1);
#pragma unroll
            for (int g8 = 0; g8 < 2; ++g8) {
              if (n0 + g8 * 16 >= p.N) break;
              uint4 stv;
              stv.x = pack_bf16(o[g8 * 8 + 0], o[g8 * 8 + 1]); stv.y = pack_bf16(o[g8 * 8 + 2], o[g8 * 8 + 3]);
              stv.z = pack_bf16(o[g8 * 8 + 4], o[g8 * 8 + 5]); stv.w = pack_bf16(o[g8 * 8 + 6], o[g8 * 8 + 7]);
              *reinterpret_cast<uint4*>(op + g8 * 8) = stv;
            }
            return;
          }
#pragma unroll
          for (int g8 = 0; g8 < 4; ++g8) {
            const int n = n0 + g8 * 8;
            if (n >= p.N) break;
            float x[8];
            affine_act8(x, v + g8 * 8, bs + g8 * 8, n);
            if (p.residual) {
              float2 f;
              f = unpack_bf16(res[g8].x); x[0] += f.x; x[1] += f.y;
              f = unpack_bf16(res[g8].y); x[2] += f.x; x[3] += f.y;
              f = unpack_bf16(res[g8].z); x[4] += f.x; x[5] += f.y;
              f = unpack_bf16(res[g8].w); x[6] += f.x; x[7] += f.y;
            }
            if (p.out_fp32) {
              float* op = reinterpret_cast<float*>(p.out) + row * p.ld_out + n;
              *reinterpret_cast<float4*>(op) = make_float4(x[0], x[1], x[2], x[3]);
              *reinterpret_cast<float4*>(op + 4) = make_float4(x[4], x[5], x[6], x[7]);
            } else {
              uint4 stv;
              stv.x = pack_bf16(x[0], x[1]); stv.y = pack_bf16(x[2], x[3]);
              stv.z = pack_bf16(x[4], x[5]); stv.w = pack_bf16(x[6], x[7]);
              *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + row * p.ld_out + n) = stv;
            }
          }
        };
        for (int c = chalf; c < nchunks; c += 4) {
          process(c, r0);
          if (p.residual && c + 4 < nchunks) load_res(c + 4, r0);
          if (c + 2 < nchunks) {
            process(c + 2, r1);
            if (p.residual && c + 6 < nchunks) load_res(c + 6, r1);
          }
        }
        release_acc();
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    if (tma_epi && lane == 0) bulk_wait<0>();     // all output boxes written before the CTA retires
  }
  tc_fence_before();
  if (csize > 1) cluster_sync(); else __syncthreads();   // no CTA may exit while peers can still multicast into it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (kPair) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------- host side
// Joint choice of the N tile and of single-CTA (128 x bn) versus CTA-pair (256 x bn, cta_group::2) tiles from a cycle
// model, fitted to a sweep of every GEMM/conv shape of the SDXL UNet tile-step (scripts/dev_gemm_tune.py):
//   per 64-deep k-block  max(tensor pipe 2*bn cycles, operand bytes into one SM / F)
//                        bytes = 128*(128 + bn) single, 128*(128 + bn/2) pair (each CTA of a pair loads half of B);
//                        F = 45 (single) / 35 (pair) B/clk/SM is what L2 delivers with all 148 SMs pulling — the UNet's
//                        shapes are bound by this operand feed, not by the MMA, which is why the pair's halved B traffic wins
//   per tile             2000 + 8*bn cycles (epilogue share that does not overlap), + 1500 for a pair (cluster hand-shakes)
// times the number of waves of tiles over the 148 SMs / 74 pairs.
struct TileChoice { int bn, cm, cn, pair; };
static int max_clusters_for(int csize) {
  static int cache[5] = {-1, -1, -1, -1, -1};
  if (csize < 1 || csize > 4) return 0;
  if (cache[csize] >= 0) return cache[csize];
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(gemm_tcgen05_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_set = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(num_sms() / csize * csize);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = 200 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gemm_tcgen05_kernel<false, 0>, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
  cache[csize] = n;
  return n;
}
static TileChoice pick_tiling(long long m_tiles, long long N, int num_kb, int sms, int force_pair) {
  TileChoice best = {32, 1, 1, 0};
  double best_cost = 1e30;
  for (int pair = 0; pair < 2; ++pair) {
    if (force_pair >= 0 && pair != force_pair) continue;
    const int units = pair ? max_clusters_for(2) : sms;
    if (units <= 0) continue;
    for (int bn = 256; bn >= 32; bn -= 32) {
      const long long n_tiles = (N + bn - 1) / bn;
      const long long work = (pair ? (m_tiles + 1) / 2 : m_tiles) * n_tiles;
      const long long waves = (work + units - 1) / units;
      const double feed = 128.0 * (128.0 + (pair ? bn / 2 : bn)) / (pair ? 35.0 : 45.0);
      const double t_kb = fmax(2.0 * bn, feed);
      const double cost = waves * (t_kb * num_kb + 2000.0 + 8.0 * bn + (pair ? 1500.0 : 0.0));
      if (cost < best_cost - 1e-9) { best_cost = cost; best = {bn, pair ? 2 : 1, 1, pair}; }
    }
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

static int gemm_impl(const fd_gemm_desc* d, fd_gemm_ln_desc* ln, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(d && d->a[0] && d->b && d->out, "fd_gemm: null operand");
  const bool ln_in = ln && ln->stats_in, ln_out = ln && ln->stats_out, gn_out = ln && ln->colstats_out;
  FD_REQUIRE(!(ln_in && ln_out) && !(gn_out && (ln_in || ln_out)), "fd_gemm_ln: one GEMM either consumes or produces one kind of statistics");
  FD_REQUIRE(!(ln_in || ln_out) || (d->mode == 0 && !d->a[1]), "fd_gemm_ln: plain single-source GEMMs only");
  FD_REQUIRE(d->mode >= 0 && d->mode <= 3, "fd_gemm: bad mode %d", d->mode);
  FD_REQUIRE(d->mode != 3 || (!d->a[1] && !d->residual && !d->out_fp32 && d->pad_lo >= 0 && d->pad_lo <= 3),
             "fd_gemm: mode 3 (3x3 conv over the 2x nearest-upsampled input, one output phase) takes one source, bf16 out, no residual");
  FD_REQUIRE(d->act >= 0 && d->act <= 5, "fd_gemm: bad act %d", d->act);
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
  p.taps = d->mode == 0 ? 1 : (d->mode == 3 ? 4 : 9);
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
    FD_REQUIRE(d->batch > 0 && H > 0 && W > 0 && (d->mode == 3 || d->pad_lo == 0 || d->pad_lo == 1), "fd_gemm: bad conv geometry");
    if (d->mode == 3) {
      p.out_h = H; p.out_w = W;          // the phase's output grid = the input grid (out pixel (2i+a, 2j+b) <-> input pixel (i, j))
    } else if (d->mode == 1) {
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
      if (d->mode == 3) {
        // phase (a, b) of conv3x3(nearest_upsample_2x(x)): the three taps of an axis fall on TWO input pixels — a = 0: {i-1, i},
        // a = 1: {i, i+1} — with the weights that share a pixel pre-summed on the host (tap t = r'*2 + s' of the [N, 4*C] operand)
        if (t >= 4) continue;
        const int a = d->pad_lo >> 1, b = d->pad_lo & 1;
        p.tap_map[t] = 0; p.tap_dy[t] = (t >> 1) - 1 + a; p.tap_dx[t] = (t & 1) - 1 + b;
      } else if (d->mode == 1) {
        p.tap_map[t] = 0; p.tap_dy[t] = r - 1; p.tap_dx[t] = s - 1;
      } else {
        const int ty = r - d->pad_lo, tx = s - d->pad_lo;
        const int py = ((ty % 2) + 2) % 2, px = ((tx % 2) + 2) % 2;
        p.tap_map[t] = py * 2 + px; p.tap_dy[t] = (ty - py) / 2; p.tap_dx[t] = (tx - px) / 2;
      }
    }
  }
  p.M = (int)M;
  p.batch = d->mode == 0 ? 1 : d->batch;
  FD_REQUIRE(M > 0 && M < (1ll << 31), "fd_gemm: bad M");
  const int sms = num_sms();
  static int env_cm = -1, env_cn = -1, env_bn = -1, env_pair = -1;
  if (env_cm < 0) {
    const char* e = getenv("FD_GEMM_CLUSTER");          // tuning override "cm,cn": TMA-multicast cluster (experiment)
    env_cm = env_cn = 0;
    if (e) sscanf(e, "%d,%d", &env_cm, &env_cn);
    const char* b = getenv("FD_GEMM_BN");
    env_bn = b ? atoi(b) : 0;
    const char* pr = getenv("FD_GEMM_PAIR");             // 0 / 1 forces 1-CTA / CTA-pair tiles
    env_pair = pr ? atoi(pr) : -1;
  }
  int want_pair = env_pair;                               // -1 auto, 0 single CTA, 1 CTA pair
  if (d->block_n & 0x1000) want_pair = 1;
  if (d->block_n & 0x2000) want_pair = 0;
  TileChoice tc = pick_tiling(p.m_tiles, d->N, p.taps * p.kb_per_tap, sms, want_pair);
  if (env_cm > 0 && env_cn > 0) {
    tc = pick_tiling(p.m_tiles, d->N, p.taps * p.kb_per_tap, sms, 0);
    tc.cm = env_cm; tc.cn = env_cn; tc.pair = 0;
  }
  if (env_bn > 0) tc.bn = env_bn;
  if (d->block_n & 0xfff) tc.bn = d->block_n & 0xfff;
  p.block_n = tc.bn; p.cm = tc.cm; p.cn = tc.cn; p.pair = tc.pair;
  FD_REQUIRE(p.block_n % 32 == 0 && p.block_n >= 32 && p.block_n <= 256, "fd_gemm: bad block_n %d", p.block_n);
  p.n_tiles = (int)((d->N + p.block_n - 1) / p.block_n);
  p.m_super = (p.m_tiles + p.cm - 1) / p.cm;
  p.n_super = (p.n_tiles + p.cn - 1) / p.cn;
  {
    static int env_gm = -1;
    if (env_gm < 0) { const char* e = getenv("FD_GEMM_GROUP_M"); env_gm = e ? atoi(e) : 0; }
    p.group_m = env_gm > 0 ? env_gm : 12;
    if (p.group_m > p.m_super) p.group_m = p.m_super;
  }
  // A part geometry (1/cn of the 128-row tile per CTA of a tile row)
  int a_box_rows = kBlockM / p.cn, a_bw = p.bw, a_bh = p.bh;
  if (p.mode != 0 && p.cn == 2) {
    if (p.bh >= 2) { a_bh = p.bh / 2; p.a_part_dy = a_bh; p.a_part_dx = 0; }
    else { a_bw = p.bw / 2; p.a_part_dy = 0; p.a_part_dx = a_bw; }
  }
  const int stage_bytes = kATileBytes + (p.pair ? p.block_n / 2 : p.block_n) * kBlockK * 2;
  // TMA epilogue whenever the output (and residual) rows can be addressed by a tensor map: bf16, 16-byte aligned rows
  p.tma_epi = !d->out_fp32 && ((uintptr_t)d->out % 16 == 0) && (d->ld_out % 8 == 0) &&
              (!d->residual || (((uintptr_t)d->residual % 16 == 0) && (d->ld_res % 8 == 0) && d->act != 2));
  {
    static int env_direct = -1;
    if (env_direct < 0) { const char* e = getenv("FD_GEMM_DIRECT_EPILOGUE"); env_direct = e ? atoi(e) : 0; }
    if (env_direct) p.tma_epi = 0;
  }
  FD_REQUIRE(d->mode != 3 || p.tma_epi, "fd_gemm: mode 3 needs the TMA epilogue (bf16 output, 16-byte aligned rows)");
  p.slot_bytes = d->act == 2 ? 1024 : 2048;
  const int staging_bytes = p.tma_epi ? 16 * p.slot_bytes : 0;
  const int ln_strip_bytes = ln_in ? 8 * 128 * (int)sizeof(float) : 0;       // colsum strips of the 8 epilogue warps, after the barriers
  const int smem_budget = 227 * 1024 - (int)sizeof(PipeBarriers) - 64 - staging_bytes - ln_strip_bytes;
  p.stages = smem_budget / stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  const size_t smem_bytes = (size_t)p.stages * stage_bytes + staging_bytes + sizeof(PipeBarriers) + ln_strip_bytes;

  // tensor maps
  if (d->mode == 0) {
    for (int s = 0; s < nsrc; ++s) {
      uint64_t dims[2] = {(uint64_t)d->a_cols[s], (uint64_t)M};
      uint64_t str[1] = {(uint64_t)d->a_ld[s] * 2};
      uint32_t box[2] = {kBlockK, (uint32_t)a_box_rows};
      if (encode_tensor_map(&p.amap[s], d->a[s], 2, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                            CU_TENSOR_MAP_SWIZZLE_128B)) return 3;
    }
  } else if (d->mode == 1 || d->mode == 3) {
    for (int s = 0; s < nsrc; ++s) {
      const uint64_t ld = (uint64_t)d->a_ld[s];
      uint64_t dims[4] = {(uint64_t)d->a_cols[s], (uint64_t)d->in_w, (uint64_t)d->in_h, (uint64_t)d->batch};
      uint64_t str[3] = {ld * 2, ld * 2 * d->in_w, ld * 2 * d->in_w * d->in_h};
      uint32_t box[4] = {kBlockK, (uint32_t)a_bw, (uint32_t)a_bh, 1};
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
        uint32_t box[4] = {kBlockK, (uint32_t)a_bw, (uint32_t)a_bh, 1};
        if (encode_tensor_map(&p.amap[py * 2 + px], base, 4, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                              CU_TENSOR_MAP_SWIZZLE_128B)) return 3;
      }
  }
  {
    uint64_t dims[2] = {(uint64_t)ktot, (uint64_t)d->N};
    uint64_t str[1] = {(uint64_t)ktot * 2};
    uint32_t box[2] = {kBlockK, (uint32_t)(p.block_n / p.cm)};
    if (encode_tensor_map(&p.bmap, d->b, 2, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                          CU_TENSOR_MAP_SWIZZLE_128B)) return 3;
  }
  if (p.tma_epi) {
    // one box = the 32 rows of an epilogue warp x one 32-column accumulator chunk (GEGLU: 16 output columns)
    const int box_w = (d->mode == 0 || p.bw >= 32) ? 32 : p.bw, box_h = 32 / box_w;       // conv: the 32 rows as a pixel rectangle of the patch
    for (int which = 0; which < 2; ++which) {
      const void* base = which == 0 ? d->out : d->residual;
      if (!base) continue;
      const uint64_t ld = (uint64_t)(which == 0 ? d->ld_out : d->ld_res);
      const uint64_t ncols = (uint64_t)((which == 0 && d->act == 2) ? d->N / 2 : d->N);
      const uint32_t box_c = (which == 0 && d->act == 2) ? 16 : 32;
      const CUtensorMapSwizzle sw = box_c == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B;
      CUtensorMap* map = which == 0 ? &p.omap : &p.rmap;
      if (d->mode == 0) {
        uint64_t dims[2] = {ncols, (uint64_t)M};
        uint64_t str[1] = {ld * 2};
        uint32_t box[2] = {box_c, 32};
        if (encode_tensor_map(map, base, 2, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, sw)) return 3;
      } else if (d->mode == 3) {
        // output phase (a, b): every second pixel of every second row of the [batch, 2H, 2W, N] image, starting at (a, b)
        const int a = d->pad_lo >> 1, b = d->pad_lo & 1;
        const uint64_t w2 = 2ull * p.out_w, h2 = 2ull * p.out_h;
        const uint8_t* pb = (const uint8_t*)base + ((uint64_t)a * w2 + b) * ld * 2;
        uint64_t dims[4] = {ncols, (uint64_t)p.out_w, (uint64_t)p.out_h, (uint64_t)d->batch};
        uint64_t str[3] = {ld * 4, ld * 4 * w2, ld * 2 * w2 * h2};
        uint32_t box[4] = {box_c, (uint32_t)box_w, (uint32_t)box_h, 1};
        if (encode_tensor_map(map, pb, 4, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, sw)) return 3;
      } else {
        uint64_t dims[4] = {ncols, (uint64_t)p.out_w, (uint64_t)p.out_h, (uint64_t)d->batch};
        uint64_t str[3] = {ld * 2, ld * 2 * p.out_w, ld * 2 * p.out_w * p.out_h};
        uint32_t box[4] = {box_c, (uint32_t)box_w, (uint32_t)box_h, 1};
        if (encode_tensor_map(map, base, 4, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, sw)) return 3;
      }
    }
  }
  p.bias = d->bias; p.rowbias = d->rowbias; p.rows_per_batch = d->rows_per_batch > 0 ? d->rows_per_batch : M;
  p.ld_rowbias = d->ld_rowbias > 0 ? d->ld_rowbias : d->N;
  FD_REQUIRE(!d->rowbias || (((uintptr_t)d->rowbias % 16) == 0 && p.ld_rowbias % 4 == 0), "fd_gemm: rowbias alignment");
  p.residual = (const __nv_bfloat16*)d->residual; p.ld_res = d->ld_res;
  p.out = d->out; p.ld_out = d->ld_out; p.alpha = d->alpha == 0.f ? 1.f : d->alpha;
  p.act = d->act; p.bias_along_m = d->bias_along_m; p.out_fp32 = d->out_fp32;

  // epilogue specialisation: 1 = TMA epilogue, no activation (bias / per-sample bias / residual optional);
  // 2 = TMA epilogue, GEGLU with column bias; 3 = TMA epilogue with an elementwise activation chosen at run time (SiLU / quick-GELU /
  // GELU / LeakyReLU: the FaithDiff stem's MLP, the condition embedding, the text encoders, BSRNet — they ran through the generic
  // kernel at a third of the specialised ones' rate); 0 = generic (every option at run time, direct stores possible)
  int epi = 0;
  if (p.tma_epi && !d->bias_along_m) {
    if (d->act == 0) epi = 1;
    else if (d->act == 2 && !d->rowbias && !d->residual) epi = 2;
    else if (d->act != 2) epi = 3;
  }
  {
    static int env_generic = -1;
    if (env_generic < 0) { const char* e = getenv("FD_GEMM_GENERIC_EPILOGUE"); env_generic = e ? atoi(e) : 0; }
    if (env_generic) epi = 0;
  }
  if (ln_in) {
    FD_REQUIRE(epi == 1 || epi == 2, "fd_gemm_ln: the consumer needs the TMA epilogue (bf16 output, act 0 or GEGLU)");
    FD_REQUIRE(ln->colsum && ln->stats_in_n > 0 && ln->stats_in_ld >= ln->stats_in_n && ln->inv_c > 0.f, "fd_gemm_ln: bad consumer arguments");
    FD_REQUIRE(d->alpha == 0.f || d->alpha == 1.f, "fd_gemm_ln: alpha is not applied on the consumer path");
    p.ln_stats_in = (const float2*)ln->stats_in; p.ln_in_ld = (int)ln->stats_in_ld; p.ln_in_n = ln->stats_in_n;
    p.ln_colsum = ln->colsum; p.ln_inv_c = ln->inv_c; p.ln_eps = ln->eps;
  }
  if (ln_out) {
    FD_REQUIRE(epi == 1 && d->N % 32 == 0, "fd_gemm_ln: the producer needs the linear TMA epilogue and N %% 32 == 0");
    FD_REQUIRE(ln->stats_out_ld >= 2 * p.n_tiles, "fd_gemm_ln: stats_out_ld %lld < %d slots", (long long)ln->stats_out_ld, 2 * p.n_tiles);
    p.ln_stats_out = (float2*)ln->stats_out; p.ln_out_ld = (int)ln->stats_out_ld;
    ln->stats_out_n = 2 * p.n_tiles;
  }
  if (gn_out) {
    FD_REQUIRE(epi == 1 && d->N % 2 == 0 && ln->colstats_ld >= d->N / 2, "fd_gemm_ln: column statistics need the linear TMA epilogue");
    p.gn_stats_out = (float2*)ln->colstats_out; p.gn_ld = (int)ln->colstats_ld;
    ln->colstats_boxes = 4 * p.m_tiles;
    ln->colstats_boxes_per_image = d->mode == 0 ? 0 : 4 * p.tiles_per_img;
  }
  using KernelFn = void (*)(GemmKParams);
  static const KernelFn kernels_gn[2] = {gemm_tcgen05_kernel<false, 1, 3>, gemm_tcgen05_kernel<true, 1, 3>};
  static const KernelFn kernels_ln[6] = {gemm_tcgen05_kernel<false, 1, 1>, gemm_tcgen05_kernel<false, 2, 1>, gemm_tcgen05_kernel<false, 1, 2>,
                                         gemm_tcgen05_kernel<true, 1, 1>,  gemm_tcgen05_kernel<true, 2, 1>,  gemm_tcgen05_kernel<true, 1, 2>};
  static const KernelFn kernels[8] = {gemm_tcgen05_kernel<false, 0>, gemm_tcgen05_kernel<false, 1>, gemm_tcgen05_kernel<false, 2>,
                                      gemm_tcgen05_kernel<false, 3>, gemm_tcgen05_kernel<true, 0>,  gemm_tcgen05_kernel<true, 1>,
                                      gemm_tcgen05_kernel<true, 2>,  gemm_tcgen05_kernel<true, 3>};
  static bool attr_set = false;
  if (!attr_set) {
    for (int i = 0; i < 8; ++i) FD_CHECK_CUDA(cudaFuncSetAttribute(kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int csize = p.cm * p.cn;
  const int num_super = p.m_super * p.n_super;
  KernelFn kernel = kernels[(p.pair ? 4 : 0) + epi];
  if (ln_in || ln_out) {
    static bool ln_attr_set = false;
    if (!ln_attr_set) {
      for (int i = 0; i < 6; ++i) FD_CHECK_CUDA(cudaFuncSetAttribute(kernels_ln[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      ln_attr_set = true;
    }
    kernel = kernels_ln[(p.pair ? 3 : 0) + (ln_out ? 2 : epi - 1)];
  }
  if (gn_out) {
    static bool gn_attr_set = false;
    if (!gn_attr_set) {
      for (int i = 0; i < 2; ++i) FD_CHECK_CUDA(cudaFuncSetAttribute(kernels_gn[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      gn_attr_set = true;
    }
    kernel = kernels_gn[p.pair ? 1 : 0];
  }
  if (csize == 1) {
    const int grid = num_super < sms ? num_super : sms;
    FD_CHECK_CUDA(launch_pdl(kernel, dim3(grid), dim3(kThreads), smem_bytes, stream, 1, p));
  } else {
    const int maxc = max_clusters_for(csize);
    FD_REQUIRE(maxc > 0, "fd_gemm: no co-resident clusters of %d CTAs available", csize);
    const int clusters = num_super < maxc ? num_super : maxc;
    FD_CHECK_CUDA(launch_pdl(kernel, dim3(clusters * csize), dim3(kThreads), smem_bytes, stream, csize, p));
  }
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int fd_gemm(const fd_gemm_desc* d, void* stream) { return gemm_impl(d, nullptr, stream); }

extern "C" int fd_gemm_ln(const fd_gemm_desc* d, fd_gemm_ln_desc* ln, void* stream) { return gemm_impl(d, ln, stream); }

extern "C" int32_t fd_sizeof_gemm_ln_desc(void) { return (int32_t)sizeof(fd_gemm_ln_desc); }
