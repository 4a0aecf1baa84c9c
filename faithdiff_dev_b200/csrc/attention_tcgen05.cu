// Flash-style attention for head_dim 64 on tcgen05 (sm_100a), bf16 in/out, fp32 softmax + accumulation.
// One CTA = one (batch, head, 128-query block); three CTAs (64-key blocks; two with 128-key blocks) co-reside per SM so one CTA's
// softmax overlaps the others' MMAs and latency-bound phases.
//   warp 0 lane 0 : TMA producer  (Q once; K and V 128-key blocks in separate 2-stage rings, 128B swizzle.  K_j is released
//                                 as soon as S_j = Q K_j^T has been computed, so K_{j+2} streams in two blocks ahead)
//   warp 1 lane 0 : MMA issuer    S_{j+1} = Q K_{j+1}^T (M128 N128 K64) -> TMEM cols [0,128) is issued as soon as the softmax
//                                 warps have copied S_j into registers, i.e. it runs under the exponentials of block j;
//                                 O += P_j V_j (M128 N64 K128, A = P from smem, B = V MN-major) accumulates in TMEM cols [128,192)
//   warps 2..5    : softmax — one query row per thread: the whole S row (128 fp32) is read from TMEM once, then max, exp2 and
//                   row sum in registers; P (bf16) goes to swizzled smem.  The running output stays in TMEM; it is rescaled
//                   in place only when some row's maximum grew by more than 2^8 since the maximum the row last committed to
//                   (the softmax is exact for any reference value, and with p <= 2^8 bf16/fp32 keep their precision), so
//                   the per-block "load O, rescale, add" of a register-resident accumulator is gone.
#include <type_traits>
#include "common.cuh"
#include "../../include/faithdiff_b200.h"

namespace fd {
extern int64_t g_launches;

struct AttnParams {
  CUtensorMap qmap, kmap, vmap;
  CUtensorMap omap;    // short-KV kernel: output as (cols, lq, batch) so a 32-row box is clipped at the end of each batch
  __nv_bfloat16* out;
  long long ldo;
  int lq, lk, heads;
  float scale_log2;
  int causal;          // short-KV kernel only: key k attends to query q iff k <= q (CLIP text encoders)
  int q_per_cta;       // short-KV kernel: consecutive 128-query blocks of one (batch, head) handled by one CTA
};

constexpr int kQ = 128, kKV = 128, kD = 64;
constexpr int kTile = kQ * kD * 2;   // 16 KB
constexpr float kRescaleThreshold = 8.f;   // log2 units

struct __align__(8) AttnBars {
  uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full, s_read, p_full, o_done;
  uint32_t tmem_base;
};

// exp2 on the FMA pipe for a PAIR of arguments (opt-in, FD_ATTN_POLY): the softmax of this kernel is paced by the 16-lane MUFU
// pipe (128 ex2 per query row and key block against 512 MMA cycles), so part of the exponentials can move to the otherwise idle
// FMA pipe: a is clamped to >= -125, split with the 1.5*2^23 magic-number rounding into n = rint(a) and f = a - n in [-0.5, 0.5],
// 2^f is a degree-3 polynomial (Lawson-fitted, max relative error 7.6e-5 = 1/50 of the bf16 rounding step P goes through),
// and n is added into the exponent field with one integer multiply-add.  6 packed fp32x2 + 2 max + 2 integer ops per pair.
__device__ __forceinline__ void exp2_pair_poly(float a0, float a1, float& e0, float& e1) {
  const uint64_t a = pack_f2(fmaxf(a0, -125.f), fmaxf(a1, -125.f));
  const uint64_t t = add_f2(a, pack_f2(12582912.f, 12582912.f));
  const uint64_t n = add_f2(t, pack_f2(-12582912.f, -12582912.f));
  const uint64_t f = fma_f2(n, pack_f2(-1.f, -1.f), a);
  uint64_t q = fma_f2(pack_f2(0.05520550534129143f, 0.05520550534129143f), f, pack_f2(0.24261397123336792f, 0.24261397123336792f));
  q = fma_f2(q, f, pack_f2(0.6932547688484192f, 0.6932547688484192f));
  q = fma_f2(q, f, pack_f2(0.9999276995658875f, 0.9999276995658875f));
  float q0, q1, t0, t1;
  unpack_f2(q, q0, q1);
  unpack_f2(t, t0, t1);
  e0 = __int_as_float(__float_as_int(t0) * 8388608 + __float_as_int(q0));     // (bits(t) << 23) + bits(2^f); bits(magic) << 23 == 0
  e1 = __int_as_float(__float_as_int(t1) * 8388608 + __float_as_int(q1));
}

// kPolyOf8: of every 8 consecutive pairs of a row, the first kPolyOf8 use exp2_pair_poly, the rest MUFU ex2 (0 = all MUFU)
#ifdef FD_ATTN_TRACE
// development only (scripts/dev_attn_trace.py): clock64 stamps of one CTA's first softmax warp and of its MMA thread, key blocks 16..31
__device__ long long g_attn_trace[16][12];
#define FD_TR(k) do { if (blockIdx.x == 5 && blockIdx.y == 1 && blockIdx.z == 2 && lane == 0 && j >= 16 && j < 32) g_attn_trace[j - 16][k] = clock64(); } while (0)
#else
#define FD_TR(k) do { } while (0)
#endif
// kKVB = keys per block.  128: two CTAs per SM (256 TMEM columns, 112 KB shared memory each).  64: THREE CTAs per SM (128 TMEM
// columns, 64 KB, <= 112 registers): a single softmax warp per scheduler keeps the MUFU pipe only 70 % busy next to its FFMA2 /
// FADD2 / F2FP companions, two reach 93 % (scripts/micro/mufu_rate.cu), and each warp sleeps ~600 cycles per block in barrier /
// TMEM-load latencies — a third resident warp per scheduler is what keeps two in their exponentials.
template <int kPolyOf8, int kKVB>
__global__ void __launch_bounds__(192, kKVB == 64 ? 3 : 2) attention_d64_kernel(const __grid_constant__ AttnParams p) {
  constexpr int kKTile = kKVB * kD * 2;                // one K or V block
  constexpr int kPH = kKVB / 64;                       // 64-key halves of P
  constexpr int kCh = kKVB / 32;                       // 32-column chunks of an S row
  constexpr int kTmemCols = kKVB == 64 ? 128 : 256;
  constexpr uint32_t kOCol = kKVB;                     // O accumulator behind the S columns
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTile;                          // 2 stages
  uint8_t* sV = sK + 2 * kKTile;                       // 2 stages
  uint8_t* sP = sV + 2 * kKTile;                       // kPH halves (64 keys each)
  AttnBars* bars = reinterpret_cast<AttnBars*>(sP + kPH * kTile);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qblk = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.lk + kKVB - 1) / kKVB;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.qmap); tma_prefetch_desc(&p.kmap); tma_prefetch_desc(&p.vmap);
    mbar_init(&bars->q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->k_full[s], 1); mbar_init(&bars->k_empty[s], 1);
      mbar_init(&bars->v_full[s], 1); mbar_init(&bars->v_empty[s], 1);
    }
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->s_read, 4);
    mbar_init(&bars->p_full, 4);
    mbar_init(&bars->o_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_base, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars->q_full, kTile);
      tma_load_2d(&p.qmap, &bars->q_full, sQ, head * kD, b * p.lq + qblk * kQ);
      auto load_k = [&](int j) {
        const int s = j & 1;
        mbar_wait(&bars->k_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->k_full[s], kKTile);
        tma_load_2d(&p.kmap, &bars->k_full[s], sK + s * kKTile, head * kD, b * p.lk + j * kKVB);
      };
      auto load_v = [&](int j) {
        const int s = j & 1;
        mbar_wait(&bars->v_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->v_full[s], kKTile);
        tma_load_2d(&p.vmap, &bars->v_full[s], sV + s * kKTile, head * kD, b * p.lk + j * kKVB);
      };
      // request order = order in which the slots become free: K_{j+1}'s slot (after S_{j-1}) before V_j's (after PV_{j-2})
      load_k(0);
      if (nkv > 1) load_k(1);
      load_v(0);
      for (int j = 1; j < nkv; ++j) {
        if (j + 1 < nkv) load_k(j + 1);
        load_v(j);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc(128, kKVB, 1, 0, 0);
      const uint32_t idesc_o = umma_idesc(128, 64, 1, 0, 1);   // B (=V) is MN-major
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
      auto issue_s = [&](int j) {
        const uint32_t aK = smem_u32(sK + (j & 1) * kKTile);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem, umma_desc_sw128(aQ + k * 32, 16, 1024), umma_desc_sw128(aK + k * 32, 16, 1024), idesc_s, k != 0);
        tc_commit(&bars->s_full);
        tc_commit(&bars->k_empty[j & 1]);
      };
      mbar_wait(&bars->q_full, 0);
      mbar_wait(&bars->k_full[0], 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        if (j + 1 < nkv) {
          // S_{j+1} overwrites S_j as soon as every softmax warp holds its S_j rows in registers
          mbar_wait(&bars->k_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
          mbar_wait(&bars->s_read, j & 1);
          tc_fence_after();
          FD_TR(8);
          issue_s(j + 1);
          FD_TR(9);
        }
        const int s = j & 1;
        mbar_wait(&bars->v_full[s], (j >> 1) & 1);
        mbar_wait(&bars->p_full, j & 1);
        tc_fence_after();
        FD_TR(10);
        const uint32_t aV = smem_u32(sV + s * kKTile);
#pragma unroll
        for (int k = 0; k < kKVB / 16; ++k)
          umma_f16(tmem + kOCol, umma_desc_sw128(aP + (k >> 2) * kTile + (k & 3) * 32, 16, 1024),
                   umma_desc_sw128(aV + k * 2048, 0, 1024), idesc_o, (j != 0) || (k != 0));
        tc_commit(&bars->v_empty[s]);
        tc_commit(&bars->o_done);
        FD_TR(11);
      }
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                       // query row within the block
    const uint32_t t_lane = tmem + ((uint32_t)(quad * 32) << 16);
    float m_used = -INFINITY, l = 0.f;                    // reference maximum the row's P and O are expressed in; row sum
    uint8_t* prow = sP + r * 128;
    const int sw = r & 7;
    const uint64_t scale2 = pack_f2(p.scale_log2, p.scale_log2);
    for (int j = 0; j < nkv; ++j) {
      if (warp == 2) FD_TR(0);
      mbar_wait(&bars->s_full, j & 1);
      tc_fence_after();
      if (warp == 2) FD_TR(1);
      uint32_t v[kKVB];
#pragma unroll
      for (int c = 0; c < kCh; ++c) tmem_ld32(t_lane + c * 32, reinterpret_cast<uint32_t(&)[32]>(v[c * 32]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->s_read);          // the tensor core may now compute S_{j+1}
      if (warp == 2) FD_TR(2);
      const int kv0 = j * kKVB;
      const bool full = kv0 + kKVB <= p.lk;               // only the last block can need masking (warp-uniform)
      float mx = m_used;
      if (full) {
        // four independent 3-input-max chains (a single chain is 64 dependent instructions of latency)
        float m1 = mx, m2 = mx, m3 = mx;
#pragma unroll
        for (int i = 0; i < kKVB; i += 8) {
          mx = max3_f(mx, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
          m1 = max3_f(m1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
          m2 = max3_f(m2, __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
          m3 = max3_f(m3, __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
        }
        mx = fmaxf(max3_f(mx, m1, m2), m3);
      } else {
#pragma unroll
        for (int i = 0; i < kKVB; ++i)
          if (kv0 + i < p.lk) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      // The reference maximum of this block is decided from the row maximum alone; the exponentials then run into REGISTERS (the 128
      // S values die as the 64 packed P words are born), and only after them does the thread wait for P_{j-1} V_{j-1} — which had the
      // whole exponential phase to finish — before it rescales O (rarely) and overwrites the single P buffer.  With the wait in front
      // of the exponentials it was 12 % of a softmax warp's time (profiles/r02_ncu_attention_4096_source.csv).
      if (warp == 2) FD_TR(3);
      float alpha = 1.f;
      bool rescale = false;
      if (j == 0) {
        m_used = mx;
      } else {
        const bool grew = (mx - m_used) * p.scale_log2 > kRescaleThreshold;
        rescale = __any_sync(0xffffffffu, grew);
        if (rescale) {
          alpha = fast_exp2((m_used - mx) * p.scale_log2);
          m_used = mx;
        }
      }
      const float nmxs = -m_used * p.scale_log2;
      const uint64_t nmxs2 = pack_f2(nmxs, nmxs);
      uint64_t lsum2 = 0ull;
      uint32_t pk[kKVB / 2];
      // p = exp2(s*scale - m*scale) -> bf16 (scale/subtract and the row sum on packed fp32x2)
      auto exp_block = [&](auto full_tag) {
        constexpr bool kFull = decltype(full_tag)::value;
#pragma unroll
        for (int c = 0; c < kCh; ++c) {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const uint64_t t = fma_f2(pack_f2(__uint_as_float(v[c * 32 + i]), __uint_as_float(v[c * 32 + i + 1])), scale2, nmxs2);
            float a0, a1;
            unpack_f2(t, a0, a1);
            if constexpr (kPolyOf8 > 0) {
              if (((i >> 1) & 7) < kPolyOf8) {          // compile-time after unrolling
                exp2_pair_poly(a0, a1, a0, a1);
              } else {
                a0 = fast_exp2(a0);
                a1 = fast_exp2(a1);
              }
            } else {
              a0 = fast_exp2(a0);      // (an f16x2 exp2 — two results per MUFU op — was measured slower, twice: the extra
              a1 = fast_exp2(a1);      //  converts cost more issue slots than the halved MUFU count saves)
            }
            if constexpr (!kFull) {
              if (kv0 + c * 32 + i >= p.lk) a0 = 0.f;
              if (kv0 + c * 32 + i + 1 >= p.lk) a1 = 0.f;
            }
            lsum2 = add_f2(lsum2, pack_f2(a0, a1));
            pk[c * 16 + (i >> 1)] = pack_bf16(a0, a1);
          }
        }
      };
      if (full) exp_block(std::true_type{}); else exp_block(std::false_type{});
      if (warp == 2) FD_TR(4);
      if (j > 0) {
        mbar_wait(&bars->o_done, (j - 1) & 1);            // P_{j-1} V_{j-1} finished: sP may be rewritten, O may be rescaled
        tc_fence_after();
        if (warp == 2) FD_TR(5);
        if (rescale) {
#pragma unroll 1
          for (int q = 0; q < 4; ++q) {
            uint32_t o[16];
            tmem_ld16(t_lane + kOCol + q * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(t_lane + kOCol + q * 16, o);
          }
          tmem_st_wait();
          l *= alpha;
        }
      }
#pragma unroll
      for (int c = 0; c < kCh; ++c) {
        uint8_t* half_base = prow + (c >> 1) * kTile;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int chunk = (c & 1) * 4 + g;
          *reinterpret_cast<uint4*>(half_base + ((chunk ^ sw) << 4)) =
              make_uint4(pk[c * 16 + g * 4], pk[c * 16 + g * 4 + 1], pk[c * 16 + g * 4 + 2], pk[c * 16 + g * 4 + 3]);
        }
      }
      float ls0, ls1;
      unpack_f2(lsum2, ls0, ls1);
      l += ls0 + ls1;
      fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full);
      if (warp == 2) FD_TR(6);
    }
    mbar_wait(&bars->o_done, (nkv - 1) & 1);
    tc_fence_after();
    const int qi = qblk * kQ + r;
    const float inv = 1.f / l;
    __nv_bfloat16* dst = p.out + ((long long)b * p.lq + qi) * p.ldo + head * kD;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld32(t_lane + kOCol + c * 32, o);
      tmem_ld_wait();
      if (qi < p.lq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(o[g * 8 + 0]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
          u.y = pack_bf16(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
          u.z = pack_bf16(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
          u.w = pack_bf16(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + c * 32 + g * 8) = u;
        }
      }
    }
    tc_fence_before();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, kTmemCols);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Keys fit one block (lk <= 128: the UNet's cross-attention over 77 text tokens).  There is no key loop to hide latencies
// in, so one CTA keeps K and V resident and streams `q_per_cta` query blocks of its (batch, head) through a pipeline:
// Q double-buffered in smem, S_{i+1} = Q_{i+1} K^T issued while block i is in its exponentials, O double-buffered in TMEM
// (cols [128,192) / [192,256)) so the normalise-and-store of block i-1 runs after P_i has been handed to the tensor core.
struct __align__(8) AttnShortBars {
  uint64_t kv_full, q_full[2], q_empty[2], s_full, s_read, p_full, o_done[2], o_read[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(192, 2) attention_d64_shortkv_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sK = smem;
  uint8_t* sV = smem + kTile;
  uint8_t* sQ = smem + 2 * kTile;        // 2 stages
  uint8_t* sP = smem + 4 * kTile;        // 2 halves (64 keys each)
  uint8_t* sO = smem + 6 * kTile;        // output staging: 4 warps x (32 rows x 128 B), 128B-swizzled, TMA-stored
  AttnShortBars* bars = reinterpret_cast<AttnShortBars*>(smem + 7 * kTile);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.y, b = blockIdx.z;
  const int nqb = (p.lq + kQ - 1) / kQ;
  const int first = blockIdx.x * p.q_per_cta;
  const int n_it = min(p.q_per_cta, nqb - first);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.qmap); tma_prefetch_desc(&p.kmap); tma_prefetch_desc(&p.vmap); tma_prefetch_desc(&p.omap);
    mbar_init(&bars->kv_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->q_full[s], 1); mbar_init(&bars->q_empty[s], 1);
      mbar_init(&bars->o_done[s], 1); mbar_init(&bars->o_read[s], 4);
    }
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->s_read, 4);
    mbar_init(&bars->p_full, 4);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_base, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  pdl_launch_dependents();
  pdl_wait();

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars->kv_full, 2 * kTile);
      tma_load_2d(&p.kmap, &bars->kv_full, sK, head * kD, b * p.lk);
      tma_load_2d(&p.vmap, &bars->kv_full, sV, head * kD, b * p.lk);
      for (int i = 0; i < n_it; ++i) {
        const int s = i & 1;
        mbar_wait(&bars->q_empty[s], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->q_full[s], kTile);
        tma_load_2d(&p.qmap, &bars->q_full[s], sQ + s * kTile, head * kD, b * p.lq + (first + i) * kQ);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc(128, ((p.lk + 31) >> 5) * 32, 1, 0, 0);   // only the 32-key chunks that hold keys
      const uint32_t idesc_o = umma_idesc(128, 64, 1, 0, 1);   // B (=V) is MN-major
      const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP);
      const int nk16 = (p.lk + 15) >> 4;                         // P V over the 16-key steps that hold keys
      auto issue_s = [&](int i) {
        mbar_wait(&bars->q_full[i & 1], (i >> 1) & 1);
        if (i > 0) mbar_wait(&bars->s_read, (i - 1) & 1);    // every softmax warp holds S_{i-1} in registers
        tc_fence_after();
        const uint32_t aQ = smem_u32(sQ + (i & 1) * kTile);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem, umma_desc_sw128(aQ + k * 32, 16, 1024), umma_desc_sw128(aK + k * 32, 16, 1024), idesc_s, k != 0);
        tc_commit(&bars->s_full);
        tc_commit(&bars->q_empty[i & 1]);
      };
      mbar_wait(&bars->kv_full, 0);
      issue_s(0);
      for (int i = 0; i < n_it; ++i) {
        if (i + 1 < n_it) issue_s(i + 1);
        mbar_wait(&bars->p_full, i & 1);
        if (i >= 2) mbar_wait(&bars->o_read[i & 1], ((i - 2) >> 1) & 1);   // O_{i-2} has been copied out of this buffer
        tc_fence_after();
        for (int k = 0; k < nk16; ++k)
          umma_f16(tmem + 128 + (i & 1) * 64, umma_desc_sw128(aP + (k >> 2) * kTile + (k & 3) * 32, 16, 1024),
                   umma_desc_sw128(aV + k * 2048, 0, 1024), idesc_o, k != 0);
        tc_commit(&bars->o_done[i & 1]);
      }
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                       // query row within the block
    const uint32_t t_lane = tmem + ((uint32_t)(quad * 32) << 16);
    uint8_t* prow = sP + r * 128;
    const int sw = r & 7;
    const uint64_t scale2 = pack_f2(p.scale_log2, p.scale_log2);
    const int nfull = p.causal ? 0 : (p.lk >> 5);         // 32-key chunks no row needs a mask for
    const int nch = (p.lk + 31) >> 5;                     // chunks holding keys at all (the rest of S is never read)
    uint8_t* stage = sO + quad * 4096 + lane * 128;       // this thread's 128-byte output row in the warp's staging box
    // normalise O_i (TMEM buffer i&1) by the row sum and hand the warp's 32 x 64 box to TMA (a row-per-thread store to global
    // memory is 32 different lines per instruction; it cost as much as the exponentials)
    auto store_block = [&](int i, float l) {
      const float inv = 1.f / l;
      if (lane == 0) bulk_wait_read<0>();                 // the previous box has left the staging buffer
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t o[32];
        tmem_ld32(t_lane + 128 + (i & 1) * 64 + c * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 u;
          u.x = pack_bf16(__uint_as_float(o[g * 8 + 0]) * inv, __uint_as_float(o[g * 8 + 1]) * inv);
          u.y = pack_bf16(__uint_as_float(o[g * 8 + 2]) * inv, __uint_as_float(o[g * 8 + 3]) * inv);
          u.z = pack_bf16(__uint_as_float(o[g * 8 + 4]) * inv, __uint_as_float(o[g * 8 + 5]) * inv);
          u.w = pack_bf16(__uint_as_float(o[g * 8 + 6]) * inv, __uint_as_float(o[g * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(stage + (((c * 4 + g) ^ sw) << 4)) = u;
        }
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars->o_read[i & 1]);
        tma_store_3d(&p.omap, sO + quad * 4096, head * kD, (first + i) * kQ + quad * 32, b);
        bulk_commit();
      }
    };
    float l_prev = 1.f;
    for (int i = 0; i < n_it; ++i) {
      mbar_wait(&bars->s_full, i & 1);
      tc_fence_after();
      uint32_t v[128];
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (c * 32 < p.lk) tmem_ld32(t_lane + c * 32, reinterpret_cast<uint32_t(&)[32]>(v[c * 32]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->s_read);          // the tensor core may now compute S_{i+1}
      // keys this row may attend to: [0, limit)
      const int limit = p.causal ? min(p.lk, (first + i) * kQ + r + 1) : p.lk;
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < nfull) {
          float m1 = mx, m2 = mx, m3 = mx;
#pragma unroll
          for (int k = c * 32; k < c * 32 + 32; k += 8) {
            mx = max3_f(mx, __uint_as_float(v[k]), __uint_as_float(v[k + 1]));
            m1 = max3_f(m1, __uint_as_float(v[k + 2]), __uint_as_float(v[k + 3]));
            m2 = max3_f(m2, __uint_as_float(v[k + 4]), __uint_as_float(v[k + 5]));
            m3 = max3_f(m3, __uint_as_float(v[k + 6]), __uint_as_float(v[k + 7]));
          }
          mx = fmaxf(max3_f(mx, m1, m2), m3);
        } else if (c < nch) {
#pragma unroll
          for (int k = c * 32; k < c * 32 + 32; ++k)
            if (k < limit) mx = fmaxf(mx, __uint_as_float(v[k]));
        }
      }
      if (i > 0) {
        mbar_wait(&bars->o_done[(i - 1) & 1], ((i - 1) >> 1) & 1);   // P_{i-1} V finished: sP may be rewritten, O_{i-1} is final
        tc_fence_after();
      }
      const float nmxs = -mx * p.scale_log2;
      const uint64_t nmxs2 = pack_f2(nmxs, nmxs);
      uint64_t lsum2 = 0ull;
      auto exp_chunk = [&](int c, auto full_tag) {
        constexpr bool kFull = decltype(full_tag)::value;
        uint32_t pk[16];
#pragma unroll
        for (int k = 0; k < 32; k += 2) {
          const uint64_t t = fma_f2(pack_f2(__uint_as_float(v[c * 32 + k]), __uint_as_float(v[c * 32 + k + 1])), scale2, nmxs2);
          float a0, a1;
          unpack_f2(t, a0, a1);
          a0 = fast_exp2(a0);
          a1 = fast_exp2(a1);
          if constexpr (!kFull) {
            if (c * 32 + k >= limit) a0 = 0.f;
            if (c * 32 + k + 1 >= limit) a1 = 0.f;
          }
          lsum2 = add_f2(lsum2, pack_f2(a0, a1));
          pk[k >> 1] = pack_bf16(a0, a1);
        }
        uint8_t* half_base = prow + (c >> 1) * kTile;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int chunk = (c & 1) * 4 + g;
          *reinterpret_cast<uint4*>(half_base + ((chunk ^ sw) << 4)) = make_uint4(pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
        }
      };
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < nfull) exp_chunk(c, std::true_type{});
        else if (c < nch) exp_chunk(c, std::false_type{});
      }
      float ls0, ls1;
      unpack_f2(lsum2, ls0, ls1);
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full);
      if (i > 0) store_block(i - 1, l_prev);              // while the tensor core computes P_i V
      l_prev = ls0 + ls1;
    }
    mbar_wait(&bars->o_done[(n_it - 1) & 1], ((n_it - 1) >> 1) & 1);
    tc_fence_after();
    store_block(n_it - 1, l_prev);
    if (lane == 0) bulk_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}
}  // namespace fd

using namespace fd;

static int attention_d64_impl(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                              int64_t ldo, int32_t batch, int32_t heads, int32_t lq, int32_t lk, float scale, int causal,
                              void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(!causal || lk <= kKV, "fd_attention_d64_causal: at most %d keys (got %d)", kKV, lk);
  FD_REQUIRE(q && k && v && out && batch > 0 && heads > 0 && lq > 0 && lk > 0, "fd_attention_d64: bad arguments");
  FD_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "fd_attention_d64: pitches must be multiples of 8");
  FD_REQUIRE(((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)v % 16) == 0 && ((uintptr_t)out % 16) == 0,
             "fd_attention_d64: pointers must be 16-byte aligned");
  AttnParams p;
  memset(&p, 0, sizeof(p));
  const void* ptr[3] = {q, k, v};
  const int64_t ld[3] = {ldq, ldk, ldv};
  const int64_t rows[3] = {(int64_t)batch * lq, (int64_t)batch * lk, (int64_t)batch * lk};
  CUtensorMap* maps[3] = {&p.qmap, &p.kmap, &p.vmap};
  for (int i = 0; i < 3; ++i) {
    uint64_t dims[2] = {(uint64_t)heads * kD, (uint64_t)rows[i]};
    uint64_t str[1] = {(uint64_t)ld[i] * 2};
    uint32_t box[2] = {kD, 128};
    if (encode_tensor_map(maps[i], ptr[i], 2, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_128B)) return 3;
  }
  p.out = (__nv_bfloat16*)out; p.ldo = ldo; p.lq = lq; p.lk = lk; p.heads = heads;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal;
  const size_t smem_bytes = 7 * kTile + sizeof(AttnBars);                    // 128-key blocks: Q + 2 K + 2 V + 2 P halves
  const size_t smem_kv64 = 4 * kTile + sizeof(AttnBars);                     // 64-key blocks: Q + 2 K/2 + 2 V/2 + 1 P half
  const size_t smem_short = 7 * kTile + sizeof(AttnShortBars);
  static int kv64 = -1;
  if (kv64 < 0) { const char* e = getenv("FD_ATTN_KV64"); kv64 = e ? atoi(e) : 1; }
  const bool use64 = kv64 && lk > kKV;
  if (use64) {            // K / V boxes of 64 keys
    for (int i = 1; i < 3; ++i) {
      uint64_t dims[2] = {(uint64_t)heads * kD, (uint64_t)rows[i]};
      uint64_t str[1] = {(uint64_t)ld[i] * 2};
      uint32_t box[2] = {kD, 64};
      if (encode_tensor_map(maps[i], ptr[i], 2, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_128B)) return 3;
    }
  }
  static bool attr = false;
  if (!attr) {
    FD_CHECK_CUDA(cudaFuncSetAttribute((attention_d64_kernel<0, 128>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    FD_CHECK_CUDA(cudaFuncSetAttribute((attention_d64_kernel<3, 128>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    FD_CHECK_CUDA(cudaFuncSetAttribute((attention_d64_kernel<0, 64>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv64));
    FD_CHECK_CUDA(cudaFuncSetAttribute((attention_d64_kernel<2, 64>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv64));
    FD_CHECK_CUDA(cudaFuncSetAttribute((attention_d64_kernel<3, 64>), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv64));
    FD_CHECK_CUDA(cudaFuncSetAttribute(attention_d64_shortkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_short));
    attr = true;
  }
  const int nqb = (lq + kQ - 1) / kQ;
  if (lk <= kKV) {
    // as few query blocks per CTA as still fit all CTAs in one wave (two CTAs per SM)
    const long long slots = 2ll * num_sms();
    int qpc = 1;
    while (qpc < nqb && (long long)((nqb + qpc - 1) / qpc) * heads * batch > slots) ++qpc;
    { const char* e = getenv("FD_ATTN_QPC"); if (e && atoi(e) > 0) qpc = atoi(e) < nqb ? atoi(e) : nqb; }   // tuning override
    p.q_per_cta = qpc;
    {
      uint64_t dims[3] = {(uint64_t)heads * kD, (uint64_t)lq, (uint64_t)batch};
      uint64_t str[2] = {(uint64_t)ldo * 2, (uint64_t)ldo * 2 * lq};
      uint32_t box[3] = {kD, 32, 1};
      if (encode_tensor_map(&p.omap, out, 3, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_128B)) return 3;
    }
    dim3 grid((nqb + qpc - 1) / qpc, heads, batch);
    FD_CHECK_CUDA(launch_pdl(attention_d64_shortkv_kernel, grid, dim3(192), smem_short, stream, 1, p));
  } else {
    dim3 grid(nqb, heads, batch);
    // FD_ATTN_POLY = 2 / 3: that many of every 8 exponential pairs on the FMA pipe (opt-in); FD_ATTN_KV64=0: 128-key blocks, 2 CTAs / SM
    static int poly = -1;
    if (poly < 0) { const char* e = getenv("FD_ATTN_POLY"); poly = e ? atoi(e) : 0; }
    if (use64) {
      if (poly == 2) FD_CHECK_CUDA(launch_pdl((attention_d64_kernel<2, 64>), grid, dim3(192), smem_kv64, stream, 1, p));
      else if (poly == 3) FD_CHECK_CUDA(launch_pdl((attention_d64_kernel<3, 64>), grid, dim3(192), smem_kv64, stream, 1, p));
      else FD_CHECK_CUDA(launch_pdl((attention_d64_kernel<0, 64>), grid, dim3(192), smem_kv64, stream, 1, p));
    } else if (poly == 3) FD_CHECK_CUDA(launch_pdl((attention_d64_kernel<3, 128>), grid, dim3(192), smem_bytes, stream, 1, p));
    else FD_CHECK_CUDA(launch_pdl((attention_d64_kernel<0, 128>), grid, dim3(192), smem_bytes, stream, 1, p));
  }
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}

#ifdef FD_ATTN_TRACE
extern "C" int fd_debug_read_attn_trace(long long* host_out) {
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(host_out, fd::g_attn_trace, sizeof(long long) * 16 * 12);
}
#endif

extern "C" int fd_attention_d64(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                                int64_t ldo, int32_t batch, int32_t heads, int32_t lq, int32_t lk, float scale, void* stream) {
  return attention_d64_impl(q, ldq, k, ldk, v, ldv, out, ldo, batch, heads, lq, lk, scale, 0, stream);
}

extern "C" int fd_attention_d64_causal(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                       void* out, int64_t ldo, int32_t batch, int32_t heads, int32_t lq, int32_t lk, float scale,
                                       void* stream) {
  return attention_d64_impl(q, ldq, k, ldk, v, ldv, out, ldo, batch, heads, lq, lk, scale, 1, stream);
}
