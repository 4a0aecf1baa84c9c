// Flash-style single-head attention with head_dim 512 on tcgen05 (sm_100a): the mid-block attention of the FaithDiff LQ encoder
// and of the SDXL VAE (FaithDiff/models/unet_2d_condition_vae_extension.py:96-106; diffusers Attention with heads = 1,
// dim_head = C = 512) at L = 4096 (512-px tile) ... 16384 (1024-px tile) tokens.  bf16 in/out, fp32 softmax and accumulation;
// no [L, L] score matrix ever reaches HBM.
//
// A 128-query x 512-column fp32 output accumulator is ALL 512 TMEM columns of an SM, which leaves no room for S.  So the output
// columns are split over TWO CTAs (blockIdx.y = half h: columns [256h, 256h+256)): each keeps its O half (256 columns) plus a
// double-buffered S tile (2 x 128 columns) in TMEM and recomputes S = Q K^T itself (1.5x the MMA work of an ideal kernel, in
// exchange for no cross-CTA softmax exchange).  Per CTA:
//   warp 0 lane 0 : TMA producer — Q once (8 d-chunks of 128 x 64, resident, 128 KB); per 128-key block the 8 K chunks then the 4
//                   V tiles of this CTA's column half stream through ONE 4-stage ring of 16 KB tiles, in the order the MMA warp
//                   consumes them: K_0 | K_1, V_0 | K_2, V_1 | ...
//   warp 1 lane 0 : MMA issuer  — S_{j+1} = Q K_{j+1}^T (32 x M128 N128 K16 into S buffer (j+1)&1) is issued BEFORE O += P_j V_j
//                   (32 x M128 N64 K16), so the tensor pipe works on S_{j+1} while the softmax warps are in block j's exponentials
//   warps 2..5    : softmax, one query row per thread, exactly the scheme of attention_d64_kernel: the S row is read from TMEM
//                   once, max / exp2 / row sum in registers, P (bf16) to swizzled smem; O stays in TMEM and is rescaled in place
//                   only when a row maximum grew by more than 2^8.
#include <type_traits>
#include "common.cuh"
#include "../../include/faithdiff_b200.h"

namespace fd {
extern int64_t g_launches;

struct Attn512Params {
  CUtensorMap qmap, kmap, vmap;      // [rows, 512] bf16 views, box {64, 128}, 128B swizzle
  __nv_bfloat16* out;
  long long ldo;
  int l;                             // tokens per sample (queries == keys)
  float scale_log2;
};

namespace a512 {
constexpr int kQ = 128, kKV = 128, kD = 512, kChunk = 64;
constexpr int kTile = kQ * kChunk * 2;       // 16 KB: 128 rows x 64 bf16
constexpr int kStages = 4;
constexpr int kDChunks = kD / kChunk;        // 8
constexpr int kVTiles = 4;                   // 256 output columns per CTA
constexpr float kRescaleThreshold = 8.f;     // log2 units
struct __align__(8) Bars {
  uint64_t q_full, full[kStages], empty[kStages], s_full[2], s_read[2], p_full, o_done;
  uint32_t tmem_base;
};
constexpr size_t kSmem = (size_t)(kDChunks + kStages + 2) * kTile + sizeof(Bars) + 1024;
}  // namespace a512

__global__ void __launch_bounds__(192, 1) attention_d512_kernel(const __grid_constant__ Attn512Params p) {
  using namespace a512;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                                   // 8 chunks
  uint8_t* sR = smem + kDChunks * kTile;                // ring
  uint8_t* sP = sR + kStages * kTile;                   // 2 halves of 64 keys
  Bars* bars = reinterpret_cast<Bars*>(sP + 2 * kTile);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qblk = blockIdx.x, half = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.l + kKV - 1) / kKV;
  const int row0 = b * p.l;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.qmap); tma_prefetch_desc(&p.kmap); tma_prefetch_desc(&p.vmap);
    mbar_init(&bars->q_full, 1);
    for (int s = 0; s < kStages; ++s) { mbar_init(&bars->full[s], 1); mbar_init(&bars->empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&bars->s_full[s], 1); mbar_init(&bars->s_read[s], 4); }
    mbar_init(&bars->p_full, 4);
    mbar_init(&bars->o_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(&bars->tmem_base, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const uint32_t tmem_o = tmem + 256;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(&bars->q_full, kDChunks * kTile);
      for (int c = 0; c < kDChunks; ++c) tma_load_2d(&p.qmap, &bars->q_full, sQ + c * kTile, c * kChunk, row0 + qblk * kQ);
      int it = 0;
      auto push = [&](const CUtensorMap* map, int col, int row) {
        const int s = it % kStages;
        if (it >= kStages) mbar_wait(&bars->empty[s], ((it / kStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->full[s], kTile);
        tma_load_2d(map, &bars->full[s], sR + s * kTile, col, row);
        ++it;
      };
      auto push_k = [&](int j) { for (int c = 0; c < kDChunks; ++c) push(&p.kmap, c * kChunk, row0 + j * kKV); };
      auto push_v = [&](int j) { for (int t = 0; t < kVTiles; ++t) push(&p.vmap, (half * kVTiles + t) * kChunk, row0 + j * kKV); };
      push_k(0);
      for (int j = 0; j < nkv; ++j) {
        if (j + 1 < nkv) push_k(j + 1);
        push_v(j);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc(128, 128, 1, 0, 0);
      const uint32_t idesc_o = umma_idesc(128, 64, 1, 0, 1);     // B (= V tile) is MN-major
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP), aR = smem_u32(sR);
      int it = 0;
      auto issue_s = [&](int j) {
        const uint32_t d = tmem + (uint32_t)(j & 1) * 128;
        for (int c = 0; c < kDChunks; ++c, ++it) {
          const int s = it % kStages;
          mbar_wait(&bars->full[s], (it / kStages) & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(d, umma_desc_sw128(aQ + c * kTile + k * 32, 16, 1024), umma_desc_sw128(aR + s * kTile + k * 32, 16, 1024), idesc_s,
                     (c | k) != 0);
          tc_commit(&bars->empty[s]);
        }
        tc_commit(&bars->s_full[j & 1]);
      };
      mbar_wait(&bars->q_full, 0);
      tc_fence_after();
      issue_s(0);
      for (int j = 0; j < nkv; ++j) {
        if (j + 1 < nkv) {
          // S_{j+1} overwrites the buffer of S_{j-1}: every softmax warp must hold S_{j-1} in registers
          if (j >= 1) mbar_wait(&bars->s_read[(j + 1) & 1], ((j - 1) >> 1) & 1);
          tc_fence_after();
          issue_s(j + 1);
        }
        mbar_wait(&bars->p_full, j & 1);
        tc_fence_after();
        for (int t = 0; t < kVTiles; ++t, ++it) {
          const int s = it % kStages;
          mbar_wait(&bars->full[s], (it / kStages) & 1);
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < 8; ++k)
            umma_f16(tmem_o + t * 64, umma_desc_sw128(aP + (k >> 2) * kTile + (k & 3) * 32, 16, 1024),
                     umma_desc_sw128(aR + s * kTile + k * 2048, 0, 1024), idesc_o, (j != 0) || (k != 0));
          tc_commit(&bars->empty[s]);
        }
        tc_commit(&bars->o_done);
      }
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                       // query row within the block
    const uint32_t t_lane = tmem + ((uint32_t)(quad * 32) << 16);
    const uint32_t o_lane = t_lane + 256;
    float m_used = -INFINITY, l = 0.f;
    uint8_t* prow = sP + r * 128;
    const int sw = r & 7;
    const uint64_t scale2 = pack_f2(p.scale_log2, p.scale_log2);
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&bars->s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      uint32_t v[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(t_lane + (j & 1) * 128 + c * 32, reinterpret_cast<uint32_t(&)[32]>(v[c * 32]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->s_read[j & 1]);
      const int kv0 = j * kKV;
      const bool full = kv0 + kKV <= p.l;                // only the last block can need masking (warp-uniform)
      float mx = m_used;
      if (full) {
        float m1 = mx, m2 = mx, m3 = mx;
#pragma unroll
        for (int i = 0; i < 128; i += 8) {
          mx = max3_f(mx, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
          m1 = max3_f(m1, __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
          m2 = max3_f(m2, __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
          m3 = max3_f(m3, __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
        }
        mx = fmaxf(max3_f(mx, m1, m2), m3);
      } else {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (kv0 + i < p.l) mx = fmaxf(mx, __uint_as_float(v[i]));
      }
      if (j == 0) {
        m_used = mx;
      } else {
        mbar_wait(&bars->o_done, (j - 1) & 1);            // P_{j-1} V_{j-1} finished: sP may be rewritten, O may be rescaled
        tc_fence_after();
        const bool grew = (mx - m_used) * p.scale_log2 > kRescaleThreshold;
        if (__any_sync(0xffffffffu, grew)) {
          const float alpha = fast_exp2((m_used - mx) * p.scale_log2);
#pragma unroll 1
          for (int q = 0; q < 16; ++q) {
            uint32_t o[16];
            tmem_ld16(o_lane + q * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st16(o_lane + q * 16, o);
          }
          tmem_st_wait();
          l *= alpha;
          m_used = mx;
        }
      }
      const float nmxs = -m_used * p.scale_log2;
      const uint64_t nmxs2 = pack_f2(nmxs, nmxs);
      uint64_t lsum2 = 0ull;
      auto exp_block = [&](auto full_tag) {
        constexpr bool kFull = decltype(full_tag)::value;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const uint64_t t = fma_f2(pack_f2(__uint_as_float(v[c * 32 + i]), __uint_as_float(v[c * 32 + i + 1])), scale2, nmxs2);
            float a0, a1;
            unpack_f2(t, a0, a1);
            a0 = fast_exp2(a0);
            a1 = fast_exp2(a1);
            if constexpr (!kFull) {
              if (kv0 + c * 32 + i >= p.l) a0 = 0.f;
              if (kv0 + c * 32 + i + 1 >= p.l) a1 = 0.f;
            }
            lsum2 = add_f2(lsum2, pack_f2(a0, a1));
            pk[i >> 1] = pack_bf16(a0, a1);
          }
          uint8_t* half_base = prow + (c >> 1) * kTile;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int chunk = (c & 1) * 4 + g;
            *reinterpret_cast<uint4*>(half_base + ((chunk ^ sw) << 4)) = make_uint4(pk[g * 4], pk[g * 4 + 1], pk[g * 4 + 2], pk[g * 4 + 3]);
          }
        }
      };
      if (full) exp_block(std::true_type{}); else exp_block(std::false_type{});
      float ls0, ls1;
      unpack_f2(lsum2, ls0, ls1);
      l += ls0 + ls1;
      fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full);
    }
    mbar_wait(&bars->o_done, (nkv - 1) & 1);
    tc_fence_after();
    const int qi = qblk * kQ + r;
    const float inv = 1.f / l;
    __nv_bfloat16* dst = p.out + ((long long)row0 + qi) * p.ldo + half * 256;
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      uint32_t o[32];
      tmem_ld32(o_lane + c * 32, o);
      tmem_ld_wait();
      if (qi < p.l) {
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
    tmem_dealloc(tmem, 512);
  }
}
}  // namespace fd

using namespace fd;

extern "C" int fd_attention_d512(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                                 int64_t ldo, int32_t batch, int32_t l, float scale, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  FD_REQUIRE(q && k && v && out && batch > 0 && l > 0, "fd_attention_d512: bad arguments");
  FD_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "fd_attention_d512: pitches must be multiples of 8");
  FD_REQUIRE(((uintptr_t)q % 16) == 0 && ((uintptr_t)k % 16) == 0 && ((uintptr_t)v % 16) == 0 && ((uintptr_t)out % 16) == 0,
             "fd_attention_d512: pointers must be 16-byte aligned");
  Attn512Params p;
  memset(&p, 0, sizeof(p));
  const void* ptr[3] = {q, k, v};
  const int64_t ld[3] = {ldq, ldk, ldv};
  CUtensorMap* maps[3] = {&p.qmap, &p.kmap, &p.vmap};
  for (int i = 0; i < 3; ++i) {
    uint64_t dims[2] = {(uint64_t)a512::kD, (uint64_t)batch * l};
    uint64_t str[1] = {(uint64_t)ld[i] * 2};
    uint32_t box[2] = {a512::kChunk, 128};
    if (encode_tensor_map(maps[i], ptr[i], 2, dims, str, box, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, CU_TENSOR_MAP_SWIZZLE_128B)) return 3;
  }
  p.out = (__nv_bfloat16*)out; p.ldo = ldo; p.l = l;
  p.scale_log2 = scale * 1.4426950408889634f;
  static bool attr = false;
  if (!attr) {
    FD_CHECK_CUDA(cudaFuncSetAttribute(attention_d512_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)a512::kSmem));
    attr = true;
  }
  dim3 grid((l + a512::kQ - 1) / a512::kQ, 2, batch);
  attention_d512_kernel<<<grid, dim3(192), a512::kSmem, stream>>>(p);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}
