// Flash-style attention for head_dim 64 on tcgen05 (sm_100a), bf16 in/out, fp32 softmax + accumulation.
// One CTA = one (batch, head, 128-query block); two CTAs co-reside per SM so one CTA's softmax overlaps the
// other's MMAs.
//   warp 0 lane 0 : TMA producer  (Q once; K/V 128-key blocks, 2-stage ring, 128B swizzle)
//   warp 1 lane 0 : MMA issuer    S = Q·K^T  (M128 N128 K64, K-major A/B)  -> TMEM cols [0,128)
//                                 O_j = P·V  (M128 N64 K128, A=P K-major from smem, B=V MN-major) -> TMEM cols [128,192)
//   warps 2..5    : softmax/epilogue — one query row per thread: TMEM->regs, online max/sum in fp32 (exp2),
//                   P (bf16) written to swizzled smem, running O kept in registers and rescaled each block.
#include <type_traits>
#include "common.cuh"
#include "../../include/faithdiff_b200.h"

namespace fd {
extern int64_t g_launches;

struct AttnParams {
  CUtensorMap qmap, kmap, vmap;
  __nv_bfloat16* out;
  long long ldo;
  int lq, lk, heads;
  float scale_log2;
};

constexpr int kQ = 128, kKV = 128, kD = 64;
constexpr int kTile = kQ * kD * 2;   // 16 KB

struct __align__(8) AttnBars {
  uint64_t q_full, kv_full[2], kv_empty[2], s_full, p_full, o_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(192, 2) attention_d64_kernel(const __grid_constant__ AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kTile;            // 2 stages
  uint8_t* sV = smem + 3 * kTile;        // 2 stages
  uint8_t* sP = smem + 5 * kTile;        // 2 halves (64 keys each)
  AttnBars* bars = reinterpret_cast<AttnBars*>(smem + 7 * kTile);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qblk = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
  const int nkv = (p.lk + kKV - 1) / kKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.qmap); tma_prefetch_desc(&p.kmap); tma_prefetch_desc(&p.vmap);
    mbar_init(&bars->q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(&bars->kv_full[s], 1); mbar_init(&bars->kv_empty[s], 1); }
    mbar_init(&bars->s_full, 1);
    mbar_init(&bars->p_full, 4);
    mbar_init(&bars->o_full, 1);
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
      mbar_arrive_expect_tx(&bars->q_full, kTile);
      tma_load_2d(&p.qmap, &bars->q_full, sQ, head * kD, b * p.lq + qblk * kQ);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        mbar_wait(&bars->kv_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&bars->kv_full[s], 2 * kTile);
        tma_load_2d(&p.kmap, &bars->kv_full[s], sK + s * kTile, head * kD, b * p.lk + j * kKV);
        tma_load_2d(&p.vmap, &bars->kv_full[s], sV + s * kTile, head * kD, b * p.lk + j * kKV);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc(128, 128, 1, 0, 0);
      const uint32_t idesc_o = umma_idesc(128, 64, 1, 0, 1);   // B (=V) is MN-major
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
      mbar_wait(&bars->q_full, 0);
      for (int j = 0; j < nkv; ++j) {
        const int s = j & 1;
        mbar_wait(&bars->kv_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t aK = smem_u32(sK + s * kTile), aV = smem_u32(sV + s * kTile);
        // S = Q K^T. (S_{j} may overwrite S_{j-1}: the softmax warps finished reading it before p_full(j-1).)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(tmem, umma_desc_sw128(aQ + k * 32, 16, 1024), umma_desc_sw128(aK + k * 32, 16, 1024), idesc_s, k != 0);
        tc_commit(&bars->s_full);
        // O_j = P V once the softmax warps have published P_j
        mbar_wait(&bars->p_full, j & 1);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          umma_f16(tmem + 128, umma_desc_sw128(aP + (k >> 2) * kTile + (k & 3) * 32, 16, 1024),
                   umma_desc_sw128(aV + k * 2048, 0, 1024), idesc_o, k != 0);
        tc_commit(&bars->kv_empty[s]);
        tc_commit(&bars->o_full);
      }
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                       // query row within the block
    const uint32_t t_lane = tmem + ((uint32_t)(quad * 32) << 16);
    uint64_t o2[32];                                      // running output row, packed fp32x2
#pragma unroll
    for (int i = 0; i < 32; ++i) o2[i] = 0ull;
    float m = -INFINITY, l = 0.f;
    uint8_t* prow = sP + r * 128;
    const int sw = r & 7;
    const uint64_t scale2 = pack_f2(p.scale_log2, p.scale_log2);
    for (int j = 0; j < nkv; ++j) {
      mbar_wait(&bars->s_full, j & 1);
      tc_fence_after();
      const int kv0 = j * kKV;
      const bool full = kv0 + kKV <= p.lk;               // only the last block can need masking
      // The block is either completely inside the key range (the common case: no index arithmetic at all) or the last,
      // partially masked one; `full` is warp-uniform, so the two variants are separate code paths rather than selects.
      float mx = m;
      float alpha, nmxs;
      uint64_t lsum2 = 0ull;
      auto softmax_block = [&](auto full_tag) {
        constexpr bool kFull = decltype(full_tag)::value;
        // pass 1: row max (3-input max halves the instruction count)
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(t_lane + c * 32, v);
          tmem_ld_wait();
          if constexpr (kFull) {
#pragma unroll
            for (int i = 0; i < 32; i += 2) mx = max3_f(mx, __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (kv0 + c * 32 + i < p.lk) mx = fmaxf(mx, __uint_as_float(v[i]));
          }
        }
        alpha = fast_exp2((m - mx) * p.scale_log2);
        nmxs = -mx * p.scale_log2;
        const uint64_t nmxs2 = pack_f2(nmxs, nmxs);
        // pass 2: p = exp2(s*scale - max*scale) -> bf16 -> swizzled smem (scale/subtract and the row sum on packed fp32x2)
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t v[32];
          tmem_ld32(t_lane + c * 32, v);
          tmem_ld_wait();
          float e[32];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const uint64_t t = fma_f2(pack_f2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), scale2, nmxs2);
            float a0, a1;
            unpack_f2(t, a0, a1);
            a0 = fast_exp2(a0);      // (an f16x2 exp2 — two results per MUFU op — was measured slower, twice: the extra
            a1 = fast_exp2(a1);      //  converts cost more issue slots than the halved MUFU count saves)
            if constexpr (!kFull) {
              if (kv0 + c * 32 + i >= p.lk) a0 = 0.f;
              if (kv0 + c * 32 + i + 1 >= p.lk) a1 = 0.f;
            }
            e[i] = a0; e[i + 1] = a1;
            lsum2 = add_f2(lsum2, pack_f2(a0, a1));
          }
          uint8_t* half_base = prow + (c >> 1) * kTile;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int chunk = (c & 1) * 4 + g;
            uint4 u;
            u.x = pack_bf16(e[g * 8 + 0], e[g * 8 + 1]); u.y = pack_bf16(e[g * 8 + 2], e[g * 8 + 3]);
            u.z = pack_bf16(e[g * 8 + 4], e[g * 8 + 5]); u.w = pack_bf16(e[g * 8 + 6], e[g * 8 + 7]);
            *reinterpret_cast<uint4*>(half_base + ((chunk ^ sw) << 4)) = u;
          }
        }
      };
      if (full) softmax_block(std::true_type{}); else softmax_block(std::false_type{});
      float ls0, ls1;
      unpack_f2(lsum2, ls0, ls1);
      l = l * alpha + (ls0 + ls1);
      m = mx;
      fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->p_full);
      // rescale the running output while the PV MMA runs; skipped when no row of the warp raised its max
      if (__any_sync(0xffffffffu, alpha != 1.f)) {
        const uint64_t al2 = pack_f2(alpha, alpha);
#pragma unroll
        for (int i = 0; i < 32; ++i) o2[i] = mul_f2(o2[i], al2);
      }
      mbar_wait(&bars->o_full, j & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t v[32];
        tmem_ld32(t_lane + 128 + c * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; i += 2)
          o2[c * 16 + (i >> 1)] = add_f2(o2[c * 16 + (i >> 1)], pack_f2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
      }
      tc_fence_before();
    }
    float o[64];
#pragma unroll
    for (int i = 0; i < 32; ++i) unpack_f2(o2[i], o[2 * i], o[2 * i + 1]);
    const int qi = qblk * kQ + r;
    if (qi < p.lq) {
      const float inv = 1.f / l;
      __nv_bfloat16* dst = p.out + ((long long)b * p.lq + qi) * p.ldo + head * kD;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint4 u;
        u.x = pack_bf16(o[g * 8 + 0] * inv, o[g * 8 + 1] * inv); u.y = pack_bf16(o[g * 8 + 2] * inv, o[g * 8 + 3] * inv);
        u.z = pack_bf16(o[g * 8 + 4] * inv, o[g * 8 + 5] * inv); u.w = pack_bf16(o[g * 8 + 6] * inv, o[g * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(dst + g * 8) = u;
      }
    }
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

extern "C" int fd_attention_d64(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                                int64_t ldo, int32_t batch, int32_t heads, int32_t lq, int32_t lk, float scale, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
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
  const size_t smem_bytes = 7 * kTile + sizeof(AttnBars);
  static bool attr = false;
  if (!attr) {
    FD_CHECK_CUDA(cudaFuncSetAttribute(attention_d64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    attr = true;
  }
  dim3 grid((lq + kQ - 1) / kQ, heads, batch);
  FD_CHECK_CUDA(launch_pdl(attention_d64_kernel, grid, dim3(192), smem_bytes, stream, 1, p));
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}
