// Per-step exchange of the overlap-region accumulation over NVLink peer memory, fused with the divide by the count map
// (opt-in, FD_P2P_EXCHANGE=1; NOT yet run on a multi-GPU box — the default multi-GPU path is one NCCL all-reduce followed by
// fd_canvas_finalize).
//
// The blend of pipeline_FaithDiff_tlc.py:230-252 is a linear weighted sum, so each rank accumulates its (tile, CFG-branch)
// units into an fp32 numerator canvas (fd_unit_accumulate).  Here that canvas lives in a cudaMalloc'd region every peer has
// mapped through CUDA IPC.  ONE kernel per rank and scheduler step then
//   (1) publishes "my canvas of step s is complete" with a release store into every peer's flag array (P2P store over NVLink),
//   (2) waits until all peers have published theirs (acquire loads on its own flags),
//   (3) reads every rank's canvas in RANK ORDER with 16-byte P2P loads, sums (the same order on every rank: all ranks hold
//       bit-identical canvases, and a given world size reproduces itself run to run), divides by the static count map and
//       writes the fp32 canvas and its bf16/fp16 copy.
// Two parity slots per region: while a slow peer still reads parity p of step s, this rank already accumulates parity p^1 of
// step s+1; it reuses parity p at step s+2, after it has seen every peer's step-(s+1) flag, which a peer only publishes once its
// step-s kernel (the reader) has completed.  Region layout: [canvas parity 0: n floats][canvas parity 1: n floats][flags].
#include "common.cuh"
#include "../../include/faithdiff_b200.h"

namespace fd {
extern int64_t g_launches;
constexpr int kPeerMaxWorld = 16;
constexpr int kPeerFlagWords = 2 * kPeerMaxWorld;

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// peer lines may sit in this SM's L1 from two steps ago (peer memory is not cached in the local L2, but it is in L1):
// volatile loads always go to the owner
__device__ __forceinline__ float4 ld_peer16(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

struct PeerRegions {
  float* base[kPeerMaxWorld];
};

__global__ void __launch_bounds__(256)
canvas_exchange_finalize_kernel(PeerRegions regions, int rank, int world, int parity, unsigned step, long long n, long long n_pad,
                                const float* __restrict__ count, float* __restrict__ out, void* __restrict__ lowp, int lowp_dtype) {
  if (threadIdx.x == 0) {
    if (blockIdx.x == 0) {
      __threadfence_system();          // the accumulate kernels before this one in the stream wrote my canvas
      for (int r = 0; r < world; ++r)
        st_release_sys(reinterpret_cast<unsigned*>(regions.base[r] + 2 * n_pad) + parity * kPeerMaxWorld + rank, step);
    }
    const unsigned* mine = reinterpret_cast<const unsigned*>(regions.base[rank] + 2 * n_pad) + parity * kPeerMaxWorld;
    const long long t0 = clock64();
    for (int r = 0; r < world; ++r) {
      while (ld_acquire_sys(mine + r) != step) {
        __nanosleep(64);
        if (clock64() - t0 > 20000000000LL) __trap();      // ~10 s: a peer never arrived — fail instead of hanging the box
      }
    }
  }
  __syncthreads();
  const long long nvec = n >> 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {                        // fixed rank order
      const float4 v = ld_peer16(regions.base[r] + (long long)parity * n_pad + i * 4);
      acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y); acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
    }
    const float4 c = __ldg(reinterpret_cast<const float4*>(count) + i);
    const float o[4] = {__fdiv_rn(acc.x, c.x), __fdiv_rn(acc.y, c.y), __fdiv_rn(acc.z, c.z), __fdiv_rn(acc.w, c.w)};
    if (out) reinterpret_cast<float4*>(out)[i] = make_float4(o[0], o[1], o[2], o[3]);
    if (lowp) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (lowp_dtype == 0) reinterpret_cast<__nv_bfloat16*>(lowp)[i * 4 + j] = __float2bfloat16_rn(o[j]);
        else if (lowp_dtype == 1) reinterpret_cast<__half*>(lowp)[i * 4 + j] = __float2half_rn(o[j]);
        else reinterpret_cast<float*>(lowp)[i * 4 + j] = o[j];
      }
    }
  }
}
}  // namespace fd

using namespace fd;

extern "C" int64_t fd_peer_region_bytes(int64_t n_floats) {
  const int64_t n_pad = (n_floats + 63) / 64 * 64;
  return 2 * n_pad * (int64_t)sizeof(float) + kPeerFlagWords * (int64_t)sizeof(unsigned);
}

extern "C" int fd_peer_alloc(void** ptr, int64_t bytes) {
  FD_REQUIRE(ptr && bytes > 0, "fd_peer_alloc: bad arguments");
  FD_CHECK_CUDA(cudaMalloc(ptr, (size_t)bytes));
  FD_CHECK_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
  FD_CHECK_CUDA(cudaDeviceSynchronize());
  return 0;
}

extern "C" int fd_peer_free(void* ptr) {
  if (ptr) FD_CHECK_CUDA(cudaFree(ptr));
  return 0;
}

extern "C" int fd_peer_ipc_handle(void* ptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  FD_REQUIRE(ptr && handle64, "fd_peer_ipc_handle: null pointer");
  FD_CHECK_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), ptr));
  return 0;
}

extern "C" int fd_peer_ipc_open(const void* handle64, void** ptr) {
  FD_REQUIRE(ptr && handle64, "fd_peer_ipc_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  FD_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

extern "C" int fd_peer_ipc_close(void* ptr) {
  if (ptr) FD_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

extern "C" int fd_canvas_exchange_finalize(const void* const* regions, int32_t rank, int32_t world, int32_t parity, uint32_t step,
                                           int64_t n, const float* count, float* canvas_f32, void* canvas_lowp, int32_t lowp_dtype,
                                           void* stream_) {
  FD_REQUIRE(regions && count && (canvas_f32 || canvas_lowp), "fd_canvas_exchange_finalize: null pointer");
  FD_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world && (parity == 0 || parity == 1) && step != 0,
             "fd_canvas_exchange_finalize: bad rank / world / parity / step (step 0 is the cleared state)");
  FD_REQUIRE(n > 0 && n % 4 == 0 && lowp_dtype >= 0 && lowp_dtype <= 2, "fd_canvas_exchange_finalize: n must be a multiple of 4");
  PeerRegions pr;
  for (int r = 0; r < kPeerMaxWorld; ++r) pr.base[r] = r < world ? (float*)regions[r] : nullptr;
  for (int r = 0; r < world; ++r) FD_REQUIRE(pr.base[r], "fd_canvas_exchange_finalize: region of rank %d is null", r);
  const long long n_pad = (n + 63) / 64 * 64;
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (long long)num_sms() * 4;          // every CTA must be resident: each one waits for the peers' flags
  if (blocks > cap) blocks = cap;
  canvas_exchange_finalize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream_>>>(pr, rank, world, parity, step, n, n_pad, count,
                                                                                         canvas_f32, canvas_lowp, lowp_dtype);
  ++g_launches;
  FD_CHECK_CUDA(cudaGetLastError());
  return 0;
}
