// Microbenchmark (B200): sustained MUFU.EX2 throughput per SM (fp32 ex2.approx.ftz), alone and next to the other instructions of
// the attention kernels' exponential loop (one F2FP bf16x2 pack per pair of exponentials; one FFMA2 + FADD2 per pair).
// Build + run: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_rate scripts/micro/mufu_rate.cu && /tmp/mufu_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int kMode>
__global__ void k(float* out, int iters, long long* cycles) {
  float x[8];
  uint32_t acc = 0;
  unsigned long long s2 = 0ull;
  for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    float y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y[i]) : "f"(x[i]));
    if (kMode >= 1) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        uint32_t h;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(y[i + 1]), "f"(y[i]));
        acc ^= h;
      }
    }
    if (kMode >= 2) {
#pragma unroll
      for (int i = 0; i < 8; i += 2) {
        unsigned long long p, q;
        asm volatile("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(y[i]), "f"(y[i + 1]));
        asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(s2) : "l"(s2), "l"(p));
        asm volatile("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(q) : "l"(p), "l"(s2));
        s2 ^= q & 1ull;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = y[i] - 1.5f;
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)acc + (float)(s2 & 0xff);
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int kMode>
void run(const char* name, int sms, float* out, long long* cyc) {
  const int iters = 20000;
  for (int threads : {128, 256, 512}) {
    k<kMode><<<sms, threads>>>(out, iters, cyc);
    k<kMode><<<sms, threads>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += (double)h[i];
    avg /= sms;
    printf("%-44s %2d warps/SMSP: %5.2f ex2 per clock per SM  (%.1f clocks per MUFU warp instruction per SMSP)\n", name, threads / 128,
           (double)threads * 8 * iters / avg, avg / ((double)threads / 128 * 8 * iters));
  }
}
int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * 1024);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  run<0>("ex2 only", sms, out, cyc);
  run<1>("ex2 + one F2FP.BF16 pack per pair", sms, out, cyc);
  run<2>("ex2 + F2FP pack + FADD2 + FFMA2 per pair", sms, out, cyc);
  return 0;
}
