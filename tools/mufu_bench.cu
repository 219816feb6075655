// Micro-benchmark: MUFU.EX2 issue rate on this GPU, alone and inside the softmax instruction mix.
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float v[16];
  for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 1e-3f + i;
  float s0 = 0, s1 = 0; unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) v[i] = ex2(v[i]);
      else {
        float e = ex2(fmaf(v[i], a, b));
        if (i & 1) s1 += e; else s0 += e;
        v[i] = e;
      }
    }
    if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 16; i += 2) { __nv_bfloat162 h = __floats2bfloat162_rn(v[i], v[i + 1]); acc ^= *reinterpret_cast<unsigned*>(&h); }
    }
  }
  float r = s0 + s1; for (int i = 0; i < 16; ++i) r += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r + acc;
}
template <int MODE> void run(const char* name, int warps_per_sm) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int threads = 32 * warps_per_sm; float* out; cudaMalloc(&out, sms * threads * 4);
  int iters = 20000; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms, threads>>>(out, 100, 0.5f, -1.f); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<sms, threads>>>(out, iters, 0.5f, -1.f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double exps = double(sms) * threads * iters * 16; int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-28s warps/SM %2d: %.2f exp/ns/SM  (%.2f exp/clk/SM at max clock %d MHz)\n", name, warps_per_sm, exps / (ms * 1e6) / sms, exps / (ms * 1e6) / sms / (clk * 1e-6), clk / 1000);
  cudaFree(out);
}
int main() {
  for (int w : {4, 8, 16, 32}) run<0>("ex2 only (dependent chains)", w);
  for (int w : {4, 8, 16, 32}) run<1>("fma+ex2+add", w);
  for (int w : {4, 8, 16, 32}) run<2>("fma+ex2+add+cvt.bf16x2", w);
  return 0;
}
