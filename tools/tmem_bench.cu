// Micro-benchmark: tcgen05.ld (32x32b.x32) throughput per SM as a function of the number of warps issuing loads.
#include <cstdio>
#include <cuda_runtime.h>
#include "../pbe_b200/csrc/ptx.cuh"
using namespace pbe;
__global__ void __launch_bounds__(512, 1) k(long long* out, float* sink, int iters, int nwarps) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tptr;
  float acc = 0.f;
  long long t0 = clock64();
  if (warp < nwarps) {
    const uint32_t addr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 128;
    for (int i = 0; i < iters; ++i) {
      uint32_t v[32];
      tmem_ld_x32(addr + (i & 3) * 32, v);
      tmem_ld_wait();
      acc += __uint_as_float(v[0]) + __uint_as_float(v[31]);
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
  if (acc == 12345.678f) sink[0] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
int main() {
  long long* out; float* sink; cudaMalloc(&out, 16); cudaMalloc(&sink, 16);
  for (int nw : {1, 2, 4, 8, 16}) {
    int iters = 4000;
    k<<<148, 512>>>(out, sink, iters, nw); cudaDeviceSynchronize();
    k<<<148, 512>>>(out, sink, iters, nw); cudaError_t e = cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    double bytes = double(nw) * iters * 32 * 32 * 4;
    printf("warps %2d: %.1f cycles per x32 load per warp, %.1f B/clk/SM  %s\n", nw, double(h) / iters, bytes / h, cudaGetErrorString(e));
  }
  return 0;
}
