// Micro-benchmark: tcgen05.mma (M=128, K=16, bf16) cycles per instruction for several N, issued back to back.
#include <cstdio>
#include <cuda_runtime.h>
#include "../pbe_b200/csrc/ptx.cuh"
using namespace pbe;
template <int N, int UNROLL>
__global__ void __launch_bounds__(64, 1) k(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint32_t tptr; __shared__ unsigned long long bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = uniform_u32(tptr);
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_bf16(128, N);
    const uint64_t adesc = umma_desc_sw128(base), bdesc = umma_desc_sw128(base + 32768);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) umma_bf16_ss_elect(tmem + (u & 1) * 256, adesc + 2u * (u & 3), bdesc + 2u * (u & 3), idesc, 1u);
    }
    umma_commit_elect(smem_u32(&bar));
    long long t1 = clock64();
    mbar_wait(smem_u32(&bar), 0);
    long long t2 = clock64();
    if (lane == 0 && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
template <int N> void run() {
  long long* out; cudaMalloc(&out, 16); int iters = 2000; constexpr int U = 4;
  cudaFuncSetAttribute(k<N, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  k<N, U><<<148, 64, 100 * 1024>>>(out, iters); cudaDeviceSynchronize();
  k<N, U><<<148, 64, 100 * 1024>>>(out, iters); 
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
  printf("N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (floor N/2 = %d)  %s\n", N, double(h[0]) / (iters * U), double(h[1]) / (iters * U), N / 2, cudaGetErrorString(e));
  cudaFree(out);
}
int main() { run<16>(); run<48>(); run<64>(); run<128>(); run<160>(); run<256>(); return 0; }
