// Microbenchmark: tensor-pipe time of one tcgen05.mma (kind::f16, M = 128, K = 16, cta_group::1) as a function
// of N, on every SM at once (one CTA per SM, operands resident in shared memory, the same accumulator).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I eb-cadrl_b200/csrc -o /tmp/mma_rate tools/mma_rate.cu
// Output: cycles per MMA for N = 16 .. 256 (decides how the value network's layer widths are padded / split).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "ebc_tc.cuh"

using namespace tc;

__global__ void __launch_bounds__(128, 1) rate_kernel(int n, int reps, int distinct_b, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  // A: 128 x 16 fp16 (2 chunks of 2048 B), B: up to 8 slabs of n x 16 fp16, all zero
  for (int i = threadIdx.x; i < (4096 + 8 * 256 * 32) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_f16(128, n, 0);
    constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);
    const uint32_t a_lo = ((smem_u32(smem) >> 4) & 0x3FFFu) | ((uint32_t)(A_CHUNK_BYTES >> 4) << 16);
    const uint32_t b0 = smem_u32(smem + 4096);
    uint32_t parity = 0;
    long long best = 1ll << 60;
    for (int trial = 0; trial < 5; ++trial) {
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
        const uint32_t b = b0 + (uint32_t)(distinct_b ? (r & 7) : 0) * (uint32_t)n * 32u;
        const uint32_t b_lo = ((b >> 4) & 0x3FFFu) | ((uint32_t)n << 16);     // LBO = n * 16 bytes
        umma_f16(tmem, a_lo, b_lo, DESC_HI, idesc, true);
      }
      umma_commit(&bar);
      mbar_wait(&bar, parity);
      parity ^= 1u;
      const long long t1 = clock64();
      if (t1 - t0 < best) best = t1 - t0;
    }
    out[blockIdx.x] = best;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long *d;
  cudaMalloc(&d, 148 * sizeof(long long));
  const int smem = 4096 + 8 * 256 * 32;
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 512;
  for (int distinct = 0; distinct < 2; ++distinct) {
    printf("distinct B slabs: %d\n", distinct);
    for (int n = 16; n <= 256; n += 8) {
      rate_kernel<<<148, 128, smem>>>(n, reps, distinct, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("N=%d: %s\n", n, cudaGetErrorString(e)); return 1; }
      long long h[148];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      long long mn = h[0], mx = h[0];
      for (int i = 1; i < 148; ++i) { if (h[i] < mn) mn = h[i]; if (h[i] > mx) mx = h[i]; }
      printf("N=%3d  cycles/MMA %.1f (max over SMs %.1f)   columns/cycle %.2f\n", n, (double)mn / reps, (double)mx / reps,
             (double)n * reps / mn);
    }
  }
  return 0;
}
