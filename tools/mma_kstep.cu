// Microbenchmark: the value network's k-step as the MMA warp issues it -- three tcgen05.mma (A_lo x B_hi, A_hi x B_lo,
// A_hi x B_hi; M = 128, N = 208, K = 16, fp16) on operands laid out like the kernel's (two A images of 53 KB, an 8-slot
// ring of 13.3 KB weight slabs), followed by two tcgen05.commit -- with nothing else running on the SM.
// Variants: commits on/off, A-operand collector reuse on/off, and a second warp group that stores to shared memory
// at the crew's rate (16-byte stores, 8 KB per k-step) to show the contention.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I eb-cadrl_b200/csrc -o tools/bin/mma_kstep tools/mma_kstep.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "ebc_tc.cuh"

using namespace tc;

constexpr int KMAX = 208, A_IMAGE = 128 * KMAX * 2, SLAB = 2 * KMAX * 32, SLOTS = 8;

__global__ void __launch_bounds__(640, 1) kstep_kernel(int n, int ksteps, int commits, int coll, int crew_stores, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  __shared__ volatile int stop;
  uint8_t *A = smem, *W = smem + 2 * A_IMAGE;
  for (int i = threadIdx.x; i < (2 * A_IMAGE + SLOTS * SLAB) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); stop = 0; }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_f16(128, n, 0);
    constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);
    const uint32_t a0 = ((smem_u32(A) >> 4) & 0x3FFFu) | ((uint32_t)(A_CHUNK_BYTES >> 4) << 16);
    const uint32_t b0 = ((smem_u32(W) >> 4) & 0x3FFFu) | ((uint32_t)n << 16);
    uint32_t parity = 0;
    long long best = 1ll << 60;
    for (int trial = 0; trial < 3; ++trial) {
      const long long t0 = clock64();
      for (int k = 0; k < ksteps; k += 8) {          // eight k-steps per trip, every address a compile-time offset
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const uint32_t a_lo = a0 + (uint32_t)h * (2 * A_CHUNK_BYTES / 16);
          const uint32_t b_lo = b0 + (uint32_t)h * (SLAB / 16);
          umma_f16<0>(tmem, a_lo + A_IMAGE / 16, b_lo, DESC_HI, idesc, true);
          if (coll) {
            umma_f16<1>(tmem, a_lo, b_lo + (uint32_t)n * 2, DESC_HI, idesc, true);
            umma_f16<3>(tmem, a_lo, b_lo, DESC_HI, idesc, true);
          } else {
            umma_f16<0>(tmem, a_lo, b_lo + (uint32_t)n * 2, DESC_HI, idesc, true);
            umma_f16<0>(tmem, a_lo, b_lo, DESC_HI, idesc, true);
          }
          // modes: 0 none | 1 one per two k-steps | 2 one per k-step | 3 three per two k-steps | 4 two per k-step
          if (commits >= 2 || (commits == 1 && (h & 1))) umma_commit(&bar[1]);
          if (commits == 4 || (commits == 3 && (h & 1))) umma_commit(&bar[2]);
        }
      }
      umma_commit(&bar[0]);
      mbar_wait(&bar[0], parity);
      parity ^= 1u;
      const long long t1 = clock64();
      if (t1 - t0 < best) best = t1 - t0;
    }
    out[blockIdx.x] = best;
    stop = 1;
  } else if (threadIdx.x >= 128 && crew_stores) {
    // 512 threads storing 16 bytes each = 8 KB per round, paced to one round per ~400 cycles like the crew
    const int t = threadIdx.x - 128;
    uint4 *dst = reinterpret_cast<uint4 *>(A) + t;
    uint32_t x = t;
    while (!stop) {
      for (int i = 0; i < crew_stores; ++i) dst[(i & 7) * 512] = make_uint4(x, x, x, x);
      ++x;
      fence_proxy_async();
      __nanosleep(100);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long *d;
  cudaMalloc(&d, 148 * sizeof(long long));
  const int smem = 2 * A_IMAGE + SLOTS * SLAB;
  cudaFuncSetAttribute(kstep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int ksteps = 256;
  for (int n = 104; n <= 208; n += 104)
    for (int commits = 0; commits <= 4; ++commits)
      for (int coll = 1; coll < 2; ++coll)
        for (int crew = 0; crew <= 4; crew += 2) {
          kstep_kernel<<<148, 640, smem>>>(n, ksteps, commits, coll, crew, d);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
          long long h[148];
          cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
          long long mn = h[0];
          for (int i = 1; i < 148; ++i) if (h[i] < mn) mn = h[i];
          printf("N=%3d commit mode=%d collector=%d crew-store rounds=%d: %.1f cycles per k-step (3 MMAs)\n", n, commits, coll, crew,
                 (double)mn / ksteps);
        }
  return 0;
}
