// Microbenchmark: how fast can ONE warp issue the value network's k-step (three tcgen05.mma with four different
// operand descriptors + two tcgen05.commit)?  Variants of the issuing code, everything else idle:
//   0: one thread of a diverged warp, descriptors in ordinary registers (R2UR before every use)
//   1: converged warp, one asm block per k-step, elect.sync inside, loop-carried descriptor words
// N = 104 makes the tensor pipe need only 3 x 88 = 264 cycles per k-step, so anything above is issue overhead.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I eb-cadrl_b200/csrc -o tools/bin/mma_issue tools/mma_issue.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "ebc_tc.cuh"

using namespace tc;

constexpr int KMAX = 208, A_IMAGE = 128 * KMAX * 2, SLAB = 2 * KMAX * 32, SLOTS = 8;

// all lanes execute; one elected lane issues the five instructions of a k-step
__device__ __forceinline__ void kstep_elect(uint32_t d, uint32_t a_hi_lo, uint32_t a_lo_lo, uint32_t b_hi_lo, uint32_t b_lo_lo,
                                            uint32_t desc_hi, uint32_t idesc, uint32_t bar1, uint32_t bar2) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t.reg .b64 ah, al, bh, bl;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "mov.b64 ah, {%1, %5};\n\t"
      "mov.b64 al, {%2, %5};\n\t"
      "mov.b64 bh, {%3, %5};\n\t"
      "mov.b64 bl, {%4, %5};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], al, bh, %6, 1;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16.collector::a::fill [%0], ah, bl, %6, 1;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16.collector::a::lastuse [%0], ah, bh, %6, 1;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%7];\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t}"
      ::"r"(d), "r"(a_hi_lo), "r"(a_lo_lo), "r"(b_hi_lo), "r"(b_lo_lo), "r"(desc_hi), "r"(idesc), "r"(bar1), "r"(bar2)
      : "memory");
}

__global__ void __launch_bounds__(128, 1) issue_kernel(int n, int ksteps, int variant, long long *out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  uint8_t *A = smem, *W = smem + 2 * A_IMAGE;
  for (int i = threadIdx.x; i < (2 * A_IMAGE + SLOTS * SLAB) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t idesc = make_idesc_f16(128, n, 0);
  constexpr uint32_t DESC_HI = (128u >> 4) | (1u << 14);
  const uint32_t a0 = ((smem_u32(A) >> 4) & 0x3FFFu) | ((uint32_t)(A_CHUNK_BYTES >> 4) << 16);
  const uint32_t b0 = ((smem_u32(W) >> 4) & 0x3FFFu) | ((uint32_t)n << 16);
  const uint32_t bar1 = smem_u32(&bar[1]), bar2 = smem_u32(&bar[2]);
  if (threadIdx.x < 32) {
    long long best = 1ll << 60;
    uint32_t parity = 0;
    for (int trial = 0; trial < 3; ++trial) {
      const long long t0 = clock64();
      if (variant == 0) {
        if (threadIdx.x == 0) {
          uint32_t s = 0, a = a0;
          for (int k = 0; k < ksteps; ++k) {
            const uint32_t b = b0 + s * (SLAB / 16);
            umma_f16<0>(tmem, a + A_IMAGE / 16, b, DESC_HI, idesc, true);
            umma_f16<1>(tmem, a, b + (uint32_t)n * 2, DESC_HI, idesc, true);
            umma_f16<3>(tmem, a, b, DESC_HI, idesc, true);
            umma_commit(&bar[1]);
            umma_commit(&bar[2]);
            s = s + 1 == SLOTS ? 0 : s + 1;
            a = (k % 13 == 12) ? a0 : a + 2 * A_CHUNK_BYTES / 16;
          }
        }
        __syncwarp();
      } else {
        uint32_t s = 0, a = a0, blk = 0;
        for (int k = 0; k < ksteps; ++k) {
          const uint32_t b = b0 + s * (SLAB / 16);
          kstep_elect(tmem, a, a + A_IMAGE / 16, b, b + (uint32_t)n * 2, DESC_HI, idesc, bar1, bar2);
          s = s + 1 == SLOTS ? 0 : s + 1;
          blk = blk + 1 == 13 ? 0 : blk + 1;
          a = blk == 0 ? a0 : a + 2 * A_CHUNK_BYTES / 16;
        }
      }
      if (threadIdx.x == 0) umma_commit(&bar[0]);
      __syncwarp();
      mbar_wait(&bar[0], parity);
      parity ^= 1u;
      const long long t1 = clock64();
      if (t1 - t0 < best) best = t1 - t0;
    }
    if (threadIdx.x == 0) out[blockIdx.x] = best;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long *d;
  cudaMalloc(&d, 148 * sizeof(long long));
  const int smem = 2 * A_IMAGE + SLOTS * SLAB;
  cudaFuncSetAttribute(issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int ksteps = 260;
  for (int n = 104; n <= 208; n += 104)
    for (int variant = 0; variant < 2; ++variant) {
      issue_kernel<<<148, 128, smem>>>(n, ksteps, variant, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; }
      long long h[148];
      cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
      long long mn = h[0];
      for (int i = 1; i < 148; ++i) if (h[i] < mn) mn = h[i];
      printf("N=%3d variant %d: %.1f cycles per k-step (3 MMAs + 2 commits)\n", n, variant, (double)mn / ksteps);
    }
  return 0;
}
