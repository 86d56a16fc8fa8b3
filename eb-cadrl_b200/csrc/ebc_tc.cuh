// ebc_tc.cuh — sm_100a tensor-core building blocks for K4 (tcgen05 + TMEM + bulk async copies).
//
// Operand layout (both A = activations and B = weights): K-major, no swizzle, the canonical UMMA
// "interleave" layout ((8,n),2):((1,SBO),LBO) in 16-byte units (cute/atom/mma_traits_sm100.hpp):
//     byte_offset(row r, element k) = (k / 8) * (R * 16) + r * 16 + (k % 8) * 2        (bf16, R rows)
// i.e. one 16-byte "k-chunk" of every row is contiguous, rows 16 B apart (a core matrix = 8 rows x 16 B
// = 128 contiguous bytes), SBO = 128 B between 8-row groups, LBO = R*16 B between the two k-chunks that
// one K=16 MMA consumes.  The epilogue writes this layout with conflict-free 16-byte stores (lane = row)
// and the host pre-packs the weights into the same image so that a slab is one cp.async.bulk.
//
// fp32-accurate mode: every fp32 value v is split into three bf16 parts v1 + v2 + v3 (24 mantissa bits);
// a product a*b is accumulated as the six MMAs a1b1 + a1b2 + a2b1 + a2b2 + a1b3 + a3b1 into one fp32 TMEM
// accumulator (dropped terms are O(2^-24)).  Fast mode: one part, one MMA (bf16 operands).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace tc {

constexpr int TILE_M = 128;          // rows per CTA tile = TMEM lanes
constexpr int A_CHUNK_BYTES = TILE_M * 16;   // one 8-element k-chunk of all 128 rows

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking test of a phase (acquire on success)
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded: a protocol bug becomes a trapped launch (cudaErrorLaunchFailure), never a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 22)) __trap();
  }
}
// the same on a 32-bit shared-memory address (no generic pointer to re-form: a generic pointer to shared memory
// costs an S2UR SR_CgaCtaId + LEA wherever the compiler re-materialises it)
__device__ __forceinline__ void mbar_wait_s(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0, ok = 0;
  while (!ok) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (++spins > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void mbar_arrive_s(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ---- async (bulk) copy global -> shared, completion on an mbarrier (UBLKCP) -----------------------
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM -------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *slot_in_smem, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {          // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 16 consecutive fp32 columns of this thread's TMEM lane (lane = 32 * (warp % 4) + laneid)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// split form: issue now, wait later (software-pipelined epilogues)
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- UMMA -----------------------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);          // start address, bits [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;    // leading byte offset, bits [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;    // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
  return d;                                            // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}
// instruction descriptor for kind::f16, bf16 x bf16 -> fp32, A and B K-major (cute::UMMA::InstrDescriptor)
// fmt: 1 = bf16, 0 = fp16 (cute::UMMA::F16F32Format)
__device__ __forceinline__ uint32_t make_idesc_f16(int M, int N, uint32_t fmt) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// One MMA issued by the calling thread, descriptors given as (low word, shared high word): the low words carry the
// start address.  Used by the microbenchmarks in tools/ (the kernels issue whole k-steps through umma_kstep below).
// COLL: use of the A-operand collector buffer -- consecutive MMAs with the same A tile (A_hi x B_lo, then A_hi x B_hi)
// read it from shared memory once: 1 = fill (keep A after this MMA), 2 = use (take A from the buffer, keep it),
// 3 = last use (take it from the buffer, then drop it), 0 = default.  SASS: UTCHMMA gdesc.A_KEEP / .A_REUSE.
template <int COLL = 0>
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                         bool accumulate) {
#define EBC_UMMA(MOD)                                                                                          \
  asm volatile(                                                                                                \
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"                                                          \
      "mov.b64 da, {%1, %3};\n\t"                                                                              \
      "mov.b64 db, {%2, %3};\n\t"                                                                              \
      "setp.ne.b32 p, %5, 0;\n\t"                                                                              \
      "tcgen05.mma.cta_group::1.kind::f16" MOD " [%0], da, db, %4, p;\n\t}"                                   \
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"((uint32_t)accumulate)                 \
      : "memory")
  if constexpr (COLL == 1) EBC_UMMA(".collector::a::fill");
  else if constexpr (COLL == 2) EBC_UMMA(".collector::a::use");
  else if constexpr (COLL == 3) EBC_UMMA(".collector::a::lastuse");
  else EBC_UMMA("");
#undef EBC_UMMA
}
// One k-step of the split product, issued by a CONVERGED warp: every lane executes the block, elect.sync picks the lane
// whose tcgen05 instructions take effect.  Measured (tools/mma_issue.cu): issued this way a k-step of three N = 104 MMAs
// and two commits takes 160 cycles (= the tensor pipe's N/2 per MMA); issued by one thread of a diverged warp, with the
// descriptors moved register -> uniform register before every use, it takes 258 (the "88 cycles per MMA" floor of
// tools/mma_rate.cu is that issue path, not the pipe).
//   d: accumulator (TMEM address); a_lo / b_lo: low descriptor words of the leading parts; a_step / b_step: distance
//   (descriptor units of 16 B) between the parts of A / B; acc: accumulate into d (false: the first MMA overwrites);
//   bar_slot: mbarrier that frees the weight slab; bar_blk: mbarrier that frees the A block (used if commit_blk).
//   The high descriptor word (SBO = 128 B, version 1) is the same for every operand: an immediate inside the block.
constexpr uint32_t DESC_HI_CONST = (128u >> 4) | (1u << 14);
template <int NSPLIT>
__device__ __forceinline__ void umma_kstep(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t a_step, uint32_t b_step,
                                           uint32_t idesc, bool acc, uint32_t bar_slot, uint32_t bar_blk, bool commit_blk) {
#define EBC_MMA(MOD, A, B, P) "@q tcgen05.mma.cta_group::1.kind::f16" MOD " [%0], " A ", " B ", %4, " P ";\n\t"
#define EBC_COMMITS                                                                                       \
  "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%5];\n\t"                    \
  "@pk tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%6];\n\t"
#define EBC_HEAD                                                                                          \
  "{\n\t.reg .pred q, pa, pt, pk;\n\t.reg .b64 a0, a1, a2, b0, b1, b2;\n\t.reg .b32 hi;\n\t"              \
  "elect.sync _|q, 0xffffffff;\n\t"                                                                      \
  "mov.b32 hi, %3;\n\t"                                                                                  \
  "setp.ne.b32 pa, %7, 0;\n\t"                                                                           \
  "setp.eq.b32 pt, 0, 0;\n\t"                                                                            \
  "setp.ne.b32 pk, %8, 0;\n\t"                                                                           \
  "and.pred pk, pk, q;\n\t"
  if constexpr (NSPLIT == 1) {
    asm volatile(EBC_HEAD
                 "mov.b64 a0, {%1, hi};\n\t"
                 "mov.b64 b0, {%2, hi};\n\t"
                 EBC_MMA("", "a0", "b0", "pa")
                 EBC_COMMITS "}"
                 ::"r"(d), "r"(a_lo), "r"(b_lo), "n"(DESC_HI_CONST), "r"(idesc), "r"(bar_slot), "r"(bar_blk),
                   "r"((uint32_t)acc), "r"((uint32_t)commit_blk)
                 : "memory");
  } else if constexpr (NSPLIT == 2) {
    // terms, smallest first: A1 B0, A0 B1, A0 B0 (the two A0 terms share one read of A0 through the collector buffer)
    asm volatile(EBC_HEAD
                 "mov.b64 a0, {%1, hi};\n\t"
                 "mov.b64 b0, {%2, hi};\n\t"
                 "mov.b64 a1, {%9, hi};\n\t"
                 "mov.b64 b1, {%10, hi};\n\t"
                 EBC_MMA("", "a1", "b0", "pa")
                 EBC_MMA(".collector::a::fill", "a0", "b1", "pt")
                 EBC_MMA(".collector::a::lastuse", "a0", "b0", "pt")
                 EBC_COMMITS "}"
                 ::"r"(d), "r"(a_lo), "r"(b_lo), "n"(DESC_HI_CONST), "r"(idesc), "r"(bar_slot), "r"(bar_blk),
                   "r"((uint32_t)acc), "r"((uint32_t)commit_blk), "r"(a_lo + a_step), "r"(b_lo + b_step)
                 : "memory");
  } else {
    // A2 B0, A1 B1, A1 B0, A0 B2, A0 B1, A0 B0
    asm volatile(EBC_HEAD
                 "mov.b64 a0, {%1, hi};\n\t"
                 "mov.b64 b0, {%2, hi};\n\t"
                 "mov.b64 a1, {%9, hi};\n\t"
                 "mov.b64 b1, {%10, hi};\n\t"
                 "mov.b64 a2, {%11, hi};\n\t"
                 "mov.b64 b2, {%12, hi};\n\t"
                 EBC_MMA("", "a2", "b0", "pa")
                 EBC_MMA(".collector::a::fill", "a1", "b1", "pt")
                 EBC_MMA(".collector::a::lastuse", "a1", "b0", "pt")
                 EBC_MMA(".collector::a::fill", "a0", "b2", "pt")
                 EBC_MMA(".collector::a::use", "a0", "b1", "pt")
                 EBC_MMA(".collector::a::lastuse", "a0", "b0", "pt")
                 EBC_COMMITS "}"
                 ::"r"(d), "r"(a_lo), "r"(b_lo), "n"(DESC_HI_CONST), "r"(idesc), "r"(bar_slot), "r"(bar_blk),
                   "r"((uint32_t)acc), "r"((uint32_t)commit_blk), "r"(a_lo + a_step), "r"(b_lo + b_step),
                   "r"(a_lo + 2 * a_step), "r"(b_lo + 2 * b_step)
                 : "memory");
  }
#undef EBC_MMA
#undef EBC_COMMITS
#undef EBC_HEAD
}

// all previously issued MMAs of this thread done -> one arrival on `bar`
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// the same for a converged warp: every lane executes it, the lane elect.sync picks commits.  The accumulator commits of
// the MMA warp use this form so that they are executed by the very lane whose MMAs they track (umma_kstep elects the
// same way) -- tcgen05.commit only covers operations initiated by the executing thread.
__device__ __forceinline__ void umma_commit_elect(uint64_t *bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar)) : "memory");
}

// one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---- fp32 -> bf16 parts --------------------------------------------------------------------------
// Operand element format of a splitting mode: NSPLIT 1 and 3 -> bf16 parts, NSPLIT 2 -> fp16 parts
// (two fp16 parts carry 22 mantissa bits; three bf16 parts carry 24).
template <int NSPLIT> struct Fmt {
  static constexpr uint32_t IDESC = 1;   // bf16
  __device__ static __forceinline__ uint32_t pack2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  }
  __device__ static __forceinline__ float lo(uint32_t pk) { return __uint_as_float(pk << 16); }
  __device__ static __forceinline__ float hi(uint32_t pk) { return __uint_as_float(pk & 0xFFFF0000u); }
  __device__ static __forceinline__ float from16(uint16_t h) { return __uint_as_float((uint32_t)h << 16); }
};
template <> struct Fmt<2> {
  static constexpr uint32_t IDESC = 0;   // fp16
  __device__ static __forceinline__ uint32_t pack2(float lo, float hi) {
    const __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<const uint32_t *>(&h);
  }
  __device__ static __forceinline__ float lo(uint32_t pk) { return __half2float(__ushort_as_half((unsigned short)(pk & 0xFFFFu))); }
  __device__ static __forceinline__ float hi(uint32_t pk) { return __half2float(__ushort_as_half((unsigned short)(pk >> 16))); }
  __device__ static __forceinline__ float from16(uint16_t h) { return __half2float(__ushort_as_half(h)); }
};

// Store 8 consecutive k-elements (k0 % 8 == 0) of row `row` into the NSPLIT A images.
// a_base: 32-bit shared-memory address, image s at a_base + s * image_bytes.  Parts are peeled two values at a time:
// p = round16(v), v -= float(p) (exact), repeat.
template <int NSPLIT>
__device__ __forceinline__ void store_a8(uint32_t a_base, uint32_t image_bytes, int row, int k0, const float (&v)[8]) {
  float r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = v[i];
#pragma unroll
  for (int s = 0; s < NSPLIT; ++s) {
    uint4 w;
    uint32_t *pw = reinterpret_cast<uint32_t *>(&w);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t pk = Fmt<NSPLIT>::pack2(r[2 * i], r[2 * i + 1]);
      pw[i] = pk;
      if (s + 1 < NSPLIT) {
        r[2 * i] -= Fmt<NSPLIT>::lo(pk);
        r[2 * i + 1] -= Fmt<NSPLIT>::hi(pk);
      }
    }
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_base + (uint32_t)s * image_bytes + (uint32_t)(k0 >> 3) * A_CHUNK_BYTES +
                                                                     (uint32_t)row * 16u),
                 "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w)
                 : "memory");
  }
}

// (a part, b part) pairs accumulated per k-step, smallest magnitude first
}  // namespace tc
