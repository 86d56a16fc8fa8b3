// ebc_value_tc.cu — K4 on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// One CTA owns a tile of 128 entity rows (whole states) and walks the whole per-entity part of
// rl/policy/sarl.py:38-82 with the activations never leaving the SM:
//   X -> mlp1.0 -> mlp1.2 (= H1) -> { mlp2.0 -> mlp2.2 (= H2),  attention.0 on [H1, mean_state(H1)]
//   -> attention.2 -> attention.4 score } -> masked softmax -> pooled H2 -> joint[state]
// and a second kernel runs mlp3 over tiles of 128 states.
//
// Warp roles (640 threads):
//   warps 0-15 (512 threads)   epilogue crew: warp w reads TMEM lane quarter (w % 4) (row = lane) and the
//                              16-column blocks b with b % 4 == w / 4, applies bias / ReLU, splits the fp32
//                              value into 16-bit parts and stores the next stage's A operand
//   warp 16 (converged)        issues the tcgen05.mma's, one converged block per k-step (elect.sync inside):
//                              A = activations in shared memory (canonical K-major layout written by the
//                              previous epilogue), B = weight slab, D = fp32 accumulator in TMEM.  Its loop
//                              (mma_warp_main) is written so that ptxas keeps it on the UNIFORM datapath
//                              (~45 uniform instructions per k-step, no R2UR): see the comment there
//   warp 17 (converged)        scout: polls the hand-over barriers and publishes how many k-steps may issue
//   warps 18-19, one lane each stream the pre-packed weight slabs L2 -> shared memory with cp.async.bulk
//                              (mbarrier full/empty ring), every other slab each
// Every stage boundary is a chase in both directions, one 16-column block (= one k-step) at a time:
//   kbar[b]   crew -> MMA warp: block b of the next A operand is written (the next stage's k-step b may issue)
//   afree[b]  MMA warp -> crew: the MMAs that read block b of the current A operand have completed
//             (tcgen05.commit after that k-step), so the following epilogue may overwrite it
// so the tensor pipe and the epilogue crew work on the same tile at the same time.  The global state of
// sarl.py:51-60 is an extra K chunk of attention.0: the crew writes the per-state mean of H1 (replicated
// over the state's rows) as an A operand and the tensor core multiplies it with the global half of the weights.
// The input X is either read from the materialised value-network input (vin) or, with vin == NULL, computed by the
// crew itself from the bound state and the 48-byte robot record K3 leaves per (episode, action) (fused input path).
// See ebc_tc.cuh for the operand layout and the fp32-accurate operand splitting.
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "ebc_internal.cuh"
#include "ebc_tc.cuh"

namespace {

using namespace tc;

constexpr int NCREW = 512;                      // epilogue threads: 4 lane quarters x 4 column groups
constexpr int NLOAD = 2;                        // weight-loader threads (one per warp), slabs dealt round-robin
constexpr int NT = NCREW + 64 + 32 * NLOAD;     // + the MMA warp (16), the scout warp (17), the loader warps (18, 19)
constexpr int MAX_SLABS = 120;                  // slabs per tile (entity program: 88, mlp3: 46)
constexpr int NCG = NCREW / TILE_M;             // column groups
constexpr int KMAX = 208;                       // widest A operand chunk kept in shared memory
constexpr int NKB = KMAX / 16;                  // 16-column blocks of an A operand
constexpr int TMEM_COLS = 512;
constexpr int MAX_STATES = 32;                  // states per entity tile (rows per state >= 4)

template <int NSPLIT> struct Cfg {
  // ring depth: a slot is busy from the issue of its bulk copy (L2 latency ~ 1000 cycles) until the MMAs that
  // read it have completed and the loader has seen that; at one k-step per ~312 cycles this needs >= 7 slots
  static constexpr int STAGES = NSPLIT == 1 ? 16 : (NSPLIT == 2 ? 8 : 3);
  // bytes of the per-state mean G[state][h1 padded]: 16 states of the widest layer (8 in the bf16x3 mode, whose
  // operand images leave less room); narrower layers fit more states per tile
  static constexpr uint32_t G_BYTES = (NSPLIT == 3 ? 8 : 16) * KMAX * 4;
  static constexpr uint32_t A_IMAGE = TILE_M * KMAX * 2;                 // bytes per split image of A
  static constexpr uint32_t A_BYTES = NSPLIT * A_IMAGE;
  static constexpr uint32_t STAGE_BYTES = NSPLIT * KMAX * 32;            // one k-step slab, all splits
  static constexpr uint32_t W_BYTES = STAGES * STAGE_BYTES;
};

struct Smem {   // offsets (bytes) into dynamic shared memory
  uint32_t a, w, g, sc, xs, tab, bars, misc, total;
};

template <int NSPLIT>
__host__ __device__ inline Smem smem_layout() {
  Smem s;
  uint32_t off = 0;
  s.a = off; off += Cfg<NSPLIT>::A_BYTES;
  s.w = off; off += Cfg<NSPLIT>::W_BYTES;
  s.g = off; off += Cfg<NSPLIT>::G_BYTES;              // G[state][k]: per-state mean of H1
  s.sc = off; off += TILE_M * 4 * NCG;                 // partial scores per column group
  s.xs = off; off += 2 * MAX_STATES * 8 * 4;            // self-state part of each state's first row, two tiles
  s.tab = off; off += MAX_SLABS * 8;                    // (offset, bytes) of every slab of the per-tile program
  s.bars = off; off += 512;                             // mbarriers + the scout's counter
  s.misc = off; off += 16 + 2 * MAX_STATES * 4;         // TMEM base address, row counts of two tiles
  s.total = off;
  return s;
}

// The pooled per-state features ("joint", the input of mlp3) travel between the two kernels in the layout the mlp3
// kernel stages them in: [tile of 128 states][k-chunk of 8 columns][state in tile][8 floats], so that its thread
// (row, chunk) reads 32 contiguous bytes and a warp 1 KB -- a row-major [state][jd] scratch costs one 32-byte sector
// per lane and load instruction there, which starves the shared-memory pipe the MMAs read their operands through.
__host__ __device__ __forceinline__ size_t joint_index(long long state, int k, int n_chunks) {
  return ((((size_t)(state >> 7)) * n_chunks + (size_t)(k >> 3)) * TILE_M + (size_t)(state & 127)) * 8 + (size_t)(k & 7);
}
static_assert(TILE_M == 128, "joint_index assumes 128 states per mlp3 tile");

__device__ __forceinline__ void crew_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NCREW) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t low_bits(int n) { return (1u << n) - 1u; }
// max(x, 0) in one instruction that propagates NaN (fmaxf would return 0)
__device__ __forceinline__ float relu_nan(float x) {
  float y;
  asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(y) : "f"(x));
  return y;
}

// Ring of weight slabs + the MMA warp's cursor; barriers shared with the crew.
template <int NSPLIT>
struct Pipe {
  uint64_t *full, *empty, *acc_bar, *a_bar, *kbar, *afree;
  uint32_t acc_s, abar_s, kbar_s, afree_s;   // the crew's barriers as 32-bit shared-memory addresses
  uint8_t *wbuf;
  const uint8_t *wpack;      // packed weights (global)
  int n_stage_slabs;         // slabs per tile
  const uint2 *tab;          // shared memory: (byte offset, bytes) of each slab of the per-tile sequence
  long long total;           // slab count over all tiles of this CTA (loader)
  uint32_t *ready;           // scout -> MMA warp: k-steps (counted over the whole launch) whose operands are in place
  uint32_t issued, seen;     // MMA warp: k-steps issued so far / the scout's count when last read
  uint32_t st;               // MMA warp: ring slot of the next slab
  bool leader;               // MMA warp: the lane that issues tcgen05.mma / tcgen05.commit
  uint32_t acc_phase, a_phase;
  uint32_t f_phase;          // crew: expected parity of every afree barrier, one bit each
  long long *trace;          // optional clock64() trace of CTA 0's crew thread 0 (EBC_TC_TRACE=1)
  int trace_pos;
  bool fine;                 // EBC_TC_TRACE=3: the MMA warp stamps EVERY k-step it issues ((clock << 16) | schedule entry << 8 | k-step)
  __device__ __forceinline__ void stamp() {
    if (trace && threadIdx.x == 0 && blockIdx.x == 0 && trace_pos < 2048) trace[trace_pos++] = clock64();
  }

  // ---- loader threads: stream every slab of every tile of this CTA through the ring.  One thread needs ~390
  //      cycles per slab (an mbarrier.try_wait costs ~170 cycles even when the slot is free, arrive.expect_tx +
  //      cp.async.bulk ~190), which is more than the 312 cycles the tensor pipe spends on a k-step: NLOAD threads
  //      in different warps take every NLOAD-th slab each --------------------------------------------------------
  __device__ void loader_loop(int which) {
    constexpr uint32_t ST = Cfg<NSPLIT>::STAGES;
    uint32_t slot = (uint32_t)which % ST, parity = 1u ^ (((uint32_t)which / ST) & 1u);   // "slot is free": passes on lap 0
    uint32_t idx = (uint32_t)which % (uint32_t)n_stage_slabs;
    for (long long i = which; i < total; i += NLOAD) {
      mbar_wait(&empty[slot], parity);
      const uint2 sl = tab[idx];
      mbar_arrive_expect_tx(&full[slot], sl.y);
      bulk_g2s(wbuf + (size_t)slot * Cfg<NSPLIT>::STAGE_BYTES, wpack + sl.x, sl.y, &full[slot]);
      slot += NLOAD;
      while (slot >= ST) { slot -= ST; parity ^= 1u; }
      idx += NLOAD;
      while (idx >= (uint32_t)n_stage_slabs) idx -= (uint32_t)n_stage_slabs;
    }
  }
  // ---- scout thread: follows the per-tile schedule one k-step at a time, waits (blocking, hardware-suspended
  //      mbarrier.try_wait) until that k-step's A block and weight slab are there, and publishes the running
  //      count of issuable k-steps.  The MMA warp then needs one shared-memory load per batch instead of one
  //      mbarrier round trip per k-step: the tensor pipe's queue is shallow, so every cycle the issuing warp
  //      spends polling is a cycle the pipe idles (measured: ~450 polling cycles per 1250 cycles of MMA work).
  __device__ void scout_loop(const TcProgram &P, long long my_tiles) {   // whole warp, converged
    constexpr uint32_t ST = Cfg<NSPLIT>::STAGES;
    constexpr uint32_t LOOK = ST < 8 ? ST : 8;        // k-steps examined per poll
    const uint32_t lane = threadIdx.x & 31;
    uint32_t count = 0, slot = 0, par = 0, kph = 0, aph = 0;
    for (long long t = 0; t < my_tiles; ++t) {
      mbar_wait(a_bar, aph);              // the tile's input operand is staged
      aph ^= 1u;
      for (int i = 0; i < P.n_sched; ++i) {
        const int ksteps = P.st[P.sched[i].stage].ksteps;
        const bool chase = P.sched[i].chase != 0;
        int ks = 0;
        uint32_t idle = 0;
        while (ks < ksteps) {
          // lanes 0-7 test kbar[ks ..], lanes 8-15 the next ring slots: one non-blocking test for all of them
          bool ok = false;
          if (lane < 8) {
            const int k = ks + (int)lane;
            ok = !chase || (k < ksteps && mbar_test(&kbar[k], (kph >> k) & 1u));
          } else if (lane < 8 + LOOK) {
            uint32_t sl = slot + lane - 8, pp = par;
            if (sl >= ST) { sl -= ST; pp ^= 1u; }
            ok = mbar_test(&full[sl], pp);
          }
          const uint32_t m = __ballot_sync(0xffffffffu, ok);
          const uint32_t both = m & (m >> 8) & 0xFFu;
          int nready = __ffs((int)~both) - 1;            // leading k-steps with both barriers complete
          if (nready > ksteps - ks) nready = ksteps - ks;
          if (nready == 0) {
            // sleep (hardware-suspended) on the first barrier that is missing, then look again
            if (chase && !(m & 1u)) mbar_try_wait(&kbar[ks], (kph >> ks) & 1u);
            else mbar_try_wait(&full[slot], par);
            if (++idle > (1u << 22)) __trap();
            continue;
          }
          idle = 0;
          if (chase) kph ^= low_bits(nready) << ks;
          ks += nready;
          slot += (uint32_t)nready;
          if (slot >= ST) { slot -= ST; par ^= 1u; }
          count += (uint32_t)nready;
          __syncwarp();                   // the observations of lanes 0-15 are ordered before the release below
          if (lane == 0) asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(smem_u32(ready)), "r"(count) : "memory");
        }
      }
    }
  }
  // ---- MMA warp: every lane runs the loops (warp-uniform values -> uniform registers, no per-instruction
  //      election loops), one elected lane issues the tcgen05 instructions ---------------------------------
  __device__ __forceinline__ uint32_t ready_count() {
    uint32_t r;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(r) : "r"(smem_u32(ready)) : "memory");
    return r;
  }
  // chase:    the crew is still writing the A operand; k-step ks is issued as soon as its two k-chunks (one
  //           16-column block of the previous stage's epilogue) have been stored and fenced by all 128 rows
  // commit_k: signal afree[ks] when the MMAs issued up to and including k-step ks have completed
  __device__ void mma_stage(const TcStage &S, uint32_t a_smem, uint32_t tmem_base, bool commit_k, int tag) {
    constexpr uint32_t ST = Cfg<NSPLIT>::STAGES;
    const uint32_t idesc = make_idesc_f16(TILE_M, S.np, Fmt<NSPLIT>::IDESC);
    const uint32_t d = tmem_base + S.acc_col;
    const uint32_t np = (uint32_t)S.np;
    const int ksteps = S.ksteps;
    // A descriptor = {start address >> 4 | LBO << 16, SBO | version}: only the low word changes (32-bit adds)
    static_assert(DESC_HI_CONST == ((128u >> 4) | (1u << 14)), "SBO = 128 B, descriptor version 1");
    // Everything the issue path needs is a running value that depends on kernel parameters and loop counters only
    // (never on what the scout published): the compiler keeps such values in uniform registers, and a k-step is a
    // dozen instructions next to its three MMAs and two commits.  That matters: this warp shares its scheduler with
    // four crew warps, and every instruction it spends per k-step is time the tensor pipe's shallow queue can run dry.
    uint32_t a_cur = ((a_smem >> 4) & 0x3FFFu) | ((uint32_t)(A_CHUNK_BYTES >> 4) << 16);
    const uint32_t b_base = ((smem_u32(wbuf) >> 4) & 0x3FFFu) | (np << 16);
    uint32_t slot = st;
    const uint32_t empty0 = smem_u32(empty), afree0 = smem_u32(afree);
    const uint32_t base = issued;                 // schedule index of this stage's first k-step
    bool acc = S.accumulate != 0;
    for (int ks = 0; ks < ksteps; ++ks) {
      if ((int)(seen - base) <= ks) {
        uint32_t idle = 0;
        while ((int)((seen = ready_count()) - base) <= ks) {
          // the scout sleeps on the barriers; this warp only watches its count (a shared-memory load per trip)
          if (++idle > (1u << 26)) __trap();   // bounded: a protocol bug traps instead of hanging
        }
        tc_fence_after();
        if (trace && !fine && leader && blockIdx.x == 0 && trace_pos < 900)   // diagnostics: (clock, cleared k-steps)
          trace[2048 + trace_pos++] = (clock64() << 16) | (long long)((seen - base - (uint32_t)ks) & 0xFFFFu);
      }
      const uint32_t b_cur = b_base + slot * (Cfg<NSPLIT>::STAGE_BYTES / 16);
      __syncwarp();      // converged: every lane runs the issue block, one elected lane's instructions take effect
      umma_kstep<NSPLIT>(d, a_cur, b_cur, Cfg<NSPLIT>::A_IMAGE / 16, np * 2, idesc, acc, empty0 + slot * 8,
                         afree0 + (uint32_t)ks * 8, commit_k);
      if (fine && leader && blockIdx.x == 0 && trace_pos < 2040)
        trace[2048 + trace_pos++] = (clock64() << 16) | (long long)((tag << 8) | ks);
      acc = true;
      a_cur += 2 * A_CHUNK_BYTES / 16;
      slot = slot + 1 == ST ? 0 : slot + 1;
    }
    st = slot;
    issued = base + (uint32_t)ksteps;
  }
  // one tile of the program: every stage of the schedule, accumulator commits where the crew waits for them
  __device__ void mma_tile(const TcProgram &P, uint32_t a_smem, uint32_t tmem_base) {
    mbar_wait(a_bar, a_phase);          // sleep until the tile's input is staged (the scout clears it right after)
    a_phase ^= 1u;
    for (int i = 0; i < P.n_sched; ++i) {
      mma_stage(P.st[P.sched[i].stage], a_smem, tmem_base, P.sched[i].commit_k != 0, i);
      if (P.sched[i].commit_acc) {
        __syncwarp();
        umma_commit_elect(acc_bar);
      }
    }
  }

  // ---- crew -------------------------------------------------------------------------------------------
  __device__ void wait_acc() {
    mbar_wait_s(acc_s, acc_phase);     // (a busy test_wait spin instead of the suspended try_wait: 0.5 % slower, measured)
    acc_phase ^= 1u;
    tc_fence_after();
    stamp();
  }
  __device__ void signal_a() {
    stamp();         // after this thread's A-operand stores / TMEM reads
    fence_proxy_async();               // generic-proxy smem writes -> async proxy (UMMA)
    tc_fence_before();                 // this thread's tcgen05.ld's are ordered before the arrive
    mbar_arrive_s(abar_s);
  }
  // block b of the A operand may be overwritten: the MMAs of the last commit_k stage that read it are done
  __device__ __forceinline__ void wait_free(int b) { mbar_wait_s(afree_s + 8u * (uint32_t)b, (f_phase >> b) & 1u); }
  // hand block b of the A operand (this thread's row) to the MMA warp
  __device__ __forceinline__ void publish(int b) {
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive_s(kbar_s + 8u * (uint32_t)b);
  }
};

// Column sums over the 32 lanes of a warp: every lane passes 16 values (columns 0..15 of its row); lane l returns the
// total of column (l >> 1) (both lanes of a pair hold it).  Transpose-reduce: 8 + 4 + 2 + 1 + 1 shuffles instead of
// 16 x 5, additions in a fixed order.
__device__ __forceinline__ float warp_colsum16(const float (&v)[16], int lane) {
  float w8[8], w4[4], w2[2];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = (lane & 16) ? v[i] : v[i + 8], keep = (lane & 16) ? v[i + 8] : v[i];
    w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = (lane & 8) ? w8[i] : w8[i + 4], keep = (lane & 8) ? w8[i + 4] : w8[i];
    w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = (lane & 4) ? w4[i] : w4[i + 2], keep = (lane & 4) ? w4[i + 2] : w4[i];
    w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  const float send = (lane & 2) ? w2[0] : w2[1], keep = (lane & 2) ? w2[1] : w2[0];
  const float x = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  return x + __shfl_xor_sync(0xffffffffu, x, 1);
}

// accumulator columns [col0, col0 + ncols) of this thread's row, this thread's column group
// -> relu(x + bias), columns >= n_real forced to 0 -> block (c / 16) of the A operand, published to the MMA warp.
// n_free: blocks below this index wait for afree (the previous commit_k stage read them).
// wait_first: the accumulator itself is only known to be complete once afree of this thread's first block has
// fired (the releasing stage was issued after the stage that produced the accumulator): wait before loading.
// With GSUM (n divides 32, so a state's rows are an aligned lane group of this warp): also reduce the
// activated fp32 values over the state's real rows with shuffles and store the mean to g_out[col * g_ld + state]
// -- the global state of sarl.py:51-60, taken from registers instead of re-reading the 16-bit images.
// g_out is [state][g_ld]: the 16 lanes of a state store 16 consecutive floats (conflict-free).
struct EpiExtra {
  int one_col = -1;          // output column forced to 1.0 (carries the next layer's bias through its GEMM), or -1
  bool wait_first = false;   // see epi_to_a
  int n = 1, row_cnt = 0;    // GSUM: rows per state, real rows of this row's state
  float *g_out = nullptr;    // GSUM: G[state][g_ld]
  int g_ld = 0, g_states = 0, g_sid = 0, g_rin = 0, g_seg = 0, g_part = 0;
};

template <int NSPLIT, bool GSUM>
__device__ __forceinline__ void epi_to_a(Pipe<NSPLIT> &pipe, uint32_t tmem_row, int cg, int col0, int ncols, int n_real,
                                         const float *__restrict__ bias, uint32_t a_base, int row, bool chase,
                                         int n_free, const EpiExtra &x = EpiExtra()) {
  const int n = x.n, row_cnt = x.row_cnt, g_ld = x.g_ld, g_states = x.g_states, g_sid = x.g_sid, g_rin = x.g_rin;
  const int g_seg = x.g_seg, g_part = x.g_part, one_col = x.one_col;
  const bool wait_first = x.wait_first;
  float *g_out = x.g_out;
  if (wait_first && 16 * cg < ncols) {
    pipe.wait_free(cg < n_free ? cg : 0);
    tc_fence_after();
  }
  for (int c = 16 * cg; c < ncols; c += 16 * NCG) {
    float v[16];
    tmem_ld16(tmem_row + col0 + c, v);
    if (bias) {              // the bias is added here unless it rode in the GEMM (TcStage::bias_k)
      const float4 *b4 = reinterpret_cast<const float4 *>(bias + c);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 bb = __ldg(b4 + q);
        v[4 * q] += bb.x; v[4 * q + 1] += bb.y; v[4 * q + 2] += bb.z; v[4 * q + 3] += bb.w;
      }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = relu_nan(v[i]);   // ReLU that keeps NaN (an overflow must stay visible)
    if (c + 16 > n_real) {   // padding columns may alias another accumulator's columns: they carry exact zeros,
#pragma unroll               // except the constant-one column that carries the next layer's bias through its GEMM
      for (int i = 0; i < 16; ++i)
        if (c + i >= n_real) v[i] = (c + i == one_col) ? 1.0f : 0.0f;
    }
    if ((c >> 4) < n_free) pipe.wait_free(c >> 4);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float u[8] = {v[8 * h], v[8 * h + 1], v[8 * h + 2], v[8 * h + 3], v[8 * h + 4], v[8 * h + 5], v[8 * h + 6], v[8 * h + 7]};
      store_a8<NSPLIT>(a_base, Cfg<NSPLIT>::A_IMAGE, row, c + 8 * h, u);
    }
    if (chase) pipe.publish(c >> 4);
    if (GSUM) {
      const int r_in = g_rin;                   // row index inside its state
      const bool real = r_in < row_cnt;         // padding rows do not enter the mean
      const float inv = row_cnt > 0 ? 1.0f / (float)row_cnt : 0.0f;
      if (n == 16) {
        // transpose-reduce over the 16 lanes of the state: 8 + 4 + 2 + 1 shuffles instead of 16 x 4;
        // lane l of the group ends with the sum of column c + l
        const int lane = threadIdx.x & 31;
        float w8[8], w4[4], w2[2];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float lo = real ? v[i] : 0.0f, hi = real ? v[i + 8] : 0.0f;
          const float send = (lane & 8) ? lo : hi, keep = (lane & 8) ? hi : lo;
          w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float send = (lane & 4) ? w8[i] : w8[i + 4], keep = (lane & 4) ? w8[i + 4] : w8[i];
          w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float send = (lane & 2) ? w4[i] : w4[i + 2], keep = (lane & 2) ? w4[i + 2] : w4[i];
          w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        const float send = (lane & 1) ? w2[0] : w2[1], keep = (lane & 1) ? w2[1] : w2[0];
        const float tot = keep + __shfl_xor_sync(0xffffffffu, send, 1);
        if (g_sid < g_states) g_out[g_sid * g_ld + c + (lane & 15)] = tot * inv;
      } else if (n >= 32) {
        // the warp's 32 rows belong to one state (a state longer than a warp leaves one partial per warp, g_part)
        const int lane = threadIdx.x & 31;
        float x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = real ? v[i] : 0.0f;
        const float tot = warp_colsum16(x, lane);
        if (!(lane & 1) && g_sid < g_states)
          g_out[((size_t)g_part * g_states + g_sid) * g_ld + c + (lane >> 1)] = tot * inv;
      } else if (32 % n == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float x = real ? v[i] : 0.0f;
          for (int o = 1; o < n; o <<= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
          if (r_in == 0 && g_sid < g_states) g_out[g_sid * g_ld + c + i] = x * inv;
        }
      } else {
        // any row count: the state's rows are one run of lanes of this warp (g_seg identifies the run; a state
        // longer than a warp leaves one partial per warp, g_part).  Segmented shuffle reduction, the run's
        // first lane stores.  The order of the additions depends on n only (see the row packing in the kernel).
        const int lane = threadIdx.x & 31;
        float x[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = real ? v[i] : 0.0f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          if (o >= n) break;          // a run holds at most min(n, 32) lanes: ceil(log2) steps reach its head (warp-uniform)
          const int s2 = __shfl_down_sync(0xffffffffu, g_seg, o);
          const bool take = lane + o < 32 && s2 == g_seg;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float y = __shfl_down_sync(0xffffffffu, x[i], o);
            if (take) x[i] += y;
          }
        }
        const int up = __shfl_up_sync(0xffffffffu, g_seg, 1);
        if ((lane == 0 || up != g_seg) && g_seg >= 0 && g_sid < g_states) {
          float4 *dst = reinterpret_cast<float4 *>(g_out + ((size_t)g_part * g_states + g_sid) * g_ld + c);
#pragma unroll
          for (int qd = 0; qd < 4; ++qd)
            dst[qd] = make_float4(x[4 * qd] * inv, x[4 * qd + 1] * inv, x[4 * qd + 2] * inv, x[4 * qd + 3] * inv);
        }
      }
    }
  }
}

// The per-state mean of H1 (G[k][state]) as the A operand of the global half of attention.0: row r carries the
// mean of its own state, so [H1 | G] . [W_local | W_global]^T accumulates in one TMEM accumulator (sarl.py:51-63).
template <int NSPLIT>
__device__ __forceinline__ void gx_to_a(Pipe<NSPLIT> &pipe, const float *G, int g_ld, int state, int cg, int h1d, int h1p,
                                        uint32_t a_base, int row, int n_free, int n_parts, int part_stride) {
  for (int c = 16 * cg; c < h1p; c += 16 * NCG) {
    float v[16];
    const float4 *g4 = reinterpret_cast<const float4 *>(G + state * g_ld + c);   // one address per state: broadcast
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 g = g4[q];
      v[4 * q] = g.x; v[4 * q + 1] = g.y; v[4 * q + 2] = g.z; v[4 * q + 3] = g.w;
    }
    for (int pt = 1; pt < n_parts; ++pt) {       // a state longer than a warp: one partial mean per warp, fixed order
      const float4 *h4 = reinterpret_cast<const float4 *>(G + (size_t)pt * part_stride + state * g_ld + c);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 g = h4[q];
        v[4 * q] += g.x; v[4 * q + 1] += g.y; v[4 * q + 2] += g.z; v[4 * q + 3] += g.w;
      }
    }
    if (c + 16 > h1d) {
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (c + i >= h1d) v[i] = 0.0f;
    }
    if ((c >> 4) < n_free) pipe.wait_free(c >> 4);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float u[8] = {v[8 * h], v[8 * h + 1], v[8 * h + 2], v[8 * h + 3], v[8 * h + 4], v[8 * h + 5], v[8 * h + 6], v[8 * h + 7]};
      store_a8<NSPLIT>(a_base, Cfg<NSPLIT>::A_IMAGE, row, c + 8 * h, u);
    }
    pipe.publish(c >> 4);
  }
}

// partial of sum_c relu(acc[c] + bias[c]) * w[c] over this thread's column group
__device__ __forceinline__ float epi_dot(uint32_t tmem_row, int cg, int col0, int ncols, const float *__restrict__ bias,
                                         const float *__restrict__ w) {
  float acc = 0.0f;
  for (int c = 16 * cg; c < ncols; c += 16 * NCG) {
    float v[16];
    tmem_ld16(tmem_row + col0 + c, v);
    if (bias) {              // unless the bias rode in the GEMM
      const float4 *b4 = reinterpret_cast<const float4 *>(bias + c);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 bb = __ldg(b4 + q);
        v[4 * q] += bb.x; v[4 * q + 1] += bb.y; v[4 * q + 2] += bb.z; v[4 * q + 3] += bb.w;
      }
    }
    // w lives in the kernel's parameter block and c is warp-uniform: constant-bank operands, no load latency
#pragma unroll
    for (int i = 0; i < 16; ++i) acc = fmaf(relu_nan(v[i]), w[c + i], acc);
  }
  return acc;
}

template <int NSPLIT>
__device__ __forceinline__ void pipe_init(Pipe<NSPLIT> &pipe, uint8_t *smem, const Smem &L, const TcProgram &P,
                                          long long my_tiles) {
  constexpr int ST = Cfg<NSPLIT>::STAGES;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bars);
  pipe.full = bars; pipe.empty = bars + ST; pipe.acc_bar = bars + 2 * ST; pipe.a_bar = bars + 2 * ST + 1;
  pipe.kbar = bars + 2 * ST + 2; pipe.afree = pipe.kbar + NKB;
  pipe.ready = reinterpret_cast<uint32_t *>(pipe.afree + NKB);
  pipe.acc_s = smem_u32(pipe.acc_bar); pipe.abar_s = smem_u32(pipe.a_bar);
  pipe.kbar_s = smem_u32(pipe.kbar); pipe.afree_s = smem_u32(pipe.afree);
  pipe.issued = 0; pipe.seen = 0;
  static_assert((2 * ST + 3 + 2 * NKB) * 8 <= 512, "barrier region");
  pipe.f_phase = 0;
  pipe.wbuf = smem + L.w; pipe.wpack = P.wpack;
  pipe.n_stage_slabs = P.n_slabs;
  uint2 *tab = reinterpret_cast<uint2 *>(smem + L.tab);
  for (int i = threadIdx.x; i < P.n_slabs; i += blockDim.x) tab[i] = make_uint2(__ldg(P.slab_off + i), __ldg(P.slab_bytes + i));
  pipe.tab = tab;
  pipe.st = 0; pipe.leader = false; pipe.acc_phase = 0; pipe.a_phase = 0;
  pipe.trace = nullptr; pipe.trace_pos = 0; pipe.fine = false;
  pipe.total = my_tiles * P.n_slabs;
  if (threadIdx.x == 0) {
    for (int i = 0; i < ST; ++i) { mbar_init(&pipe.full[i], 1); mbar_init(&pipe.empty[i], 1); }
    mbar_init(pipe.acc_bar, 1);
    mbar_init(pipe.a_bar, NCREW);
    *pipe.ready = 0;
    for (int i = 0; i < NKB; ++i) { mbar_init(&pipe.kbar[i], TILE_M); mbar_init(&pipe.afree[i], 1); }
    fence_barrier_init();
  }
}

// ---- the MMA warp ------------------------------------------------------------------------------------------
// Everything this loop touches is WARP-UNIFORM BY CONSTRUCTION and visibly so to the compiler: shared-memory
// addresses derived from the kernel's dynamic shared-memory symbol, stage fields read at constant offsets of the
// kernel parameter block (the schedule loop is fully unrolled over P.issue[]), the TMEM base broadcast from lane 0.
// ptxas then keeps the descriptors, the ring slot and the loop counters in UNIFORM registers and a k-step is ~35
// uniform-datapath instructions around its three UTCHMMA and two UTCBAR.  The first version of this loop (Pipe::
// mma_stage, values reached through generic pointers in a struct and dynamically indexed parameters) compiled to ~80
// vector instructions per k-step, 21 of them R2UR moves -- and this warp shares its scheduler with four busy crew
// warps, so the issue path, not the tensor pipe, set the floor: ~343 cycles per k-step whatever N (tools/trace_mma.py:
// the N = 104 k-steps of mlp2.2 took as long as the N = 208 ones, whose three MMAs need 313 cycles of pipe).
constexpr int MAX_SCHED = 12;
template <int NSPLIT>
__device__ __forceinline__ void mma_warp_main(const TcProgram &P, const uint8_t *smem, uint32_t tmem_base_lane,
                                              long long my_tiles) {
  constexpr uint32_t ST = Cfg<NSPLIT>::STAGES;
  const Smem L = smem_layout<NSPLIT>();
  const uint32_t sbase = smem_u32(smem);
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, tmem_base_lane, 0);
  const uint32_t bars = sbase + L.bars;                       // same order as pipe_init
  const uint32_t empty0 = bars + 8u * ST, acc_bar = bars + 16u * ST, a_bar = acc_bar + 8u;
  const uint32_t afree0 = a_bar + 8u + 8u * NKB, ready = afree0 + 8u * NKB;
  const uint32_t a_desc0 = (((sbase + L.a) >> 4) & 0x3FFFu) | ((uint32_t)(A_CHUNK_BYTES >> 4) << 16);
  const uint32_t w_desc0 = ((sbase + L.w) >> 4) & 0x3FFFu;
  uint32_t slot = 0, issued = 0, seen = 0, a_phase = 0;
  for (long long t = 0; t < my_tiles; ++t) {
    {   // sleep until the tile's input is staged (the scout clears it right after)
      uint32_t spins = 0, ok = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(a_bar), "r"(a_phase) : "memory");
        if (++spins > (1u << 22)) __trap();
      }
      a_phase ^= 1u;
    }
#pragma unroll
    for (int i = 0; i < MAX_SCHED; ++i) {
      if (i < P.n_sched) {
        const TcIssue &E = P.issue[i];
        const uint32_t d = tmem_base + (uint32_t)E.acc_col;
        const uint32_t np = (uint32_t)E.np, idesc = E.idesc;
        const uint32_t b_base = w_desc0 | (np << 16);
        const bool commit_k = E.commit_k != 0;
        const int ksteps = E.ksteps;
        const uint32_t base = issued;             // schedule index of this entry's first k-step
        uint32_t a_cur = a_desc0;
        bool acc = E.accumulate != 0;
        for (int ks = 0; ks < ksteps; ++ks) {
          if ((int)(seen - base) <= ks) {
            // the scout sleeps on the barriers; this warp only watches its count (a shared-memory load per trip)
            uint32_t idle = 0;
            do {
              asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(seen) : "r"(ready) : "memory");
              if (++idle > (1u << 26)) __trap();   // bounded: a protocol bug traps instead of hanging
            } while ((int)(seen - base) <= ks);
            tc_fence_after();
          }
          umma_kstep<NSPLIT>(d, a_cur, b_base + slot * (Cfg<NSPLIT>::STAGE_BYTES / 16), Cfg<NSPLIT>::A_IMAGE / 16, np * 2, idesc,
                             acc, empty0 + slot * 8u, afree0 + (uint32_t)ks * 8u, commit_k);
          acc = true;
          a_cur += 2 * A_CHUNK_BYTES / 16;
          slot = slot + 1 == ST ? 0 : slot + 1;
        }
        issued = base + (uint32_t)ksteps;
        if (E.commit_acc) {
          asm volatile(
              "{\n\t.reg .pred q;\n\t"
              "elect.sync _|q, 0xffffffff;\n\t"
              "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
              ::"r"(acc_bar) : "memory");
        }
      }
    }
  }
}

struct TcEntityParams {
  TcProgram prog;
  const float *vin;          // materialised rows, or NULL: the fused input path below
  // fused input: the crew builds its k-chunk of every rotated row (rl/policy/cadrl.py:236-337) from the
  // bound state and the robot record ebc_lookahead left per (episode, action); the n x D input never touches HBM
  const float4 *rec;         // [n_states * 3]: {out0..3}, {out4, out5, robot px, py}, {cos, sin, -, -}
  const float4 *hum_pv, *hum_gr, *stat;
  const float2 *hum_nv;
  const uint8_t *hum_type;
  int Hmax, Smax;
  double dt;
  const int32_t *row_count, *hum_count, *stat_count;
  int n_actions;
  long long n_states;
  int n, ts, D;
  float *joint;
  float *attn;     // [n_states][n] softmax attention weights (ebc_set_attention_output), or null
  int jd;          // self_dim + H2
  int jch;         // k-chunks of a joint row: (jd + 7) / 8
  int self_dim;
  long long *trace;
  int trace_fine;
};

template <int NSPLIT>
__global__ void __launch_bounds__(NT, 1) tc_entity_kernel(const TcEntityParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const Smem L = smem_layout<NSPLIT>();
  uint8_t *A = smem + L.a;
  float *G = reinterpret_cast<float *>(smem + L.g);     // G[state][KMAX]
  float *SC = reinterpret_cast<float *>(smem + L.sc), *XS = reinterpret_cast<float *>(smem + L.xs);
  uint32_t &tmem_slot = *reinterpret_cast<uint32_t *>(smem + L.misc);     // no static shared memory: the dynamic
  int (*cnt2)[MAX_STATES] = reinterpret_cast<int (*)[MAX_STATES]>(smem + L.misc + 16);   // region needs every byte
  const TcProgram &P = p.prog;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler

  Pipe<NSPLIT> pipe;
  // Tile counts in 32-bit unsigned arithmetic (launch_tc bounds n_states below 2^31).  Not a micro-optimisation: a
  // 64-bit division by a run-time value compiles to a subroutine whose result ptxas no longer treats as warp-uniform,
  // and with a "non-uniform" trip count every loop of the MMA warp fell back to vector registers + R2UR (see
  // mma_warp_main).
  const long long n_tiles = (long long)(((uint32_t)p.n_states + (uint32_t)p.ts - 1u) / (uint32_t)p.ts);
  const long long my_tiles = (long long)(((uint32_t)n_tiles - blockIdx.x + gridDim.x - 1u) / gridDim.x);   // grid <= n_tiles
  pipe_init<NSPLIT>(pipe, smem, L, P, my_tiles);
  pipe.trace = p.trace;
  pipe.fine = p.trace != nullptr && p.trace_fine != 0;
  if (warp == 0) tmem_alloc(&tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t a_smem = smem_u32(A);

  if (warp == NCREW / 32) {
    // =================================== MMA warp (converged) ===================================
    // schedule (built on the host, P.sched): mlp1.0 halves | commit #1 | mlp1.2 K chunks chasing the wide-half
    // epilogues | commit #2 | mlp2.0 chasing the H1 epilogue, attention.0 local (H1 complete), attention.0 global
    // chasing the G operand | commit #3 | mlp2.2 chasing the T2 epilogue, attention.2 chasing the U epilogue | commit #4
    if (p.trace == nullptr) {
      mma_warp_main<NSPLIT>(P, smem, tmem_slot, my_tiles);
    } else {          // EBC_TC_TRACE: the instrumented (slower) form of the same loop
      pipe.leader = elect_one();
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) pipe.mma_tile(P, a_smem, tmem_base);
    }
  } else if (warp == NCREW / 32 + 1) {
    pipe.scout_loop(P, my_tiles);
  } else if (warp >= NCREW / 32 + 2) {
    if ((tid & 31) == 0) pipe.loader_loop(warp - (NCREW / 32 + 2));
  } else {
    // ======================================= epilogue crew =======================================
    const int row = ((warp & 3) << 5) | (tid & 31);   // TMEM lane quarter of this warp
    const int cg = warp >> 2;                         // column group
    const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int n = p.n, ts = p.ts, D = p.D;
    const int h1d = P.h1d, h2d = P.h2d;
    const int h1b = P.st[ST_L2].ksteps;               // 16-column blocks of H1 (= k-steps of its readers)
    const int g_ld = 16 * h1b;                        // row stride of G
    // this thread's 8 input values of a tile (k-chunk cg of row `row`), fetched one tile ahead so that the
    // HBM latency of the value-network input hides behind the previous tile's tail
    // Row -> (state, row inside the state).  States are packed so that the position of a state's rows relative
    // to the warp boundaries depends on n only: floor(32 / n) whole states per warp when n <= 32 (the remaining
    // lanes of the warp are padding rows), a whole number of warps per state otherwise.  Every per-state
    // reduction then adds in an order that does not depend on where the state sits in the batch, and results
    // are bit-identical under any permutation of the episodes.  For n | 32 this is the plain row = s n + r.
    const int spw = n <= 32 ? 32 / n : 0, wps = (n + 31) >> 5;     // states per warp / warps per state
    int my_sid, my_rin;
    bool my_row;
    if (n <= 32) {
      const int sl = (row & 31) / n;
      my_sid = (row >> 5) * spw + sl;
      my_rin = (row & 31) - sl * n;
      my_row = sl < spw;
    } else {
      my_sid = (row >> 5) / wps;
      my_rin = ((row >> 5) % wps) * 32 + (row & 31);
      my_row = my_rin < n;
    }
    my_row = my_row && my_sid < ts;
    const int one0 = P.st[ST_L0A].bias_k;             // constant-one input column: mlp1.0's bias rides in its GEMM
    float xu[8];
    int cnt_next = 0;                                 // threads 0-31: row count of state tid of the next tile
    auto load_x = [&](long long t) {
      const long long t0 = t * ts;
      const int tstates = (int)min((long long)ts, p.n_states - t0);
      if (p.vin) {
        const float *src = p.vin + ((size_t)(t0 + my_sid) * n + my_rin) * D;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = 8 * cg + j;
          xu[j] = k == one0 ? 1.0f : ((my_row && my_sid < tstates && k < D) ? __ldg(src + k) : 0.0f);
        }
      } else {
        // Fused input: this thread's k-chunk of its rotated row, computed exactly like K3's rotate_row (separate
        // roundings: the _rn intrinsics are never contracted), so that the values equal the materialised ones bit for bit.
        float (&o)[8] = xu;
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.0f;
        if (my_row && my_sid < tstates && cg < 3) {
          const long long gs = t0 + my_sid;
          const int e = (int)((uint32_t)gs / (uint32_t)p.n_actions);      // (n_states < 2^31: launch_tc)
          // every load that does not depend on the entity counts is issued first, the human slot speculatively (always
          // in bounds): one L2 round trip for a human row, two for a static disc, instead of three dependent ones
          const int H = __ldg(p.hum_count + e), S = __ldg(p.stat_count + e);
          const float4 *rc = p.rec + (size_t)gs * 3;
          const float4 ra = __ldg(rc), rb = __ldg(rc + 1), rcs = __ldg(rc + 2);
          const size_t h = (size_t)e * p.Hmax + min(my_rin, p.Hmax - 1);
          const float4 pv = __ldg(p.hum_pv + h);
          const float2 nv = __ldg(p.hum_nv + h);
          const float hr = __ldg(p.hum_gr + h).w;
          const int hty = __ldg(p.hum_type + h);
          if (my_rin < H + S) {
            float px1, py1, vx1 = 0.0f, vy1 = 0.0f, r1;
            int ty = 3;                                    // EBC_ADULT_STATIC (env.py:457-458)
            if (my_rin < H) {                              // the human advanced by its ORCA action (agent.py:80-93)
              px1 = (float)__dadd_rn((double)pv.x, __dmul_rn((double)nv.x, p.dt));
              py1 = (float)__dadd_rn((double)pv.y, __dmul_rn((double)nv.y, p.dt));
              vx1 = nv.x; vy1 = nv.y;
              r1 = hr;
              ty = hty;
            } else {
              const float4 sd = __ldg(p.stat + (size_t)e * p.Smax + (my_rin - H));
              px1 = sd.x; py1 = sd.y; r1 = sd.z;
            }
            if (cg == 0) {
              const float ddx = __fsub_rn(px1, rb.z), ddy = __fsub_rn(py1, rb.w);
              o[0] = ra.x; o[1] = ra.y; o[2] = ra.z; o[3] = ra.w; o[4] = rb.x; o[5] = rb.y;
              o[6] = __fadd_rn(__fmul_rn(ddx, rcs.x), __fmul_rn(ddy, rcs.y));
              o[7] = __fsub_rn(__fmul_rn(ddy, rcs.x), __fmul_rn(ddx, rcs.y));
            } else if (cg == 1) {
              const float ax = __fsub_rn(rb.z, px1), ay = __fsub_rn(rb.w, py1);
              o[0] = __fadd_rn(__fmul_rn(vx1, rcs.x), __fmul_rn(vy1, rcs.y));
              o[1] = __fsub_rn(__fmul_rn(vy1, rcs.x), __fmul_rn(vx1, rcs.y));
              o[2] = r1;
              o[3] = __fsqrt_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)));
              o[4] = __fadd_rn(ra.w, r1);
              if (D == 17) { o[5] = ty == 0 ? 1.0f : 0.0f; o[6] = ty == 1 ? 1.0f : 0.0f; o[7] = ty == 2 ? 1.0f : 0.0f; }
            } else if (D == 17) {
              o[0] = ty == 3 ? 1.0f : 0.0f;
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (8 * cg + j == one0) xu[j] = 1.0f;
      }
      if (tid < MAX_STATES) {
        int c = 0;
        if (tid < tstates) {
          if (p.row_count) c = __ldg(p.row_count + t0 + tid);
          else {
            const int e = (int)((uint32_t)(t0 + tid) / (uint32_t)p.n_actions);
            c = __ldg(p.hum_count + e) + __ldg(p.stat_count + e);
          }
          c = min(max(c, 0), n);
        }
        cnt_next = c;
      }
    };
    // X -> A (K padded to 32): thread (row, cg) converts k-chunk cg of the values prefetched by load_x; the
    // per-tile side data (row counts, self states) is double-buffered because the next tile's input is staged
    // -- and its mlp1.0 runs -- while this tile's softmax and pooling are still in progress
    float *XS2 = XS;
    auto stage_x = [&](int buf, long long t) {
      const int tstates = (int)min((long long)ts, p.n_states - t * ts);
      if (tid < MAX_STATES) cnt2[buf][tid] = cnt_next;
      store_a8<NSPLIT>(a_smem, Cfg<NSPLIT>::A_IMAGE, row, 8 * cg, xu);
      if (cg == 0 && my_row && my_sid < tstates && my_rin == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k)        // (constant indices: xu stays in registers)
          if (k < p.self_dim) XS2[buf * MAX_STATES * 8 + my_sid * 8 + k] = xu[k];
      }
      pipe.signal_a();
    };
    const bool early_x = P.early_x != 0;
    int cur = 0;
    bool staged = false;
    if (blockIdx.x < n_tiles) load_x(blockIdx.x);
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long s0 = tile * ts;
      const int ns = (int)min((long long)ts, p.n_states - s0);
      const bool was_early = staged;     // staged during the previous tile's tail: a crew_sync has passed since
      if (!staged) stage_x(cur, tile);
      staged = false;
      int *cnt = cnt2[cur];
      float *XS = XS2 + cur * MAX_STATES * 8;
      // ---- mlp1.0 -> mlp1.2, K chunked by the wide halves; the second half overwrites the blocks of the first
      //      as mlp1.2's first K chunk releases them --------------------------------------------------------
      pipe.wait_acc();                                                             // #1
#pragma unroll
      for (int h = 0; h < 2; ++h) {        // (unrolled: the stage fields become constant-bank operands, no LDC chain)
        if (h < P.n_wide) {
          const TcStage &W = P.st[ST_L0A + h];
          const int n_free = h ? P.st[ST_L1A + (h ? h - 1 : 0)].ksteps : 0;
          EpiExtra ex;
          ex.one_col = P.st[ST_L1A + h].bias_k;     // this half is K chunk h of mlp1.2
          epi_to_a<NSPLIT, false>(pipe, tmem_row, cg, W.acc_col, W.np, W.n_real, W.bias_k >= 0 ? nullptr : P.bias[0] + W.n_lo,
                                  a_smem, row, true, n_free, ex);
          if (h) pipe.f_phase ^= low_bits(n_free);
          pipe.stamp();
        }
      }
      // ---- H1 -> A; mlp2.0 and attention.0 (local half) share it -------------------------------------
      pipe.wait_acc();                                                             // #2
      const int st_of_row = min(my_sid, ts - 1);
      {
        const TcStage &S = P.st[ST_L1A];
        const bool b_in = S.bias_k >= 0 || (P.n_wide > 1 && P.st[ST_L1B].bias_k >= 0);   // bias rode in one K chunk
        EpiExtra ex;
        ex.one_col = P.st[ST_L2].bias_k;          // H1 is the A operand of mlp2.0 and attention.0
        if (P.with_global) {
          if (!was_early) crew_sync();   // cnt[] of this tile is visible
          // the state mean of H1 is taken here, from the fp32 registers of the epilogue: a segmented shuffle reduction
          // per warp; a state longer than a warp (n > 32) leaves one partial per warp, added in a fixed order below
          ex.n = n; ex.row_cnt = cnt[st_of_row]; ex.g_out = G; ex.g_ld = g_ld; ex.g_states = ts;
          ex.g_sid = my_sid; ex.g_rin = my_rin; ex.g_seg = my_row ? my_sid : -1 - (row >> 5);
          ex.g_part = n <= 32 ? 0 : (row >> 5) % wps;
          epi_to_a<NSPLIT, true>(pipe, tmem_row, cg, S.acc_col, S.np, S.n_real, b_in ? nullptr : P.bias[1], a_smem, row, true, 0, ex);
        } else {
          epi_to_a<NSPLIT, false>(pipe, tmem_row, cg, S.acc_col, S.np, S.n_real, b_in ? nullptr : P.bias[1], a_smem, row, true, 0, ex);
        }
      }
      pipe.stamp();
      // ---- global state: G -> A as the second K chunk of attention.0 (sarl.py:51-63) -----------------------
      if (P.with_global) {
        if (n <= 32) {
          __syncwarp();    // G[., state] for this thread's blocks was written by this warp
          gx_to_a<NSPLIT>(pipe, G, g_ld, st_of_row, cg, h1d, 16 * h1b, a_smem, row, h1b, 1, 0);
        } else {
          crew_sync();     // the partial means of the state's other warps
          gx_to_a<NSPLIT>(pipe, G, g_ld, st_of_row, cg, h1d, 16 * h1b, a_smem, row, h1b, wps, ts * g_ld);
        }
        pipe.f_phase ^= low_bits(h1b);                  // attention.0 (local half) released every H1 block
        pipe.stamp();
      }
      // ---- T2 = relu(mlp2.0) -> A; mlp2.2 chases it --------------------------------------------------------
      {
        const TcStage &S = P.st[ST_L2];
        EpiExtra ex;
        ex.one_col = P.st[ST_L3].bias_k;
        ex.wait_first = true;
        epi_to_a<NSPLIT, false>(pipe, tmem_row, cg, S.acc_col, S.np, S.n_real, S.bias_k >= 0 ? nullptr : P.bias[2], a_smem, row, true,
                                h1b, ex);
        pipe.f_phase ^= low_bits(h1b);                  // attention.0's last K chunk released its blocks
      }
      pipe.stamp();
      crew_sync();       // every read of the mlp2.0 accumulator is done: attention.2 will reuse its columns
      pipe.wait_acc();                                                             // #3: attention.0 complete
      // ---- U = relu(attention.0) -> A; attention.2 chases it ----------------------------------------------
      {
        const TcStage &S = P.st[ST_L4];
        const int n_free = P.st[ST_L3].ksteps;
        EpiExtra ex;
        ex.one_col = P.st[ST_L5].bias_k;
        epi_to_a<NSPLIT, false>(pipe, tmem_row, cg, S.acc_col, S.np, S.n_real, S.bias_k >= 0 ? nullptr : P.bias[4], a_smem, row, true,
                                n_free, ex);
        pipe.f_phase ^= low_bits(n_free);
      }
      pipe.stamp();
      if (tile + gridDim.x < n_tiles) load_x(tile + gridDim.x);   // next tile's input, consumed after the tail
      pipe.wait_acc();                                                             // #4
      // ---- attention.4 score, masked softmax, pooling (sarl.py:64-78) --------------------------------------
      {
        const TcStage &S = P.st[ST_L5];
        SC[cg * TILE_M + row] = epi_dot(tmem_row, cg, S.acc_col, S.np, S.bias_k >= 0 ? nullptr : P.bias[5], P.w6c);
      }
      pipe.stamp();
      crew_sync();
      pipe.stamp();
      if (early_x && tile + gridDim.x < n_tiles) {
        // every read of the attention.2 accumulator is done and no MMA reads the operand images any more: hand
        // the next tile's input over now, its mlp1.0 (columns [0, 304)) runs under this tile's softmax / pooling
        // (which read the mlp2.2 accumulator at [400, 512) only)
        stage_x(cur ^ 1, tile + gridDim.x);
        staged = true;
      }
      if (n == 16) {
        // a state = 16 aligned lanes of this warp: softmax and pooling stay in registers (shuffles), the
        // pooled feature goes straight to the joint row -- no scratch, no further block synchronisation
        const int lane = tid & 31, st_row = row >> 4;
        const int c_real = cnt[min(st_row, ts - 1)];
        const float sc = ((SC[row] + SC[TILE_M + row]) + (SC[2 * TILE_M + row] + SC[3 * TILE_M + row])) + P.b6;
        const bool real = st_row < ns && (row & 15) < c_real;
        float e = (real && sc != 0.0f) ? expf(sc) : 0.0f;
        float sum = e;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float wrow = e / sum;          // NaN like the reference when every score is exactly 0
        if (p.attn && cg == 0 && st_row < ns) p.attn[(size_t)(s0 + st_row) * 16 + (row & 15)] = real ? wrow : 0.0f;
        const TcStage &S = P.st[ST_L3];
        for (int c = 16 * cg; c < S.np; c += 16 * NCG) {
          float v[16];
          tmem_ld16(tmem_row + S.acc_col + c, v);
          if (S.bias_k < 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float4 bb = __ldg(reinterpret_cast<const float4 *>(P.bias[3] + c) + q);
              v[4 * q] += bb.x; v[4 * q + 1] += bb.y; v[4 * q + 2] += bb.z; v[4 * q + 3] += bb.w;
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = (real && c + i < h2d) ? v[i] * wrow : 0.0f;
          float w8[8], w4[4], w2[2];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float send = (lane & 8) ? v[i] : v[i + 8], keep = (lane & 8) ? v[i + 8] : v[i];
            w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float send = (lane & 4) ? w8[i] : w8[i + 4], keep = (lane & 4) ? w8[i + 4] : w8[i];
            w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
          }
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const float send = (lane & 2) ? w4[i] : w4[i + 2], keep = (lane & 2) ? w4[i + 2] : w4[i];
            w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
          }
          const float send = (lane & 1) ? w2[0] : w2[1], keep = (lane & 1) ? w2[1] : w2[0];
          const float tot = keep + __shfl_xor_sync(0xffffffffu, send, 1);
          const int col = c + (lane & 15);
          if (st_row < ns && col < h2d) p.joint[joint_index(s0 + st_row, p.self_dim + col, p.jch)] = tot;
        }
        if (tid < ns * p.self_dim) {
          const int s = tid / p.self_dim, k = tid % p.self_dim;
          p.joint[joint_index(s0 + s, k, p.jch)] = XS[s * 8 + k];
        }
      } else if (n <= 32) {
        // a state = one run of n lanes of this warp (n does not divide 32: the warp's last lanes are padding rows).
        // Softmax denominator and pooled features by segmented shuffle reductions whose order depends on n only;
        // the run's first lane stores the pooled block straight to the joint row -- no scratch, no block barrier.
        // (Same additions as the scratch path below, which adds its single per-warp partial to 0.)
        const int lane = tid & 31;
        const int sid = my_row ? my_sid : -1 - (row >> 5);         // padding rows: a run of their own
        const bool in_tile = my_row && my_sid < ns;
        const bool real = in_tile && my_rin < cnt[min(my_sid, MAX_STATES - 1)];
        const int head_lane = my_row ? lane - my_rin : lane;
        const bool head = lane == head_lane;
        const TcStage &S = P.st[ST_L3];
        const float sc = ((SC[row] + SC[TILE_M + row]) + (SC[2 * TILE_M + row] + SC[3 * TILE_M + row])) + P.b6;
        const float e = (real && sc != 0.0f) ? expf(sc) : 0.0f;
        float es = e;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          if (o >= n) break;
          const float y = __shfl_down_sync(0xffffffffu, es, o);
          const int s2 = __shfl_down_sync(0xffffffffu, sid, o);
          if (lane + o < 32 && s2 == sid) es += y;
        }
        const float sum = __shfl_sync(0xffffffffu, es, head_lane);
        const float wrow = e / sum;          // NaN like the reference when every score of the state is exactly 0
        if (p.attn && cg == 0 && in_tile) p.attn[(size_t)(s0 + my_sid) * n + my_rin] = real ? wrow : 0.0f;
        for (int c = 16 * cg; c < S.np; c += 16 * NCG) {
          float v[16];
          tmem_ld16(tmem_row + S.acc_col + c, v);
          if (S.bias_k < 0) {
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
              const float4 bb = __ldg(reinterpret_cast<const float4 *>(P.bias[3] + c) + qd);
              v[4 * qd] += bb.x; v[4 * qd + 1] += bb.y; v[4 * qd + 2] += bb.z; v[4 * qd + 3] += bb.w;
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = (real && c + i < h2d) ? v[i] * wrow : 0.0f;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            if (o >= n) break;
            const int s2 = __shfl_down_sync(0xffffffffu, sid, o);
            const bool take = lane + o < 32 && s2 == sid;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float y = __shfl_down_sync(0xffffffffu, v[i], o);
              if (take) v[i] += y;
            }
          }
          if (head && in_tile) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c + i < h2d) p.joint[joint_index(s0 + my_sid, p.self_dim + c + i, p.jch)] = v[i];
          }
        }
        if (tid < ns * p.self_dim) {
          const int s = tid / p.self_dim, k = tid % p.self_dim;
          p.joint[joint_index(s0 + s, k, p.jch)] = XS[s * 8 + k];
        }
      } else {
        // states longer than a warp: a state's rows are a contiguous run of rows, one run of lanes per warp.
        // Sums over a state = segmented shuffle reduction inside each warp, one partial per (lane quarter, state)
        // in shared memory, added in a fixed order (deterministic, no atomics).  Scratch aliases the A region,
        // which is free: every MMA that read it has completed.
        const int lane = tid & 31, q = warp & 3;
        const int sid = my_row ? my_sid : -1 - (row >> 5);         // padding rows: a run of their own
        const int r_in = my_rin;
        const bool in_tile = my_row && my_sid < ns;
        const bool real = in_tile && r_in < cnt[min(my_sid, MAX_STATES - 1)];
        const int sid_up = __shfl_up_sync(0xffffffffu, sid, 1);
        const bool head = lane == 0 || sid_up != sid;              // first lane of this state's run in the warp
        // (behind the four k-chunks of the first image that the next tile's input occupies)
        float *PSUM = reinterpret_cast<float *>(A + 4 * A_CHUNK_BYTES);   // [4][MAX_STATES]
        float *PJ = PSUM + 4 * MAX_STATES;                         // [4][ts][np3]
        const TcStage &S = P.st[ST_L3];
        const float sc = ((SC[row] + SC[TILE_M + row]) + (SC[2 * TILE_M + row] + SC[3 * TILE_M + row])) + P.b6;
        const float e = (real && sc != 0.0f) ? expf(sc) : 0.0f;
        {
          float es = e;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            if (o >= n) break;
            const float y = __shfl_down_sync(0xffffffffu, es, o);
            const int s2 = __shfl_down_sync(0xffffffffu, sid, o);
            if (lane + o < 32 && s2 == sid) es += y;
          }
          if (cg == 0 && head && in_tile) PSUM[q * MAX_STATES + sid] = es;
        }
        crew_sync();
        const int q_lo = n <= 32 ? my_sid / spw : my_sid * wps;    // lane quarters the state spans
        const int q_hi = n <= 32 ? q_lo : q_lo + wps - 1;
        float sum = 0.0f;
        if (in_tile)
          for (int qq = q_lo; qq <= q_hi; ++qq) sum += PSUM[qq * MAX_STATES + sid];
        const float wrow = e / sum;          // NaN like the reference when every score of the state is exactly 0
        if (p.attn && cg == 0 && in_tile) p.attn[(size_t)(s0 + my_sid) * n + r_in] = real ? wrow : 0.0f;
        for (int c = 16 * cg; c < S.np; c += 16 * NCG) {
          float v[16];
          tmem_ld16(tmem_row + S.acc_col + c, v);
          if (S.bias_k < 0) {
#pragma unroll
            for (int qd = 0; qd < 4; ++qd) {
              const float4 bb = __ldg(reinterpret_cast<const float4 *>(P.bias[3] + c) + qd);
              v[4 * qd] += bb.x; v[4 * qd + 1] += bb.y; v[4 * qd + 2] += bb.z; v[4 * qd + 3] += bb.w;
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = (real && c + i < h2d) ? v[i] * wrow : 0.0f;
          // the warp's 32 rows belong to one state: one transpose-reduce, lane pair l holds the sum of column l >> 1
          const float tot = warp_colsum16(v, lane);
          const int sid_w = (row >> 5) / wps;
          if (!(lane & 1) && sid_w < ns) PJ[((size_t)q * ts + sid_w) * S.np + c + (lane >> 1)] = tot;
        }
        crew_sync();
        for (int i = tid; i < ns * p.jd; i += NCREW) {
          const int s = i / p.jd, k = i % p.jd;
          float v;
          if (k < p.self_dim) v = XS[s * 8 + k];
          else {
            v = 0.0f;
            const int col = k - p.self_dim;
            const int a_lo = n <= 32 ? s / spw : s * wps, a_hi = n <= 32 ? a_lo : a_lo + wps - 1;
            for (int qq = a_lo; qq <= a_hi; ++qq) v += PJ[((size_t)qq * ts + s) * S.np + col];
          }
          p.joint[joint_index(s0 + s, k, p.jch)] = v;
        }
      }
      pipe.stamp();
      tc_fence_before();
      crew_sync();   // scratch (aliases A) / SC / XS / cnt are rewritten by the next tile
      pipe.stamp();
      cur ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

struct TcMlp3Params {
  TcProgram prog;
  const float *joint;
  float *values;
  long long n_states;
  int jd, jch;
  long long *trace;          // EBC_TC_TRACE=2: clock64 stamps of CTA 0's crew thread 0
};

template <int NSPLIT>
__global__ void __launch_bounds__(NT, 1) tc_mlp3_kernel(const TcMlp3Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const Smem L = smem_layout<NSPLIT>();
  uint8_t *A = smem + L.a;
  float *SC = reinterpret_cast<float *>(smem + L.sc);
  uint32_t &tmem_slot = *reinterpret_cast<uint32_t *>(smem + L.misc);
  const TcProgram &P = p.prog;
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  Pipe<NSPLIT> pipe;
  const long long n_tiles = (long long)(((uint32_t)p.n_states + (uint32_t)TILE_M - 1u) / (uint32_t)TILE_M);
  const long long my_tiles = (long long)(((uint32_t)n_tiles - blockIdx.x + gridDim.x - 1u) / gridDim.x);   // grid <= n_tiles
  pipe_init<NSPLIT>(pipe, smem, L, P, my_tiles);
  pipe.trace = p.trace;
  if (warp == 0) tmem_alloc(&tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const uint32_t a_smem = smem_u32(A);

  if (warp == NCREW / 32) {
    if (p.trace == nullptr) {
      mma_warp_main<NSPLIT>(P, smem, tmem_slot, my_tiles);
    } else {          // EBC_TC_TRACE: the instrumented (slower) form of the same loop
      pipe.leader = elect_one();
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) pipe.mma_tile(P, a_smem, tmem_base);
    }
  } else if (warp == NCREW / 32 + 1) {
    pipe.scout_loop(P, my_tiles);
  } else if (warp >= NCREW / 32 + 2) {
    if ((tid & 31) == 0) pipe.loader_loop(warp - (NCREW / 32 + 2));
  } else {
    const int row = ((warp & 3) << 5) | (tid & 31);
    const int cg = warp >> 2;
    const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int jd = p.jd, kp = P.st[ST_L0A].ksteps * 16, one0 = P.st[ST_L0A].bias_k;
    // joint row -> A (K padded to a multiple of 16), k-chunks interleaved over the column groups.  The first XC
    // chunks of a thread are fetched one tile ahead (their L2 / HBM latency hides behind the previous tile's last
    // layer) and handed over as soon as that tile's output accumulator has been read.
    constexpr int XC = 4;
    float xu[XC][8];
    const int jch = p.jch;
    auto load_chunk = [&](long long t, int c, float (&u)[8]) {     // k-chunk c of this thread's row of tile t
      const bool live = t * TILE_M + row < p.n_states && c < jch;
      float4 lo = make_float4(0.0f, 0.0f, 0.0f, 0.0f), hi = lo;
      if (live) {
        const float4 *src = reinterpret_cast<const float4 *>(p.joint + joint_index(t * TILE_M + row, 8 * c, jch));
        lo = __ldg(src); hi = __ldg(src + 1);
      }
      const float v[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) u[j] = 8 * c + j == one0 ? 1.0f : ((live && 8 * c + j < jd) ? v[j] : 0.0f);
    };
    auto load_x = [&](long long t) {
#pragma unroll
      for (int i = 0; i < XC; ++i) load_chunk(t, cg + NCG * i, xu[i]);
    };
    auto stage_x = [&](long long t) {
#pragma unroll
      for (int i = 0; i < XC; ++i) {
        const int k0 = 8 * (cg + NCG * i);
        if (k0 < kp) store_a8<NSPLIT>(a_smem, Cfg<NSPLIT>::A_IMAGE, row, k0, xu[i]);
      }
      for (int c = cg + NCG * XC; 8 * c < kp; c += NCG) {      // joint rows wider than 128 columns
        float u[8];
        load_chunk(t, c, u);
        store_a8<NSPLIT>(a_smem, Cfg<NSPLIT>::A_IMAGE, row, 8 * c, u);
      }
      pipe.signal_a();
    };
    bool staged = false;
    if (blockIdx.x < n_tiles) load_x(blockIdx.x);
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long s0 = tile * TILE_M;
      const bool more = tile + gridDim.x < n_tiles;
      if (!staged) stage_x(tile);
      staged = false;
      pipe.wait_acc();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h < P.n_wide) {
          const TcStage &W = P.st[ST_L0A + h];
          const int n_free = h ? P.st[ST_L1A + (h ? h - 1 : 0)].ksteps : 0;
          EpiExtra ex;
          ex.one_col = P.st[ST_L1A + h].bias_k;
          epi_to_a<NSPLIT, false>(pipe, tmem_row, cg, W.acc_col, W.np, W.n_real, W.bias_k >= 0 ? nullptr : P.bias[0] + W.n_lo,
                                  a_smem, row, true, n_free, ex);
          if (h) pipe.f_phase ^= low_bits(n_free);
        }
      }
      pipe.wait_acc();
      {
        const TcStage &S = P.st[ST_L1A];
        const bool b_in = S.bias_k >= 0 || (P.n_wide > 1 && P.st[ST_L1B].bias_k >= 0);
        EpiExtra ex;
        ex.one_col = P.st[ST_L2].bias_k;
        epi_to_a<NSPLIT, false>(pipe, tmem_row, cg, S.acc_col, S.np, S.n_real, b_in ? nullptr : P.bias[1], a_smem, row, true, 0, ex);
      }
      pipe.stamp();
      if (more) load_x(tile + gridDim.x);
      pipe.wait_acc();
      {
        const TcStage &S = P.st[ST_L2];
        SC[cg * TILE_M + row] = epi_dot(tmem_row, cg, S.acc_col, S.np, S.bias_k >= 0 ? nullptr : P.bias[2], P.w6c);
      }
      pipe.stamp();
      crew_sync();
      if (more) {    // every MMA of this tile has completed and its output accumulator has been read
        stage_x(tile + gridDim.x);
        staged = true;
      }
      if (tid < TILE_M && s0 + tid < p.n_states)
        p.values[s0 + tid] = ((SC[tid] + SC[TILE_M + tid]) + (SC[2 * TILE_M + tid] + SC[3 * TILE_M + tid])) + P.b6;
      tc_fence_before();
      crew_sync();   // SC is rewritten by the next tile; the next A stores follow every TMEM read above
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}


// ---------------------------------------------------------------------------------------------------
// Host: pack weights into slab images and build the two per-tile programs.
// ---------------------------------------------------------------------------------------------------
inline int pad16(int x) { return (x + 15) / 16 * 16; }

struct Packer {
  int nsplit;
  std::vector<uint8_t> bytes;
  std::vector<uint32_t> slab_off, slab_bytes;

  static uint16_t round16(float f, int fp16) {
    if (!fp16) return bf16_rn(f);
    const __half h = __float2half_rn(f);
    uint16_t u;
    memcpy(&u, &h, 2);
    return u;
  }
  static float back16(uint16_t u, int fp16) {
    if (!fp16) return bf16_to_f(u);
    __half h;
    memcpy(&h, &u, 2);
    return __half2float(h);
  }
  static uint16_t bf16_rn(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
    const uint32_t r = 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)((u + r) >> 16);
  }
  static float bf16_to_f(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
  }
  // GEMM stage over W[n_lo:n_lo+np (zero past n_hi)][k_lo:k_lo+16*ksteps (zero past k_hi)], W is [out][ld]
  // bias / bias_k: row bias_k (stage-local k) of the weights carries the bias (its A column is a constant one)
  void add_stage(const float *W, int ld, int n_lo, int n_hi, int np, int k_lo, int k_hi, int ksteps, int k_col_off,
                 const float *bias = nullptr, int bias_k = -1) {
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint32_t off = (uint32_t)bytes.size();
      const uint32_t sz = (uint32_t)nsplit * (uint32_t)np * 32u;
      bytes.resize(off + sz, 0);
      for (int c = 0; c < 2; ++c)
        for (int nn = 0; nn < np; ++nn)
          for (int j = 0; j < 8; ++j) {
            const int n = n_lo + nn, k = k_lo + ks * 16 + c * 8 + j;
            float v = (n < n_hi && k < k_hi) ? W[(size_t)n * ld + k_col_off + k] : 0.0f;
            if (bias && ks * 16 + c * 8 + j == bias_k) v = n < n_hi ? bias[n] : 0.0f;
            for (int s = 0; s < nsplit; ++s) {
              const uint16_t h = round16(v, nsplit == 2);
              v -= back16(h, nsplit == 2);
              uint8_t *dst = bytes.data() + off + (size_t)s * np * 32 + (size_t)c * np * 16 + (size_t)nn * 16 + j * 2;
              memcpy(dst, &h, 2);
            }
          }
      slab_off.push_back(off);
      slab_bytes.push_back(sz);
    }
  }
};

template <int NSPLIT>
int launch_tc(ebc_sim *s, const float *vin, int64_t n_states, const int32_t *row_count, float *values, cudaStream_t stream) {
  if (n_states <= 0) return EBC_OK;     // an empty batch launches nothing (like the FFMA path)
  if (n_states >= (1ll << 31) - 256) return ebc_fail(s, EBC_ERR_INVALID, "tensor-core value path: at most 2^31 - 256 states per call");
  const Smem L = smem_layout<NSPLIT>();
  if ((int)L.total > s->max_smem_optin)
    return ebc_fail(s, EBC_ERR_INVALID, "tensor-core value path needs %u B of shared memory", L.total);
  const int n = s->cfg.max_humans + s->cfg.max_statics;
  TcEntityParams p;
  p.prog = s->tc[NSPLIT - 1].entity;
  p.vin = vin; p.row_count = row_count; p.hum_count = s->st.hum_count; p.stat_count = s->st.stat_count;
  p.rec = s->d_la_rec;
  p.hum_pv = reinterpret_cast<const float4 *>(s->st.hum_pv); p.hum_gr = reinterpret_cast<const float4 *>(s->st.hum_gr);
  p.stat = reinterpret_cast<const float4 *>(s->st.stat); p.hum_nv = reinterpret_cast<const float2 *>(s->st.hum_nv);
  p.hum_type = s->st.hum_type; p.Hmax = s->cfg.max_humans; p.Smax = s->cfg.max_statics; p.dt = s->cfg.time_step;
  p.n_actions = s->cfg.n_actions; p.n_states = n_states; p.n = n; p.D = s->net.D;
  // states per 128-row tile: whole states per warp (n <= 32) or whole warps per state (see the kernel)
  int ts = n <= 32 ? 4 * (32 / n) : 4 / ((n + 31) / 32);
  if (ts < 1) return ebc_fail(s, EBC_ERR_INVALID, "tensor-core value path: %d rows per state do not fit a 128-row tile", n);
  if (ts > MAX_STATES) ts = MAX_STATES;
  if (s->net.with_global) {                      // G[state][h1 padded] must fit its shared-memory region
    const int g_cap = (int)(Cfg<NSPLIT>::G_BYTES / (4u * (uint32_t)p.prog.st[ST_L2].ksteps * 16u));
    if (ts > g_cap) ts = g_cap;
  }
  // states longer than a warp pool through a scratch that aliases the operand images: [4][32] + [4][ts][np3] floats
  // (one image minus the four k-chunks that hold the next tile's input)
  while (n > 32 && ts > 1 && 512u + 16u * (uint32_t)ts * (uint32_t)p.prog.st[ST_L3].np > Cfg<NSPLIT>::A_IMAGE - 4u * A_CHUNK_BYTES) --ts;
  p.ts = ts;
  p.joint = s->d_joint; p.jd = s->net.self_dim + s->net.l[3].out; p.self_dim = s->net.self_dim;
  p.attn = s->attn_out;
  p.jch = (p.jd + 7) / 8;
  p.trace = nullptr; p.trace_fine = 0;
  TcMlp3Params q;
  q.trace = nullptr;
  if (const char *tr = getenv("EBC_TC_TRACE")) {      // 1: the entity kernel, 2: the mlp3 kernel
    if (!s->d_trace) { cudaMalloc(&s->d_trace, 4096 * sizeof(long long)); }
    cudaMemsetAsync(s->d_trace, 0, 4096 * sizeof(long long), stream);
    if (tr[0] == '2') q.trace = s->d_trace; else p.trace = s->d_trace;
    p.trace_fine = tr[0] == '3';
  }
  q.prog = s->tc[NSPLIT - 1].mlp3;
  q.joint = s->d_joint; q.values = values; q.n_states = n_states; q.jd = p.jd; q.jch = p.jch;
  cudaFuncSetAttribute(tc_entity_kernel<NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
  cudaFuncSetAttribute(tc_mlp3_kernel<NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total);
  const long long tiles_a = (n_states + ts - 1) / ts, tiles_b = (n_states + TILE_M - 1) / TILE_M;
  const int grid_a = (int)(tiles_a < s->sm_count ? tiles_a : s->sm_count);
  const int grid_b = (int)(tiles_b < s->sm_count ? tiles_b : s->sm_count);
  tc_entity_kernel<NSPLIT><<<grid_a, NT, L.total, stream>>>(p);
  int rc = ebc_check_launch(s, "tc_entity_kernel");
  if (rc) return rc;
  tc_mlp3_kernel<NSPLIT><<<grid_b, NT, L.total, stream>>>(q);
  return ebc_check_launch(s, "tc_mlp3_kernel");
}

}  // namespace

// Build both programs (entity part, mlp3) for one operand-splitting mode.  Returns 0 if the network's
// shape fits the tensor-core path, 1 if it does not (caller keeps the FFMA path), < 0 on error.
int ebc_tc_prepare(ebc_sim *s, const ebc_weights *w, int mode_index, int nsplit) {
  const ebc_linear *m10 = &w->mlp1[0], *m12 = &w->mlp1[1], *m20 = &w->mlp2[0], *m22 = &w->mlp2[1];
  const ebc_linear *a0 = &w->attention[0], *a2 = &w->attention[1], *a4 = &w->attention[2];
  const ebc_linear *p0 = &w->mlp3[0], *p2 = &w->mlp3[1], *p4 = &w->mlp3[2], *p6 = &w->mlp3[3];
  const int h1 = m12->out_dim;
  auto fits = [&](int wide_out, int mid, int last_in) {
    return pad16(wide_out) <= 2 * KMAX && pad16(mid) <= KMAX && pad16(last_in) <= KMAX;
  };
  if (!fits(m10->out_dim, h1, h1) || pad16(m20->out_dim) > KMAX || pad16(m22->out_dim) > KMAX ||
      pad16(a0->out_dim) > KMAX || pad16(a2->out_dim) > KMAX || m10->in_dim > 32 ||
      !fits(p0->out_dim, p2->out_dim, p4->out_dim) || pad16(p0->in_dim) > KMAX ||
      w->self_state_dim > 8)
    return 1;
  TcPrograms &T = s->tc[mode_index];
  std::vector<float> fl;   // biases, w6, wg (fp32 side data)
  auto push_f = [&](const float *src, int nreal, int npad) {
    const size_t off = fl.size();
    fl.resize(off + npad, 0.0f);
    for (int i = 0; i < nreal; ++i) fl[off + i] = src[i];
    return off;
  };
  auto wide_split = [&](int out, int &np0, int &np1) {
    const int op = pad16(out);
    if (op <= KMAX) { np0 = op; np1 = 0; }
    else { np0 = pad16((op / 2 + 15) / 16 * 16); if (np0 > KMAX) np0 = KMAX; np1 = op - np0; }
  };
  // ---- a chain  IN -> wide (ReLU) -> mid (ReLU) -> ...  shared by both kernels -----------------------------
  auto build_front = [&](Packer &pk, TcProgram &P, const ebc_linear *wide, const ebc_linear *mid, int in_pad,
                         size_t &b_wide, size_t &b_mid) {
    int np0, np1;
    wide_split(wide->out_dim, np0, np1);
    P.n_wide = np1 ? 2 : 1;
    const int midp = pad16(mid->out_dim);
    const int col_mid = TMEM_COLS - midp;
    if (np0 + np1 > col_mid) return 1;
    int lo = 0;
    for (int h = 0; h < P.n_wide; ++h) {
      const int np = h ? np1 : np0;
      TcStage &S = P.st[ST_L0A + h];
      S.np = np; S.ksteps = in_pad / 16; S.acc_col = lo; S.accumulate = 0; S.n_lo = lo;
      S.n_real = wide->out_dim - lo < np ? wide->out_dim - lo : np;
      S.bias_k = wide->in_dim < in_pad ? wide->in_dim : -1;        // a spare K column carries the bias
      pk.add_stage(wide->weight, wide->in_dim, lo, wide->out_dim, np, 0, wide->in_dim, S.ksteps, 0, wide->bias, S.bias_k);
      lo += np;
    }
    lo = 0;
    for (int h = 0; h < P.n_wide; ++h) {
      const int kp = h ? np1 : np0;
      TcStage &S = P.st[ST_L1A + h];
      S.np = midp; S.ksteps = kp / 16; S.acc_col = col_mid; S.accumulate = h; S.n_lo = 0; S.n_real = mid->out_dim;
      // the first padding column of the wide layer (global K index wide->out_dim), in whichever K chunk holds it
      const int bk = mid->in_dim < np0 + np1 ? mid->in_dim - lo : -1;
      S.bias_k = bk >= 0 && bk < kp ? bk : -1;
      pk.add_stage(mid->weight, mid->in_dim, 0, mid->out_dim, midp, lo, mid->in_dim, S.ksteps, 0, mid->bias, S.bias_k);
      lo += kp;
    }
    b_wide = push_f(wide->bias, wide->out_dim, np0 + np1);
    b_mid = push_f(mid->bias, mid->out_dim, midp);
    return 0;
  };
  auto simple_stage = [&](Packer &pk, TcStage &S, const ebc_linear *l, int in_real, int in_pad, int acc_col, int k_col_off,
                          bool with_bias = true) {
    S.np = pad16(l->out_dim); S.ksteps = in_pad / 16; S.acc_col = acc_col; S.accumulate = 0; S.n_lo = 0;
    S.n_real = l->out_dim;
    S.bias_k = (with_bias && in_real < in_pad) ? in_real : -1;
    pk.add_stage(l->weight, l->in_dim, 0, l->out_dim, S.np, 0, in_real, S.ksteps, k_col_off, l->bias, S.bias_k);
  };

  Packer pe; pe.nsplit = nsplit;
  Packer pm; pm.nsplit = nsplit;
  TcProgram E, M;
  memset(&E, 0, sizeof(E));
  memset(&M, 0, sizeof(M));
  size_t be0, be1, bm0, bm1;
  if (build_front(pe, E, m10, m12, 32, be0, be1)) return 1;
  const int h1p = pad16(h1);
  // TMEM columns of the second half of the chain.  mlp2.0 (L2), attention.0 (L4) and mlp2.2 (L3) are alive
  // together (L3 chases the epilogue of L2 while L4 waits for its own epilogue); attention.2 (L5) reuses L2's
  // columns.  When the padded widths do not fit in 512 columns the accumulators are packed at their REAL
  // widths rounded to 8: a stage's zero-weight padding columns then alias the first columns of its neighbour,
  // which is harmless because (i) the MMA warp issues L2, L4, L3 in that order, so a stage's (zero) padding
  // writes land before its neighbour's first MMA, and (ii) every epilogue forces padding columns to zero.
  const int np2 = pad16(m20->out_dim), np4 = pad16(a0->out_dim), np3 = pad16(m22->out_dim), np5 = pad16(a2->out_dim);
  int c4 = np2, c3 = np2 + np4;
  if (c3 + np3 > TMEM_COLS) {
    c4 = (m20->out_dim + 7) / 8 * 8;
    c3 = c4 + (a0->out_dim + 7) / 8 * 8;
    if (c3 + np3 > TMEM_COLS) return 1;
  }
  c3 = TMEM_COLS - np3;   // as high as it goes (>= the minimum above): the next tile's mlp1.0 starts under this tile's pooling
  if (np5 > c3) return 1;
  // slab order = issue order of the MMA warp: L2, L4, L4G, L3, L5
  simple_stage(pe, E.st[ST_L2], m20, h1, h1p, 0, 0);                               // mlp2.0
  simple_stage(pe, E.st[ST_L4], a0, h1, h1p, c4, 0);                               // attention.0, local half
  if (w->with_global_state) {                                                      // attention.0, global half
    simple_stage(pe, E.st[ST_L4G], a0, h1, h1p, c4, h1, false);   // the local half already carries the bias
    E.st[ST_L4G].accumulate = 1;
  }
  simple_stage(pe, E.st[ST_L3], m22, m20->out_dim, pad16(m20->out_dim), c3, 0);    // mlp2.2
  simple_stage(pe, E.st[ST_L5], a2, a0->out_dim, pad16(a0->out_dim), 0, 0);        // attention.2
  const size_t be2 = push_f(m20->bias, m20->out_dim, E.st[ST_L2].np);
  const size_t be3 = push_f(m22->bias, m22->out_dim, E.st[ST_L3].np);
  const size_t be4 = push_f(a0->bias, a0->out_dim, E.st[ST_L4].np);
  const size_t be5 = push_f(a2->bias, a2->out_dim, E.st[ST_L5].np);
  const size_t we6 = push_f(a4->weight, a4->in_dim, E.st[ST_L5].np);
  {
    auto add = [](TcProgram &Q, int stage, int chase, int commit_k, int commit_acc) {
      TcSched &e = Q.sched[Q.n_sched++];
      e.stage = (uint8_t)stage; e.chase = (uint8_t)chase; e.commit_k = (uint8_t)commit_k; e.commit_acc = (uint8_t)commit_acc;
    };
    for (int h = 0; h < E.n_wide; ++h) add(E, ST_L0A + h, 0, 0, h + 1 == E.n_wide);
    for (int h = 0; h < E.n_wide; ++h) add(E, ST_L1A + h, 1, h + 1 < E.n_wide, h + 1 == E.n_wide);
    add(E, ST_L2, 1, 0, 0);
    add(E, ST_L4, 0, 1, !w->with_global_state);
    if (w->with_global_state) add(E, ST_L4G, 1, 1, 1);
    add(E, ST_L3, 1, 1, 0);
    add(E, ST_L5, 1, 0, 1);
  }
  if (build_front(pm, M, p0, p2, pad16(p0->in_dim), bm0, bm1)) return 1;
  {
    auto add = [](TcProgram &Q, int stage, int chase, int commit_k, int commit_acc) {
      TcSched &e = Q.sched[Q.n_sched++];
      e.stage = (uint8_t)stage; e.chase = (uint8_t)chase; e.commit_k = (uint8_t)commit_k; e.commit_acc = (uint8_t)commit_acc;
    };
    for (int h = 0; h < M.n_wide; ++h) add(M, ST_L0A + h, 0, 0, h + 1 == M.n_wide);
    for (int h = 0; h < M.n_wide; ++h) add(M, ST_L1A + h, 1, h + 1 < M.n_wide, h + 1 == M.n_wide);
    add(M, ST_L2, 1, 0, 1);
  }
  simple_stage(pm, M.st[ST_L2], p4, p2->out_dim, pad16(p2->out_dim), 0, 0);        // mlp3.4
  const size_t bm2 = push_f(p4->bias, p4->out_dim, M.st[ST_L2].np);
  const size_t wm6 = push_f(p6->weight, p6->in_dim, M.st[ST_L2].np);

  // the schedule as the MMA warp reads it: one flat record per entry, instruction descriptor precomputed
  for (TcProgram *Q : {&E, &M})
    for (int i = 0; i < Q->n_sched; ++i) {
      const TcStage &S = Q->st[Q->sched[i].stage];
      TcIssue &I = Q->issue[i];
      I.np = S.np; I.ksteps = S.ksteps; I.acc_col = S.acc_col; I.accumulate = S.accumulate;
      I.commit_k = Q->sched[i].commit_k; I.commit_acc = Q->sched[i].commit_acc;
      const uint32_t fmt = nsplit == 2 ? 0u : 1u;           // fp16 parts in the two-part mode, bf16 otherwise
      I.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(S.np >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
      I.pad = 0;
    }
  // ---- upload ---------------------------------------------------------------------------------------
  const size_t n_e = pe.slab_off.size(), n_m = pm.slab_off.size();
  if (n_e > (size_t)MAX_SLABS || n_m > (size_t)MAX_SLABS) return 1;
  const size_t bytes_total = pe.bytes.size() + pm.bytes.size() + fl.size() * 4 + (n_e + n_m) * 8 + 1024;
  uint8_t *d = nullptr;
  cudaError_t err = cudaMalloc(&d, bytes_total);
  if (err != cudaSuccess) return ebc_fail(s, EBC_ERR_NOMEM, "cudaMalloc tc weights: %s", cudaGetErrorString(err));
  size_t off = 0;
  auto up = [&](const void *src, size_t nbytes) {
    const size_t o = off;
    if (nbytes) cudaMemcpy(d + o, src, nbytes, cudaMemcpyHostToDevice);
    off += (nbytes + 127) / 128 * 128;
    return o;
  };
  const size_t o_pe = up(pe.bytes.data(), pe.bytes.size());
  const size_t o_pm = up(pm.bytes.data(), pm.bytes.size());
  const size_t o_fl = up(fl.data(), fl.size() * 4);
  const size_t o_eo = up(pe.slab_off.data(), n_e * 4), o_eb = up(pe.slab_bytes.data(), n_e * 4);
  const size_t o_mo = up(pm.slab_off.data(), n_m * 4), o_mb = up(pm.slab_bytes.data(), n_m * 4);
  if (cudaGetLastError() != cudaSuccess) { cudaFree(d); return ebc_fail(s, EBC_ERR_CUDA, "tc weight upload failed"); }
  const float *dfl = reinterpret_cast<const float *>(d + o_fl);
  E.wpack = d + o_pe; E.n_slabs = (int)n_e;
  E.slab_off = reinterpret_cast<const uint32_t *>(d + o_eo); E.slab_bytes = reinterpret_cast<const uint32_t *>(d + o_eb);
  E.bias[0] = dfl + be0; E.bias[1] = dfl + be1; E.bias[2] = dfl + be2; E.bias[3] = dfl + be3;
  E.bias[4] = dfl + be4; E.bias[5] = dfl + be5; E.w6 = dfl + we6; E.b6 = a4->bias[0];
  static_assert(sizeof(E.w6c) / sizeof(float) >= KMAX, "w6c holds the widest last-layer input");
  for (int i = 0; i < KMAX; ++i) E.w6c[i] = i < a4->in_dim ? a4->weight[i] : 0.0f;
  E.with_global = w->with_global_state ? 1 : 0; E.h1d = h1; E.h2d = m22->out_dim;
  // the next tile's mlp1.0 may overwrite columns [0, wide) while this tile's pooling still reads the mlp2.2 accumulator
  E.early_x = (E.st[ST_L0A].np + (E.n_wide > 1 ? E.st[ST_L0B].np : 0)) <= c3 ? 1 : 0;
  M.wpack = d + o_pm; M.n_slabs = (int)n_m;
  M.slab_off = reinterpret_cast<const uint32_t *>(d + o_mo); M.slab_bytes = reinterpret_cast<const uint32_t *>(d + o_mb);
  M.bias[0] = dfl + bm0; M.bias[1] = dfl + bm1; M.bias[2] = dfl + bm2; M.w6 = dfl + wm6; M.b6 = p6->bias[0];
  for (int i = 0; i < KMAX; ++i) M.w6c[i] = i < p6->in_dim ? p6->weight[i] : 0.0f;
  if (T.slab) cudaFree(T.slab);
  T.slab = d;
  T.entity = E;
  T.mlp3 = M;
  T.ready = 1;
  return 0;
}

void ebc_tc_release(ebc_sim *s) {
  if (s->d_trace) cudaFree(s->d_trace);
  s->d_trace = nullptr;
  for (int i = 0; i < 3; ++i) {
    if (s->tc[i].slab) cudaFree(s->tc[i].slab);
    s->tc[i].slab = nullptr;
    s->tc[i].ready = 0;
  }
}

int ebc_launch_value_tc(ebc_sim *s, int mode, const float *vin, int64_t n_states, const int32_t *row_count,
                        float *values, cudaStream_t stream) {
  if (mode == EBC_VALUE_TC_BF16) return launch_tc<1>(s, vin, n_states, row_count, values, stream);
  if (mode == EBC_VALUE_TC_FP16X2) return launch_tc<2>(s, vin, n_states, row_count, values, stream);
  return launch_tc<3>(s, vin, n_states, row_count, values, stream);
}
