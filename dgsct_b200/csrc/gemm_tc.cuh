// bf16 tensor-core GEMM for sm_100a: TMA (cp.async.bulk.tensor) -> 128B/64B-swizzled shared memory ->
// tcgen05.mma (UMMA, cta_group::1, M=128) with the fp32 accumulator in TMEM -> tcgen05.ld epilogue.
//
//   D[b](m, n) = sum_k A[b](m, k) * B[b](n, k)          m < M, n < N, k < K, b < batch
//
// Either operand may be K-major (rows of K, e.g. activations (tokens, C) and nn.Linear weights (out, in)) or
// MN-major (rows of M/N, e.g. other[b] (No, Co) used as the B operand of the token resampler); the choice is a
// run-time field of the shared-memory and instruction descriptors.  Ragged edges are handled by TMA
// out-of-bounds zero fill on loads and by predication in the epilogue functor.
//
// One CTA computes one 128 x BN output tile: warp 0 (one lane) is the TMA producer, warp 1 (one lane) issues
// the MMAs, then all four warps drain their 32 TMEM lanes through the epilogue functor.  Several CTAs are
// resident per SM so the epilogue of one tile overlaps the main loop of another.
#pragma once
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace dgsct {
namespace tc {

constexpr int BM = 128;
constexpr int MAX_STAGES = 8;

struct Params {
  CUtensorMap mapA, mapB;
  int M, N, K, batch;
  int a_mn_major, b_mn_major;   // 0: K-major, 1: MN-major
  int a_sw, b_sw;               // swizzle span in bytes (128 or 64)
  int a_batched, b_batched;     // tensor-map coordinate 2 is the batch index (else 0)
  int BN;                       // tile N per frame (multiple of 16, <= 256)
  int fpack;                    // frames packed side by side along the instruction N (MN-major batched B, shared A)
  int kb_frames;                // >0: a tile accumulates over this many consecutive frames (K extended over the batch:
                                // weight gradients sum_b A_b . B_b^T); the epilogue then receives the frame-GROUP index
  int kflat;                    // weight gradients with both operands token-major (MN-major) and frames contiguous: the sum over
                                // frames IS a longer K, so (batch, K) is flattened into one K extent cut into `batch` equal chunks
                                // of K rows (no per-frame padding to BK: 36 / 144-token frames wasted 44 % / 25 % of every tile);
                                // the k coordinate of a load is then chunk * K + k0 and the tensor maps are 2-D
  int BK;                       // K elements per pipeline stage (64 or 32)
  int stages;
  int a_stage_bytes, b_stage_bytes;
  int a_boxes, b_boxes, a_box_bytes, b_box_bytes;
  uint32_t tmem_cols;
  int acc_bufs;                 // TMEM accumulators per CTA: 2 = epilogue overlaps the next tile's MMAs, 1 = single
  int ctas_per_sm;              // 1: one big-smem CTA per SM; 2: two half-smem CTAs per SM
  int b_resident;               // shared (unbatched) B small enough to stay in shared memory for the CTA's lifetime: it is
                                // loaded once (all n-blocks x k-blocks) and the ring streams A only
  int stage_out;                // short-K products: functors with chunk32() write their tile through a per-warp shared-memory
                                // staging buffer (this many bytes per epilogue warp, 0 = off) so that global stores / residual
                                // loads are row-contiguous 64-128 B segments
  int stats_cols;               // > 0: the functor reduces per-column statistics (BatchNorm moments): 2 x stats_cols floats of
                                // shared memory hold the CTA's partial sums, flushed with ONE global atomic per column per CTA
                                // (one per warp per 32 rows put 5760 same-address fp64 atomics on every column at stage 0)
  int rows_per_frame;           // > 0: A is (batch*rows_per_frame, K) contiguous and B is shared, so the frames are flattened
                                // into one M extent (no per-frame tile padding); the epilogue recovers (b, m) per row
  int* err_flag;                // device int set to non-zero when a bounded wait expires
  int debug_mode;               // bring-up only: 1 = epilogue skips the TMEM loads, 2 = skips loads and functor
};

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// bounded wait: ~2 s of SM clocks, then flag the error and let every role drain (no hang, no trap)
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int* err_flag) {
  if (mbar_try_wait(bar, parity)) return true;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      if (err_flag) { *reinterpret_cast<volatile int*>(err_flag) = 1; __threadfence_system(); }
      return false;
    }
  }
  return true;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&a)[16], float (&b)[16]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) { a[i] = __uint_as_float(r[i]); b[i] = __uint_as_float(r[16 + i]); }
}

// functors that must see a whole accumulator row in one thread declare `static constexpr bool kWholeRow = true`
template <typename T, typename = void>
struct epi_whole_row { static constexpr bool value = false; };
template <typename T>
struct epi_whole_row<T, decltype((void)T::kWholeRow)> { static constexpr bool value = T::kWholeRow; };

// functors that accept whole 32-column blocks and a staging buffer declare chunk32(...) (see tc_epilogues.cuh)
template <typename T, typename = void>
struct epi_has_chunk32 { static constexpr bool value = false; };
template <typename T>
struct epi_has_chunk32<T, decltype((void)T::kHasChunk32)> { static constexpr bool value = T::kHasChunk32; };
// functors that take both 16-column halves of a 32-column block at once declare pair(...): every global load of the block is
// issued before the first dependent use (read-modify-write epilogues were one exposed L2 round trip per 16 columns)
template <typename T, typename = void>
struct epi_has_pair { static constexpr bool value = false; };
template <typename T>
struct epi_has_pair<T, decltype((void)T::kHasPair)> { static constexpr bool value = T::kHasPair; };
// functors that keep per-CTA column statistics in shared memory (Params::stats_cols) declare kBlockStats and have a `blk`
// pointer member plus flush_block(blk, N, tid, nthreads)
template <typename T, typename = void>
struct epi_block_stats { static constexpr bool value = false; };
template <typename T>
struct epi_block_stats<T, decltype((void)T::kBlockStats)> { static constexpr bool value = T::kBlockStats; };
// functors whose chunk() tolerates a different frame index b in every lane (no warp-wide per-frame reduction)
template <typename T, typename = void>
struct epi_row_independent { static constexpr bool value = false; };
template <typename T>
struct epi_row_independent<T, decltype((void)T::kRowIndependent)> { static constexpr bool value = T::kRowIndependent; };

// staging bytes a functor needs per epilogue warp: 32 rows x 64 B (bf16 blocks) or x 128 B (fp32 blocks)
template <typename T, typename = void>
struct epi_stage_bytes { static constexpr int value = 0; };
template <typename T>
struct epi_stage_bytes<T, decltype((void)T::kStageBytes)> { static constexpr int value = T::kHasChunk32 ? T::kStageBytes : 0; };

// shared-memory matrix descriptor (tcgen05): start address, leading/stride byte offsets (16-B units),
// descriptor version 1, swizzle mode in bits [61,64)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, int sw_bytes) {
  const uint64_t layout = sw_bytes == 128 ? 2ull : (sw_bytes == 64 ? 4ull : (sw_bytes == 32 ? 6ull : 0ull));
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (layout << 61);
}
// instruction descriptor: bf16 x bf16 -> fp32, M=128
__host__ __device__ inline uint32_t make_idesc(int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// sum over the 32 lanes of each of 16 per-lane values in 16 shuffles (recursive halving); on return lane l
// holds the total of column (l >> 1) in v[0]
__device__ __forceinline__ float warp_colsum16(float (&v)[16]) {
  const int lane = threadIdx.x & 31;
  {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float send = up ? v[i] : v[i + 8];
      const float keep = up ? v[i + 8] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = up ? v[i] : v[i + 4];
      const float keep = up ? v[i + 4] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = up ? v[i] : v[i + 2];
      const float keep = up ? v[i + 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool up = lane & 2;
    const float send = up ? v[0] : v[1];
    const float keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return v[0];
}
// column owned by `lane` after warp_colsum16: bit4 -> +8, bit3 -> +4, bit2 -> +2, bit1 -> +1
__device__ __forceinline__ int colsum16_owner_col(int lane) {
  return ((lane & 16) ? 8 : 0) + ((lane & 8) ? 4 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0);
}

// ---- the kernel ----------------------------------------------------------------------------------------
// Persistent, warp-specialised: grid = #SMs (or fewer tiles); every CTA walks tiles t = blockIdx.x, +gridDim.x, ...
//   warp 0 (one lane)  TMA producer: runs ahead over (tile, k-block), bounded by the smem ring
//   warp 1 (one lane)  MMA issuer: accumulates tile i into TMEM accumulator (i & 1)
//   warps 2..5         epilogue: drain accumulator (i & 1) through the functor while the MMAs of tile i+1 run
// EPI interface:
//   void begin(int b, int m, bool row_ok)                       once per thread per tile (row m of batch b)
//   void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N)   16 consecutive columns, warp-convergent
//   void end(int b, int m, bool row_ok)
constexpr int EPI_WARPS = 8;   // two warps per TMEM lane group: a lone warp per scheduler cannot hide its own latencies
constexpr int THREADS = 64 + 32 * EPI_WARPS;

template <typename EPI>
__global__ void __launch_bounds__(THREADS, 2) gemm_tc_kernel(const __grid_constant__ Params p, EPI epi_in) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // layout: [ring: stages x (A [+ B])] [resident B: n_blocks x num_k x b_stage_bytes] [barriers] [output staging]
  const int stage_bytes = p.a_stage_bytes + (p.b_resident ? 0 : p.b_stage_bytes);
  const int num_k = (p.K + p.BK - 1) / p.BK;
  const int n_blocks = (p.N + p.BN - 1) / p.BN, m_blocks = (p.M + BM - 1) / BM;
  uint8_t* bres = smem + (size_t)p.stages * stage_bytes;
  const size_t bres_bytes = p.b_resident ? (size_t)n_blocks * num_k * p.b_stage_bytes : 0;
  uint64_t* bars = reinterpret_cast<uint64_t*>(bres + bres_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + MAX_STAGES;
  uint64_t* tmem_full_bar = bars + 2 * MAX_STAGES;        // [2]
  uint64_t* tmem_empty_bar = bars + 2 * MAX_STAGES + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);
  uint64_t* bres_bar = bars + 2 * MAX_STAGES + 5;
  uint8_t* out_stage = reinterpret_cast<uint8_t*>(bars) + 256;   // 16-byte aligned (bars is 1024-aligned)
  float* blk_stats = reinterpret_cast<float*>(out_stage + (size_t)EPI_WARPS * p.stage_out);
  for (int i = threadIdx.x; i < 2 * p.stats_cols; i += THREADS) blk_stats[i] = 0.f;

  const int tiles_per_batch = n_blocks * m_blocks;
  const int fpack = p.fpack;                              // frames per tile (1 unless packed)
  const int kbf = p.kb_frames;
  const int num_tiles = tiles_per_batch * (kbf > 0 ? (p.batch + kbf - 1) / kbf : (p.batch + fpack - 1) / fpack);
  const int BNI = p.BN * fpack;                           // instruction N

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.mapA);
    prefetch_tmap(&p.mapB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(smem_u32(&tmem_full_bar[a]), 1);
      mbar_init(smem_u32(&tmem_empty_bar[a]), EPI_WARPS);
    }
    mbar_init(smem_u32(bres_bar), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t acc_stride = p.acc_bufs == 2 ? p.tmem_cols / 2 : 0;   // columns between the two accumulators
  const int acc_bufs = p.acc_bufs;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      const int a_atom = p.a_sw / 2, b_atom = p.b_sw / 2;   // elements per swizzle row
      const uint32_t tx_bytes = (uint32_t)(p.a_boxes * p.a_box_bytes + (p.b_resident ? 0 : fpack * p.b_boxes * p.b_box_bytes));
      bool ok = true;
      if (p.b_resident) {                                   // the whole (shared) B operand, once
        const uint32_t rb = smem_u32(bres_bar);
        mbar_expect_tx(rb, (uint32_t)(n_blocks * num_k * p.b_boxes * p.b_box_bytes));
        for (int j = 0; j < n_blocks; ++j)
          for (int kb = 0; kb < num_k; ++kb) {
            const uint32_t dst = smem_u32(bres + (size_t)(j * num_k + kb) * p.b_stage_bytes);
            if (p.b_mn_major) {
              for (int q = 0; q < p.b_boxes; ++q)
                tma_load_3d(dst + q * p.b_box_bytes, &p.mapB, j * p.BN + q * b_atom, kb * p.BK, 0, rb);
            } else {
              tma_load_3d(dst, &p.mapB, kb * p.BK, j * p.BN, 0, rb);
            }
          }
      }
      // The single-thread role loops are on the critical path of every k-block (scripts/gemm_pipe_probe.py: with neither
      // loads nor MMAs the barrier ping-pong alone cost 0.3-0.6 us per k-block): no run-time division / modulo per k-block,
      // ring position and phase are counters, shared-memory addresses are base + stage * stride.
      const uint32_t smem_base = smem_u32(smem);
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      const bool one_group = (num_tiles == tiles_per_batch);
      int st_i = 0; uint32_t st_ph = 0;                      // ring slot and its phase
      for (int t = blockIdx.x; t < num_tiles && ok; t += gridDim.x) {
        int grp = 0, r = t;
        if (!one_group) { grp = t / tiles_per_batch; r = t - grp * tiles_per_batch; }
        int mi = r, ni = 0;
        if (n_blocks > 1) { mi = r / n_blocks; ni = r - mi * n_blocks; }
        const int bfirst = kbf > 0 ? grp * kbf : grp * fpack;
        const int nfr = kbf > 0 ? min(kbf, p.batch - bfirst) : 1;
        const int m0 = mi * BM, n0 = ni * p.BN;
        int b = bfirst, k0 = p.kflat ? bfirst * p.K : 0;      // (kflat: absolute row of the flattened K extent)
        const int kend = p.kflat ? k0 + p.K : p.K;
        const int total = nfr * num_k;
        for (int kk = 0; kk < total; ++kk) {
          ok = mbar_wait(empty0 + 8u * st_i, st_ph ^ 1, p.err_flag);
          if (!ok) break;
          const uint32_t fb = full0 + 8u * st_i;
          if (p.debug_mode & 8) {                          // bring-up: no loads, the stage is "full" at once
            mbar_arrive(fb);
          } else {
            mbar_expect_tx(fb, tx_bytes);
            const uint32_t sa = smem_base + (uint32_t)st_i * (uint32_t)stage_bytes;
            const uint32_t sb = sa + p.a_stage_bytes;
            if (p.a_mn_major) {
              for (int q = 0; q < p.a_boxes; ++q)
                tma_load_3d(sa + q * p.a_box_bytes, &p.mapA, m0 + q * a_atom, k0, p.a_batched ? b : 0, fb);
            } else {
              tma_load_3d(sa, &p.mapA, k0, m0, p.a_batched ? b : 0, fb);
            }
            if (p.b_resident) {
            } else if (p.b_mn_major) {
              for (int f = 0; f < fpack; ++f)       // a frame past the batch is out of bounds for TMA: zero fill
                for (int q = 0; q < p.b_boxes; ++q)
                  tma_load_3d(sb + (f * p.b_boxes + q) * p.b_box_bytes, &p.mapB, n0 + q * b_atom, k0, p.b_batched ? b + f : 0, fb);
            } else {
              tma_load_3d(sb, &p.mapB, k0, n0, p.b_batched ? b : 0, fb);
            }
          }
          if (++st_i == p.stages) { st_i = 0; st_ph ^= 1; }
          k0 += p.BK;
          if (k0 >= kend) { k0 = 0; ++b; }                    // next frame of the group (weight gradients)
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      const uint32_t idesc = make_idesc(BNI, p.a_mn_major, p.b_mn_major);
      const uint32_t a_lbo = p.a_mn_major ? (uint32_t)p.a_box_bytes : 16u;
      const uint32_t a_sbo = 8u * (uint32_t)p.a_sw;
      const uint32_t a_kstep = p.a_mn_major ? 16u * (uint32_t)p.a_sw : 32u;   // bytes per UMMA_K (16 elements)
      const uint32_t b_lbo = p.b_mn_major ? (uint32_t)p.b_box_bytes : 16u;
      const uint32_t b_sbo = 8u * (uint32_t)p.b_sw;
      const uint32_t b_kstep = p.b_mn_major ? 16u * (uint32_t)p.b_sw : 32u;
      const int ksteps = p.BK / 16;
      // descriptors: everything but the 14-bit start-address field is constant, and the field is linear in the byte
      // address (>> 4; shared memory is < 256 KB so it never carries out), so a stage / k-step is one 64-bit add
      const uint32_t smem_base = smem_u32(smem);
      const uint64_t adesc0 = make_smem_desc(smem_base, a_lbo, a_sbo, p.a_sw);
      const uint64_t bdesc0 = make_smem_desc(p.b_resident ? smem_u32(bres) : smem_base + (uint32_t)p.a_stage_bytes, b_lbo, b_sbo, p.b_sw);
      const uint64_t stage_step = (uint64_t)((uint32_t)stage_bytes >> 4);
      const uint64_t a_kadd = (uint64_t)(a_kstep >> 4), b_kadd = (uint64_t)(b_kstep >> 4);
      const uint64_t bres_step = (uint64_t)((uint32_t)p.b_stage_bytes >> 4);
      const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
      const uint32_t tfull0 = smem_u32(&tmem_full_bar[0]), tempty0 = smem_u32(&tmem_empty_bar[0]);
      const bool one_group = (num_tiles == tiles_per_batch);
      int st_i = 0; uint32_t st_ph = 0;
      int ti = 0;
      bool ok = true;
      if (p.b_resident) ok = mbar_wait(smem_u32(bres_bar), 0, p.err_flag);
      for (int t = blockIdx.x; t < num_tiles && ok; t += gridDim.x, ++ti) {
        const int acc = acc_bufs == 2 ? (ti & 1) : 0;
        const uint32_t acc_ph = acc_bufs == 2 ? ((ti >> 1) & 1) : (ti & 1);
        ok = mbar_wait(tempty0 + 8u * acc, acc_ph ^ 1, p.err_flag);   // epilogue drained this accumulator
        if (!ok) break;
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * acc_stride;
        int grp = 0, r = t;
        if (!one_group) { grp = t / tiles_per_batch; r = t - grp * tiles_per_batch; }
        int nblk = 0;
        if (n_blocks > 1) nblk = r - (r / n_blocks) * n_blocks;
        const int nfr = kbf > 0 ? min(kbf, p.batch - grp * kbf) : 1;
        const int total = nfr * num_k;
        int kb_in = 0;                                        // k-block inside the current frame (resident-B slot)
        const uint64_t bres_tile = bdesc0 + (uint64_t)(nblk * num_k) * bres_step;
        for (int kb = 0; kb < total; ++kb) {
          ok = mbar_wait(full0 + 8u * st_i, st_ph, p.err_flag);
          if (!ok) break;
          tc_fence_after();
          if (p.debug_mode & 4) {                             // bring-up: no MMAs, software release
            mbar_arrive(empty0 + 8u * st_i);
          } else {
            uint64_t ad = adesc0 + (uint64_t)st_i * stage_step;
            uint64_t bd = p.b_resident ? bres_tile + (uint64_t)kb_in * bres_step : bdesc0 + (uint64_t)st_i * stage_step;
            for (int k = 0; k < ksteps; ++k) {
              umma_bf16(tmem_d, ad, bd, idesc, (kb | k) ? 1u : 0u);
              ad += a_kadd; bd += b_kadd;
            }
            umma_commit(empty0 + 8u * st_i);   // frees the smem stage once these MMAs have read it
          }
          if (++st_i == p.stages) { st_i = 0; st_ph ^= 1; }
          if (++kb_in == num_k) kb_in = 0;
        }
        if (ok) {
          if (p.debug_mode & 4) mbar_arrive(tfull0 + 8u * acc);
          else umma_commit(tfull0 + 8u * acc);   // accumulator complete
        }
      }
    }
  } else {
    // ===== epilogue warps: warp w owns TMEM lanes [32 (w % 4), +32); the two warps of a lane group take
    // alternating 32-column blocks (functors that need a whole row in one thread set kWholeRow) =====
    EPI epi = epi_in;
    if constexpr (epi_block_stats<EPI>::value) epi.blk = p.stats_cols ? blk_stats : nullptr;
    const int lane_grp = warp & 3;
    const int half = (warp - 2) >> 2;
    const bool whole_row = epi_whole_row<EPI>::value;
    int ti = 0;
    bool ok = true;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++ti) {
      int grp = 0, r = t;
      if (num_tiles != tiles_per_batch) { grp = t / tiles_per_batch; r = t - grp * tiles_per_batch; }
      int mi = r, ni = 0;
      if (n_blocks > 1) { mi = r / n_blocks; ni = r - mi * n_blocks; }
      const int b0 = kbf > 0 ? grp : grp * fpack;
      const int m0 = mi * BM, n0 = ni * p.BN;
      const int acc = acc_bufs == 2 ? (ti & 1) : 0;
      const uint32_t acc_ph = acc_bufs == 2 ? ((ti >> 1) & 1) : (ti & 1);
      if (ok) ok = mbar_wait(smem_u32(&tmem_full_bar[acc]), acc_ph, p.err_flag);
      ok = __all_sync(0xffffffffu, ok);
      if (!ok) break;
      tc_fence_after();
      int m = m0 + lane_grp * 32 + lane, bq = b0;
      const bool row_ok = m < p.M;
      if (p.rows_per_frame > 0) {                         // (b, m) of this lane's row inside the flattened M extent
        bq = row_ok ? m / p.rows_per_frame : 0;           // rows past the end keep a valid frame index for loads
        if (row_ok) m -= bq * p.rows_per_frame;
      }
      const uint32_t lane_base = tmem_base + (uint32_t)acc * acc_stride + ((uint32_t)(lane_grp * 32) << 16);
      if ((p.debug_mode & 3) < 2 && !(whole_row && half == 1)) {
        epi.begin(bq, m, row_ok);
        const int c_begin = whole_row ? 0 : half * 32, c_step = whole_row ? 32 : 64;
        for (int c = c_begin; c < BNI; c += c_step) {
          float v0[16], v1[16];
          const bool two = (c + 32 <= BNI);
          if ((p.debug_mode & 3) == 0) {
            if (two) tmem_ld32(lane_base + (uint32_t)c, v0, v1);
            else tmem_ld16(lane_base + (uint32_t)c, v0);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) { v0[j] = (float)(c + j); v1[j] = (float)(c + 16 + j); }
          }
          if constexpr (epi_has_chunk32<EPI>::value) {
            if (p.stage_out && fpack == 1) {
              epi.chunk32(bq, m, n0 + c, v0, v1, two, row_ok, p.N, out_stage + (size_t)(warp - 2) * p.stage_out);
              continue;
            }
          }
          if constexpr (epi_has_pair<EPI>::value) {
            if (fpack == 1) {
              epi.pair(bq, m, n0 + c, v0, v1, two, row_ok, p.N);
              continue;
            }
          }
          if (fpack == 1) {
            epi.chunk(bq, m, n0 + c, v0, row_ok, p.N);
            if (two) epi.chunk(bq, m, n0 + c + 16, v1, row_ok, p.N);
          } else {                               // packed frames: 16-column chunks never straddle a frame (BN % 16 == 0)
            const int f0 = c / p.BN, f1 = (c + 16) / p.BN;
            epi.chunk(b0 + f0, m, c - f0 * p.BN, v0, row_ok && (b0 + f0 < p.batch), p.N);
            if (two) epi.chunk(b0 + f1, m, c + 16 - f1 * p.BN, v1, row_ok && (b0 + f1 < p.batch), p.N);
          }
        }
        epi.end(bq, m, row_ok);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[acc]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
  if constexpr (epi_block_stats<EPI>::value) {
    if (p.stats_cols) epi_in.flush_block(blk_stats, p.N, (int)threadIdx.x, THREADS);
  }
}

// ---- host side ---------------------------------------------------------------------------------------------
struct OperandDesc {
  const void* ptr;        // bf16
  int mn_major;           // 0: (rows = M|N, cols = K) K contiguous; 1: (rows = K, cols = M|N) MN contiguous
  long long ld;           // elements between consecutive rows
  long long batch_stride; // elements between batches; 0 => shared
  int rows, cols;         // logical extents of the stored matrix: rows x cols (cols contiguous)
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();
int* tc_err_flag_device_ptr();

// Fills `p` (tensor maps, tile recipe) for D = A . B^T.  Returns DGSCT_OK, or DGSCT_ERR_UNSUPPORTED when the
// operands cannot be described to TMA (alignment) so the caller can take the SIMT kernel instead.
int make_params(Params& p, const OperandDesc& A, const OperandDesc& B, int M, int N, int K, int batch, int BN_hint = 0,
                int stages_hint = 0, bool sum_over_batch = false, bool allow_flatten = false, int stage_bytes_per_warp = 0,
                int block_stats_cols = 0);
size_t smem_bytes(const Params& p);
int num_sms();
bool operands_ok(const OperandDesc& A, const OperandDesc& B, int K);

template <typename EPI>
inline cudaError_t launch(const Params& p, EPI epi, cudaStream_t st) {
  static PerDeviceOnce attr;                 // the > 48 KB dynamic shared-memory opt-in is a per-device function attribute
  const size_t smem = smem_bytes(p);
  if (attr.first()) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
  }
  const long long tiles = (long long)ceil_div(p.N, p.BN) * ceil_div(p.M, BM) *
                          (p.kb_frames > 0 ? ceil_div(p.batch, p.kb_frames) : ceil_div(p.batch, p.fpack));
  // (measured in round 2: capping a GEMM at 75 / 50 / 35 % of the CTA slots to leave SMs to the neighbour streams costs
  //  +3 / +6 / +13 ms per train step -- the small products are latency-bound and want every CTA they can get)
  const long long slots = (long long)num_sms() * p.ctas_per_sm;
  const int grid = (int)(tiles < slots ? tiles : slots);
  gemm_tc_kernel<EPI><<<grid, THREADS, smem, st>>>(p, epi);
  return cudaGetLastError();
}

}  // namespace tc
}  // namespace dgsct
