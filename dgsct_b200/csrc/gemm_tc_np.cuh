// Non-persistent variant of the tcgen05 GEMM (one CTA per tile, 128 threads, single TMEM accumulator, several
// CTAs per SM).  Kept for A/B measurements against the persistent kernel (dgsct_debug_gemm_bf16 use_tc=2).
#pragma once
#include "gemm_tc.cuh"

namespace dgsct {
namespace tc {

// ---- the kernel ----------------------------------------------------------------------------------------
// EPI interface:
//   void begin(int b, int m, bool row_ok)                       once per thread (row m of batch b)
//   void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N)   16 consecutive columns, warp-convergent
//   void end(int b, int m, bool row_ok)
template <typename EPI>
__global__ void __launch_bounds__(128) gemm_tc_np_kernel(const __grid_constant__ Params p, EPI epi) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stage_bytes = p.a_stage_bytes + p.b_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + MAX_STAGES;
  uint64_t* tmem_full_bar = bars + 2 * MAX_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 1);

  const int n_blk = blockIdx.x, m_blk = blockIdx.y, b = blockIdx.z;
  const int m0 = m_blk * BM, n0 = n_blk * p.BN;
  const int num_k = (p.K + p.BK - 1) / p.BK;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.mapA);
    prefetch_tmap(&p.mapB);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(tmem_full_bar), 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    const int a_atom = p.a_sw / 2, b_atom = p.b_sw / 2;   // elements per swizzle row
    bool ok = true;
    for (int kb = 0; kb < num_k && ok; ++kb) {
      const int s = kb % p.stages;
      const uint32_t ph = (kb / p.stages) & 1;
      ok = mbar_wait(smem_u32(&empty_bar[s]), ph ^ 1, p.err_flag);
      if (!ok) break;
      const uint32_t fb = smem_u32(&full_bar[s]);
      mbar_expect_tx(fb, (uint32_t)(p.a_boxes * p.a_box_bytes + p.b_boxes * p.b_box_bytes));
      const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
      const uint32_t sb = sa + p.a_stage_bytes;
      const int k0 = kb * p.BK;
      if (p.a_mn_major) {
        for (int j = 0; j < p.a_boxes; ++j)
          tma_load_3d(sa + j * p.a_box_bytes, &p.mapA, m0 + j * a_atom, k0, p.a_batched ? b : 0, fb);
      } else {
        tma_load_3d(sa, &p.mapA, k0, m0, p.a_batched ? b : 0, fb);
      }
      if (p.b_mn_major) {
        for (int j = 0; j < p.b_boxes; ++j)
          tma_load_3d(sb + j * p.b_box_bytes, &p.mapB, n0 + j * b_atom, k0, p.b_batched ? b : 0, fb);
      } else {
        tma_load_3d(sb, &p.mapB, k0, n0, p.b_batched ? b : 0, fb);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    const uint32_t idesc = make_idesc(p.BN, p.a_mn_major, p.b_mn_major);
    // per-operand descriptor recipe
    const uint32_t a_lbo = p.a_mn_major ? (uint32_t)p.a_box_bytes : 16u;
    const uint32_t a_sbo = 8u * (uint32_t)p.a_sw;
    const uint32_t a_kstep = p.a_mn_major ? 16u * (uint32_t)p.a_sw : 32u;   // bytes per UMMA_K (16 elements)
    const uint32_t b_lbo = p.b_mn_major ? (uint32_t)p.b_box_bytes : 16u;
    const uint32_t b_sbo = 8u * (uint32_t)p.b_sw;
    const uint32_t b_kstep = p.b_mn_major ? 16u * (uint32_t)p.b_sw : 32u;
    bool ok = true;
    for (int kb = 0; kb < num_k && ok; ++kb) {
      const int s = kb % p.stages;
      const uint32_t ph = (kb / p.stages) & 1;
      ok = mbar_wait(smem_u32(&full_bar[s]), ph, p.err_flag);
      if (!ok) break;
      tc_fence_after();
      const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
      const uint32_t sb = sa + p.a_stage_bytes;
      const int ksteps = p.BK / 16;
      for (int k = 0; k < ksteps; ++k) {
        const uint64_t ad = make_smem_desc(sa + k * a_kstep, a_lbo, a_sbo, p.a_sw);
        const uint64_t bd = make_smem_desc(sb + k * b_kstep, b_lbo, b_sbo, p.b_sw);
        umma_bf16(tmem_base, ad, bd, idesc, (kb | k) ? 1u : 0u);
      }
      umma_commit(smem_u32(&empty_bar[s]));   // frees the smem stage once these MMAs have read it
    }
    umma_commit(smem_u32(tmem_full_bar));     // accumulator complete
  }
  __syncwarp();

  // ===== epilogue: all 4 warps, warp w owns TMEM lanes [32w, 32w+32) =====
  const bool acc_ok = mbar_wait(smem_u32(tmem_full_bar), 0, p.err_flag);
  __syncwarp();
  tc_fence_after();
  const int m = m0 + warp * 32 + lane;
  const bool row_ok = acc_ok && (m < p.M);
  epi.begin(b, m, row_ok);
  const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
  for (int c = 0; c < p.BN; c += 16) {
    float v[16];
    tmem_ld16(lane_base + (uint32_t)c, v);
    epi.chunk(b, m, n0 + c, v, row_ok, p.N);
  }
  epi.end(b, m, row_ok);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}


template <typename EPI>
inline cudaError_t launch_np(Params p, EPI epi, cudaStream_t st, int smem_budget_kb) {
  static bool attr_set = false;
  if (p.fpack != 1) return cudaErrorInvalidValue;   // the A/B kernel does not implement frame packing
  const int stage_bytes = p.a_stage_bytes + p.b_stage_bytes;
  int stages = (smem_budget_kb * 1024) / stage_bytes;
  if (stages < 2) stages = 2;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  p.stages = stages;
  uint32_t cols = 32;
  while ((int)cols < p.BN) cols <<= 1;
  p.tmem_cols = cols;
  const size_t smem = smem_bytes(p);
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_np_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  dim3 grid(ceil_div(p.N, p.BN), ceil_div(p.M, BM), p.batch);
  gemm_tc_np_kernel<EPI><<<grid, 128, smem, st>>>(p, epi);
  return cudaGetLastError();
}

}  // namespace tc
}  // namespace dgsct
