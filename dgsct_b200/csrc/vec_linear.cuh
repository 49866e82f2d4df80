// Lean linear layer for the per-frame VECTOR chains of the adapter (channel gate, audio queries and their backward:
// reference DG-SCT/AVE/nets/net_trans.py:592-598, 602-603): out[b, n] = act(sum_k in[b, k] W(n, k) + bias[n]) with only BT rows
// (one per frame).  These products are a few MFLOP behind 1-5 MB of weights; on the persistent tcgen05 kernel they are one
// 128-row tile walked by a handful of CTAs through 12-48 dependent k-blocks (12-14 us per launch, ncu round 2).  Here every CTA
// owns EIGHT output columns, its eight warps split K, operand fragments come straight from global memory (no staging: each
// element is used once per CTA), partial sums meet in shared memory: N / 8 CTAs, a few microseconds.  Used where it wins (narrow
// rows, see vec_linear_ok).
#pragma once
#include "attn_fused.cuh"
#include "epilogues.cuh"

namespace dgsct {
namespace vl {

struct VecLinArgs {
  const bf16* in; int ld_in;        // (M, K) bf16, row pitch ld_in (even)
  const bf16* W; int ld_w;          // w_trans == 0: (N, K) rows of K ; w_trans == 1: (K, N) rows of N
  const float* bias;                // (N) fp32 or NULL
  float* out32; int ld32;           // fp32 output or NULL
  bf16* out16; int ld16;            // bf16 output or NULL
  int M, N, K, act, w_trans, accumulate;   // accumulate: out32 += (atomic, other kernels add into the same buffer)
};

constexpr int VL_MT = 8;            // row tiles of 16 per CTA pass (128 rows); grid.y walks further groups

__global__ void __launch_bounds__(256) vec_linear_kernel(const VecLinArgs a) {
  __shared__ float red[8][VL_MT * 4][32];         // [warp][accumulator register][lane]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * 8, m_base = blockIdx.y * (VL_MT * 16);
  const int mt_count = min(VL_MT, (a.M - m_base + 15) >> 4);
  const int ksteps = (a.K + 15) >> 4;
  const int kper = (ksteps + 7) >> 3;
  const int ks0 = warp * kper, ks1 = min(ksteps, ks0 + kper);
  float acc[VL_MT][4];
#pragma unroll
  for (int i = 0; i < VL_MT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int n = n0 + g;
  const bool n_ok = n < a.N;
  for (int ks = ks0; ks < ks1; ++ks) {
    const int k = ks * 16 + 2 * t;
    uint32_t b0 = 0, b1 = 0;
    if (n_ok) {
      if (!a.w_trans) {
        const bf16* w = a.W + (size_t)n * a.ld_w;
        if (k + 1 < a.K) b0 = *reinterpret_cast<const uint32_t*>(w + k);
        else if (k < a.K) b0 = (uint32_t)__bfloat16_as_ushort(w[k]);
        if (k + 9 < a.K) b1 = *reinterpret_cast<const uint32_t*>(w + k + 8);
        else if (k + 8 < a.K) b1 = (uint32_t)__bfloat16_as_ushort(w[k + 8]);
      } else {
        const bf16* w = a.W + n;
        const uint32_t w0 = k < a.K ? (uint32_t)__bfloat16_as_ushort(w[(size_t)k * a.ld_w]) : 0u;
        const uint32_t w1 = k + 1 < a.K ? (uint32_t)__bfloat16_as_ushort(w[(size_t)(k + 1) * a.ld_w]) : 0u;
        const uint32_t w2 = k + 8 < a.K ? (uint32_t)__bfloat16_as_ushort(w[(size_t)(k + 8) * a.ld_w]) : 0u;
        const uint32_t w3 = k + 9 < a.K ? (uint32_t)__bfloat16_as_ushort(w[(size_t)(k + 9) * a.ld_w]) : 0u;
        b0 = w0 | (w1 << 16);
        b1 = w2 | (w3 << 16);
      }
    }
#pragma unroll
    for (int mt = 0; mt < VL_MT; ++mt) {
      if (mt < mt_count) {
        const int r0 = m_base + mt * 16 + g, r1 = r0 + 8;
        uint32_t af[4] = {0u, 0u, 0u, 0u};
        const bf16* p0 = a.in + (size_t)r0 * a.ld_in + k;
        const bf16* p1 = a.in + (size_t)r1 * a.ld_in + k;
        if (r0 < a.M) {
          if (k + 1 < a.K) af[0] = *reinterpret_cast<const uint32_t*>(p0); else if (k < a.K) af[0] = (uint32_t)__bfloat16_as_ushort(p0[0]);
          if (k + 9 < a.K) af[2] = *reinterpret_cast<const uint32_t*>(p0 + 8); else if (k + 8 < a.K) af[2] = (uint32_t)__bfloat16_as_ushort(p0[8]);
        }
        if (r1 < a.M) {
          if (k + 1 < a.K) af[1] = *reinterpret_cast<const uint32_t*>(p1); else if (k < a.K) af[1] = (uint32_t)__bfloat16_as_ushort(p1[0]);
          if (k + 9 < a.K) af[3] = *reinterpret_cast<const uint32_t*>(p1 + 8); else if (k + 8 < a.K) af[3] = (uint32_t)__bfloat16_as_ushort(p1[8]);
        }
        fa::mma16816(acc[mt], af, b0, b1);
      }
    }
  }
#pragma unroll
  for (int mt = 0; mt < VL_MT; ++mt)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[warp][mt * 4 + j][lane] = acc[mt][j];
  __syncthreads();
  // warp w finishes row tile w: sum of the eight K slices, bias, activation, store (rows g / g+8, columns 2t / 2t+1 of the CTA's 8)
  if (warp < mt_count) {
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) s += red[w][warp * 4 + j][lane];
      v[j] = s;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = m_base + warp * 16 + g + (j >> 1) * 8, c = n0 + 2 * t + (j & 1);
      if (r < a.M && c < a.N) {
        float x = v[j] + (a.bias ? a.bias[c] : 0.f);
        if (a.act == ACT_RELU) x = fmaxf(x, 0.f);
        else if (a.act == ACT_SIGMOID) x = sigmoidf_(x);
        if (a.out32) {
          if (a.accumulate) atomicAdd(a.out32 + (size_t)r * a.ld32 + c, x);
          else a.out32[(size_t)r * a.ld32 + c] = x;
        }
        if (a.out16) a.out16[(size_t)r * a.ld16 + c] = __float2bfloat16_rn(x);
      }
    }
  }
}

// Measured on B200 (round 2, scripts/gpu_r2r.sh, per call at BT = 80): against the tcgen05 kernel the three forward linears
// take 45 vs 57 us at C = 192, 52 vs 60 at C = 384, but 70 vs 65 at C = 768 and 102 vs 76 at C = 1536 (every CTA re-reads the
// whole (BT, K) input), and the transposed-weight products of the backward lose at every width -- so: K-major weights with
// K <= 384 only.  DGSCT_VEC_LEAN=0 sends everything back to the tcgen05 kernel.
inline bool vec_linear_enabled() {
  static const bool on = [] { const char* e = getenv("DGSCT_VEC_LEAN"); return !(e && e[0] == '0'); }();
  return on;
}
inline bool vec_linear_ok(const VecLinArgs& a) {
  return vec_linear_enabled() && !a.w_trans && a.K <= 384 && a.M >= 1 && a.N >= 1 && a.K >= 1 && (a.ld_in % 2) == 0 &&
         ((uintptr_t)a.in & 3) == 0 && (a.ld_w % 2) == 0 && ((uintptr_t)a.W & 3) == 0;
}
inline cudaError_t launch_vec_linear(const VecLinArgs& a, cudaStream_t st) {
  vec_linear_kernel<<<dim3((a.N + 7) / 8, (a.M + VL_MT * 16 - 1) / (VL_MT * 16)), 256, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace vl
}  // namespace dgsct
