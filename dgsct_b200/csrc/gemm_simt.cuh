// Generic fp32-accumulate SIMT GEMM with functor prologue/epilogue.
//
// This is the exact-fp32 path of the engine (io_dtype fp32: parity with the fp32 reference to ~1e-6) and
// the bring-up/debug twin of the tcgen05 kernels in gemm_tc.cuh (bf16 path).  One launch computes, for
// every batch b and group gi,   D[b,gi](m,n) = sum_k  ax(A[b,gi](m,k)) * B[b,gi](k,n)   and hands 4x4
// register micro-tiles to the epilogue functor.  Operands are addressed by element strides so the same
// kernel serves the NT (nn.Linear), NN and TN products of the adapter.
#pragma once
#include "common.cuh"

namespace dgsct {

struct GemmOperand {
  const void* p;
  long long s_row;    // A: stride of m ; B: stride of n
  long long s_k;      // stride along the contraction
  long long s_batch;  // per-frame stride (0 = shared)
  long long s_group;  // per-group offset
  int is_bf16;
};

struct GemmShape {
  int M, N, K, batch, groups;
  int ksplit;   // > 1: the contraction is split over this many CTAs per tile (grid.z); the epilogue must then ADD ATOMICALLY
                // into a zero-initialised (or accumulating) output.  Used for the skinny per-frame-vector products
                // (M = BT rows, K = C up to 1536) whose 64x64 tiling yields only a few dozen CTAs.
};

constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 32, SG_THREADS = 256;

struct AxIdentity {
  __device__ __forceinline__ float operator()(int, int, int, float v) const { return v; }
};

template <typename AX, typename EPI>
__global__ void __launch_bounds__(SG_THREADS) gemm_simt_kernel(GemmShape sh, GemmOperand A, GemmOperand B, AX ax, EPI epi) {
  __shared__ float As[SG_BK][SG_BM + 4];
  __shared__ float Bs[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int ks = sh.ksplit > 1 ? sh.ksplit : 1;
  const int z = blockIdx.z / ks, kz = blockIdx.z % ks;
  const int b = z / sh.groups, gi = z % sh.groups;
  const int kchunk = ((sh.K + ks - 1) / ks + SG_BK - 1) / SG_BK * SG_BK;
  const int k_begin = kz * kchunk, k_end = min(sh.K, k_begin + kchunk);
  const int m0 = blockIdx.y * SG_BM, n0 = blockIdx.x * SG_BN;
  const char* Ap = reinterpret_cast<const char*>(A.p);
  const char* Bp = reinterpret_cast<const char*>(B.p);
  const long long a_off = (long long)b * A.s_batch + (long long)gi * A.s_group;
  const long long b_off = (long long)b * B.s_batch + (long long)gi * B.s_group;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const bool a_kfast = (A.s_k == 1), b_kfast = (B.s_k == 1);
  for (int k0 = k_begin; k0 < k_end; k0 += SG_BK) {
    // ---- stage A tile (BM x BK) -> As[k][m]
#pragma unroll
    for (int i = 0; i < (SG_BM * SG_BK) / SG_THREADS; ++i) {
      int e = tid + i * SG_THREADS;
      int mm, kk;
      if (a_kfast) { kk = e % SG_BK; mm = e / SG_BK; } else { mm = e % SG_BM; kk = e / SG_BM; }
      int gm = m0 + mm, gk = k0 + kk;
      float v = 0.f;
      if (gm < sh.M && gk < k_end) {
        v = ld_act(Ap, (size_t)(a_off + (long long)gm * A.s_row + (long long)gk * A.s_k), A.is_bf16);
        v = ax(b, gi, gk, v);
      }
      As[kk][mm] = v;
    }
#pragma unroll
    for (int i = 0; i < (SG_BN * SG_BK) / SG_THREADS; ++i) {
      int e = tid + i * SG_THREADS;
      int nn, kk;
      if (b_kfast) { kk = e % SG_BK; nn = e / SG_BK; } else { nn = e % SG_BN; kk = e / SG_BN; }
      int gn = n0 + nn, gk = k0 + kk;
      float v = 0.f;
      if (gn < sh.N && gk < k_end)
        v = ld_act(Bp, (size_t)(b_off + (long long)gn * B.s_row + (long long)gk * B.s_k), B.is_bf16);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    __syncthreads();
  }
  epi(b, gi, m0 + ty * 4, n0 + tx * 4, acc, sh.M, sh.N);
}

template <typename AX, typename EPI>
inline cudaError_t launch_gemm_simt(const GemmShape& sh, const GemmOperand& A, const GemmOperand& B, AX ax, EPI epi,
                                    cudaStream_t st) {
  if (sh.M <= 0 || sh.N <= 0 || sh.batch <= 0) return cudaSuccess;
  dim3 grid(ceil_div(sh.N, SG_BN), ceil_div(sh.M, SG_BM), sh.batch * sh.groups * (sh.ksplit > 1 ? sh.ksplit : 1));
  gemm_simt_kernel<AX, EPI><<<grid, SG_THREADS, 0, st>>>(sh, A, B, ax, epi);
  return cudaGetLastError();
}

// ---- reductions usable from epilogues (thread layout: tx = tid&15 columns, ty = tid>>4 rows) -----
// sum over the 16 threads that share a row group (same ty): lanes differ in the low 4 bits
__device__ __forceinline__ float rowgroup_sum(float v) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// sum over the two row groups that live in one warp (lanes l and l^16 share tx)
__device__ __forceinline__ float colpair_sum(float v) { return v + __shfl_xor_sync(0xffffffffu, v, 16); }

}  // namespace dgsct
