// Prologue / epilogue functors for gemm_simt_kernel.  Thread layout inside a 64x64 tile: each thread owns a
// 4x4 micro-tile at rows m..m+3, columns n..n+3 (n is local to the group; global column = gi*N + n).
#pragma once
#include "gemm_simt.cuh"

namespace dgsct {

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_SIGMOID = 2 };

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.f);
  if (act == ACT_SIGMOID) return 1.f / (1.f + expf(-v));
  return v;
}

// out[b, m, gi*N+n] = act(alpha*acc + bias_m[m] + bias_n[col] + add_mn[m, col]); optional per-frame column
// sums (colsum[b, col] += sum_m v) and global per-column first/second moments (BatchNorm batch statistics).
struct EpiBiasStore {
  void* out; int out_bf16; long long ld, sb;
  const float* bias_m; const float* bias_n; const float* add_mn; long long add_ld;
  float alpha; int act;
  float* colsum; int ncols_total;
  double* stat_sum; double* stat_sq;
  __device__ __forceinline__ void operator()(int b, int gi, int m, int n, float (&acc)[4][4], int M, int N) const {
    float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int gm = m + i;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ln = n + j, col = gi * N + ln;
        if (gm < M && ln < N) {
          float v = alpha * acc[i][j];
          if (bias_m) v += bias_m[gm];
          if (bias_n) v += bias_n[col];
          if (add_mn) v += add_mn[(long long)gm * add_ld + col];
          v = apply_act(v, act);
          if (out) st_act(out, (size_t)((long long)b * sb + (long long)gm * ld + col), v, out_bf16);
          cs[j] += v; cq[j] += v * v;
        }
      }
    }
    if (colsum || stat_sum) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float s = colpair_sum(cs[j]);
        const float q = colpair_sum(cq[j]);
        const int ln = n + j, col = gi * N + ln;
        if ((threadIdx.x & 16) == 0 && ln < N) {
          if (colsum) atomicAdd(&colsum[(long long)b * ncols_total + col], s);
          if (stat_sum) { atomicAdd(&stat_sum[col], (double)s); atomicAdd(&stat_sq[col], (double)q); }
        }
      }
    }
  }
};

// out[b,m,n] = resid[b,m,n] + gate[0]*acc        (x + gate_av * x_res, net_trans.py:589)
struct EpiResidual {
  void* out; const void* resid; int is_bf16; long long ld, sb; const float* gate;
  __device__ __forceinline__ void operator()(int b, int gi, int m, int n, float (&acc)[4][4], int M, int N) const {
    const float g = gate[0];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gm = m + i, gn = n + j;
        if (gm < M && gn < N) {
          const size_t idx = (size_t)((long long)b * sb + (long long)gm * ld + gn);
          st_act(out, idx, ld_act(resid, idx, is_bf16) + g * acc[i][j], is_bf16);
        }
      }
  }
};

// colsum[b,n] += sum_m relu(acc + bias[n]) * mul[b,n]     (channel-gate query, net_trans.py:594-595)
struct EpiReluMulColsum {
  const float* bias; const float* mul; float* colsum; int ncols; int mul_ld;
  __device__ __forceinline__ void operator()(int b, int gi, int m, int n, float (&acc)[4][4], int M, int N) const {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n + j;
      float s = 0.f;
      if (gn < N) {
        const float bj = bias[gn];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (m + i < M) s += fmaxf(acc[i][j] + bj, 0.f);
        if (mul) s *= mul[(long long)b * mul_ld + gn];
      }
      s = colpair_sum(s);
      if ((threadIdx.x & 16) == 0 && gn < N) atomicAdd(&colsum[(long long)b * ncols + gn], s);
    }
  }
};

// rowsum[b,m] += sum_n relu(acc + bias[n]) * mul[b,n] * w[n]   (spatial-gate logit, net_trans.py:602-605)
struct EpiSpatialRowsum {
  const float* bias; const float* mul; const float* w; float* rowsum; int mul_ld; int nrows;
  __device__ __forceinline__ void operator()(int b, int gi, int m, int n, float (&acc)[4][4], int M, int N) const {
    float f[4], bj[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n + j;
      const bool ok = gn < N;
      f[j] = ok ? mul[(long long)b * mul_ld + gn] * w[gn] : 0.f;
      bj[j] = ok ? bias[gn] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) s += fmaxf(acc[i][j] + bj[j], 0.f) * f[j];
      s = rowgroup_sum(s);
      if ((threadIdx.x & 15) == 0 && m + i < M) atomicAdd(&rowsum[(long long)b * nrows + m + i], s);
    }
  }
};

// A-operand prologues
struct AxOnePlusCol {   // a * (1 + m[b, k])                  (net_trans.py:598)
  const float* mvec; int ncols;
  __device__ __forceinline__ float operator()(int b, int, int k, float v) const { return v * (1.f + mvec[(long long)b * ncols + k]); }
};
struct AxAffineRelu {   // relu(a * scale[gi*Kg + k] + shift[...])   (bn1 + ReLU, net_trans.py:636-638)
  const float* scale; const float* shift; int kg;
  __device__ __forceinline__ float operator()(int, int gi, int k, float v) const {
    const int c = gi * kg + k;
    return fmaxf(fmaf(v, scale[c], shift[c]), 0.f);
  }
};

}  // namespace dgsct
