// Epilogue functors for gemm_tc_kernel.  A thread owns one accumulator row (TMEM lane) and receives it in
// chunks of 16 consecutive columns; `chunk` is always called by all 32 lanes of a warp (shuffles are legal).
//
// Keep these bodies SMALL: the first version used one run-time-generic functor (activation switch, optional
// bias/transposed/tap paths per element) whose 16x-unrolled body was ~32 KB of SASS; the epilogue warps then ran
// at ~0.1 IPC on instruction-cache misses and every product was epilogue-bound (profiles/r01_gemm_ablation.txt).
// Everything optional is a template parameter or hoisted out of the per-element loop.  Register arrays are
// only ever indexed with compile-time constants (a run-time index silently moves them to local memory).
#pragma once
#include "epilogues.cuh"
#include "gemm_tc.cuh"

namespace dgsct {
namespace tc {

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

__device__ __forceinline__ void load16(const float* __restrict__ p, float (&d)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + j);
    d[4 * j] = t.x; d[4 * j + 1] = t.y; d[4 * j + 2] = t.z; d[4 * j + 3] = t.w;
  }
}
// 16 floats starting at p[0]; elements >= nvalid read as 0 (ragged right edge or unaligned base)
__device__ __forceinline__ void load16_guard(const float* __restrict__ p, float (&d)[16], int nvalid) {
  if (nvalid == 16 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) { load16(p, d); return; }
#pragma unroll
  for (int j = 0; j < 16; ++j) d[j] = j < nvalid ? p[j] : 0.f;   // static indices: keep d[] in registers
}

template <bool OUT_BF16>
__device__ __forceinline__ void store_row16(void* out, size_t idx, const float (&v)[16], int nvalid) {
  if (OUT_BF16) {
    bf16* o = reinterpret_cast<bf16*>(out) + idx;
    if (nvalid == 16 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
      uint4 a, b;
      a.x = pack_bf16x2(v[0], v[1]); a.y = pack_bf16x2(v[2], v[3]); a.z = pack_bf16x2(v[4], v[5]); a.w = pack_bf16x2(v[6], v[7]);
      b.x = pack_bf16x2(v[8], v[9]); b.y = pack_bf16x2(v[10], v[11]); b.z = pack_bf16x2(v[12], v[13]); b.w = pack_bf16x2(v[14], v[15]);
      reinterpret_cast<uint4*>(o)[0] = a;
      reinterpret_cast<uint4*>(o)[1] = b;
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nvalid) o[j] = __float2bfloat16_rn(v[j]);
    }
  } else {
    float* o = reinterpret_cast<float*>(out) + idx;
    if (nvalid == 16 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(o)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nvalid) o[j] = v[j];
    }
  }
}

__device__ __forceinline__ int nvalid16(int n0, int N) { const int r = N - n0; return r < 0 ? 0 : (r > 16 ? 16 : r); }

// out[b, m, n] = act(alpha*acc + bias_m[m] + bias_n[n])  (+ per-frame column sums, + global fp64 column moments)
template <bool OUT_BF16, int ACT, bool COLSUM, bool STATS>
struct EpiStore {
  void* out; long long ld, sb;
  const float* bias_m; const float* bias_n;
  float alpha;
  float* colsum; int ncols_total;
  double* stat_sum; double* stat_sq;
  __device__ __forceinline__ void begin(int, int, bool) {}
  __device__ __forceinline__ void end(int, int, bool) {}
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
    const int nvalid = nvalid16(n0, N);
    const float bm = (bias_m && row_ok) ? bias_m[m] : 0.f;
    if (bias_n) {
      float bn[16];
      load16_guard(bias_n + n0, bn, nvalid);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaf(alpha, v[j], bn[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] *= alpha;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float x = v[j] + bm;
      if (ACT == ACT_RELU) x = fmaxf(x, 0.f);
      if (ACT == ACT_SIGMOID) x = __frcp_rn(1.f + __expf(-x));
      v[j] = x;
    }
    if (out && row_ok && nvalid > 0) store_row16<OUT_BF16>(out, (size_t)((long long)b * sb + (long long)m * ld + n0), v, nvalid);
    if (COLSUM || STATS) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = (row_ok && j < nvalid) ? v[j] : 0.f;
      float q[16];
      if (STATS) {
#pragma unroll
        for (int j = 0; j < 16; ++j) q[j] = v[j] * v[j];
      }
      const float s = warp_colsum16(v);
      const int lane = threadIdx.x & 31;
      const int col = n0 + colsum16_owner_col(lane);
      const bool owner = ((lane & 1) == 0) && col < N;
      if (COLSUM && owner) atomicAdd(&colsum[(long long)b * ncols_total + col], s);
      if (STATS) {
        const float s2 = warp_colsum16(q);
        if (owner) { atomicAdd(&stat_sum[col], (double)s); atomicAdd(&stat_sq[col], (double)s2); }
      }
    }
  }
};

// latent-token update written token-major: accumulator is D(c, k); out[b, k, c] = acc + tok[k, c]  (bf16 + fp32)
struct EpiTokUpdate {
  bf16* out; float* out32; const float* tok; int C, K;
  __device__ __forceinline__ void begin(int, int, bool) {}
  __device__ __forceinline__ void end(int, int, bool) {}
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
    if (!row_ok) return;
    const int nvalid = nvalid16(n0, N);
    const long long base = (long long)b * K * C + m;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < nvalid) {
        const long long idx = base + (long long)(n0 + j) * C;      // lanes = consecutive c: coalesced
        const float x = v[j] + tok[(long long)(n0 + j) * C + m];
        out[idx] = __float2bfloat16_rn(x);
        out32[idx] = x;
      }
    }
  }
};

// out[b,m,n] = resid[b,m,n] + gate[0]*acc      (bf16 in / bf16 out)
struct EpiResidual {
  bf16* out; const bf16* resid; long long ld, sb; const float* gate; float g;
  __device__ __forceinline__ void begin(int, int, bool) { g = gate[0]; }
  __device__ __forceinline__ void end(int, int, bool) {}
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
    const int nvalid = nvalid16(n0, N);
    if (!row_ok || nvalid <= 0) return;
    const size_t idx = (size_t)((long long)b * sb + (long long)m * ld + n0);
    if (nvalid == 16 && ((idx * 2) & 15) == 0) {
      const uint4* r = reinterpret_cast<const uint4*>(resid + idx);
      const uint4 ra = r[0], rb = r[1];
      const uint32_t w[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
        v[2 * j] = fmaf(g, v[2 * j], __bfloat162float(t.x));
        v[2 * j + 1] = fmaf(g, v[2 * j + 1], __bfloat162float(t.y));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nvalid) v[j] = fmaf(g, v[j], __bfloat162float(resid[idx + j]));
    }
    store_row16<true>(out, idx, v, nvalid);
  }
};

// colsum[b,n] += mul[b,n] * sum_m relu(acc + bias[n])
struct EpiReluMulColsum {
  const float* bias; const float* mul; float* colsum; int ncols; int mul_ld;
  __device__ __forceinline__ void begin(int, int, bool) {}
  __device__ __forceinline__ void end(int, int, bool) {}
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
    const int nvalid = nvalid16(n0, N);
    float bn[16];
    load16_guard(bias + n0, bn, nvalid);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = (row_ok && j < nvalid) ? fmaxf(v[j] + bn[j], 0.f) : 0.f;
    const float s = warp_colsum16(v);
    const int lane = threadIdx.x & 31;
    const int col = n0 + colsum16_owner_col(lane);
    if (((lane & 1) == 0) && col < N) atomicAdd(&colsum[(long long)b * ncols + col], mul ? s * mul[(long long)b * mul_ld + col] : s);
  }
};

// rowsum[b,m] += sum_n relu(acc + bias[n]) * mul[b,n] * w[n]
struct EpiSpatialRowsum {
  const float* bias; const float* mul; const float* w; float* rowsum; int mul_ld; int nrows; float acc;
  __device__ __forceinline__ void begin(int, int, bool) { acc = 0.f; }
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
    const int nvalid = nvalid16(n0, N);
    float bn[16], f[16], ww[16];
    load16_guard(bias + n0, bn, nvalid);
    load16_guard(mul + (long long)b * mul_ld + n0, f, nvalid);
    load16_guard(w + n0, ww, nvalid);
#pragma unroll
    for (int j = 0; j < 16; ++j) acc = fmaf(fmaxf(v[j] + bn[j], 0.f), f[j] * ww[j], acc);
  }
  __device__ __forceinline__ void end(int b, int m, bool row_ok) {
    if (row_ok) atomicAdd(&rowsum[(long long)b * nrows + m], acc);
  }
};

// Q[b,m,0:32] = softmax_k(acc[0:32])    (N == 32, one N-block), bf16
struct EpiRowSoftmax32 {
  static constexpr bool kWholeRow = true;
  bf16* out; long long ld, sb; float lo[16];
  __device__ __forceinline__ void begin(int, int, bool) {}
  __device__ __forceinline__ void end(int, int, bool) {}
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
    if (n0 == 0) {
#pragma unroll
      for (int j = 0; j < 16; ++j) lo[j] = v[j];
      return;
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 16; ++j) mx = fmaxf(mx, fmaxf(lo[j], v[j]));
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) { lo[j] = __expf(lo[j] - mx); v[j] = __expf(v[j] - mx); s += lo[j] + v[j]; }
    const float inv = __frcp_rn(s);
#pragma unroll
    for (int j = 0; j < 16; ++j) { lo[j] *= inv; v[j] *= inv; }
    if (row_ok) {
      const size_t idx = (size_t)((long long)b * sb + (long long)m * ld);
      store_row16<true>(out, idx, lo, 16);
      store_row16<true>(out, idx + 16, v, 16);
    }
  }
};

// zr[b,m,n] = relu(acc*scale[n] + shift[n]) (bf16) ; optional fp32 tap of acc        (bn1 + ReLU folded)
struct EpiAffineRelu {
  bf16* zr; long long ldz; int rows_per_batch; const float* scale; const float* shift; float* ztap; int Cd;
  __device__ __forceinline__ void begin(int, int, bool) {}
  __device__ __forceinline__ void end(int, int, bool) {}
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
    const int nvalid = nvalid16(n0, N);
    if (!row_ok || nvalid <= 0) return;
    const long long row = (long long)b * rows_per_batch + m;
    if (ztap) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nvalid) ztap[row * Cd + n0 + j] = v[j];
    }
    float sc[16], sh[16];
    load16_guard(scale + n0, sc, nvalid);
    load16_guard(shift + n0, sh, nvalid);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(fmaf(v[j], sc[j], sh[j]), 0.f);
    store_row16<true>(zr, (size_t)(row * ldz + n0), v, nvalid);
  }
};

}  // namespace tc
}  // namespace dgsct
