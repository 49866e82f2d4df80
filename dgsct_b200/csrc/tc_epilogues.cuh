// Epilogue functors for gemm_tc_kernel.  A thread owns one accumulator row (TMEM lane) and receives it in
// chunks of 16 consecutive columns; `chunk` is always called by all 32 lanes of a warp (shuffles are legal).
#pragma once
#include "epilogues.cuh"
#include "gemm_tc.cuh"

namespace dgsct {
namespace tc {

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

__device__ __forceinline__ void store_row16(void* out, int out_bf16, size_t idx, const float (&v)[16], int nvalid) {
  if (out_bf16) {
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

// out[b, m, n] = act(alpha*acc + bias_m[m] + bias_n[n])   (+ per-frame column sums, + global column moments)
// transposed: out[b, n, m] = acc + add_t[n, m]            (latent-token update written token-major)
struct EpiStore {
  void* out; int out_bf16; long long ld, sb;
  const float* bias_m; const float* bias_n; float alpha; int act;
  float* colsum; int ncols_total;
  double* stat_sum; double* stat_sq;
  int transposed; const float* add_t; long long add_ld;
  float* out2; long long ld2, sb2;     // optional second (fp32) copy, same orientation as `out`
  __device__ __forceinline__ void begin(int, int, bool) {}
  __device__ __forceinline__ void end(int, int, bool) {}
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
    const int nvalid = N - n0 < 16 ? (N - n0 < 0 ? 0 : N - n0) : 16;
    const float bm = (bias_m && row_ok) ? bias_m[m] : 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float x = alpha * v[j] + bm;
      if (bias_n && j < nvalid) x += bias_n[n0 + j];
      v[j] = (row_ok && j < nvalid) ? apply_act(x, act) : 0.f;
    }
    if (row_ok && nvalid > 0) {
      if (!transposed) {
        if (out) store_row16(out, out_bf16, (size_t)((long long)b * sb + (long long)m * ld + n0), v, nvalid);
        if (out2) store_row16(out2, 0, (size_t)((long long)b * sb2 + (long long)m * ld2 + n0), v, nvalid);
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < nvalid) {
            const float x = v[j] + (add_t ? add_t[(long long)(n0 + j) * add_ld + m] : 0.f);
            if (out) st_act(out, (size_t)((long long)b * sb + (long long)(n0 + j) * ld + m), x, out_bf16);
            if (out2) out2[(long long)b * sb2 + (long long)(n0 + j) * ld2 + m] = x;
          }
      }
    }
    if (colsum || stat_sum) {
      float q[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) q[j] = v[j] * v[j];
      const float s = warp_colsum16(v);
      const int lane = threadIdx.x & 31;
      const int col = n0 + colsum16_owner_col(lane);
      const bool owner = ((lane & 1) == 0) && col < N;
      if (colsum && owner) atomicAdd(&colsum[(long long)b * ncols_total + col], s);
      if (stat_sum) {
        const float s2 = warp_colsum16(q);
        if (owner) { atomicAdd(&stat_sum[col], (double)s); atomicAdd(&stat_sq[col], (double)s2); }
      }
    }
  }
};

// out[b,m,n] = resid[b,m,n] + gate[0]*acc      (bf16 in / bf16 out)
struct EpiResidual {
  void* out; const void* resid; int is_bf16; long long ld, sb; const float* gate; float g;
  __device__ __forceinline__ void begin(int, int, bool) { g = gate[0]; }
  __device__ __forceinline__ void end(int, int, bool) {}
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
    const int nvalid = N - n0 < 16 ? (N - n0 < 0 ? 0 : N - n0) : 16;
    if (!row_ok || nvalid <= 0) return;
    const size_t idx = (size_t)((long long)b * sb + (long long)m * ld + n0);
    if (is_bf16 && nvalid == 16 && ((idx * 2) & 15) == 0) {
      const uint4* r = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(resid) + idx);
      uint4 ra = r[0], rb = r[1];
      const uint32_t w[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
        v[2 * j] = __bfloat162float(t.x) + g * v[2 * j];
        v[2 * j + 1] = __bfloat162float(t.y) + g * v[2 * j + 1];
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nvalid) v[j] = ld_act(resid, idx + j, is_bf16) + g * v[j];
    }
    store_row16(out, is_bf16, idx, v, nvalid);
  }
};

// colsum[b,n] += sum_m relu(acc + bias[n]) * mul[b,n]
struct EpiReluMulColsum {
  const float* bias; const float* mul; float* colsum; int ncols;
  __device__ __forceinline__ void begin(int, int, bool) {}
  __device__ __forceinline__ void end(int, int, bool) {}
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const bool ok = row_ok && (n0 + j < N);
      v[j] = ok ? fmaxf(v[j] + bias[n0 + j], 0.f) : 0.f;
    }
    const float s = warp_colsum16(v);
    const int lane = threadIdx.x & 31;
    const int col = n0 + colsum16_owner_col(lane);
    if (((lane & 1) == 0) && col < N) atomicAdd(&colsum[(long long)b * ncols + col], s * mul[(long long)b * ncols + col]);
  }
};

// rowsum[b,m] += sum_n relu(acc + bias[n]) * mul[b,n] * w[n]
struct EpiSpatialRowsum {
  const float* bias; const float* mul; const float* w; float* rowsum; int ncols; int nrows; float acc;
  __device__ __forceinline__ void begin(int, int, bool) { acc = 0.f; }
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (n0 + j < N) acc += fmaxf(v[j] + bias[n0 + j], 0.f) * (mul[(long long)b * ncols + n0 + j] * w[n0 + j]);
  }
  __device__ __forceinline__ void end(int b, int m, bool row_ok) {
    if (row_ok) atomicAdd(&rowsum[(long long)b * nrows + m], acc);
  }
};

// Q[b,m,0:32] = softmax_k(acc[0:32])    (N == 32, one N-block)
struct EpiRowSoftmax32 {
  void* out; int out_bf16; long long ld, sb; float lo[16];
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
    const float inv = 1.f / s;
#pragma unroll
    for (int j = 0; j < 16; ++j) { lo[j] *= inv; v[j] *= inv; }
    if (row_ok) {
      const size_t idx = (size_t)((long long)b * sb + (long long)m * ld);
      store_row16(out, out_bf16, idx, lo, 16);
      store_row16(out, out_bf16, idx + 16, v, 16);
    }
  }
};

}  // namespace tc
}  // namespace dgsct
