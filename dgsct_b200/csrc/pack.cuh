// bf16 operand copies of the fp32 master weights for the tensor-core path ("packed weights").
#pragma once
#include "common.cuh"

namespace dgsct {

// byte offsets of each packed matrix inside the packed-weights buffer; `ld*` are row pitches in elements
struct PackPlan {
  size_t res, fc, tok, v1, v2, down, up, total;
  size_t cat1, bott, catt, cat1_bias;    // [a1_w; a2_w] (C+H, C), bottleneck (H, C), channel gate (C, H), [a1_b; a2_b] fp32
  int ld_res, ld_fc, ld_c, ld_up, ld_h;
};

inline int pad8(int x) { return (x + 7) / 8 * 8; }

inline PackPlan make_pack_plan(const dgsct_desc& d) {
  PackPlan p{};
  const size_t N = d.N, C = d.C, K = d.K, Cd = d.C / d.r, H = d.C / 2;
  p.ld_res = pad8(d.No); p.ld_fc = pad8(d.Co); p.ld_c = pad8(d.C); p.ld_up = pad8((int)Cd); p.ld_h = pad8((int)H);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  p.res = take(N * p.ld_res * 2);
  p.fc = take(C * p.ld_fc * 2);
  p.tok = take(K * p.ld_c * 2);
  p.v1 = take(C * p.ld_c * 2);
  p.v2 = take(H * p.ld_c * 2);
  p.down = take(Cd * p.ld_c * 2);
  p.up = take(C * p.ld_up * 2);
  p.cat1 = take((C + H) * p.ld_c * 2);
  p.bott = take(H * p.ld_c * 2);
  p.catt = take(C * p.ld_h * 2);
  p.cat1_bias = take((C + H) * 4);
  p.total = off;
  return p;
}

// dst (rows, ld_dst) bf16 <- src (rows, cols) fp32, pad columns zeroed
__global__ void pack_rows_kernel(const float* __restrict__ src, int rows, int cols, bf16* __restrict__ dst, int ld_dst) {
  const size_t n = (size_t)rows * ld_dst;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / ld_dst), c = (int)(i % ld_dst);
    dst[i] = __float2bfloat16_rn(c < cols ? src[(size_t)r * cols + c] : 0.f);
  }
}

// grouped 1x1 conv weight (out, in/g) -> dense block-diagonal (out, ld_dst) bf16
__global__ void pack_grouped_kernel(const float* __restrict__ src, int out_ch, int in_ch, int groups, bf16* __restrict__ dst,
                                    int ld_dst) {
  const size_t n = (size_t)out_ch * ld_dst;
  const int ipg = in_ch / groups, opg = out_ch / groups;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int o = (int)(i / ld_dst), c = (int)(i % ld_dst);
    float v = 0.f;
    if (c < in_ch && c / ipg == o / opg) v = src[(size_t)o * ipg + (c - (o / opg) * ipg)];
    dst[i] = __float2bfloat16_rn(v);
  }
}

// dst bf16 = alpha * src fp32 (per-frame vectors feeding the small tensor-core linears)
__global__ void scale_to_bf16_kernel(const float* __restrict__ src, float alpha, bf16* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16_rn(alpha * src[i]);
}

// q[b,c] = aq1[b,c] * r1sum[b,c] * inv_n  (fp32 and/or bf16 copies)      (net_trans.py:595)
__global__ void channel_q_kernel(const float* __restrict__ aq1, int aq_ld, const float* __restrict__ r1sum, float inv_n,
                                 float* __restrict__ q32, bf16* __restrict__ q16, int BT, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BT * C) return;
  const int b = i / C, c = i % C;
  const float v = aq1[(size_t)b * aq_ld + c] * r1sum[i] * inv_n;
  if (q32) q32[i] = v;
  if (q16) q16[i] = __float2bfloat16_rn(v);
}

__global__ void concat2_f32_kernel(const float* __restrict__ a, int na, const float* __restrict__ b, int nb, float* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < na) dst[i] = a[i];
  else if (i < na + nb) dst[i] = b[i - na];
}

// out (rows, C) bf16 = x1 * (1 + m[b, c])          (net_trans.py:598, A operand of the spatial-gate GEMM)
__global__ void scale_cols_kernel(const bf16* __restrict__ x1, const float* __restrict__ m, bf16* __restrict__ out,
                                  long long rows, int N, int C) {
  const long long total2 = rows * C / 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total2; i += (long long)gridDim.x * blockDim.x) {
    const long long e = i * 2;
    const long long row = e / C;
    const int c = (int)(e % C);
    const int b = (int)(row / N);
    const __nv_bfloat162 v = reinterpret_cast<const __nv_bfloat162*>(x1)[i];
    const float2 mm = *reinterpret_cast<const float2*>(m + (size_t)b * C + c);
    reinterpret_cast<__nv_bfloat162*>(out)[i] =
        __floats2bfloat162_rn(__bfloat162float(v.x) * (1.f + mm.x), __bfloat162float(v.y) * (1.f + mm.y));
  }
}

// zr (rows, ldz) bf16 = relu(z * scale[c] + shift[c])      (bn1 + ReLU, net_trans.py:636-638)
__global__ void bn_relu_kernel(const float* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                               bf16* __restrict__ zr, long long rows, int Cd, int ldz) {
  const long long total = rows * Cd;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / Cd;
    const int c = (int)(i % Cd);
    zr[row * ldz + c] = __float2bfloat16_rn(fmaxf(fmaf(z[i], scale[c], shift[c]), 0.f));
  }
}

// softmax over tokens (dim 1) of L (BT, N, K) fp32 -> P (BT, N, K) bf16; one block per frame, coalesced
__global__ void col_softmax_bf16_kernel(const float* __restrict__ L, bf16* __restrict__ P, int N, int K) {
  extern __shared__ float sm[];          // [groups][K] partials
  const int b = blockIdx.x;
  const int groups = blockDim.x / K;
  const int k = threadIdx.x % K, g = threadIdx.x / K;
  const bool active = g < groups;
  const float* base = L + (size_t)b * N * K;
  float mx = -INFINITY;
  if (active)
    for (int n = g; n < N; n += groups) mx = fmaxf(mx, base[(size_t)n * K + k]);
  if (active) sm[g * K + k] = mx;
  __syncthreads();
  if (active) for (int j = 0; j < groups; ++j) mx = fmaxf(mx, sm[j * K + k]);
  __syncthreads();
  float s = 0.f;
  if (active)
    for (int n = g; n < N; n += groups) s += __expf(base[(size_t)n * K + k] - mx);
  if (active) sm[g * K + k] = s;
  __syncthreads();
  if (active) { s = 0.f; for (int j = 0; j < groups; ++j) s += sm[j * K + k]; }
  const float inv = 1.f / s;
  bf16* pb = P + (size_t)b * N * K;
  if (active)
    for (int n = g; n < N; n += groups) pb[(size_t)n * K + k] = __float2bfloat16_rn(__expf(base[(size_t)n * K + k] - mx) * inv);
}

}  // namespace dgsct
