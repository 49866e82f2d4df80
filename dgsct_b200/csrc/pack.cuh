// bf16 operand copies of the fp32 master weights for the tensor-core path ("packed weights").
#pragma once
#include "common.cuh"

namespace dgsct {

// byte offsets of each packed matrix inside the packed-weights buffer; `ld*` are row pitches in elements
struct PackPlan {
  size_t res, fc, tok, v1, v2, down, up, total;
  size_t cat1, bott, catt, cat1_bias;    // [a1_w; a2_w] (C+H, C), bottleneck (H, C), channel gate (C, H), [a1_b; a2_b] fp32
  size_t fc_s;                           // row sums of the bf16 fc.weight copy (C) fp32: rank-1 bias factor of the re-associated resampler
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
  p.fc_s = take(C * 4);
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

// ---- all packed operands of one adapter in ONE launch ------------------------------------------------------------------
// Training re-packs after every optimizer step; as twelve launches per adapter that was 576 tiny kernels at the head of the 48
// forward calls.  grid = (blocks, segments): blockIdx.y selects the segment, the x blocks stride over its elements.
struct PackSeg {
  const float* src; const float* src2;   // src2: second source of a concatenation (kind 2)
  void* dst;
  int kind;                              // 0: rows fp32 -> bf16 (pad columns zeroed); 1: grouped 1x1 conv -> dense block-diagonal bf16;
                                         // 2: fp32 concat [src (rows) ; src2 (cols)] -> fp32
                                         // 3: dst[r] = sum_c bf16(src[r, c])  (fp32; a warp per row)
  int rows, cols, ld, groups;
};
constexpr int PACK_MAX_SEGS = 13;
struct PackTable { PackSeg seg[PACK_MAX_SEGS]; int n; };

__global__ void __launch_bounds__(256) pack_all_kernel(const PackTable t) {
  const PackSeg& g = t.seg[blockIdx.y];
  const size_t stride = (size_t)gridDim.x * blockDim.x, i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g.kind == 0) {
    if (!g.src) return;
    bf16* dst = reinterpret_cast<bf16*>(g.dst);
    const size_t n = (size_t)g.rows * g.ld;
    for (size_t i = i0; i < n; i += stride) {
      const int r = (int)(i / g.ld), c = (int)(i % g.ld);
      dst[i] = __float2bfloat16_rn(c < g.cols ? g.src[(size_t)r * g.cols + c] : 0.f);
    }
  } else if (g.kind == 1) {
    bf16* dst = reinterpret_cast<bf16*>(g.dst);
    const size_t n = (size_t)g.rows * g.ld;
    const int ipg = g.cols / g.groups, opg = g.rows / g.groups;
    for (size_t i = i0; i < n; i += stride) {
      const int o = (int)(i / g.ld), c = (int)(i % g.ld);
      float v = 0.f;
      if (c < g.cols && c / ipg == o / opg) v = g.src[(size_t)o * ipg + (c - (o / opg) * ipg)];
      dst[i] = __float2bfloat16_rn(v);
    }
  } else if (g.kind == 2) {
    float* dst = reinterpret_cast<float*>(g.dst);
    for (size_t i = i0; i < (size_t)(g.rows + g.cols); i += stride) dst[i] = i < (size_t)g.rows ? g.src[i] : g.src2[i - g.rows];
  } else {
    if (!g.src) return;
    float* dst = reinterpret_cast<float*>(g.dst);
    const int lane = threadIdx.x & 31;
    for (size_t r = i0 >> 5; r < (size_t)g.rows; r += stride >> 5) {
      float acc = 0.f;
      for (int c = lane; c < g.cols; c += 32) acc += __bfloat162float(__float2bfloat16_rn(g.src[r * g.cols + c]));
      acc = warp_sum(acc);
      if (lane == 0) dst[r] = acc;
    }
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

// s[c] = sum_co W[c, co] over the bf16 operand copy of fc.weight (C, ld): the column factor of the rank-1 bias term
// b_res[n] * s[c] when the resampler is re-associated as W_res (O W_fc^T) (DGSCT_FC_FIRST); one warp per row
__global__ void fc_rowsum_kernel(const bf16* __restrict__ W, int C, int Co, int ld, float* __restrict__ s) {
  const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (c >= C) return;
  float acc = 0.f;
  for (int j = lane; j < Co; j += 32) acc += __bfloat162float(W[(size_t)c * ld + j]);
  acc = warp_sum(acc);
  if (lane == 0) s[c] = acc;
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
// 16-byte version (C % 8 == 0): two vectors in flight per thread
__global__ void __launch_bounds__(256) scale_cols_vec_kernel(const bf16* __restrict__ x1, const float* __restrict__ m,
                                                             bf16* __restrict__ out, long long rows, int N, int C) {
  const int nvec = C >> 3;
  const long long total = rows * nvec;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 2 * stride) {
    const long long i1 = i0 + stride;
    const bool two = i1 < total;
    const uint4 u0 = __ldg(reinterpret_cast<const uint4*>(x1) + i0);
    const uint4 u1 = two ? __ldg(reinterpret_cast<const uint4*>(x1) + i1) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (h == 1 && !two) break;
      const long long i = h ? i1 : i0;
      const uint4 u = h ? u1 : u0;
      const long long row = i / nvec;
      const int c = (int)(i - row * nvec) * 8;
      const int b = (int)(row / N);
      const float4 ma = __ldg(reinterpret_cast<const float4*>(m + (size_t)b * C + c));
      const float4 mb = __ldg(reinterpret_cast<const float4*>(m + (size_t)b * C + c) + 1);
      const float mm[8] = {ma.x, ma.y, ma.z, ma.w, mb.x, mb.y, mb.z, mb.w};
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
      uint32_t o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&w[q]);
        const __nv_bfloat162 r = __floats2bfloat162_rn(__bfloat162float(t.x) * (1.f + mm[2 * q]),
                                                       __bfloat162float(t.y) * (1.f + mm[2 * q + 1]));
        o[q] = *reinterpret_cast<const uint32_t*>(&r);
      }
      reinterpret_cast<uint4*>(out)[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}
inline void launch_scale_cols(const bf16* x1, const float* m, bf16* out, long long rows, int N, int C, cudaStream_t st) {
  if (C % 8 == 0) {
    const long long total = rows * (C / 8);
    long long blocks = (total + 511) / 512;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    if (blocks < 1) blocks = 1;
    scale_cols_vec_kernel<<<(unsigned)blocks, 256, 0, st>>>(x1, m, out, rows, N, C);
  } else {
    scale_cols_kernel<<<2048, 256, 0, st>>>(x1, m, out, rows, N, C);
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

// bn1 finalize + bn1 + ReLU in one launch (train-mode forward): every block derives the per-channel affine from the fp64 batch
// moments into shared memory (C/r <= a few hundred channels: cheaper than waiting for a one-block kernel to publish them),
// block 0 also publishes scale / shift / mean / rstd for the backward and updates the running statistics.
// Same arithmetic as bn_finalize_kernel (rowwise.cuh).
__global__ void __launch_bounds__(256) bn_finalize_relu_kernel(const float* __restrict__ z, bf16* __restrict__ zr, long long rows, int Cd, int ldz,
                                                               int use_bn, int train, const double* __restrict__ sum, const double* __restrict__ sq,
                                                               double count, const float* __restrict__ w, const float* __restrict__ bias,
                                                               float* __restrict__ rm, float* __restrict__ rv, float eps, float momentum,
                                                               float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_out,
                                                               float* __restrict__ rstd_out, int update_running) {
  extern __shared__ float bn_sm[];                 // [2][Cd]
  float* s_sc = bn_sm; float* s_sh = bn_sm + Cd;
  for (int c = threadIdx.x; c < Cd; c += blockDim.x) {
    float sc = 1.f, sh = 0.f, mo = 0.f, ro = 1.f;
    if (use_bn) {
      double mean, var;
      if (train) {
        mean = sum[c] / count;
        var = sq[c] / count - mean * mean;
        if (var < 0.0) var = 0.0;
        if (update_running && blockIdx.x == 0) {
          const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
          rm[c] = (float)((1.0 - momentum) * rm[c] + momentum * mean);
          rv[c] = (float)((1.0 - momentum) * rv[c] + momentum * unb);
        }
      } else {
        mean = rm[c];
        var = rv[c];
      }
      const double rstd = 1.0 / sqrt(var + (double)eps);
      const double scd = (double)w[c] * rstd;
      sc = (float)scd; sh = (float)((double)bias[c] - mean * scd); mo = (float)mean; ro = (float)rstd;
    }
    s_sc[c] = sc; s_sh[c] = sh;
    if (blockIdx.x == 0) { scale[c] = sc; shift[c] = sh; mean_out[c] = mo; rstd_out[c] = ro; }
  }
  __syncthreads();
  const long long total = rows * Cd;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / Cd;
    const int c = (int)(i % Cd);
    zr[row * ldz + c] = __float2bfloat16_rn(fmaxf(fmaf(z[i], s_sc[c], s_sh[c]), 0.f));
  }
}

// softmax over tokens (dim 1) of L (BT, N, K) fp32 -> P (BT, N, K) bf16                       (net_trans.py:576)
// Two kernels over a (row-chunk, frame) grid so that a frame is not confined to one block (BT = 80 frames would leave
// half of the 148 SMs idle): (1) per-chunk column maxima and exp-sums, (2) combine the <= 32 partials of the frame
// (redundantly per block -- a few hundred flops) and write the normalised probabilities.
constexpr int CSM_MAX_CHUNKS = 32;
__host__ __device__ inline int csm_chunk_rows(int N) {           // rows per chunk: >= 64, at most CSM_MAX_CHUNKS chunks
  int rows = (N + CSM_MAX_CHUNKS - 1) / CSM_MAX_CHUNKS;
  if (rows < 64) rows = 64;
  return rows;
}
__host__ __device__ inline int csm_num_chunks(int N) { const int r = csm_chunk_rows(N); return (N + r - 1) / r; }

// part (BT, S, 2, K): [.,.,0,k] = max over the chunk, [.,.,1,k] = sum exp(l - max)
__global__ void __launch_bounds__(256) col_softmax_partial_kernel(const float* __restrict__ L, float* __restrict__ part, int N, int K) {
  extern __shared__ float sm[];          // [2][groups][K]
  const int s = blockIdx.x, b = blockIdx.y, S = gridDim.x;
  const int groups = blockDim.x / K;
  const int k = threadIdx.x % K, g = threadIdx.x / K;
  const bool active = g < groups;
  const int rows = csm_chunk_rows(N);
  const int n0 = s * rows, n1 = min(N, n0 + rows);
  const float* base = L + (size_t)b * N * K;
  float mx = -INFINITY, sum = 0.f;
  if (active) {
    for (int n = n0 + g; n < n1; n += groups) {                  // online max / sum
      const float l = base[(size_t)n * K + k];
      if (l > mx) { sum *= __expf(mx - l); mx = l; }
      sum += __expf(l - mx);
    }
    sm[g * K + k] = mx;
    sm[(groups + g) * K + k] = sum;
  }
  __syncthreads();
  if (active && g == 0) {
    float M = -INFINITY;
    for (int j = 0; j < groups; ++j) M = fmaxf(M, sm[j * K + k]);
    float Z = 0.f;
    for (int j = 0; j < groups; ++j) {
      const float mj = sm[j * K + k];
      if (mj > -INFINITY) Z += sm[(groups + j) * K + k] * __expf(mj - M);
    }
    float* o = part + ((size_t)b * S + s) * 2 * K;
    o[k] = M;
    o[K + k] = Z;
  }
}
// (P has row pitch ldp >= K: pitches are padded to 16 bytes so that TMA can describe P for any K, e.g. AVQA's K = 2)
__global__ void __launch_bounds__(256) col_softmax_apply_kernel(const float* __restrict__ L, const float* __restrict__ part,
                                                                bf16* __restrict__ P, int N, int K, int ldp) {
  const int s = blockIdx.x, b = blockIdx.y, S = gridDim.x;
  const int groups = blockDim.x / K;
  const int k = threadIdx.x % K, g = threadIdx.x / K;
  if (g >= groups) return;
  const float* pb = part + (size_t)b * S * 2 * K;
  float M = -INFINITY;
  for (int j = 0; j < S; ++j) M = fmaxf(M, pb[(size_t)j * 2 * K + k]);
  float Z = 0.f;
  for (int j = 0; j < S; ++j) {
    const float mj = pb[(size_t)j * 2 * K + k];
    if (mj > -INFINITY) Z += pb[(size_t)j * 2 * K + K + k] * __expf(mj - M);
  }
  const float inv = 1.f / Z;
  const int rows = csm_chunk_rows(N);
  const int n0 = s * rows, n1 = min(N, n0 + rows);
  const float* base = L + (size_t)b * N * K;
  bf16* out = P + (size_t)b * N * ldp;
  for (int n = n0 + g; n < n1; n += groups)
    out[(size_t)n * ldp + k] = __float2bfloat16_rn(__expf(base[(size_t)n * K + k] - M) * inv);
}

}  // namespace dgsct
