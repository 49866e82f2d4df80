// Non-GEMM kernels of the adapter: softmaxes, gate combine + LayerNorm, BatchNorm finalize, output LN.
// All statistics are fp32 (BatchNorm moments fp64); reductions use warp shuffles.
#pragma once
#include "common.cuh"

namespace dgsct {

// softmax over n (dim 1) of L (BT, N, K) fp32, in place; one block per (k, b)   (net_trans.py:576)
__global__ void col_softmax_kernel(float* __restrict__ L, int N, int K) {
  __shared__ float red[32];
  const int k = blockIdx.x, b = blockIdx.y;
  float* base = L + (size_t)b * N * K + k;
  float mx = -INFINITY;
  for (int n = threadIdx.x; n < N; n += blockDim.x) mx = fmaxf(mx, base[(size_t)n * K]);
  mx = block_max(mx, red);
  float s = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) s += expf(base[(size_t)n * K] - mx);
  s = block_sum(s, red);
  const float inv = 1.f / s;
  for (int n = threadIdx.x; n < N; n += blockDim.x) base[(size_t)n * K] = expf(base[(size_t)n * K] - mx) * inv;
}

// softmax over the last dim of L (rows, K) fp32, in place; one warp per row   (net_trans.py:585)
__global__ void row_softmax_kernel(float* __restrict__ L, long long rows, int K) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float* p = L + row * K;
  float mx = -INFINITY;
  for (int k = lane; k < K; k += 32) mx = fmaxf(mx, p[k]);
  mx = warp_max(mx);
  float s = 0.f;
  for (int k = lane; k < K; k += 32) s += expf(p[k] - mx);
  s = warp_sum(s);
  const float inv = 1.f / s;
  for (int k = lane; k < K; k += 32) p[k] = expf(p[k] - mx) * inv;
}

// per frame: s = ssum + bias; sig = sigmoid(s); smap = softmax_n(tanh(s))   (net_trans.py:605-608)
__global__ void spatial_map_kernel(const float* __restrict__ ssum, const float* __restrict__ s_bias, int N,
                                   float* __restrict__ sig, void* __restrict__ smap, int smap_bf16,
                                   float* __restrict__ s_tap) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const float bias = s_bias ? s_bias[0] : 0.f;
  const float* sp = ssum + (size_t)b * N;
  float z = 0.f;   // tanh is bounded in [-1,1]: exp needs no max subtraction
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float s = sp[n] + bias;
    z += expf(tanhf(s));
  }
  z = block_sum(z, red);
  const float inv = 1.f / z;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float s = sp[n] + bias;
    sig[(size_t)b * N + n] = 1.f / (1.f + expf(-s));
    st_act(smap, (size_t)b * N + n, expf(tanhf(s)) * inv, smap_bf16);
    if (s_tap) s_tap[(size_t)b * N + n] = s;
  }
}

struct GateLnArgs {
  const void* x1; int in_bf16;           // (rows, C)
  void* y; int out_bf16;                 // (rows, C)
  const float* m;                        // (BT, C)   channel gate
  const float* sig;                      // (rows)    sigmoid spatial gate
  const float* tau;                      // (BT) or NULL
  const float* ln_w; const float* ln_b;  // NULL => no LayerNorm
  float alpha, beta, gamma, eps;
  int N, C; long long rows;
  float* y_tap;
};

// y = LN_before( x1 * (alpha*m + beta*sig + gamma*tau + 1 - alpha) ); one warp per token row, the row is
// staged in shared memory (C floats per warp).   (net_trans.py:611-613, 626-627; few-shot :537-549)
__global__ void gate_ln_kernel(GateLnArgs a) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= a.rows) return;
  float* buf = smem + (size_t)warp * a.C;
  const int b = (int)(row / a.N);
  const float base = a.beta * a.sig[row] + (a.tau ? a.gamma * a.tau[b] : 0.f) + 1.f - a.alpha;
  const float* mrow = a.m + (size_t)b * a.C;
  float sum = 0.f;
  for (int c = lane; c < a.C; c += 32) {
    const float v = ld_act(a.x1, (size_t)row * a.C + c, a.in_bf16) * (a.alpha * mrow[c] + base);
    buf[c] = v;
    sum += v;
  }
  if (a.ln_w) {
    sum = warp_sum(sum);
    const float mean = sum / a.C;
    float sq = 0.f;
    for (int c = lane; c < a.C; c += 32) { const float d = buf[c] - mean; sq += d * d; }
    sq = warp_sum(sq);
    const float rstd = rsqrtf(sq / a.C + a.eps);
    for (int c = lane; c < a.C; c += 32) {
      const float v = (buf[c] - mean) * rstd * a.ln_w[c] + a.ln_b[c];
      st_act(a.y, (size_t)row * a.C + c, v, a.out_bf16);
      if (a.y_tap) a.y_tap[(size_t)row * a.C + c] = v;
    }
  } else {
    for (int c = lane; c < a.C; c += 32) {
      st_act(a.y, (size_t)row * a.C + c, buf[c], a.out_bf16);
      if (a.y_tap) a.y_tap[(size_t)row * a.C + c] = buf[c];
    }
  }
}

// BatchNorm statistics -> per-channel affine (scale, shift); updates running stats in train mode
// (net_trans.py:636,643: momentum 0.1, unbiased running variance).
__global__ void bn_finalize_kernel(int C, int use_bn, int train, const double* __restrict__ sum,
                                   const double* __restrict__ sq, double count, const float* __restrict__ w,
                                   const float* __restrict__ bias, float* __restrict__ rm, float* __restrict__ rv,
                                   float eps, float momentum, float* __restrict__ scale, float* __restrict__ shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (!use_bn) { scale[c] = 1.f; shift[c] = 0.f; return; }
  double mean, var;
  if (train) {
    mean = sum[c] / count;
    var = sq[c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
    rm[c] = (float)((1.0 - momentum) * rm[c] + momentum * mean);
    rv[c] = (float)((1.0 - momentum) * rv[c] + momentum * unb);
  } else {
    mean = rm[c];
    var = rv[c];
  }
  const double sc = (double)w[c] / sqrt(var + (double)eps);
  scale[c] = (float)sc;
  shift[c] = (float)((double)bias[c] - mean * sc);
}

struct OutLnArgs {
  const void* u; int in_bf16;            // (rows, C) up-sampler output (pre-BN)
  void* out; int out_bf16;
  const float* scale; const float* shift;  // bn2 folded affine
  const float* ln_w; const float* ln_b;    // NULL => no ln_post
  const float* gate;                       // NULL => no gate
  int gate_before_ln; float eps; int C; long long rows;
};

// out = gate * LN_post(bn2(u))   (AVE order, net_trans.py:667-672)  or  LN_post(gate * bn2(u))  (AVS order)
__global__ void out_ln_kernel(OutLnArgs a) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= a.rows) return;
  float* buf = smem + (size_t)warp * a.C;
  const float g = a.gate ? a.gate[0] : 1.f;
  const float gpre = a.gate_before_ln ? g : 1.f, gpost = a.gate_before_ln ? 1.f : g;
  float sum = 0.f;
  for (int c = lane; c < a.C; c += 32) {
    const float v = fmaf(ld_act(a.u, (size_t)row * a.C + c, a.in_bf16), a.scale[c], a.shift[c]) * gpre;
    buf[c] = v;
    sum += v;
  }
  if (a.ln_w) {
    sum = warp_sum(sum);
    const float mean = sum / a.C;
    float sq = 0.f;
    for (int c = lane; c < a.C; c += 32) { const float d = buf[c] - mean; sq += d * d; }
    sq = warp_sum(sq);
    const float rstd = rsqrtf(sq / a.C + a.eps);
    for (int c = lane; c < a.C; c += 32)
      st_act(a.out, (size_t)row * a.C + c, ((buf[c] - mean) * rstd * a.ln_w[c] + a.ln_b[c]) * gpost, a.out_bf16);
  } else {
    for (int c = lane; c < a.C; c += 32) st_act(a.out, (size_t)row * a.C + c, buf[c] * gpost, a.out_bf16);
  }
}

// dtype-converting copy into an fp32 tap buffer
__global__ void tap_copy_kernel(const void* __restrict__ src, int is_bf16, float* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = ld_act(src, i, is_bf16);
}

}  // namespace dgsct
