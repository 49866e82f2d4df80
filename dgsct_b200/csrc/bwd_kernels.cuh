// Element-wise / reduction kernels of the adapter backward (reference: autograd of net_trans.py:552-674).
// Activation-sized gradients use the activation dtype (fp32 or bf16); every statistic and every parameter
// gradient is accumulated in fp32.  Column reductions are staged in shared memory per block and flushed with
// one global atomicAdd per column per block.
#pragma once
#include <stdlib.h>
#include "common.cuh"

namespace dgsct {

// ---- (1) output stage: out = gate * LN_post(B2)  [AVE]   or   out = LN_post(gate * B2)  [AVS],  B2 = Uo*sc2+sh2 ----
struct BwdOutArgs {
  const void* dout; int dout_bf16;          // (R, C)
  const float* uo;                           // (R, C) fp32 pre-BN2 up-sampler output
  const float* scale; const float* shift;    // folded bn2 affine used by the forward
  const float* ln_w; const float* ln_b;      // NULL => no ln_post
  const float* gate;                         // NULL => no gate
  int gate_before_ln; float eps; int C; long long rows;
  void* db2; int db2_bf16;                   // (R, C) gradient w.r.t. B2
  float* d_lnw; float* d_lnb; float* d_gate; // accumulated (may be NULL)
};

__global__ void bwd_out_kernel(BwdOutArgs a) {
  extern __shared__ float sm[];               // [2][C] column partials + [warps][C] row staging (x-hat) + [warps][C] (dy)
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* col_w = sm; float* col_b = sm + a.C;
  float* xh = sm + 2 * a.C + (size_t)warp * 2 * a.C; float* dy = xh + a.C;
  for (int c = threadIdx.x; c < 2 * a.C; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  const float g = a.gate ? a.gate[0] : 1.f;
  float gate_acc = 0.f;
  for (long long row = (long long)blockIdx.x * warps + warp; row < a.rows; row += (long long)gridDim.x * warps) {
    const float gpre = a.gate_before_ln ? g : 1.f, gpost = a.gate_before_ln ? 1.f : g;
    float sum = 0.f;
    for (int c = lane; c < a.C; c += 32) {
      const float b2 = fmaf(a.uo[row * a.C + c], a.scale[c], a.shift[c]);
      const float v = b2 * gpre;            // LN input
      xh[c] = v; sum += v;
      dy[c] = ld_act(a.dout, (size_t)(row * a.C + c), a.dout_bf16);
    }
    if (a.ln_w) {
      const float mean = warp_sum(sum) / a.C;
      float sq = 0.f;
      for (int c = lane; c < a.C; c += 32) { const float dlt = xh[c] - mean; sq += dlt * dlt; }
      const float rstd = rsqrtf(warp_sum(sq) / a.C + a.eps);
      // LN output L = xhat*w + b ; out = gpost * L
      float s1 = 0.f, s2 = 0.f;
      for (int c = lane; c < a.C; c += 32) {
        const float xhat = (xh[c] - mean) * rstd;
        const float dL = dy[c] * gpost;
        if (!a.gate_before_ln && a.gate) gate_acc += dy[c] * fmaf(xhat, a.ln_w[c], a.ln_b[c]);   // AVE order: d_gate = sum dOut * L
        atomicAdd(&col_w[c], dL * xhat);
        atomicAdd(&col_b[c], dL);
        const float dxh = dL * a.ln_w[c];
        xh[c] = xhat; dy[c] = dxh;
        s1 += dxh; s2 += dxh * xhat;
      }
      s1 = warp_sum(s1) / a.C; s2 = warp_sum(s2) / a.C;
      for (int c = lane; c < a.C; c += 32) {
        const float dv = rstd * (dy[c] - s1 - xh[c] * s2);      // d(LN input)
        // AVS order: LN input = g * B2  ->  dB2 = g * dv ; d_gate += dv * B2
        if (a.gate_before_ln && a.gate) {
          const float b2 = fmaf(a.uo[row * a.C + c], a.scale[c], a.shift[c]);
          gate_acc += dv * b2;
        }
        st_act(a.db2, (size_t)(row * a.C + c), dv * gpre, a.db2_bf16);
      }
    } else {
      // no ln_post: out = g * B2 in both orders
      for (int c = lane; c < a.C; c += 32) {
        const float b2 = fmaf(a.uo[row * a.C + c], a.scale[c], a.shift[c]);
        if (a.gate) gate_acc += dy[c] * b2;
        st_act(a.db2, (size_t)(row * a.C + c), dy[c] * g, a.db2_bf16);
      }
    }
  }
  __syncthreads();
  if (a.ln_w) {
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
      if (a.d_lnw) atomicAdd(&a.d_lnw[c], col_w[c]);
      if (a.d_lnb) atomicAdd(&a.d_lnb[c], col_b[c]);
    }
  }
  if (a.gate && a.d_gate) {
    gate_acc = warp_sum(gate_acc);
    if (lane == 0 && gate_acc != 0.f) atomicAdd(a.d_gate, gate_acc);
  }
}

// ---- column reductions over rows: s1[c] += sum_r a[r,c] ; s2[c] += sum_r a[r,c]*b[r,c] (b optional) ----------
__global__ void col_reduce_kernel(const void* __restrict__ a, int a_bf16, long long a_ld, const void* __restrict__ b, int b_bf16,
                                  long long b_ld, long long rows, int C, float* __restrict__ s1, float* __restrict__ s2) {
  __shared__ float r1[8][33], r2[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float p1 = 0.f, p2 = 0.f;
  if (c < C)
    for (long long r = (long long)blockIdx.y * 8 + ry; r < rows; r += (long long)gridDim.y * 8) {
      const float va = ld_act(a, (size_t)(r * a_ld + c), a_bf16);
      p1 += va;
      if (b) p2 += va * ld_act(b, (size_t)(r * b_ld + c), b_bf16);
    }
  r1[ry][cx] = p1; r2[ry][cx] = p2;
  __syncthreads();
  if (ry == 0 && c < C) {
    for (int j = 1; j < 8; ++j) { p1 += r1[j][cx]; p2 += r2[j][cx]; }
    if (s1) atomicAdd(&s1[c], p1);
    if (s2 && b) atomicAdd(&s2[c], p2);
  }
}

// ---- BatchNorm backward coefficients: dx = A*dy + B*x + Cc ; also dW, dB -------------------------------------
// train: xhat = (x-mean)*rstd, dw = sum dy*xhat, db = sum dy, dx = w*rstd*(dy - db/R - xhat*dw/R)
// eval (running stats): dx = w*rstd*dy.  No BN: A = 1.
__global__ void bn_bwd_coeffs_kernel(int C, int use_bn, int train, const float* __restrict__ sum_dy, const float* __restrict__ sum_dyx,
                                     const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ w,
                                     double count, float* __restrict__ A, float* __restrict__ B, float* __restrict__ Cc,
                                     float* __restrict__ dW, float* __restrict__ dB) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (!use_bn) { A[c] = 1.f; B[c] = 0.f; Cc[c] = 0.f; return; }
  const float mu = mean[c], rs = rstd[c];
  const float db = sum_dy[c];
  const float dw = (sum_dyx[c] - mu * db) * rs;       // sum dy*(x-mu)*rstd
  if (dW) atomicAdd(&dW[c], dw);
  if (dB) atomicAdd(&dB[c], db);
  const float wr = w[c] * rs;
  if (train) {
    const float inv = (float)(1.0 / count);
    A[c] = wr;
    B[c] = -wr * rs * dw * inv;
    Cc[c] = -wr * db * inv + wr * rs * mu * dw * inv;
  } else {
    A[c] = wr; B[c] = 0.f; Cc[c] = 0.f;
  }
}

// g[r,c] = A[c]*g[r,c] + B[c]*x[r,c] + Cc[c]   (in place on g), optional ReLU mask from z = x*sc+sh > 0 applied FIRST:
// used as (a) bn2 backward on dB2, (b) relu+bn1 backward on dZr.
__global__ void affine_combine_kernel(void* __restrict__ g, int g_bf16, long long g_ld, const float* __restrict__ x, long long x_ld,
                                      const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ Cc,
                                      long long rows, int C) {
  const long long total = rows * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C; const int c = (int)(i % C);
    const float v = ld_act(g, (size_t)(r * g_ld + c), g_bf16);
    st_act(g, (size_t)(r * g_ld + c), fmaf(A[c], v, fmaf(B[c], x[r * x_ld + c], Cc[c])), g_bf16);
  }
}
// g[r,c] *= 1[z[r,c]*sc[c]+sh[c] > 0]      (ReLU backward with the pre-activation recomputed from Z)
__global__ void relu_mask_kernel(void* __restrict__ g, int g_bf16, long long g_ld, const float* __restrict__ z, long long z_ld,
                                 const float* __restrict__ sc, const float* __restrict__ sh, long long rows, int C) {
  const long long total = rows * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / C; const int c = (int)(i % C);
    if (fmaf(z[r * z_ld + c], sc[c], sh[c]) <= 0.f) st_act(g, (size_t)(r * g_ld + c), 0.f, g_bf16);
  }
}

// ---- (4) LN_before + gate combine backward --------------------------------------------------------------------
struct BwdGateLnArgs {
  const void* dy; int dy_bf16;               // (R, C) gradient w.r.t. Y (LN_before output)
  const void* x1; int x1_bf16;               // (R, C)
  const float* m; const float* sig; const float* tau;   // (BT,C), (R), (BT)|NULL
  const float* ln_w;                         // NULL => no ln_before
  float alpha, beta, gamma, eps; int N, C; long long rows;
  void* dx1; int dx1_bf16;                   // (R, C)  written:  dYg * gmul
  float* dm_acc;                             // (BT, C) += alpha * sum_n dYg*X1
  float* dsig;                               // (R)     =  beta * sum_c dYg*X1
  float* dtau;                               // (BT)    += gamma * sum_{n,c} dYg*X1   (NULL if unused)
  float* d_lnw; float* d_lnb;
};

__global__ void bwd_gate_ln_kernel(BwdGateLnArgs a) {
  extern __shared__ float sm[];               // [2][C] column partials + per warp [2][C]
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* col_w = sm; float* col_b = sm + a.C;
  float* yg = sm + 2 * a.C + (size_t)warp * 2 * a.C; float* dv = yg + a.C;
  for (int c = threadIdx.x; c < 2 * a.C; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  for (long long row = (long long)blockIdx.x * warps + warp; row < a.rows; row += (long long)gridDim.x * warps) {
    const int b = (int)(row / a.N);
    const float base = a.beta * a.sig[row] + (a.tau ? a.gamma * a.tau[b] : 0.f) + 1.f - a.alpha;
    const float* mrow = a.m + (size_t)b * a.C;
    float sum = 0.f;
    for (int c = lane; c < a.C; c += 32) {
      const float v = ld_act(a.x1, (size_t)(row * a.C + c), a.x1_bf16) * fmaf(a.alpha, mrow[c], base);
      yg[c] = v; sum += v;
      dv[c] = ld_act(a.dy, (size_t)(row * a.C + c), a.dy_bf16);
    }
    if (a.ln_w) {
      const float mean = warp_sum(sum) / a.C;
      float sq = 0.f;
      for (int c = lane; c < a.C; c += 32) { const float dlt = yg[c] - mean; sq += dlt * dlt; }
      const float rstd = rsqrtf(warp_sum(sq) / a.C + a.eps);
      float s1 = 0.f, s2 = 0.f;
      for (int c = lane; c < a.C; c += 32) {
        const float xhat = (yg[c] - mean) * rstd;
        atomicAdd(&col_w[c], dv[c] * xhat);
        atomicAdd(&col_b[c], dv[c]);
        const float dxh = dv[c] * a.ln_w[c];
        yg[c] = xhat; dv[c] = dxh;
        s1 += dxh; s2 += dxh * xhat;
      }
      s1 = warp_sum(s1) / a.C; s2 = warp_sum(s2) / a.C;
      for (int c = lane; c < a.C; c += 32) dv[c] = rstd * (dv[c] - s1 - yg[c] * s2);   // dYg
    }
    float rowdot = 0.f;
    for (int c = lane; c < a.C; c += 32) {
      const float x1 = ld_act(a.x1, (size_t)(row * a.C + c), a.x1_bf16);
      const float dgm = dv[c] * x1;
      rowdot += dgm;
      atomicAdd(&a.dm_acc[(size_t)b * a.C + c], a.alpha * dgm);
      st_act(a.dx1, (size_t)(row * a.C + c), dv[c] * fmaf(a.alpha, mrow[c], base), a.dx1_bf16);
    }
    rowdot = warp_sum(rowdot);
    if (lane == 0) {
      a.dsig[row] = a.beta * rowdot;
      if (a.dtau) atomicAdd(&a.dtau[b], a.gamma * rowdot);
    }
  }
  __syncthreads();
  if (a.ln_w)
    for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
      if (a.d_lnw) atomicAdd(&a.d_lnw[c], col_w[c]);
      if (a.d_lnb) atomicAdd(&a.d_lnb[c], col_b[c]);
    }
}

// ---- (5) spatial map backward: ds[b,n] from dSmap (softmax(tanh s)) and dsig (sigmoid s) ; d_bs += sum ds -------
__global__ void bwd_spatial_map_kernel(const float* __restrict__ ssum, const float* __restrict__ s_bias, int N,
                                       const void* __restrict__ dsmap, int dsmap_bf16, const float* __restrict__ dsig,
                                       float* __restrict__ ds, float* __restrict__ d_sbias) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const float bias = s_bias[0];
  const float* sp = ssum + (size_t)b * N;
  float z = 0.f, dot = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float e = expf(tanhf(sp[n] + bias));
    z += e;
    if (dsmap) dot += e * ld_act(dsmap, (size_t)b * N + n, dsmap_bf16);
  }
  z = block_sum(z, red);
  dot = block_sum(dot, red) / z;            // sum_n y*dy
  float acc = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) {
    const float s = sp[n] + bias, t = tanhf(s), sg = 1.f / (1.f + expf(-s));
    float g = dsig[(size_t)b * N + n] * sg * (1.f - sg);
    if (dsmap) {
      const float y = expf(t) / z;
      g += y * (ld_act(dsmap, (size_t)b * N + n, dsmap_bf16) - dot) * (1.f - t * t);
    }
    ds[(size_t)b * N + n] = g;
    acc += g;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0 && d_sbias) atomicAdd(d_sbias, acc);
}

// z2 (R,H) holds the pre-bias spatial query; in place -> dz2 = ds*ws*aq2*1[z2+b>0]; reductions:
//   d_ws[j] += sum ds*aq2*relu ; d_aq2[b,j] += ws[j]*sum_n ds*relu ; d_b2[j] += sum dz2
__global__ void bwd_spatial_mask_kernel(void* __restrict__ z2, int bf, int H, int N, const float* __restrict__ bias2,
                                        const float* __restrict__ ds, const float* __restrict__ ws, const float* __restrict__ aq2,
                                        int aq_ld, float* __restrict__ d_ws, float* __restrict__ d_aq2, int daq_ld,
                                        float* __restrict__ d_b2) {
  // grid (ceil(H/32), BT, row_splits) ; block (32, 8)
  __shared__ float r_ws[8][33], r_aq[8][33], r_b[8][33];
  const int jx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + jx, b = blockIdx.y;
  float p_ws = 0.f, p_aq = 0.f, p_b = 0.f;
  if (j < H) {
    const float w = ws[j], a2 = aq2[(size_t)b * aq_ld + j], bj = bias2[j];
    for (int n = blockIdx.z * 8 + ry; n < N; n += gridDim.z * 8) {
      const size_t idx = ((size_t)b * N + n) * H + j;
      const float z = ld_act(z2, idx, bf) + bj;
      const float r = fmaxf(z, 0.f), d = ds[(size_t)b * N + n];
      const float dz = z > 0.f ? d * w * a2 : 0.f;
      st_act(z2, idx, dz, bf);
      p_ws += d * a2 * r; p_aq += d * r; p_b += dz;
    }
    p_aq *= w;
  }
  r_ws[ry][jx] = p_ws; r_aq[ry][jx] = p_aq; r_b[ry][jx] = p_b;
  __syncthreads();
  if (ry == 0 && j < H) {
    for (int t = 1; t < 8; ++t) { p_ws += r_ws[t][jx]; p_aq += r_aq[t][jx]; p_b += r_b[t][jx]; }
    atomicAdd(&d_ws[j], p_ws);
    atomicAdd(&d_aq2[(size_t)b * daq_ld + j], p_aq);
    atomicAdd(&d_b2[j], p_b);
  }
}

// dXc (R,C) given: dX1 += dXc*(1+m) ; dm_acc[b,c] += sum_n dXc*X1
__global__ void bwd_xc_kernel(const void* __restrict__ dxc, void* __restrict__ dx1, const void* __restrict__ x1, int bf, int C, int N,
                              const float* __restrict__ m, float* __restrict__ dm_acc) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx, b = blockIdx.y;
  float p = 0.f;
  if (c < C) {
    const float opm = 1.f + m[(size_t)b * C + c];
    for (int n = blockIdx.z * 8 + ry; n < N; n += gridDim.z * 8) {
      const size_t idx = ((size_t)b * N + n) * C + c;
      const float d = ld_act(dxc, idx, bf);
      st_act(dx1, idx, ld_act(dx1, idx, bf) + d * opm, bf);
      p += d * ld_act(x1, idx, bf);
    }
  }
  red[ry][cx] = p;
  __syncthreads();
  if (ry == 0 && c < C) {
    for (int t = 1; t < 8; ++t) p += red[t][cx];
    atomicAdd(&dm_acc[(size_t)b * C + c], p);
  }
}

// z1 (R,C) pre-bias channel query, in place -> dz1 = dr1[b,c]*1[z1+b>0] ; d_b1[c] += sum dz1
__global__ void bwd_channel_mask_kernel(void* __restrict__ z1, int bf, int C, int N, const float* __restrict__ bias1,
                                        const float* __restrict__ dr1, float* __restrict__ d_b1) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx, b = blockIdx.y;
  float p = 0.f;
  if (c < C) {
    const float bj = bias1[c], d = dr1[(size_t)b * C + c];
    for (int n = blockIdx.z * 8 + ry; n < N; n += gridDim.z * 8) {
      const size_t idx = ((size_t)b * N + n) * C + c;
      const float dz = (ld_act(z1, idx, bf) + bj) > 0.f ? d : 0.f;
      st_act(z1, idx, dz, bf);
      p += dz;
    }
  }
  red[ry][cx] = p;
  __syncthreads();
  if (ry == 0 && c < C) {
    for (int t = 1; t < 8; ++t) p += red[t][cx];
    atomicAdd(&d_b1[c], p);
  }
}

// q = aq1 * r1sum / N  =>  d aq1 = dq * r1sum / N ,  d r1 = dq * aq1 / N   (both products of one pass; they were two launches)
__global__ void dq_split_kernel(const float* __restrict__ dq, const float* __restrict__ r1sum, const float* __restrict__ aq1, int aq_ld,
                                float inv_n, float* __restrict__ daq1, int daq_ld, float* __restrict__ dr1, int BT, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= BT * C) return;
  const int b = i / C, c = i - b * C;
  const float g = dq[i] * inv_n;
  daq1[(size_t)b * daq_ld + c] = g * r1sum[i];
  dr1[i] = g * aq1[(size_t)b * aq_ld + c];
}

// ---- small per-frame vector ops (BT x width) ---------------------------------------------------------------------
enum { VOP_SIGMOID_BWD = 0, VOP_RELU_BWD = 1, VOP_MUL_SCALE = 2 };
// out[i] = op(g[i], y[i]) : sigmoid bwd g*y*(1-y) ; relu bwd g*1[y>0] ; mul-scale g*y*alpha
// (out16: optional bf16 copy with pitch out16_ld, the tensor-core operand of the per-frame-vector products)
__global__ void vec_op_kernel(int op, const float* g, int g_ld, const float* __restrict__ y, int y_ld, float alpha,
                              float* out, int out_ld, int rows, int cols, bf16* __restrict__ out16 = nullptr, int out16_ld = 0,
                              const bf16* __restrict__ y16 = nullptr) {       // y16: read y as bf16 (same pitch) instead of fp32
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int r = i / cols, c = i % cols;
  const float gv = g[(size_t)r * g_ld + c];
  const float yv = y16 ? __bfloat162float(y16[(size_t)r * y_ld + c]) : y[(size_t)r * y_ld + c];
  float v;
  if (op == VOP_SIGMOID_BWD) v = gv * yv * (1.f - yv);
  else if (op == VOP_RELU_BWD) v = yv > 0.f ? gv : 0.f;
  else v = gv * yv * alpha;
  out[(size_t)r * out_ld + c] = v;
  if (out16) out16[(size_t)r * out16_ld + c] = __float2bfloat16_rn(v);
}

// ---- token attention backward helpers ----------------------------------------------------------------------------
// rows of (R,K): dL2 = Q*(g*dQraw - sum_k g*dQraw*Q) in place on dq ; d_gate_av += sum Q*dQraw
// (ld = row pitch of dq and q, >= K)
__global__ void row_softmax_bwd_kernel(void* __restrict__ dq, const void* __restrict__ q, int bf, long long rows, int K, int ld,
                                       const float* __restrict__ gate_av, float* __restrict__ d_gate_av) {
  const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float g = gate_av[0];
  float gacc = 0.f;
  for (long long row = (long long)blockIdx.x * warps + warp; row < rows; row += (long long)gridDim.x * warps) {
    float dot = 0.f;
    for (int k = lane; k < K; k += 32) dot += ld_act(dq, (size_t)(row * ld + k), bf) * ld_act(q, (size_t)(row * ld + k), bf);
    dot = warp_sum(dot);
    gacc += (lane == 0) ? dot : 0.f;
    for (int k = lane; k < K; k += 32) {
      const float qq = ld_act(q, (size_t)(row * ld + k), bf);
      st_act(dq, (size_t)(row * ld + k), g * qq * (ld_act(dq, (size_t)(row * ld + k), bf) - dot), bf);
    }
  }
  if (lane == 0 && d_gate_av && gacc != 0.f) atomicAdd(d_gate_av, gacc);
}

// per frame, columns of (N,K): dL1 = P*(dP - sum_n dP*P)   (dP fp32 in, dL1 act dtype out)
// Two kernels over a (row-chunk, frame) grid like the forward's column softmax (a frame per block would leave half of
// the SMs idle at BT = 80): per-chunk partial dot products, then combine + apply.
// (dP is fp32 with pitch K; P and dL1 have pitch ldp >= K)
__global__ void __launch_bounds__(256) col_softmax_bwd_partial_kernel(const float* __restrict__ dP, const void* __restrict__ P, int p_bf,
                                                                      float* __restrict__ part, int N, int K, int ldp) {
  extern __shared__ float sm[];             // [groups][K]
  const int s_ = blockIdx.x, b = blockIdx.y, S = gridDim.x;
  const int groups = blockDim.x / K;
  const int k = threadIdx.x % K, g = threadIdx.x / K;
  const bool active = g < groups;
  const int rows = csm_chunk_rows(N);
  const int n0 = s_ * rows, n1 = min(N, n0 + rows);
  const size_t base = (size_t)b * N * K, basep = (size_t)b * N * ldp;
  float acc = 0.f;
  if (active) {
    for (int n = n0 + g; n < n1; n += groups) acc += dP[base + (size_t)n * K + k] * ld_act(P, basep + (size_t)n * ldp + k, p_bf);
    sm[g * K + k] = acc;
  }
  __syncthreads();
  if (active && g == 0) {
    float t = 0.f;
    for (int j = 0; j < groups; ++j) t += sm[j * K + k];
    part[((size_t)b * S + s_) * K + k] = t;
  }
}
__global__ void __launch_bounds__(256) col_softmax_bwd_apply_kernel(const float* __restrict__ dP, const void* __restrict__ P, int p_bf,
                                                                    const float* __restrict__ part, void* __restrict__ dL1, int out_bf,
                                                                    int N, int K, int ldp) {
  const int s_ = blockIdx.x, b = blockIdx.y, S = gridDim.x;
  const int groups = blockDim.x / K;
  const int k = threadIdx.x % K, g = threadIdx.x / K;
  if (g >= groups) return;
  float t = 0.f;
  for (int j = 0; j < S; ++j) t += part[((size_t)b * S + j) * K + k];
  const int rows = csm_chunk_rows(N);
  const int n0 = s_ * rows, n1 = min(N, n0 + rows);
  const size_t base = (size_t)b * N * K, basep = (size_t)b * N * ldp;
  for (int n = n0 + g; n < n1; n += groups) {
    const size_t i = base + (size_t)n * K + k, ip = basep + (size_t)n * ldp + k;
    st_act(dL1, ip, ld_act(P, ip, p_bf) * (dP[i] - t), out_bf);
  }
}


// vectorised in-place g = A[c]*g + B[c]*x + Cc[c] for bf16 g (pitch g_ld) and fp32 x (pitch x_ld); C % 8 == 0, 16-byte rows
__global__ void __launch_bounds__(256) affine_combine_vec_kernel(bf16* __restrict__ g, long long g_ld, const float* __restrict__ x,
                                                                 long long x_ld, const float* __restrict__ A, const float* __restrict__ B,
                                                                 const float* __restrict__ Cc, long long rows, int C) {
  const int nvec = C >> 3;
  const long long total = rows * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nvec;
    const int c = (int)(i - r * nvec) * 8;
    const uint4 raw = *reinterpret_cast<const uint4*>(g + r * g_ld + c);
    float v[8], xv[8], a[8], bb[8], cc[8];
    unpack8(raw, v);
    ldf8(x + r * x_ld + c, xv);
    ldf8(A + c, a); ldf8(B + c, bb); ldf8(Cc + c, cc);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(a[j], v[j], fmaf(bb[j], xv[j], cc[j]));
    *reinterpret_cast<uint4*>(g + r * g_ld + c) = pack8(v);
  }
}
inline void launch_affine_combine(void* g, int g_bf16, long long g_ld, const float* x, long long x_ld, const float* A, const float* B,
                                  const float* Cc, long long rows, int C, int ew_grid, cudaStream_t st);

// BatchNorm backward in ONE pass: the per-channel coefficients of bn_bwd_coeffs_kernel are a handful of flops from five C-sized
// vectors, so every thread derives them for its 8 channels instead of waiting for a 1-block kernel to publish them; the threads
// of row 0 also add the parameter gradients (each channel exactly once).  g = A g + B x + Cc in place (bf16 g, fp32 x).
__global__ void __launch_bounds__(256) bn_bwd_apply_vec_kernel(bf16* __restrict__ g, long long g_ld, const float* __restrict__ x, long long x_ld,
                                                               const float* __restrict__ sum_dy, const float* __restrict__ sum_dyx,
                                                               const float* __restrict__ mean, const float* __restrict__ rstd,
                                                               const float* __restrict__ w, float inv_count, int train,
                                                               float* __restrict__ dW, float* __restrict__ dB, long long rows, int C) {
  const int nvec = C >> 3;
  const long long total = rows * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / nvec;
    const int c = (int)(i - r * nvec) * 8;
    const uint4 raw = *reinterpret_cast<const uint4*>(g + r * g_ld + c);
    float v[8], xv[8], sdy[8], sdx[8], mu[8], rs[8], ww[8];
    unpack8(raw, v);
    ldf8(x + r * x_ld + c, xv);
    ldf8(sum_dy + c, sdy); ldf8(sum_dyx + c, sdx); ldf8(mean + c, mu); ldf8(rstd + c, rs); ldf8(w + c, ww);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float db = sdy[j];
      const float dw = (sdx[j] - mu[j] * db) * rs[j];
      const float wr = ww[j] * rs[j];
      if (r == 0) {
        if (dW) atomicAdd(&dW[c + j], dw);
        if (dB) atomicAdd(&dB[c + j], db);
      }
      const float a = wr;
      const float b = train ? -wr * rs[j] * dw * inv_count : 0.f;
      const float cc = train ? (-wr * db * inv_count + wr * rs[j] * mu[j] * dw * inv_count) : 0.f;
      v[j] = fmaf(a, v[j], fmaf(b, xv[j], cc));
    }
    *reinterpret_cast<uint4*>(g + r * g_ld + c) = pack8(v);
  }
}
// true when the fused pass ran (bf16 g, 8-channel vectors, 16-byte aligned rows); else the caller runs coeffs + combine
inline bool launch_bn_bwd_apply(void* g, int g_bf16, long long g_ld, const float* x, long long x_ld, const float* sum_dy, const float* sum_dyx,
                                const float* mean, const float* rstd, const float* w, double count, int train, float* dW, float* dB,
                                long long rows, int C, cudaStream_t st) {
  if (!(g_bf16 && C % 8 == 0 && g_ld % 8 == 0 && x_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(x) & 15) == 0 && ((reinterpret_cast<uintptr_t>(sum_dy) | reinterpret_cast<uintptr_t>(sum_dyx) |
        reinterpret_cast<uintptr_t>(mean) | reinterpret_cast<uintptr_t>(rstd) | reinterpret_cast<uintptr_t>(w)) & 15) == 0))
    return false;
  const long long total = rows * (C / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 16) blocks = 148LL * 16;
  if (blocks < 1) blocks = 1;
  bn_bwd_apply_vec_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<bf16*>(g), g_ld, x, x_ld, sum_dy, sum_dyx, mean, rstd, w,
                                                            (float)(1.0 / count), train, dW, dB, rows, C);
  return true;
}

// out[k,c] += sum_b in[b,k,c]     (gradient of the shared latent tokens through T' = T + ...)
// out[i] += sum_b in[b, i]  (i < per_frame = K*C).  With d_fcb != NULL the same sums also feed the gradient of fc.bias:
//   d fc_b[c] = sum_{b,n} dV[b,n,c]  with  dV = P dT' + dL1 T + dobar/N .
// Column sums over n of the three terms: sum_n P[n,k] = 1 (softmax over n), sum_n dL1[n,k] = 0 (softmax backward), and the
// broadcast term gives dobar, hence EXACTLY  d fc_b[c] = sum_b ( sum_k dT'[b,k,c] + dobar[b,c] )  -- no pass over the
// (BT,N,C) tensor, and none of its bf16 rounding (summing 10^5 rounded rows left fc.bias at 0.6-0.8 rel-L2 in round 1).
__global__ void sum_frames_kernel(const float* __restrict__ in, int BT, size_t per_frame, float* __restrict__ out,
                                  float* __restrict__ d_fcb = nullptr, int C = 0, const float* __restrict__ dobar = nullptr) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_frame) return;
  float s = 0.f;
  for (int b = 0; b < BT; ++b) s += in[(size_t)b * per_frame + i];
  atomicAdd(&out[i], s);          // (atomic: runs on the auxiliary stream beside the kernels that add the other half of d tokens)
  if (d_fcb) {
    const int c = (int)(i % (size_t)C);
    if (i < (size_t)C)                      // first token row also carries sum_b dobar[b, c]
      for (int b = 0; b < BT; ++b) s += dobar[(size_t)b * C + c];
    atomicAdd(&d_fcb[c], s);
  }
}

// fp32 -> act dtype copy with scale (operand copies for the tensor-core products)
__global__ void cast_scale_kernel(const float* __restrict__ src, float alpha, void* __restrict__ dst, int dst_bf16, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    st_act(dst, i, alpha * src[i], dst_bf16);
}

// d_bres[n] += sum_{b,c'} dU[b,n,c']
__global__ void rowsum_frames_kernel(const void* __restrict__ du, int bf, int BT, int N, int Co, float* __restrict__ out) {
  const int n = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  __shared__ float red[32];
  float s = 0.f;
  for (int b = warp; b < BT; b += warps)
    for (int c = lane; c < Co; c += 32) s += ld_act(du, ((size_t)b * N + n) * Co + c, bf);
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(&out[n], s);
}

template <int LPR>
__device__ __forceinline__ float rowgroups_sum(float v) {      // sum over the 32/LPR row groups of a warp (same columns)
#pragma unroll
  for (int o = LPR; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- re-associated resampler: both bias-related reductions of dV in ONE pass -------------------------------------------------
//   d_bres[n] += sum_{b,c} dV[b,n,c] * s[c]          (s = row sums of fc.weight)
//   t[c]      += sum_{b,n} dV[b,n,c] * b_res[n]      (rank-1 part of d fc.weight, added by add_col_broadcast_kernel)
// One warp per row, lanes stride over 16-byte vectors; the column sums stay in registers across the rows a warp visits (<= 6
// vectors per lane: C <= 1536) and are flushed once per warp.  Replaces two kernels that each read dV completely.
template <int LPR, int VPL>
__global__ void __launch_bounds__(256) dv_bias_sums_kernel(const bf16* __restrict__ dv, long long rows, int N, int C,
                                                           const float* __restrict__ s, const float* __restrict__ b_res,
                                                           float* __restrict__ d_bres, float* __restrict__ t) {
  extern __shared__ float blk_t[];               // [C] block partial of t (one global atomic per column per BLOCK)
  constexpr int RPW = 32 / LPR;                  // rows per warp iteration (LPR lanes share a row, as in rowwise.cuh)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane % LPR, rsub = lane / LPR, nvec = C >> 3;
  for (int c = threadIdx.x; c < C; c += blockDim.x) blk_t[c] = 0.f;
  __syncthreads();
  float sv[VPL][8], acc[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = sub + LPR * i;
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc[i][j] = 0.f; sv[i][j] = (vi < nvec) ? s[vi * 8 + j] : 0.f; }
  }
  const long long step = (long long)gridDim.x * 8 * RPW;
  for (long long base = ((long long)blockIdx.x * 8 + warp) * RPW; base < rows; base += step) {
    const long long row = base + rsub;
    const bool rv = row < rows;
    const int n = rv ? (int)(row % N) : 0;
    const float br = rv ? b_res[n] : 0.f;
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = sub + LPR * i;
      if (rv && vi < nvec) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(dv + row * C + vi * 8));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[q]);
          const float a = __bfloat162float(h.x), b = __bfloat162float(h.y);
          dot = fmaf(a, sv[i][2 * q], fmaf(b, sv[i][2 * q + 1], dot));
          acc[i][2 * q] = fmaf(a, br, acc[i][2 * q]);
          acc[i][2 * q + 1] = fmaf(b, br, acc[i][2 * q + 1]);
        }
      }
    }
    dot = group_sum<LPR>(dot);
    if (rv && sub == 0) atomicAdd(&d_bres[n], dot);
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = sub + LPR * i;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float v = rowgroups_sum<LPR>(acc[i][j]);
      if (rsub == 0 && vi < nvec) atomicAdd(&blk_t[vi * 8 + j], v);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(&t[c], blk_t[c]);
}
struct DvBiasLaunch {
  const bf16* p; long long rows; int N, C; const float* s; const float* b_res; float* d_bres; float* t; cudaStream_t st; bool* ok;
  template <int LPR, int VPL> void run() {
    if (VPL > 6) { *ok = false; return; }
    const long long rows_per_block = 8LL * (32 / LPR);
    long long blocks = (rows + 4 * rows_per_block - 1) / (4 * rows_per_block);        // >= 4 iterations per block
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    dv_bias_sums_kernel<LPR, (VPL > 6 ? 1 : VPL)><<<(unsigned)blocks, 256, (size_t)C * 4, st>>>(p, rows, N, C, s, b_res, d_bres, t);
    *ok = true;
  }
};
// returns false when the shape is outside the vector kernel's range: the caller runs the two scalar kernels
inline bool launch_dv_bias_sums(const void* dv, int bf, long long rows, int N, int C, const float* s, const float* b_res, float* d_bres,
                                float* t, cudaStream_t st) {
  if (!bf || C % 8 || C > 1536) return false;
  bool ok = false;
  if (!dispatch_row_geometry(C / 8, DvBiasLaunch{reinterpret_cast<const bf16*>(dv), rows, N, C, s, b_res, d_bres, t, st, &ok})) return false;
  return ok;
}

// ---- re-associated resampler (DGSCT_FC_FIRST): the rank-1 bias term b_res[n] * s[c] of V --------------------------------
// d_bres[n] += sum_{b,c} dV[b,n,c] * s[c]
__global__ void rowsum_frames_weighted_kernel(const void* __restrict__ dv, int bf, int BT, int N, int C, const float* __restrict__ s,
                                              float* __restrict__ out) {
  const int n = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  __shared__ float red[32];
  float acc = 0.f;
  for (int b = warp; b < BT; b += warps)
    for (int c = lane; c < C; c += 32) acc += ld_act(dv, ((size_t)b * N + n) * C + c, bf) * s[c];
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(&out[n], acc);
}
// t[c] += sum_{b,n} dV[b,n,c] * w[n]          (grid (ceil(C/32), row splits), block 256 = 32 columns x 8 row lanes)
__global__ void col_reduce_rowweight_kernel(const void* __restrict__ a, int bf, long long rows, int N, int C,
                                            const float* __restrict__ w, float* __restrict__ t) {
  __shared__ float r1[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float p1 = 0.f;
  if (c < C)
    for (long long r = (long long)blockIdx.y * 8 + ry; r < rows; r += (long long)gridDim.y * 8)
      p1 += ld_act(a, (size_t)(r * C + c), bf) * w[(int)(r % N)];
  r1[ry][cx] = p1;
  __syncthreads();
  if (ry == 0 && c < C) {
    for (int j = 1; j < 8; ++j) p1 += r1[j][cx];
    atomicAdd(&t[c], p1);
  }
}
// g[c, co] += t[c] for every co      (d fc.weight through s = rowsum(W_fc))
__global__ void add_col_broadcast_kernel(float* __restrict__ g, const float* __restrict__ t, int C, int Co) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C * Co) g[i] += t[i / Co];
}

// zr (rows, ldz) fp32 = relu(z*scale+shift)
__global__ void bn_relu_f32_kernel(const float* __restrict__ z, const float* __restrict__ scale, const float* __restrict__ shift,
                                   float* __restrict__ zr, long long rows, int Cd, int ldz) {
  const long long total = rows * Cd;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / Cd; const int c = (int)(i % Cd);
    zr[row * ldz + c] = fmaxf(fmaf(z[i], scale[c], shift[c]), 0.f);
  }
}
__global__ void one_plus_kernel(const float* __restrict__ m, float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = 1.f + m[i];
}
// tau[b] = sigmoid(wt . obar[b] + bt): d_pre = (dtau_gate[b] + dtmap[b]) * tau (1-tau)
__global__ void temporal_bwd_kernel(const float* __restrict__ dtau, const float* __restrict__ dtmap, const float* __restrict__ tau,
                                    const float* __restrict__ obar, const float* __restrict__ tw, int C, float* __restrict__ dobar,
                                    float* __restrict__ d_tw, float* __restrict__ d_tb) {
  const int b = blockIdx.x;
  const float t = tau[b];
  const float g = ((dtau ? dtau[b] : 0.f) + (dtmap ? dtmap[b] : 0.f)) * t * (1.f - t);
  if (g == 0.f) return;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    atomicAdd(&d_tw[c], g * obar[(size_t)b * C + c]);
    dobar[(size_t)b * C + c] += g * tw[c];
  }
  if (threadIdx.x == 0) atomicAdd(d_tb, g);
}

// dense block-diagonal (out, in) fp32 gradient -> grouped conv layout (out, in/g): grad[o, j] += dense[o, (o/opg)*ipg + j]
__global__ void ungroup_grad_kernel(const float* __restrict__ dense, int out_ch, int in_ch, int groups, int ld_dense,
                                    float* __restrict__ grad) {
  const int ipg = in_ch / groups, opg = out_ch / groups;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out_ch * ipg) return;
  const int o = i / ipg, j = i % ipg;
  grad[i] += dense[(size_t)o * ld_dense + (o / opg) * ipg + j];
}
// grouped conv weight (out, in/g) fp32 -> dense block-diagonal fp32 (out, ld)
__global__ void densify_grouped_kernel(const float* __restrict__ src, int out_ch, int in_ch, int groups, float* __restrict__ dst, int ld) {
  const int ipg = in_ch / groups, opg = out_ch / groups;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= out_ch * ld) return;
  const int o = i / ld, c = i % ld;
  dst[i] = (c < in_ch && c / ipg == o / opg) ? src[(size_t)o * ipg + (c - (o / opg) * ipg)] : 0.f;
}


// =====================================================================================================================
// Vectorised (bf16, 16-byte) versions of the two LayerNorm-backward row kernels.  Same row geometry as rowwise.cuh
// (LPR lanes per row, VPL vectors of 8 channels per lane), so a lane owns FIXED columns for all of its rows: the column
// reductions (d ln weight / bias, the BatchNorm-backward sums, the channel-gate gradient) are accumulated in registers
// over the block's rows and flushed once per block -- the scalar kernels issue one shared-memory (and, for the channel
// gate, one GLOBAL) atomic per element.
// =====================================================================================================================

// out = gate * LN_post(B2) [AVE] or LN_post(gate * B2) [AVS], B2 = uo*scale + shift.  Also accumulates the bn2-backward
// column sums colA[c] += sum dB2, colB[c] += sum dB2*uo (NULL to skip) so that no separate reduction pass is needed.
// SMACC (wide rows, VPL > 4: C = 1536): the per-lane column accumulators (4 x VPL x 8 registers) would not fit next to the row
// data, so every element adds straight into the block's shared-memory partials instead (such shapes have few rows; round 1
// sent them to the scalar kernel: 71 / 122 us for a 9 MB tensor).
template <int LPR, int VPL, bool SMACC = (VPL > 4)>
__global__ void __launch_bounds__(256, (VPL <= 1 ? 2 : 1)) bwd_out_vec_kernel(BwdOutArgs a, float* __restrict__ colA, float* __restrict__ colB) {
  extern __shared__ float sm[];                 // [4][C] block partials: d_lnw, d_lnb, colA, colB
  constexpr int AV = SMACC ? 1 : VPL;
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane % LPR, rsub = lane / LPR;
  const int C = a.C, nvec = C >> 3;
  const float invC = 1.f / (float)C;
  for (int c = threadIdx.x; c < 4 * C; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  const bf16* dout = reinterpret_cast<const bf16*>(a.dout);
  bf16* db2 = reinterpret_cast<bf16*>(a.db2);
  const float g = a.gate ? a.gate[0] : 1.f;
  const float gpre = a.gate_before_ln ? g : 1.f, gpost = a.gate_before_ln ? 1.f : g;
  float cw[AV][8], cb[AV][8], ca[AV][8], cq[AV][8];
#pragma unroll
  for (int i = 0; i < AV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { cw[i][j] = 0.f; cb[i][j] = 0.f; ca[i][j] = 0.f; cq[i][j] = 0.f; }
  float gate_acc = 0.f;
  const long long step = (long long)gridDim.x * 8 * RPW;
  // software pipeline (see bwd_gate_ln_vec_kernel): next row group's loads are in flight during this group's reductions
  float4 nu0[VPL], nu1[VPL];
  uint4 ndo[VPL];
  auto fetch = [&](long long base_) {
    const long long r_ = base_ + rsub;
    const bool ok = base_ < a.rows && r_ < a.rows;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = sub + LPR * i;
      if (ok && vi < nvec) {
        const float4* up = reinterpret_cast<const float4*>(a.uo + r_ * C + vi * 8);
        nu0[i] = __ldg(up); nu1[i] = __ldg(up + 1);
        ndo[i] = __ldg(reinterpret_cast<const uint4*>(dout + r_ * C + vi * 8));
      } else {
        nu0[i] = make_float4(0.f, 0.f, 0.f, 0.f); nu1[i] = nu0[i]; ndo[i] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  };
  fetch(((long long)blockIdx.x * 8 + warp) * RPW);
  for (long long base = ((long long)blockIdx.x * 8 + warp) * RPW; base < a.rows; base += step) {
    const long long row = base + rsub;
    const bool rv = row < a.rows;
    float v[VPL][8], dy[VPL][8], ucur[VPL][8];
    uint4 rawd[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      ucur[i][0] = nu0[i].x; ucur[i][1] = nu0[i].y; ucur[i][2] = nu0[i].z; ucur[i][3] = nu0[i].w;
      ucur[i][4] = nu1[i].x; ucur[i][5] = nu1[i].y; ucur[i][6] = nu1[i].z; ucur[i][7] = nu1[i].w;
      rawd[i] = ndo[i];
    }
    fetch(base + step);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = sub + LPR * i;
      if (rv && vi < nvec) {
        float sc[8], sh[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = ucur[i][j];
        const uint4 raw = rawd[i];
        ldf8(a.scale + vi * 8, sc);
        ldf8(a.shift + vi * 8, sh);
        unpack8(raw, dy[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[i][j] = fmaf(v[i][j], sc[j], sh[j]) * gpre; sum += v[i][j]; }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[i][j] = 0.f; dy[i][j] = 0.f; }
      }
    }
    float rstd = 1.f, s1 = 0.f, s2 = 0.f;
    if (a.ln_w) {
      const float mean = group_sum<LPR>(sum) * invC;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i)
        if (sub + LPR * i < nvec) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float dlt = v[i][j] - mean; sq += dlt * dlt; }
        }
      rstd = rsqrtf(group_sum<LPR>(sq) * invC + a.eps);
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int vi = sub + LPR * i;
        if (rv && vi < nvec) {
          float w[8], bb[8];
          ldf8(a.ln_w + vi * 8, w);
          ldf8(a.ln_b + vi * 8, bb);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xhat = (v[i][j] - mean) * rstd;
            const float dL = dy[i][j] * gpost;
            if (!a.gate_before_ln && a.gate) gate_acc += dy[i][j] * fmaf(xhat, w[j], bb[j]);   // AVE: d_gate = sum dOut * L
            if (SMACC) { atomicAdd(&sm[vi * 8 + j], dL * xhat); atomicAdd(&sm[C + vi * 8 + j], dL); }
            else { cw[SMACC ? 0 : i][j] += dL * xhat; cb[SMACC ? 0 : i][j] += dL; }
            const float dxh = dL * w[j];
            s1 += dxh; s2 += dxh * xhat;
            v[i][j] = xhat; dy[i][j] = dxh;
          }
        }
      }
      s1 = group_sum<LPR>(s1) * invC;
      s2 = group_sum<LPR>(s2) * invC;
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = sub + LPR * i;
      if (rv && vi < nvec) {
        float u[8], o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) u[j] = ucur[i][j];    // (register copy of the row: no second global read)
        if (a.ln_w) {
          if (a.gate_before_ln && a.gate) {
            float sc[8], sh[8];
            ldf8(a.scale + vi * 8, sc);
            ldf8(a.shift + vi * 8, sh);
#pragma unroll
            for (int j = 0; j < 8; ++j) gate_acc += rstd * (dy[i][j] - s1 - v[i][j] * s2) * fmaf(u[j], sc[j], sh[j]);   // AVS: dv * B2
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = rstd * (dy[i][j] - s1 - v[i][j] * s2) * gpre;
        } else {
          if (a.gate) {
            float sc[8], sh[8];
            ldf8(a.scale + vi * 8, sc);
            ldf8(a.shift + vi * 8, sh);
#pragma unroll
            for (int j = 0; j < 8; ++j) gate_acc += dy[i][j] * fmaf(u[j], sc[j], sh[j]);
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = dy[i][j] * g;
        }
        *reinterpret_cast<uint4*>(db2 + row * C + vi * 8) = pack8(o);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (SMACC) { atomicAdd(&sm[2 * C + vi * 8 + j], o[j]); atomicAdd(&sm[3 * C + vi * 8 + j], o[j] * u[j]); }
          else { ca[SMACC ? 0 : i][j] += o[j]; cq[SMACC ? 0 : i][j] += o[j] * u[j]; }
        }
      }
    }
  }
  // flush: row groups of the warp -> block partials in shared memory -> one global atomic per column per block
#pragma unroll
  for (int i = 0; i < AV; ++i) {
    if (SMACC) break;
    const int vi = sub + LPR * i;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float w_ = rowgroups_sum<LPR>(cw[i][j]), b_ = rowgroups_sum<LPR>(cb[i][j]);
      const float a_ = rowgroups_sum<LPR>(ca[i][j]), q_ = rowgroups_sum<LPR>(cq[i][j]);
      if (rsub == 0 && vi < nvec) {
        const int c = vi * 8 + j;
        atomicAdd(&sm[c], w_); atomicAdd(&sm[C + c], b_); atomicAdd(&sm[2 * C + c], a_); atomicAdd(&sm[3 * C + c], q_);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (a.ln_w) {
      if (a.d_lnw) atomicAdd(&a.d_lnw[c], sm[c]);
      if (a.d_lnb) atomicAdd(&a.d_lnb[c], sm[C + c]);
    }
    if (colA) atomicAdd(&colA[c], sm[2 * C + c]);
    if (colB) atomicAdd(&colB[c], sm[3 * C + c]);
  }
  if (a.gate && a.d_gate) {
    gate_acc = warp_sum(gate_acc);
    if (lane == 0 && gate_acc != 0.f) atomicAdd(a.d_gate, gate_acc);
  }
}

// LN_before + gate-combine backward; grid (row chunks, BT): a block stays inside one frame, so the channel-gate gradient
// dm[b, c] is a per-lane register sum as well.
template <int LPR, int VPL, bool SMACC = (VPL > 4)>
__global__ void __launch_bounds__(256, (VPL <= 2 ? 2 : 1)) bwd_gate_ln_vec_kernel(BwdGateLnArgs a, int rows_per_chunk) {
  extern __shared__ float sm[];                 // [3][C] block partials: d_lnw, d_lnb, dm
  constexpr int AV = SMACC ? 1 : VPL;           // (SMACC: see bwd_out_vec_kernel)
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane % LPR, rsub = lane / LPR;
  const int C = a.C, nvec = C >> 3, b = blockIdx.y;
  const float invC = 1.f / (float)C;
  for (int c = threadIdx.x; c < 3 * C; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  const bf16* x1 = reinterpret_cast<const bf16*>(a.x1);
  const bf16* dyp = reinterpret_cast<const bf16*>(a.dy);
  bf16* dx1 = reinterpret_cast<bf16*>(a.dx1);
  const float* mrow = a.m + (size_t)b * C;
  const float tau_term = a.tau ? a.gamma * a.tau[b] : 0.f;
  float cw[AV][8], cb[AV][8], cm[AV][8];
#pragma unroll
  for (int i = 0; i < AV; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) { cw[i][j] = 0.f; cb[i][j] = 0.f; cm[i][j] = 0.f; }
  float tau_acc = 0.f;
  const int n_begin = blockIdx.x * rows_per_chunk;
  const int n_end = min(a.N, n_begin + rows_per_chunk);
  // Software pipeline: the two 16-byte loads per vector of the NEXT row group are issued before the current group's
  // shuffle reductions (8-16 warps x one row group each is < 24 KB in flight per SM: latency-bound at ~1 TB/s without it).
  uint4 nx[VPL], nd[VPL];
  auto fetch = [&](int nb_) {
    const int n_ = nb_ + rsub;
    const bool ok = nb_ < n_end && n_ < n_end;
    const long long r_ = (long long)b * a.N + n_;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = sub + LPR * i;
      if (ok && vi < nvec) {
        nx[i] = __ldg(reinterpret_cast<const uint4*>(x1 + r_ * C + vi * 8));
        nd[i] = __ldg(reinterpret_cast<const uint4*>(dyp + r_ * C + vi * 8));
      } else {
        nx[i] = make_uint4(0u, 0u, 0u, 0u); nd[i] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  };
  fetch(n_begin + warp * RPW);
  for (int nb = n_begin + warp * RPW; nb < n_end; nb += 8 * RPW) {
    const int n = nb + rsub;
    const bool rv = n < n_end;
    const long long row = (long long)b * a.N + n;
    const float gbase = rv ? (a.beta * a.sig[row] + tau_term + 1.f - a.alpha) : 0.f;
    float v[VPL][8], dv[VPL][8];
    uint4 raw[VPL], rawd[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) { raw[i] = nx[i]; rawd[i] = nd[i]; }
    fetch(nb + 8 * RPW);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = sub + LPR * i;
      if (rv && vi < nvec) {
        const uint4 rd = rawd[i];
        float mm[8];
        ldf8(mrow + vi * 8, mm);
        unpack8(raw[i], v[i]);
        unpack8(rd, dv[i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[i][j] *= fmaf(a.alpha, mm[j], gbase); sum += v[i][j]; }
      } else {
        raw[i] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[i][j] = 0.f; dv[i][j] = 0.f; }
      }
    }
    if (a.ln_w) {
      const float mean = group_sum<LPR>(sum) * invC;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i)
        if (sub + LPR * i < nvec) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float dlt = v[i][j] - mean; sq += dlt * dlt; }
        }
      const float rstd = rsqrtf(group_sum<LPR>(sq) * invC + a.eps);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const int vi = sub + LPR * i;
        if (rv && vi < nvec) {
          float w[8];
          ldf8(a.ln_w + vi * 8, w);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float xhat = (v[i][j] - mean) * rstd;
            if (SMACC) { atomicAdd(&sm[vi * 8 + j], dv[i][j] * xhat); atomicAdd(&sm[C + vi * 8 + j], dv[i][j]); }
            else { cw[SMACC ? 0 : i][j] += dv[i][j] * xhat; cb[SMACC ? 0 : i][j] += dv[i][j]; }
            const float dxh = dv[i][j] * w[j];
            s1 += dxh; s2 += dxh * xhat;
            v[i][j] = xhat; dv[i][j] = dxh;
          }
        }
      }
      s1 = group_sum<LPR>(s1) * invC;
      s2 = group_sum<LPR>(s2) * invC;
#pragma unroll
      for (int i = 0; i < VPL; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) dv[i][j] = rstd * (dv[i][j] - s1 - v[i][j] * s2);      // dYg
    }
    float rowdot = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = sub + LPR * i;
      if (rv && vi < nvec) {
        float xr[8], mm[8], o[8];
        unpack8(raw[i], xr);
        ldf8(mrow + vi * 8, mm);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float dgm = dv[i][j] * xr[j];
          rowdot += dgm;
          if (SMACC) atomicAdd(&sm[2 * C + vi * 8 + j], dgm);
          else cm[SMACC ? 0 : i][j] += dgm;
          o[j] = dv[i][j] * fmaf(a.alpha, mm[j], gbase);
        }
        *reinterpret_cast<uint4*>(dx1 + row * C + vi * 8) = pack8(o);
      }
    }
    rowdot = group_sum<LPR>(rowdot);
    if (rv && sub == 0) {
      a.dsig[row] = a.beta * rowdot;
      tau_acc += rowdot;
    }
  }
#pragma unroll
  for (int i = 0; i < AV; ++i) {
    if (SMACC) break;
    const int vi = sub + LPR * i;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float w_ = rowgroups_sum<LPR>(cw[i][j]), b_ = rowgroups_sum<LPR>(cb[i][j]), m_ = rowgroups_sum<LPR>(cm[i][j]);
      if (rsub == 0 && vi < nvec) {
        const int c = vi * 8 + j;
        atomicAdd(&sm[c], w_); atomicAdd(&sm[C + c], b_); atomicAdd(&sm[2 * C + c], m_);
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    if (a.ln_w) {
      if (a.d_lnw) atomicAdd(&a.d_lnw[c], sm[c]);
      if (a.d_lnb) atomicAdd(&a.d_lnb[c], sm[C + c]);
    }
    atomicAdd(&a.dm_acc[(size_t)b * C + c], a.alpha * sm[2 * C + c]);
  }
  if (a.dtau) {
    tau_acc = warp_sum(tau_acc);
    if (lane == 0 && tau_acc != 0.f) atomicAdd(&a.dtau[b], a.gamma * tau_acc);
  }
}

struct BwdOutLaunch {
  const BwdOutArgs& a; float* colA; float* colB; cudaStream_t st; bool* ok;
  template <int LPR, int VPL> void run() {
    if (VPL > 6) { *ok = false; return; }
    const long long rows_per_block = 8LL * (32 / LPR);
    long long blocks = (a.rows + rows_per_block - 1) / rows_per_block;
    if (blocks > 2 * 148) blocks = 2 * 148;              // every block resident (1 or 2 per SM): one flush per block
    bwd_out_vec_kernel<LPR, (VPL > 6 ? 1 : VPL)><<<(unsigned)blocks, 256, (size_t)4 * a.C * 4, st>>>(a, colA, colB);
    *ok = true;
  }
};
// returns false when the caller must run the scalar kernel (+ its own column reduction)
inline bool launch_bwd_out_vec(const BwdOutArgs& a, float* colA, float* colB, cudaStream_t st) {
  if (a.C % 8 || !a.dout_bf16 || !a.db2_bf16 || a.C > 2048) return false;
  bool ok = false;
  static const bool heavy = getenv("DGSCT_ROW_HEAVY") != nullptr;       // round-1 geometry (A/B switch)
  if (!heavy && a.C <= 1024) { if (!dispatch_row_geometry_light(a.C / 8, BwdOutLaunch{a, colA, colB, st, &ok})) return false; }
  else if (!dispatch_row_geometry(a.C / 8, BwdOutLaunch{a, colA, colB, st, &ok})) return false;
  return ok;
}

struct BwdGateLnLaunch {
  const BwdGateLnArgs& a; int BT; cudaStream_t st; bool* ok;
  template <int LPR, int VPL> void run() {
    if (VPL > 6) { *ok = false; return; }
    const int rows_per_iter = 8 * (32 / LPR);
    int chunks = (a.N + rows_per_iter - 1) / rows_per_iter;          // at least one block iteration per chunk
    const int want = (4 * 148 + BT - 1) / BT;                        // ~4 blocks per SM overall
    if (chunks > want) chunks = want;
    if (chunks < 1) chunks = 1;
    int rpc = (a.N + chunks - 1) / chunks;
    rpc = (rpc + rows_per_iter - 1) / rows_per_iter * rows_per_iter;
    chunks = (a.N + rpc - 1) / rpc;
    bwd_gate_ln_vec_kernel<LPR, (VPL > 6 ? 1 : VPL)><<<dim3(chunks, BT), 256, (size_t)3 * a.C * 4, st>>>(a, rpc);
    *ok = true;
  }
};
inline bool launch_bwd_gate_ln_vec(const BwdGateLnArgs& a, int BT, cudaStream_t st) {
  if (a.C % 8 || !a.dy_bf16 || !a.x1_bf16 || !a.dx1_bf16 || a.C > 2048) return false;
  bool ok = false;
  static const bool heavy = getenv("DGSCT_ROW_HEAVY") != nullptr;
  if (!heavy && a.C <= 1024) { if (!dispatch_row_geometry_light(a.C / 8, BwdGateLnLaunch{a, BT, st, &ok})) return false; }
  else if (!dispatch_row_geometry(a.C / 8, BwdGateLnLaunch{a, BT, st, &ok})) return false;
  return ok;
}

inline void launch_affine_combine(void* g, int g_bf16, long long g_ld, const float* x, long long x_ld, const float* A, const float* B,
                                  const float* Cc, long long rows, int C, int ew_grid, cudaStream_t st) {
  if (g_bf16 && C % 8 == 0 && g_ld % 8 == 0 && x_ld % 4 == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const long long total = rows * (C / 8);
    long long blocks = (total + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    if (blocks < 1) blocks = 1;
    affine_combine_vec_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<bf16*>(g), g_ld, x, x_ld, A, B, Cc, rows, C);
  } else {
    affine_combine_kernel<<<ew_grid, 256, 0, st>>>(g, g_bf16, g_ld, x, x_ld, A, B, Cc, rows, C);
  }
}

}  // namespace dgsct
