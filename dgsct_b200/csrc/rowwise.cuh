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
  const float* ssum; const float* s_bias; // non-NULL ssum: the gate is derived here as sigmoid(ssum[row] + s_bias[0]) (same expression
                                         // as spatial_map_kernel, which then runs beside this kernel instead of in front of it)
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
  const float sgv = a.ssum ? 1.f / (1.f + expf(-(a.ssum[row] + (a.s_bias ? a.s_bias[0] : 0.f)))) : a.sig[row];
  const float base = a.beta * sgv + (a.tau ? a.gamma * a.tau[b] : 0.f) + 1.f - a.alpha;
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
                                   float eps, float momentum, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_out, float* __restrict__ rstd_out, int update_running) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (!use_bn) { scale[c] = 1.f; shift[c] = 0.f; mean_out[c] = 0.f; rstd_out[c] = 1.f; return; }
  double mean, var;
  if (train) {
    mean = sum[c] / count;
    var = sq[c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double unb = count > 1.0 ? var * count / (count - 1.0) : var;
    if (update_running) {
      rm[c] = (float)((1.0 - momentum) * rm[c] + momentum * mean);
      rv[c] = (float)((1.0 - momentum) * rv[c] + momentum * unb);
    }
  } else {
    mean = rm[c];
    var = rv[c];
  }
  const double rstd = 1.0 / sqrt(var + (double)eps);
  const double sc = (double)w[c] * rstd;
  mean_out[c] = (float)mean;
  rstd_out[c] = (float)rstd;
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
  const void* resid;                       // non-NULL: out = resid + adapter output (the caller's residual add, net_trans.py:895-906,
                                           // folded into this pass by the pair entry point; same dtype as out)
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
      st_act(a.out, (size_t)row * a.C + c, ((buf[c] - mean) * rstd * a.ln_w[c] + a.ln_b[c]) * gpost +
             (a.resid ? ld_act(a.resid, (size_t)row * a.C + c, a.out_bf16) : 0.f), a.out_bf16);
  } else {
    for (int c = lane; c < a.C; c += 32)
      st_act(a.out, (size_t)row * a.C + c, buf[c] * gpost + (a.resid ? ld_act(a.resid, (size_t)row * a.C + c, a.out_bf16) : 0.f), a.out_bf16);
  }
}

// dtype-converting copy into an fp32 tap buffer
__global__ void tap_copy_kernel(const void* __restrict__ src, int is_bf16, float* __restrict__ dst, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = ld_act(src, i, is_bf16);
}

}  // namespace dgsct

// ---- vectorised (16-byte) row kernels for the bf16 path ---------------------------------------------------
// One warp per token row, each lane owns VPL vectors of 8 consecutive channels kept in registers; statistics by
// warp shuffles.  Requires C % 8 == 0 and 16-byte aligned rows.
namespace dgsct {

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
    f[2 * j] = __bfloat162float(t.x); f[2 * j + 1] = __bfloat162float(t.y);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162 t;
  t = __floats2bfloat162_rn(f[0], f[1]); u.x = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(f[2], f[3]); u.y = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(f[4], f[5]); u.z = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(f[6], f[7]); u.w = *reinterpret_cast<uint32_t*>(&t);
  return u;
}
__device__ __forceinline__ void ldf8(const float* __restrict__ p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// Row geometry: a token row of C channels is nvec = C/8 16-byte vectors.  LPR lanes share one row (LPR a power of two
// <= 32), each lane owns VPL vectors (lane sub, sub+LPR, ...), so a warp processes 32/LPR rows at once with every lane
// busy and VPL independent 16-byte loads in flight per lane -- narrow rows (C = 96: 12 vectors) no longer leave 20 of 32
// lanes idle.  Statistics are reduced with xor-shuffles inside the LPR-lane group.
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int LPR, int VPL>
__global__ void __launch_bounds__(256) gate_ln_vec_kernel(GateLnArgs a) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR, rsub = lane / LPR;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int nvec = a.C >> 3;
  const float invC = 1.f / (float)a.C;
  const bf16* x1 = reinterpret_cast<const bf16*>(a.x1);
  bf16* y = reinterpret_cast<bf16*>(a.y);
  for (long long base = warp0 * RPW; base < a.rows; base += nwarps * RPW) {
    const long long row = base + rsub;
    const bool rv = row < a.rows;
    const int b = rv ? (int)(row / a.N) : 0;
    const float sgv = !rv ? 0.f : (a.ssum ? 1.f / (1.f + expf(-(a.ssum[row] + (a.s_bias ? a.s_bias[0] : 0.f)))) : a.sig[row]);
    const float gbase = rv ? (a.beta * sgv + (a.tau ? a.gamma * a.tau[b] : 0.f) + 1.f - a.alpha) : 0.f;
    float v[VPL][8];
    uint4 raw[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = sub + LPR * i;
      raw[i] = (rv && vi < nvec) ? __ldg(reinterpret_cast<const uint4*>(x1 + row * a.C + vi * 8)) : make_uint4(0, 0, 0, 0);
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = sub + LPR * i;
      unpack8(raw[i], v[i]);
      if (rv && vi < nvec) {
        float mm[8];
        ldf8(a.m + (size_t)b * a.C + vi * 8, mm);
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[i][j] *= fmaf(a.alpha, mm[j], gbase); sum += v[i][j]; }
      }
    }
    float mean = 0.f, rstd = 1.f;
    if (a.ln_w) {
      mean = group_sum<LPR>(sum) * invC;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i)
        if (sub + LPR * i < nvec) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float dlt = v[i][j] - mean; sq += dlt * dlt; }
        }
      rstd = rsqrtf(group_sum<LPR>(sq) * invC + a.eps);
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = sub + LPR * i;
      if (rv && vi < nvec) {
        if (a.ln_w) {
          float w[8], bb[8];
          ldf8(a.ln_w + vi * 8, w);
          ldf8(a.ln_b + vi * 8, bb);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] = fmaf((v[i][j] - mean) * rstd, w[j], bb[j]);
        }
        *reinterpret_cast<uint4*>(y + row * a.C + vi * 8) = pack8(v[i]);
        if (a.y_tap) {
#pragma unroll
          for (int j = 0; j < 8; ++j) a.y_tap[row * a.C + vi * 8 + j] = v[i][j];
        }
      }
    }
  }
}

template <int LPR, int VPL>
__global__ void __launch_bounds__(256) out_ln_vec_kernel(OutLnArgs a) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR, rsub = lane / LPR;
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const int nvec = a.C >> 3;
  const float invC = 1.f / (float)a.C;
  const float* u = reinterpret_cast<const float*>(a.u);
  bf16* out = reinterpret_cast<bf16*>(a.out);
  const float g = a.gate ? a.gate[0] : 1.f;
  const float gpre = a.gate_before_ln ? g : 1.f, gpost = a.gate_before_ln ? 1.f : g;
  for (long long base = warp0 * RPW; base < a.rows; base += nwarps * RPW) {
    const long long row = base + rsub;
    const bool rv = row < a.rows;
    float v[VPL][8];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = sub + LPR * i;
      if (rv && vi < nvec) {
        ldf8(u + row * a.C + vi * 8, v[i]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
      }
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = sub + LPR * i;
      if (rv && vi < nvec) {
        float sc[8], sh[8];
        ldf8(a.scale + vi * 8, sc);
        ldf8(a.shift + vi * 8, sh);
#pragma unroll
        for (int j = 0; j < 8; ++j) { v[i][j] = fmaf(v[i][j], sc[j], sh[j]) * gpre; sum += v[i][j]; }
      }
    }
    float mean = 0.f, rstd = 1.f;
    if (a.ln_w) {
      mean = group_sum<LPR>(sum) * invC;
      float sq = 0.f;
#pragma unroll
      for (int i = 0; i < VPL; ++i)
        if (sub + LPR * i < nvec) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float dlt = v[i][j] - mean; sq += dlt * dlt; }
        }
      rstd = rsqrtf(group_sum<LPR>(sq) * invC + a.eps);
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = sub + LPR * i;
      if (rv && vi < nvec) {
        if (a.ln_w) {
          float w[8], bb[8];
          ldf8(a.ln_w + vi * 8, w);
          ldf8(a.ln_b + vi * 8, bb);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] = fmaf((v[i][j] - mean) * rstd, w[j], bb[j]) * gpost;
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] *= gpost;
        }
        if (a.resid) {
          float rr[8];
          unpack8(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(a.resid) + row * a.C + vi * 8)), rr);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[i][j] += rr[j];
        }
        *reinterpret_cast<uint4*>(out + row * a.C + vi * 8) = pack8(v[i]);
      }
    }
  }
}

// (LPR, VPL) for a row of nvec 16-byte vectors: an exact power-of-two split with 3-4 vectors per lane when one exists
// (every Swin-V2 / HTS-AT / CLIP width does), otherwise a whole warp per row with ceil(nvec/32) guarded vectors.
struct RowGeom { int lpr, vpl; };
inline RowGeom row_geometry(int nvec) {
  const int cand[] = {4, 3, 6, 8, 2, 1};
  for (int vpl : cand) {
    if (nvec % vpl) continue;
    const int lpr = nvec / vpl;
    if (lpr <= 32 && lpr >= 1 && (lpr & (lpr - 1)) == 0) return RowGeom{lpr, vpl};
  }
  return RowGeom{32, (nvec + 31) / 32};
}

// Invokes f.template run<LPR, VPL>() for the geometry of `nvec`; false when no instantiation covers it.
template <typename F>
inline bool dispatch_row_geometry(int nvec, F&& f) {
  const RowGeom g = row_geometry(nvec);
#define DGSCT_ROW_CASE(L_, V_) if (g.lpr == L_ && g.vpl == V_) { f.template run<L_, V_>(); return true; }
  DGSCT_ROW_CASE(1, 1) DGSCT_ROW_CASE(1, 2) DGSCT_ROW_CASE(1, 3) DGSCT_ROW_CASE(1, 4) DGSCT_ROW_CASE(1, 6)
  DGSCT_ROW_CASE(2, 1) DGSCT_ROW_CASE(2, 3) DGSCT_ROW_CASE(2, 4)
  DGSCT_ROW_CASE(4, 3) DGSCT_ROW_CASE(4, 4) DGSCT_ROW_CASE(8, 3) DGSCT_ROW_CASE(8, 4) DGSCT_ROW_CASE(16, 3) DGSCT_ROW_CASE(16, 4)
  DGSCT_ROW_CASE(32, 1) DGSCT_ROW_CASE(32, 2) DGSCT_ROW_CASE(32, 3) DGSCT_ROW_CASE(32, 4) DGSCT_ROW_CASE(32, 5) DGSCT_ROW_CASE(32, 6)
  DGSCT_ROW_CASE(32, 7) DGSCT_ROW_CASE(32, 8)
#undef DGSCT_ROW_CASE
  return false;
}

// "Light" geometry for the register-heavy backward row kernels: as few vectors per lane as possible (a whole warp per row once
// the row has more than 16 vectors), trading idle lanes (C = 192 / 384: 25 %) for ~half the registers per thread, i.e. two
// resident blocks per SM instead of one (the column accumulators alone are 3-4 x VPL x 8 registers).
inline RowGeom row_geometry_light(int nvec) {
  int lpr = 32;
  while (lpr > 1 && lpr / 2 >= nvec) lpr >>= 1;
  return RowGeom{lpr, (nvec + lpr - 1) / lpr};
}
template <typename F>
inline bool dispatch_row_geometry_light(int nvec, F&& f) {
  const RowGeom g = row_geometry_light(nvec);
#define DGSCT_ROW_CASE(L_, V_) if (g.lpr == L_ && g.vpl == V_) { f.template run<L_, V_>(); return true; }
  DGSCT_ROW_CASE(1, 1) DGSCT_ROW_CASE(2, 1) DGSCT_ROW_CASE(4, 1) DGSCT_ROW_CASE(8, 1) DGSCT_ROW_CASE(16, 1)
  DGSCT_ROW_CASE(32, 1) DGSCT_ROW_CASE(32, 2) DGSCT_ROW_CASE(32, 3) DGSCT_ROW_CASE(32, 4)
#undef DGSCT_ROW_CASE
  return false;
}

inline unsigned row_grid(long long rows, int lpr) {
  const long long rows_per_block = 8LL * (32 / lpr);
  const long long blocks_needed = (rows + rows_per_block - 1) / rows_per_block;
  const long long cap = 148LL * 16;
  return (unsigned)(blocks_needed < cap ? (blocks_needed < 1 ? 1 : blocks_needed) : cap);
}

struct GateLnLaunch {
  const GateLnArgs& a; cudaStream_t st;
  template <int LPR, int VPL> void run() { gate_ln_vec_kernel<LPR, VPL><<<row_grid(a.rows, LPR), 256, 0, st>>>(a); }
};
struct OutLnLaunch {
  const OutLnArgs& a; cudaStream_t st;
  template <int LPR, int VPL> void run() { out_ln_vec_kernel<LPR, VPL><<<row_grid(a.rows, LPR), 256, 0, st>>>(a); }
};

// launch the vectorised kernel when the row width allows it; return false if the caller must use the scalar one
inline bool launch_gate_ln_vec(const GateLnArgs& a, cudaStream_t st) {
  if (a.C % 8 || !a.in_bf16 || !a.out_bf16 || a.C > 2048) return false;
  return dispatch_row_geometry(a.C / 8, GateLnLaunch{a, st});
}
inline bool launch_out_ln_vec(const OutLnArgs& a, cudaStream_t st) {
  if (a.C % 8 || a.in_bf16 || !a.out_bf16 || a.C > 2048) return false;
  return dispatch_row_geometry(a.C / 8, OutLnLaunch{a, st});
}

// ---- sparse token resampler (AVS-S4 bicubic stencil, PVT_AVSModel.py:190-197) --------------------------------------------
// out[b, r, :] = sum_j val[j] * in[b, col[j], :]  for the CSR row r (<= 16 taps): one thread per 16-byte vector of an output
// row, fp32 coefficients and accumulation.  HBM-bound: reads <= 16 (L2-resident) input rows per output row, writes once.
template <bool BF16>
__global__ void __launch_bounds__(256) spmm_frames_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                          const float* __restrict__ val, const void* __restrict__ in,
                                                          void* __restrict__ out, int rows_out, int rows_in, int Cc, int nvec) {
  constexpr int VE = BF16 ? 8 : 4;               // elements per 16-byte vector
  const int b = blockIdx.y;
  const long long total = (long long)rows_out * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / nvec), v = (int)(i - (long long)r * nvec);
    float acc[VE];
#pragma unroll
    for (int e = 0; e < VE; ++e) acc[e] = 0.f;
    const int j0 = rowptr[r], j1 = rowptr[r + 1];
    for (int j = j0; j < j1; ++j) {
      const float w = val[j];
      const size_t off = ((size_t)b * rows_in + col[j]) * Cc + (size_t)v * VE;
      if (BF16) {
        const uint4 u = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(in) + off);
        const uint32_t q[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&q[e]);
          acc[2 * e] = fmaf(w, __bfloat162float(t.x), acc[2 * e]);
          acc[2 * e + 1] = fmaf(w, __bfloat162float(t.y), acc[2 * e + 1]);
        }
      } else {
        const float4 u = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(in) + off);
        acc[0] = fmaf(w, u.x, acc[0]); acc[1] = fmaf(w, u.y, acc[1]); acc[2] = fmaf(w, u.z, acc[2]); acc[3] = fmaf(w, u.w, acc[3]);
      }
    }
    const size_t o = ((size_t)b * rows_out + r) * Cc + (size_t)v * VE;
    if (BF16) {
      uint4 u;
      __nv_bfloat162 t;
      t = __floats2bfloat162_rn(acc[0], acc[1]); u.x = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2bfloat162_rn(acc[2], acc[3]); u.y = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2bfloat162_rn(acc[4], acc[5]); u.z = *reinterpret_cast<uint32_t*>(&t);
      t = __floats2bfloat162_rn(acc[6], acc[7]); u.w = *reinterpret_cast<uint32_t*>(&t);
      *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(out) + o) = u;
    } else {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + o) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
  }
}

// Cc must be a multiple of 8 (bf16) / 4 (fp32): every width of the reference geometries is.  Returns false otherwise.
inline bool launch_spmm_frames(const int32_t* rowptr, const int32_t* col, const float* val, const void* in, void* out, int is_bf16,
                               int BT, int rows_out, int rows_in, int Cc, cudaStream_t st) {
  const int ve = is_bf16 ? 8 : 4;
  if (Cc % ve) return false;
  const int nvec = Cc / ve;
  const long long total = (long long)rows_out * nvec;
  int gx = (int)((total + 255) / 256);
  const int cap = (148 * 8 + BT - 1) / BT;        // ~8 CTAs per SM over the whole (gx, BT) grid
  if (gx > cap) gx = cap < 1 ? 1 : cap;
  if (is_bf16) spmm_frames_kernel<true><<<dim3(gx, BT), 256, 0, st>>>(rowptr, col, val, in, out, rows_out, rows_in, Cc, nvec);
  else spmm_frames_kernel<false><<<dim3(gx, BT), 256, 0, st>>>(rowptr, col, val, in, out, rows_out, rows_in, Cc, nvec);
  return true;
}

}  // namespace dgsct
