// what sits directly behind the adapters in the task networks.
//   * token pooling with the LAST adapters' spatial maps:  f = bmm(spatial_att_maps (BT,1,N), f (BT,N,C))
//     (DG-SCT/AVE/nets/net_trans.py:922,924) -- this is what makes the adapter's second return value load-bearing;
//   * the sigmoid gates at the end of TemporalAttention.forward (net_trans.py:243-249; same idiom in the S4 copy,
//     PVT_AVSModel.py:546-580):  ag = sigmoid(a_kv wa + ba), vg = sigmoid(v_kv wv + bv), av = ag * vg,
//     vq' = vq + ag * vq * gamma,  aq' = aq + vg * aq * gamma.
// Both are HBM-bound row kernels (one warp per row, shuffle reductions), forward and backward; the LSTM / transformer part of
// TemporalAttention stays PyTorch.
#include "common.cuh"
#include "profile.h"

namespace dgsct {

// ---- token pooling ------------------------------------------------------------------------------------------------
// out[b, c] = sum_n map[b, n] * f[b, n, c]        grid (ceil(C / 256), BT): a thread owns one channel, coalesced over c
__global__ void __launch_bounds__(256) pool_fwd_kernel(const void* __restrict__ map, const void* __restrict__ f, void* __restrict__ out,
                                                       int N, int C, int bf) {
  const int b = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
  extern __shared__ float smap[];
  for (int n = threadIdx.x; n < N; n += blockDim.x) smap[n] = ld_act(map, (size_t)b * N + n, bf);
  __syncthreads();
  if (c >= C) return;
  float acc = 0.f;
  for (int n = 0; n < N; ++n) acc = fmaf(smap[n], ld_act(f, ((size_t)b * N + n) * C + c, bf), acc);
  st_act(out, (size_t)b * C + c, acc, bf);
}
// d_f[b,n,c] = map[b,n] * d_out[b,c] ; d_map[b,n] = sum_c d_out[b,c] * f[b,n,c]       one warp per (b, n) row
__global__ void __launch_bounds__(256) pool_bwd_kernel(const void* __restrict__ map, const void* __restrict__ f,
                                                       const void* __restrict__ d_out, void* __restrict__ d_map, void* __restrict__ d_f,
                                                       long long rows, int N, int C, int bf) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += (long long)gridDim.x * wpb) {
    const long long b = row / N;
    const float w = ld_act(map, (size_t)row, bf);
    float dot = 0.f;
    for (int c = lane; c < C; c += 32) {
      const float g = ld_act(d_out, (size_t)b * C + c, bf);
      dot = fmaf(g, ld_act(f, (size_t)row * C + c, bf), dot);
      if (d_f) st_act(d_f, (size_t)row * C + c, w * g, bf);
    }
    dot = warp_sum(dot);
    if (lane == 0 && d_map) st_act(d_map, (size_t)row, dot, bf);
  }
}

// ---- TemporalAttention gates ------------------------------------------------------------------------------------------
// one warp per (t, b) row of D channels
__global__ void __launch_bounds__(256) tgate_fwd_kernel(const void* __restrict__ a_kv, const void* __restrict__ v_kv,
                                                        const void* __restrict__ vq, const void* __restrict__ aq,
                                                        const float* __restrict__ wa, const float* __restrict__ ba,
                                                        const float* __restrict__ wv, const float* __restrict__ bv, float gamma,
                                                        void* __restrict__ vq_out, void* __restrict__ aq_out, void* __restrict__ av_gate,
                                                        float* __restrict__ gates /* (rows, 2) fp32: saved for the backward */,
                                                        int rows, int D, int bf) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
    float sa = 0.f, sv = 0.f;
    for (int c = lane; c < D; c += 32) {
      sa = fmaf(ld_act(a_kv, (size_t)row * D + c, bf), wa[c], sa);
      sv = fmaf(ld_act(v_kv, (size_t)row * D + c, bf), wv[c], sv);
    }
    const float ag = sigmoidf_(warp_sum(sa) + ba[0]), vg = sigmoidf_(warp_sum(sv) + bv[0]);
    const float fv = 1.f + gamma * ag, fa = 1.f + gamma * vg;
    for (int c = lane; c < D; c += 32) {
      st_act(vq_out, (size_t)row * D + c, ld_act(vq, (size_t)row * D + c, bf) * fv, bf);
      st_act(aq_out, (size_t)row * D + c, ld_act(aq, (size_t)row * D + c, bf) * fa, bf);
    }
    if (lane == 0) {
      st_act(av_gate, (size_t)row, ag * vg, bf);
      if (gates) { gates[2 * row] = ag; gates[2 * row + 1] = vg; }
    }
  }
}
// backward; parameter gradients (d_wa, d_ba, d_wv, d_bv) are ADDED (fp32 atomics), like every dgsct gradient
__global__ void __launch_bounds__(256) tgate_bwd_kernel(const void* __restrict__ a_kv, const void* __restrict__ v_kv,
                                                        const void* __restrict__ vq, const void* __restrict__ aq,
                                                        const float* __restrict__ wa, const float* __restrict__ wv, float gamma,
                                                        const float* __restrict__ gates, const void* __restrict__ d_vq_out,
                                                        const void* __restrict__ d_aq_out, const void* __restrict__ d_av,
                                                        void* __restrict__ d_a_kv, void* __restrict__ d_v_kv, void* __restrict__ d_vq,
                                                        void* __restrict__ d_aq, float* __restrict__ d_wa, float* __restrict__ d_ba,
                                                        float* __restrict__ d_wv, float* __restrict__ d_bv, int rows, int D, int bf) {
  extern __shared__ float sm[];                  // [2][D] block partials of d_wa, d_wv ; [2] of d_ba, d_bv
  for (int c = threadIdx.x; c < 2 * D + 2; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int row = blockIdx.x * wpb + (threadIdx.x >> 5); row < rows; row += gridDim.x * wpb) {
    const float ag = gates[2 * row], vg = gates[2 * row + 1];
    const float fv = 1.f + gamma * ag, fa = 1.f + gamma * vg;
    float dv = 0.f, da = 0.f;
    for (int c = lane; c < D; c += 32) {
      const float gv = ld_act(d_vq_out, (size_t)row * D + c, bf), ga = ld_act(d_aq_out, (size_t)row * D + c, bf);
      dv = fmaf(gv, ld_act(vq, (size_t)row * D + c, bf), dv);
      da = fmaf(ga, ld_act(aq, (size_t)row * D + c, bf), da);
      st_act(d_vq, (size_t)row * D + c, gv * fv, bf);
      st_act(d_aq, (size_t)row * D + c, ga * fa, bf);
    }
    dv = warp_sum(dv); da = warp_sum(da);
    const float dav = d_av ? ld_act(d_av, (size_t)row, bf) : 0.f;
    const float d_ag = gamma * dv + dav * vg, d_vg = gamma * da + dav * ag;
    const float d_sa = d_ag * ag * (1.f - ag), d_sv = d_vg * vg * (1.f - vg);
    for (int c = lane; c < D; c += 32) {
      const float xa = ld_act(a_kv, (size_t)row * D + c, bf), xv = ld_act(v_kv, (size_t)row * D + c, bf);
      st_act(d_a_kv, (size_t)row * D + c, d_sa * wa[c], bf);
      st_act(d_v_kv, (size_t)row * D + c, d_sv * wv[c], bf);
      atomicAdd(&sm[c], d_sa * xa);
      atomicAdd(&sm[D + c], d_sv * xv);
    }
    if (lane == 0) { atomicAdd(&sm[2 * D], d_sa); atomicAdd(&sm[2 * D + 1], d_sv); }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    if (d_wa) atomicAdd(&d_wa[c], sm[c]);
    if (d_wv) atomicAdd(&d_wv[c], sm[D + c]);
  }
  if (threadIdx.x == 0) {
    if (d_ba) atomicAdd(d_ba, sm[2 * D]);
    if (d_bv) atomicAdd(d_bv, sm[2 * D + 1]);
  }
}

}  // namespace dgsct

using namespace dgsct;

extern "C" {

int dgsct_pool_tokens_fwd(const void* spatial_map, const void* f, void* out, int BT, int N, int C, int io_dtype, void* stream) {
  if (!spatial_map || !f || !out) return DGSCT_ERR_NULL;
  if (BT <= 0 || N <= 0 || C <= 0 || (io_dtype != DGSCT_F32 && io_dtype != DGSCT_BF16) || (size_t)N * 4 > 96 * 1024) return DGSCT_ERR_BAD_DESC;
  int rc = check_arch();
  if (rc != DGSCT_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  StageScope _sc(ST_MISC, st); count_launch();
  static PerDeviceOnce attr;
  if (attr.first()) DGSCT_CUDA_CHECK(cudaFuncSetAttribute(pool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  pool_fwd_kernel<<<dim3(ceil_div(C, 256), BT), 256, (size_t)N * 4, st>>>(spatial_map, f, out, N, C, io_dtype == DGSCT_BF16);
  DGSCT_LAUNCH_CHECK("pool_fwd");
  return DGSCT_OK;
}

int dgsct_pool_tokens_bwd(const void* spatial_map, const void* f, const void* d_out, void* d_map, void* d_f, int BT, int N, int C,
                          int io_dtype, void* stream) {
  if (!spatial_map || !f || !d_out || (!d_map && !d_f)) return DGSCT_ERR_NULL;
  if (BT <= 0 || N <= 0 || C <= 0 || (io_dtype != DGSCT_F32 && io_dtype != DGSCT_BF16)) return DGSCT_ERR_BAD_DESC;
  int rc = check_arch();
  if (rc != DGSCT_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const long long rows = (long long)BT * N;
  long long blocks = (rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  StageScope _sc(ST_MISC, st); count_launch();
  pool_bwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(spatial_map, f, d_out, d_map, d_f, rows, N, C, io_dtype == DGSCT_BF16);
  DGSCT_LAUNCH_CHECK("pool_bwd");
  return DGSCT_OK;
}

int dgsct_temporal_gates_fwd(const void* a_kv, const void* v_kv, const void* vq, const void* aq, const float* wa, const float* ba,
                             const float* wv, const float* bv, float gamma, void* vq_out, void* aq_out, void* av_gate, float* gates,
                             int rows, int D, int io_dtype, void* stream) {
  if (!a_kv || !v_kv || !vq || !aq || !wa || !ba || !wv || !bv || !vq_out || !aq_out || !av_gate) return DGSCT_ERR_NULL;
  if (rows <= 0 || D <= 0 || (io_dtype != DGSCT_F32 && io_dtype != DGSCT_BF16)) return DGSCT_ERR_BAD_DESC;
  int rc = check_arch();
  if (rc != DGSCT_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  StageScope _sc(ST_MISC, st); count_launch();
  tgate_fwd_kernel<<<ceil_div(rows, 8), 256, 0, st>>>(a_kv, v_kv, vq, aq, wa, ba, wv, bv, gamma, vq_out, aq_out, av_gate, gates, rows, D,
                                                     io_dtype == DGSCT_BF16);
  DGSCT_LAUNCH_CHECK("tgate_fwd");
  return DGSCT_OK;
}

int dgsct_temporal_gates_bwd(const void* a_kv, const void* v_kv, const void* vq, const void* aq, const float* wa, const float* wv,
                             float gamma, const float* gates, const void* d_vq_out, const void* d_aq_out, const void* d_av_gate,
                             void* d_a_kv, void* d_v_kv, void* d_vq, void* d_aq, float* d_wa, float* d_ba, float* d_wv, float* d_bv,
                             int rows, int D, int io_dtype, void* stream) {
  if (!a_kv || !v_kv || !vq || !aq || !wa || !wv || !gates || !d_vq_out || !d_aq_out || !d_a_kv || !d_v_kv || !d_vq || !d_aq)
    return DGSCT_ERR_NULL;
  if (rows <= 0 || D <= 0 || D > 8192 || (io_dtype != DGSCT_F32 && io_dtype != DGSCT_BF16)) return DGSCT_ERR_BAD_DESC;
  int rc = check_arch();
  if (rc != DGSCT_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  int blocks = ceil_div(rows, 8);
  if (blocks > 148) blocks = 148;
  StageScope _sc(ST_MISC, st); count_launch();
  tgate_bwd_kernel<<<blocks, 256, (size_t)(2 * D + 2) * 4, st>>>(a_kv, v_kv, vq, aq, wa, wv, gamma, gates, d_vq_out, d_aq_out, d_av_gate,
                                                                   d_a_kv, d_v_kv, d_vq, d_aq, d_wa, d_ba, d_wv, d_bv, rows, D,
                                                                   io_dtype == DGSCT_BF16);
  DGSCT_LAUNCH_CHECK("tgate_bwd");
  return DGSCT_OK;
}

}  // extern "C"
