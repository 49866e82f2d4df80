// Optimizer step over the FLAT gradient buckets of the data-parallel wrapper (dgsct_b200/ddp.py): one fused Adam kernel per
// bucket instead of torch's multi-tensor apply over ~1400 parameter tensors (3.6 ms for 5.3 GB of traffic in round 1, 4.5x
// its HBM floor).  Reference: torch.optim.Adam(lr, betas, eps, weight_decay) as used by DG-SCT/AVE/main_trans.py:276.
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// HBM-bound: reads p, g, m, v and writes p, m, v = 28 B per parameter with fp32 gradients (26 B with bf16 gradients).
#include "common.cuh"
#include "profile.h"

namespace dgsct {

__global__ void adam_tick_kernel(int* step) { if (threadIdx.x == 0 && blockIdx.x == 0) *step += 1; }

template <bool G_BF16>
__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const void* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, size_t n4, size_t n, float lr, float b1, float b2,
                                                        float eps, float wd, float gscale, const int* __restrict__ step) {
  const float t = (float)(*step);
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  auto upd = [&](float& pp, float gg, float& mm, float& vv) {
    gg *= gscale;
    if (wd != 0.f) gg = fmaf(wd, pp, gg);
    mm = fmaf(b1, mm, (1.f - b1) * gg);
    vv = fmaf(b2, vv, (1.f - b2) * gg * gg);
    pp -= step_size * mm / (sqrtf(vv) * inv_sqrt_bc2 + eps);
  };
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
    float g0, g1, g2, g3;
    if (G_BF16) {
      const uint2 u = reinterpret_cast<const uint2*>(g)[i];
      const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
      g0 = __bfloat162float(a.x); g1 = __bfloat162float(a.y); g2 = __bfloat162float(b.x); g3 = __bfloat162float(b.y);
    } else {
      const float4 gg = reinterpret_cast<const float4*>(g)[i];
      g0 = gg.x; g1 = gg.y; g2 = gg.z; g3 = gg.w;
    }
    upd(pp.x, g0, mm.x, vv.x); upd(pp.y, g1, mm.y, vv.y); upd(pp.z, g2, mm.z, vv.z); upd(pp.w, g3, mm.w, vv.w);
    reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
  }
  // ragged tail (< 4 elements)
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float gg = G_BF16 ? __bfloat162float(reinterpret_cast<const bf16*>(g)[i]) : reinterpret_cast<const float*>(g)[i];
    upd(p[i], gg, m[i], v[i]);
  }
}

// dst (bf16) = scale * src (fp32): the communication copy of a gradient bucket
__global__ void __launch_bounds__(256) cast_bucket_kernel(const float* __restrict__ src, bf16* __restrict__ dst, size_t n4, size_t n,
                                                          float scale) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 s = reinterpret_cast<const float4*>(src)[i];
    __nv_bfloat162 a = __floats2bfloat162_rn(s.x * scale, s.y * scale), b = __floats2bfloat162_rn(s.z * scale, s.w * scale);
    uint2 u; u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    reinterpret_cast<uint2*>(dst)[i] = u;
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __float2bfloat16_rn(src[i] * scale);
}

}  // namespace dgsct

using namespace dgsct;

extern "C" {

int dgsct_adam_tick(int* step_dev, void* stream) {
  if (!step_dev) return DGSCT_ERR_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  StageScope _sc(ST_ADAM, st); count_launch();
  adam_tick_kernel<<<1, 32, 0, st>>>(step_dev);
  DGSCT_LAUNCH_CHECK("adam_tick");
  return DGSCT_OK;
}

int dgsct_adam_flat(float* p, const void* g, int g_is_bf16, float* m, float* v, size_t n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, float grad_scale, const int* step_dev, void* stream) {
  if (!p || !g || !m || !v || !step_dev) return DGSCT_ERR_NULL;
  if (((uintptr_t)p | (uintptr_t)m | (uintptr_t)v) & 15 || ((uintptr_t)g & (g_is_bf16 ? 7 : 15))) return DGSCT_ERR_WORKSPACE;
  if (n == 0) return DGSCT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n4 = n / 4;
  size_t blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  StageScope _sc(ST_ADAM, st); count_launch();
  if (g_is_bf16) adam_flat_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, n4, n, lr, beta1, beta2, eps, weight_decay, grad_scale, step_dev);
  else adam_flat_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(p, g, m, v, n4, n, lr, beta1, beta2, eps, weight_decay, grad_scale, step_dev);
  DGSCT_LAUNCH_CHECK("adam_flat");
  return DGSCT_OK;
}

int dgsct_cast_bucket_bf16(const float* src, void* dst_bf16, size_t n, float scale, void* stream) {
  if (!src || !dst_bf16) return DGSCT_ERR_NULL;
  if (((uintptr_t)src & 15) || ((uintptr_t)dst_bf16 & 7)) return DGSCT_ERR_WORKSPACE;
  if (n == 0) return DGSCT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n4 = n / 4;
  size_t blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  StageScope _sc(ST_ADAM, st); count_launch();
  cast_bucket_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, (bf16*)dst_bf16, n4, n, scale);
  DGSCT_LAUNCH_CHECK("cast_bucket");
  return DGSCT_OK;
}

}  // extern "C"
