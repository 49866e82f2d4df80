// Host side of the tcgen05 GEMM: TMA tensor-map construction and tile recipe selection.
#include "gemm_tc.cuh"

#include <stdlib.h>

#include <mutex>

namespace dgsct {
namespace tc {

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

__device__ int g_tc_err_flag = 0;

int* tc_err_flag_device_ptr() {
  static thread_local int* ptr = nullptr;
  static thread_local int dev = -1;
  int cur = 0;
  cudaGetDevice(&cur);
  if (!ptr || cur != dev) {
    void* p = nullptr;
    if (cudaGetSymbolAddress(&p, g_tc_err_flag) != cudaSuccess) return nullptr;
    ptr = reinterpret_cast<int*>(p);
    dev = cur;
  }
  return ptr;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

bool operands_ok(const OperandDesc& A, const OperandDesc& B, int K) {
  auto ok = [](const OperandDesc& o) {
    return o.ptr && aligned16(o.ptr) && (o.ld * 2) % 16 == 0 && (o.batch_stride * 2) % 16 == 0 && o.ld >= o.cols;
  };
  return ok(A) && ok(B) && K > 0 && get_encode_fn() != nullptr;
}

static int round_up(int x, int a) { return (x + a - 1) / a * a; }

static CUtensorMapSwizzle sw_enum(int sw) {
  return sw == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (sw == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

static int encode(CUtensorMap* map, const OperandDesc& o, int batch, int box0, int box1, int sw) {
  const bool batched = o.batch_stride != 0;
  cuuint64_t dims[3] = {(cuuint64_t)o.cols, (cuuint64_t)o.rows, (cuuint64_t)(batched ? batch : 1)};
  cuuint64_t strides[2] = {(cuuint64_t)o.ld * 2, (cuuint64_t)(batched ? o.batch_stride * 2 : (long long)o.rows * o.ld * 2)};
  if (strides[1] % 16) strides[1] = (strides[1] + 15) / 16 * 16;
  cuuint32_t box[3] = {(cuuint32_t)box0, (cuuint32_t)box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(o.ptr), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, sw_enum(sw), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? DGSCT_OK : DGSCT_ERR_UNSUPPORTED;
}

size_t smem_bytes(const Params& p) { return 1024 + (size_t)p.stages * (p.a_stage_bytes + p.b_stage_bytes) + 256; }

int num_sms() {
  static thread_local int n = 0, dev = -1;
  int cur = 0;
  cudaGetDevice(&cur);
  if (n == 0 || cur != dev) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, cur);
    dev = cur;
    if (n <= 0) n = 148;
  }
  return n;
}

int make_params(Params& p, const OperandDesc& A, const OperandDesc& B, int M, int N, int K, int batch, int BN_hint,
                int stages_hint, bool sum_over_batch) {
  if (!operands_ok(A, B, K)) return DGSCT_ERR_UNSUPPORTED;
  p.M = M; p.N = N; p.K = K; p.batch = batch;
  p.a_mn_major = A.mn_major; p.b_mn_major = B.mn_major;
  p.a_batched = A.batch_stride != 0; p.b_batched = B.batch_stride != 0;
  // tile N
  int BN;
  if (BN_hint > 0) BN = BN_hint;
  else if (N <= 256) BN = round_up(N, 16);
  else { const int nb = (N + 255) / 256; BN = round_up((N + nb - 1) / nb, 16); }
  if (BN < 16 || BN > 256 || BN % 16) return DGSCT_ERR_UNSUPPORTED;
  p.BN = BN;
  p.BK = (K > 32) ? 64 : 32;
  // swizzle spans: K-major rows hold BK elements; MN-major rows hold one atom of the MN extent
  p.a_sw = A.mn_major ? 128 : p.BK * 2;
  p.b_sw = B.mn_major ? ((BN % 64 == 0) ? 128 : ((BN % 32 == 0) ? 64 : 32)) : p.BK * 2;
  const int a_atom = p.a_sw / 2, b_atom = p.b_sw / 2;
  if (A.mn_major) { p.a_boxes = BM / a_atom; p.a_box_bytes = p.BK * p.a_sw; }
  else { p.a_boxes = 1; p.a_box_bytes = BM * p.BK * 2; }
  if (B.mn_major) { p.b_boxes = (BN + b_atom - 1) / b_atom; p.b_box_bytes = p.BK * p.b_sw; }
  else { p.b_boxes = 1; p.b_box_bytes = BN * p.BK * 2; }
  // frame packing: when every frame's N fits several times in one instruction (N <= 128) and the A tile is shared,
  // put f frames side by side (each frame's swizzle atoms are separate TMA boxes, so the layout is uniform)
  p.fpack = 1;
  p.kb_frames = 0;
  if (!sum_over_batch && B.mn_major && p.b_batched && !p.a_batched && N <= 128 && N == BN && batch > 1 && BN % b_atom == 0 && !getenv("DGSCT_NO_FPACK")) {
    int f = 256 / BN;
    if (f > batch) f = batch;
    if (f > 4) f = 4;
    p.fpack = f < 1 ? 1 : f;
  }
  p.a_stage_bytes = round_up(p.a_boxes * p.a_box_bytes, 1024);
  p.b_stage_bytes = round_up(p.fpack * p.b_boxes * p.b_box_bytes, 1024);
  const int stage_bytes = p.a_stage_bytes + p.b_stage_bytes;
  const int num_k = (K + p.BK - 1) / p.BK;
  // Long-K products (main loop dominated, L2-latency sensitive) run two half-smem CTAs per SM: two independent
  // TMA/MMA streams per SM hide each other's barrier round trips (968 vs 713 TFLOP/s on the stage-0 resampler).
  // Short-K products keep one big-smem CTA per SM with a double-buffered accumulator.
  if (sum_over_batch) {
    // frames per tile: enough frame groups to fill ~2 tiles per SM, the rest of the batch is accumulated in the tile
    const long long tiles = (long long)((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    long long groups = (2LL * num_sms() + tiles - 1) / tiles;
    if (groups < 1) groups = 1;
    if (groups > batch) groups = batch;
    p.kb_frames = (int)((batch + groups - 1) / groups);
  }
  const long long k_iters = (long long)num_k * (p.kb_frames > 0 ? p.kb_frames : 1);
  const char* force = getenv("DGSCT_TC_CTAS");
  p.ctas_per_sm = force ? atoi(force) : ((k_iters >= 8 && 2 * stage_bytes <= 104 * 1024) ? 2 : 1);
  if (p.ctas_per_sm != 2) p.ctas_per_sm = 1;
  const int budget = p.ctas_per_sm == 2 ? 104 * 1024 : 200 * 1024;
  int stages = stages_hint > 0 ? stages_hint : budget / stage_bytes;
  if (stages < 2) stages = 2;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  p.stages = stages;
  if (smem_bytes(p) > 227 * 1024) return DGSCT_ERR_UNSUPPORTED;
  const int BNI = BN * p.fpack;
  const int max_cols = p.ctas_per_sm == 2 ? 256 : 512;
  p.acc_bufs = (2 * BNI <= max_cols) ? 2 : 1;
  uint32_t cols = 32;
  while ((int)cols < p.acc_bufs * BNI) cols <<= 1;
  p.tmem_cols = cols;
  p.err_flag = tc_err_flag_device_ptr();
  p.debug_mode = 0;
  int rc;
  if (A.mn_major) rc = encode(&p.mapA, A, batch, a_atom, p.BK, p.a_sw);
  else rc = encode(&p.mapA, A, batch, p.BK, BM, p.a_sw);
  if (rc) return rc;
  if (B.mn_major) rc = encode(&p.mapB, B, batch, b_atom, p.BK, p.b_sw);
  else rc = encode(&p.mapB, B, batch, p.BK, BN, p.b_sw);
  return rc;
}

}  // namespace tc
}  // namespace dgsct
