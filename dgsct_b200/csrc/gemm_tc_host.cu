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

// The bounded mbarrier waits of the GEMM raise this flag instead of hanging.  It lives in host-mapped (zero-copy, portable)
// memory so that every entry point of the library can test it WITHOUT synchronising: a pipeline timeout in an earlier call
// turns into DGSCT_ERR_TIMEOUT on the next call (sticky, like a CUDA error) instead of DGSCT_OK with garbage.
int* tc_err_flag_device_ptr() { return timeout_flag_device_ptr(); }

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

size_t smem_bytes(const Params& p) {
  const int num_k = (p.K + p.BK - 1) / p.BK, n_blocks = (p.N + p.BN - 1) / p.BN;
  const size_t ring = (size_t)p.stages * (p.a_stage_bytes + (p.b_resident ? 0 : p.b_stage_bytes));
  const size_t bres = p.b_resident ? (size_t)n_blocks * num_k * p.b_stage_bytes : 0;
  return 1024 + ring + bres + 256 + (size_t)EPI_WARPS * p.stage_out + (size_t)2 * p.stats_cols * 4;
}

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

int make_params(Params& p, const OperandDesc& A_in, const OperandDesc& B_in, int M, int N, int K, int batch, int BN_hint,
                int stages_hint, bool sum_over_batch, bool allow_flatten, int stage_bytes_per_warp, int block_stats_cols) {
  if (!operands_ok(A_in, B_in, K)) return DGSCT_ERR_UNSUPPORTED;
  OperandDesc A = A_in, B = B_in;
  // Weight gradients sum_b A_b^T B_b with token-major operands whose frames are contiguous: one flattened K extent of batch*K
  // rows cut into equal chunks (see Params::kflat).  Chosen so that every CTA slot gets a chunk of >= 8 k-blocks.
  p.kflat = 0;
  int kflat_chunk = 0, kflat_groups = 0;
  if (sum_over_batch && A.mn_major && B.mn_major && batch > 1 && A.batch_stride == (long long)K * A.ld &&
      B.batch_stride == (long long)K * B.ld && (long long)K * batch < (1LL << 30) && !getenv("DGSCT_NO_KFLAT")) {
    const long long rows = (long long)K * batch;
    const int bk = 64;
    const long long nk = (rows + bk - 1) / bk;
    int bn = BN_hint > 0 ? BN_hint : (N <= 256 ? round_up(N, 16) : round_up((N + (N + 255) / 256 - 1) / ((N + 255) / 256), 16));
    const long long tiles = (long long)((M + BM - 1) / BM) * ((N + bn - 1) / bn);
    long long groups = (2LL * num_sms() + tiles - 1) / tiles;
    if (groups > nk / 8) groups = nk / 8;
    if (groups < 1) groups = 1;
    const long long kbpg = (nk + groups - 1) / groups;
    groups = (nk + kbpg - 1) / kbpg;
    kflat_chunk = (int)(kbpg * bk);
    kflat_groups = (int)groups;
    A.rows = (int)rows; A.batch_stride = 0;
    B.rows = (int)rows; B.batch_stride = 0;
    K = kflat_chunk;
    batch = kflat_groups;
    p.kflat = 1;
  }
  // Frames flattened into one M extent: A[b] (M, K) rows are contiguous across frames and B is shared, so
  // (batch, M) is just (batch*M) rows -- no 128-row tile padding per frame (N = 36/64/144 tokens at the deep stages).
  p.rows_per_frame = 0;
  if (allow_flatten && !A.mn_major && A.batch_stride != 0 && B.batch_stride == 0 && A.batch_stride == (long long)M * A.ld &&
      !sum_over_batch && batch > 1 && (long long)M * batch < (1LL << 30) && !getenv("DGSCT_NO_FLATTEN")) {
    p.rows_per_frame = M;
    A.rows = M * batch;
    A.batch_stride = 0;
    M = M * batch;
    batch = 1;
  }
  p.M = M; p.N = N; p.K = K; p.batch = batch;
  p.a_mn_major = A.mn_major; p.b_mn_major = B.mn_major;
  p.a_batched = A.batch_stride != 0; p.b_batched = B.batch_stride != 0;
  // tile N
  int BN;
  if (BN_hint > 0) BN = BN_hint;
  else if (N <= 256) BN = round_up(N, 16);
  else { const int nb = (N + 255) / 256; BN = round_up((N + nb - 1) / nb, 16); }
  {   // Medium-K products take tiles of at most 128 columns: twice the CTAs of a 256-column tile and a double-buffered
      // accumulator (epilogue under the next main loop).  Measured on the train step (B200, round 2): cap 256 -> 56.4 ms,
      // 192 -> 56.8, 128 -> 54.3, 96 -> 54.2, 64 -> 55.2.  DGSCT_TC_BN_MAX overrides (256 = the old rule).
    static const int bn_max = [] { const char* e = getenv("DGSCT_TC_BN_MAX"); const int v = e ? atoi(e) : 128; return (v >= 32 && v <= 256) ? v / 16 * 16 : 128; }();
    if (BN_hint <= 0 && !sum_over_batch && BN > bn_max && (K + 63) / 64 <= 32) {
      const int nb = (N + bn_max - 1) / bn_max;
      BN = round_up((N + nb - 1) / nb, 16);
    }
  }
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
  const int num_k = (K + p.BK - 1) / p.BK;
  const int n_blocks = (N + BN - 1) / BN;
  if (p.kflat) {
    p.kb_frames = 1;              // one K chunk per tile; the epilogue receives the chunk index and adds atomically
  } else if (sum_over_batch) {
    // frames per tile: enough frame groups to fill ~2 tiles per SM, the rest of the batch is accumulated in the tile
    const long long tiles = (long long)((M + BM - 1) / BM) * n_blocks;
    long long groups = (2LL * num_sms() + tiles - 1) / tiles;
    if (groups < 1) groups = 1;
    if (groups > batch) groups = batch;
    p.kb_frames = (int)((batch + groups - 1) / groups);
  }
  const long long k_iters = (long long)num_k * (p.kb_frames > 0 ? p.kb_frames : 1);
  const bool short_k = k_iters < 8;
  const int BNI = BN * p.fpack;
  // Short-K products are bound by shared-memory fill traffic and by the epilogue's global stores, not by the MMAs:
  //  * a shared B operand that fits stays resident for the CTA's lifetime (it used to be re-streamed for every tile);
  //  * the output tile is staged through shared memory for row-contiguous stores;
  //  * two CTAs per SM (when both keep a double-buffered accumulator) hide each other's TMA round trips.
  p.stage_out = (short_k && p.fpack == 1 && !getenv("DGSCT_NO_STAGE_OUT")) ? stage_bytes_per_warp : 0;
  p.stats_cols = block_stats_cols > 0 ? N : 0;
  const size_t out_stage = (size_t)EPI_WARPS * p.stage_out + (size_t)2 * p.stats_cols * 4;
  const size_t bres_total = (size_t)n_blocks * num_k * p.b_stage_bytes;
  const bool bres_possible = !p.b_batched && !sum_over_batch && p.fpack == 1 && !getenv("DGSCT_NO_BRES");
  const char* force = getenv("DGSCT_TC_CTAS");
  const size_t cta_budget[3] = {0, 220 * 1024, 110 * 1024};
  auto plan = [&](int ctas, bool bres, int& stages_out) {
    const size_t fixed = 1024 + 256 + out_stage + (bres ? bres_total : 0);
    if (fixed >= cta_budget[ctas]) return false;
    const size_t ring_stage = (size_t)p.a_stage_bytes + (bres ? 0 : p.b_stage_bytes);
    long long st = (long long)((cta_budget[ctas] - fixed) / ring_stage);
    if (st > MAX_STAGES) st = MAX_STAGES;
    stages_out = (int)st;
    return st >= 2;
  };
  int ctas = 1, stages = 0;
  bool bres = false;
  if (short_k) {
    // measured (scripts/gemm_skinny.py): a second CTA per SM is worth more than a deeper ring or a resident B, so it
    // comes first; B stays resident whenever it fits next to >= 3 A stages
    const bool two_ok = BNI <= 128;                      // 2 x (double-buffered accumulator <= 256 columns) per SM
    int st = 0;
    if (two_ok && bres_possible && bres_total <= 48 * 1024 && plan(2, true, st) && st >= 3) { ctas = 2; bres = true; stages = st; }
    else if (two_ok && plan(2, false, st)) { ctas = 2; stages = st; }
    else if (bres_possible && bres_total <= 112 * 1024 && plan(1, true, st) && st >= 3) { ctas = 1; bres = true; stages = st; }
    else if (plan(1, false, st)) { ctas = 1; stages = st; }
    else return DGSCT_ERR_UNSUPPORTED;
  } else {
    // Long-K products (main loop dominated): two half-smem CTAs per SM when two stages each fit -- two independent
    // TMA/MMA streams hide each other's barrier round trips (968 vs 713 TFLOP/s on the stage-0 resampler).
    int st = 0;
    const size_t stage_bytes = (size_t)p.a_stage_bytes + p.b_stage_bytes;
    // (tried in round 2: one big-smem CTA per SM with the deepest ring for grids that fit in one wave -- 56.4 ms against
    //  55.5 ms per train step without it; two half-smem CTAs per SM stay the rule)
    if (2 * stage_bytes <= 104 * 1024 && plan(2, false, st)) { ctas = 2; stages = st; }
    else if (plan(1, false, st)) { ctas = 1; stages = st; }
    else return DGSCT_ERR_UNSUPPORTED;
  }
  if (force) {
    const int f = atoi(force);
    int st = 0;
    if ((f == 1 || f == 2) && f != ctas && plan(f, bres && (f == 1 || bres_total <= 48 * 1024), st)) {
      bres = bres && (f == 1 || bres_total <= 48 * 1024);
      ctas = f; stages = st;
    }
  }
  if (stages_hint > 0 && stages_hint < stages) stages = stages_hint < 2 ? 2 : stages_hint;
  p.ctas_per_sm = ctas;
  p.b_resident = bres ? 1 : 0;
  p.stages = stages;
  if (smem_bytes(p) > 227 * 1024) return DGSCT_ERR_UNSUPPORTED;
  const int max_cols = p.ctas_per_sm == 2 ? 256 : 512;
  p.acc_bufs = (2 * BNI <= max_cols) ? 2 : 1;
  uint32_t cols = 32;
  while ((int)cols < p.acc_bufs * BNI) cols <<= 1;
  p.tmem_cols = cols;
  p.err_flag = tc_err_flag_device_ptr();
  p.debug_mode = 0;
  if (sum_over_batch && getenv("DGSCT_DW_EPI_SKIP")) p.debug_mode = 2;   // measurement only: weight-gradient main loop without its epilogue
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
