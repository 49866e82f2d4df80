// Debug / bring-up entry points of the C ABI (used by tests/test_gpu_gemm_tc.py).
#include "epilogues.cuh"
#include "tc_epilogues.cuh"

using namespace dgsct;

extern "C" {

// D[b] (M,N) fp32 = A[b] . B[b]^T with bf16 operands.  K-major operand: element (row, k) at row*ld + k;
// MN-major operand: element (row, k) at k*ld + row.  batch strides in elements (0 = shared).
int dgsct_debug_gemm_bf16(const void* A, const void* B, float* D, int M, int N, int K, int batch, int a_mn_major,
                          int b_mn_major, long long lda, long long ldb, long long sa, long long sb, int use_tc,
                          int bn_hint, int stages_hint, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (!A || !B || !D) return DGSCT_ERR_NULL;
  if (use_tc) {
    tc::OperandDesc a{A, a_mn_major, lda, sa, a_mn_major ? K : M, a_mn_major ? M : K};
    tc::OperandDesc b{B, b_mn_major, ldb, sb, b_mn_major ? K : N, b_mn_major ? N : K};
    tc::Params p;
    // (staging sized for the fp32 functor; the bf16 one needs half of it)
    int rc = tc::make_params(p, a, b, M, N, K, batch, bn_hint & 0xFFFF, stages_hint, false, true, 4096);
    if (rc) return rc;
    // bn_hint bit 16: store D as bf16 into the same buffer (timing experiments); bits 20..21: epilogue ablation
    const int out_bf16 = (bn_hint >> 16) & 1;
    p.debug_mode = (bn_hint >> 20) & 15;   // bits 0-1: epilogue ablation, bit 2: no MMAs, bit 3: no TMA loads
    cudaError_t err;
    if (out_bf16) {
      tc::EpiStore<true, ACT_NONE, false, false> e{};
      e.alpha = 1.f; e.out = D; e.ld = N; e.sb = (long long)M * N;
      if ((bn_hint >> 24) & 1) e.out = nullptr;   // timing experiment: no global stores
      if ((bn_hint >> 25) & 1) e.sb = 0;          // timing experiment: every frame writes the same (L2-resident) tile
      err = tc::launch(p, e, st);
    } else {
      tc::EpiStore<false, ACT_NONE, false, false> e{};
      e.alpha = 1.f; e.out = D; e.ld = N; e.sb = (long long)M * N;
      err = tc::launch(p, e, st);
    }
    if (err != cudaSuccess) { set_cuda_error(err, "debug_gemm_tc"); return DGSCT_ERR_CUDA; }
    return DGSCT_OK;
  }
  GemmShape sh{M, N, K, batch, 1};
  GemmOperand a{A, a_mn_major ? 1 : lda, a_mn_major ? lda : 1, sa, 0, 1};
  GemmOperand b{B, b_mn_major ? 1 : ldb, b_mn_major ? ldb : 1, sb, 0, 1};
  EpiBiasStore e{};
  e.out = D; e.out_bf16 = 0; e.ld = N; e.sb = (long long)M * N; e.alpha = 1.f;
  cudaError_t err = launch_gemm_simt(sh, a, b, AxIdentity{}, e, st);
  if (err != cudaSuccess) { set_cuda_error(err, "debug_gemm_simt"); return DGSCT_ERR_CUDA; }
  return DGSCT_OK;
}

// value of the sticky pipeline-timeout flag (0 = healthy); synchronises the current device first so that every kernel
// launched so far has had the chance to raise it; `reset` clears it
int dgsct_tc_error_flag(int reset) {
  if (cudaDeviceSynchronize() != cudaSuccess) return -1;
  return timeout_flag_host_value(reset);
}

}  // extern "C"
