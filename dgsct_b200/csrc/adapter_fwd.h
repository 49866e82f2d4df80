// Internal interface between the C ABI (api.cu) and the forward pipelines.
#pragma once
#include "common.cuh"

namespace dgsct {

// byte offsets into the caller-provided workspace (all 256-B aligned)
struct FwdPlan {
  size_t bufA, bufV, bufX1, bufL, Tp, Z;
  size_t acc_begin, osum, qsum, ssum, bn1_sum, bn1_sq, bn2_sum, bn2_sq, fa_ticket, acc_end;
  size_t aq1, aq2, h, m, tau, sig, bn1_scale, bn1_shift, bn2_scale, bn2_shift;
  size_t csm_part;                                // column-softmax partials (BT, chunks, 2, K) fp32
  size_t fc_s;                                    // row sums of fc.weight (C) fp32 (re-associated resampler, DGSCT_FC_FIRST)
  size_t Pb, TpB, Zr, Uo32, packed, aq, vec_bf, vec_bf_o;   // tensor-core path: bf16 P, bf16 latent tokens, bf16 relu(bn1(Z)), fp32 up-sampler
                                                  // output, packed weights, [aq1 | aq2] fp32 (BT, C+H), bf16 per-frame vectors
  // Regions that alias others unless the plan is built with keep = true (training forward: everything the
  // backward needs survives the call): Y (else the V buffer), Q fp32 / bf16 (else the P buffers), Xc (else U).
  size_t Ybuf, Qf, Qb, XcBuf;
  size_t r1sum, qvec;                             // sum_n relu(z1) (BT,C) ; q = aq1 * r1sum / N (BT,C) fp32
  size_t bn1_mean, bn1_rstd, bn2_mean, bn2_rstd;  // statistics used by the two BatchNorms (batch or running)
  // fused latent-token attention (attn_fused.cuh): split partials (m, l, O), per-chunk reference maxima and the P rescale
  // factors; fa_ticket (BT counters) lives in the zero-initialised region
  size_t fa_part_o, fa_part_ml, fa_mref, fa_pfac, fa_stat;   // fa_stat (BT, 2, 32): final max / sum of the token softmax (backward)
  size_t total;
};

int validate_desc(const dgsct_desc* d);
// Re-associated token resampler V = W_res (O W_fc^T) + b_res (x) rowsum(W_fc) + b_fc instead of (W_res O + b_res) W_fc^T + b_fc:
// when C < C' the N x N' contraction runs over C instead of C' columns (half the FLOPs of the forward product and of the
// dother / dW_res products for the audio direction).  Tensor-core path only.  Measured on B200 in round 2 (forward 14.45 ->
// 13.83 ms, train step 79.7 -> 78.1 ms) and promoted to the default; DGSCT_NO_FC_FIRST=1 restores the reference order.
bool use_fc_first(const dgsct_desc& d, const dgsct_params& P);
FwdPlan make_fwd_plan(const dgsct_desc& d, bool keep = false);

// bf16 tensor-core pipeline (tcgen05 GEMMs); `packed` = pre-packed bf16 weights or NULL (packed per call)
int adapter_fwd_tc(const dgsct_desc& d, const dgsct_params& P, const void* packed, const void* x, const void* other,
                   void* out, void* smap, float* tmap, const dgsct_taps* taps, char* ws, cudaStream_t st, bool keep = false,
                   const void* resid = nullptr);
int pack_params(const dgsct_desc& d, const dgsct_params& P, char* packed, cudaStream_t st);
size_t packed_bytes(const dgsct_desc& d);

int adapter_fwd_simt(const dgsct_desc& d, const dgsct_params& P, const void* x, const void* other, void* out,
                     void* smap, float* tmap, const dgsct_taps* taps, char* ws, cudaStream_t st, bool keep = false,
                     const void* resid = nullptr);

// `saved` = workspace of a forward run with the keep plan (DGSCT_F_KEEP), or NULL to recompute the forward inside `ws`
size_t bwd_workspace_bytes(const dgsct_desc& d, bool have_saved = false);
int adapter_bwd(const dgsct_desc& d, const dgsct_params& P, const void* packed, const void* x, const void* other,
                const void* dout, const void* dsmap, const float* dtmap, void* dx, void* dother, const dgsct_grads& G,
                char* ws, cudaStream_t st, const char* saved = nullptr);

}  // namespace dgsct
