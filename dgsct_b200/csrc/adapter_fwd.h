// Internal interface between the C ABI (api.cu) and the forward pipelines.
#pragma once
#include "common.cuh"

namespace dgsct {

// byte offsets into the caller-provided workspace (all 256-B aligned)
struct FwdPlan {
  size_t bufA, bufV, bufX1, bufL, Tp, Z;
  size_t acc_begin, osum, qsum, ssum, bn1_sum, bn1_sq, bn2_sum, bn2_sq, acc_end;
  size_t aq1, aq2, h, m, tau, sig, bn1_scale, bn1_shift, bn2_scale, bn2_shift;
  size_t Pb, TpB, Zr, Uo32, packed, aq, vec_bf;   // aq: [aq1 | aq2] fp32 (BT, C+H); vec_bf: bf16 per-frame vectors   // tensor-core path: bf16 P/Q, bf16 latent tokens, bf16 relu(bn1(Z)), packed weights
  size_t total;
};

int validate_desc(const dgsct_desc* d);
FwdPlan make_fwd_plan(const dgsct_desc& d);

// bf16 tensor-core pipeline (tcgen05 GEMMs); `packed` = pre-packed bf16 weights or NULL (packed per call)
int adapter_fwd_tc(const dgsct_desc& d, const dgsct_params& P, const void* packed, const void* x, const void* other,
                   void* out, void* smap, float* tmap, const dgsct_taps* taps, char* ws, cudaStream_t st);
int pack_params(const dgsct_desc& d, const dgsct_params& P, char* packed, cudaStream_t st);
size_t packed_bytes(const dgsct_desc& d);

int adapter_fwd_simt(const dgsct_desc& d, const dgsct_params& P, const void* x, const void* other, void* out,
                     void* smap, float* tmap, const dgsct_taps* taps, char* ws, cudaStream_t st);

}  // namespace dgsct
