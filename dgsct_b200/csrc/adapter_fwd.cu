// VisualAdapter.forward as a sequence of CUDA kernels (reference: DG-SCT/AVE/nets/net_trans.py:552-674).
// This file is the orchestration + the exact-fp32 SIMT path; the bf16 tensor-core kernels (gemm_tc.cuh) are
// substituted for the large products when io_dtype is bf16.
#include "adapter_fwd.h"

#include <stdlib.h>

#include "epilogues.cuh"
#include "pack.cuh"
#include "profile.h"
#include "tc_epilogues.cuh"
#include "rowwise.cuh"
#include "side_stream.h"
#include "attn_fused.cuh"
#include "vec_linear.cuh"

namespace dgsct {

int validate_desc(const dgsct_desc* d) {
  if (!d) return DGSCT_ERR_NULL;
  if (d->BT <= 0 || d->C <= 0 || d->N <= 0 || d->Co <= 0 || d->No <= 0 || d->K <= 0 || d->r <= 0 || d->g <= 0)
    return DGSCT_ERR_BAD_DESC;
  if (d->C % 2 || d->C % d->r || d->C % d->g || (d->C / d->r) % d->g) return DGSCT_ERR_BAD_DESC;
  if (d->io_dtype != DGSCT_F32 && d->io_dtype != DGSCT_BF16) return DGSCT_ERR_BAD_DESC;
  if ((d->flags & DGSCT_F_TEMPORAL) && (d->T <= 0 || d->BT % d->T)) return DGSCT_ERR_BAD_DESC;
  return DGSCT_OK;
}

bool use_fc_first(const dgsct_desc& d, const dgsct_params& P) {
  if (P.rs_rowptr) return false;            // sparse (stencil) resampler: U = M . other is a gather, keep the reference order
  return getenv("DGSCT_NO_FC_FIRST") == nullptr && d.io_dtype == DGSCT_BF16 && !(d.flags & DGSCT_F_FORCE_SIMT) && d.C < d.Co &&
         d.C % 8 == 0 && d.Co % 8 == 0;
}

FwdPlan make_fwd_plan(const dgsct_desc& d, bool keep) {
  FwdPlan p{};
  const size_t BT = d.BT, N = d.N, C = d.C, Co = d.Co, K = d.K, Cd = d.C / d.r, H = d.C / 2;
  const size_t es = d.io_dtype == DGSCT_BF16 ? 2 : 4;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  const size_t No = d.No;
  size_t bufA_elems = BT * N * (Co > C ? Co : C);   // U (BT,N,Co) then up-sampler output (BT,N,C)
  if (BT * No * C > bufA_elems) bufA_elems = BT * No * C;   // or O W_fc^T (BT,No,C) of the re-associated resampler
  p.bufA = take(bufA_elems * es);
  p.bufV = take(BT * N * C * es);                   // V then Y
  p.bufX1 = take(BT * N * C * es);
  p.bufL = take(BT * N * K * 4);                    // token logits -> probabilities (P then Q), fp32
  p.Tp = take(BT * K * C * 4);
  p.Z = take(BT * N * Cd * 4);
  p.acc_begin = off;                                // ---- zero-initialised accumulators
  p.osum = take(BT * C * 4);
  p.qsum = take(BT * C * 4);
  p.r1sum = p.qsum;                                 // same accumulator: sum_n relu(z1), unscaled
  p.ssum = take(BT * N * 4);
  p.bn1_sum = take(Cd * 8); p.bn1_sq = take(Cd * 8);
  p.bn2_sum = take(C * 8); p.bn2_sq = take(C * 8);
  p.fa_ticket = take(BT * 4);
  p.acc_end = off;
  p.aq1 = take(BT * C * 4); p.aq2 = take(BT * H * 4); p.h = take(BT * H * 4); p.m = take(BT * C * 4);
  p.tau = take(BT * 4); p.sig = take(BT * N * 4);
  p.bn1_scale = take(Cd * 4); p.bn1_shift = take(Cd * 4);
  p.bn2_scale = take(C * 4); p.bn2_shift = take(C * 4);
  p.bn1_mean = take(Cd * 4); p.bn1_rstd = take(Cd * 4); p.bn2_mean = take(C * 4); p.bn2_rstd = take(C * 4);
  p.qvec = take(BT * C * 4);
  p.Ybuf = p.bufV; p.Qf = p.bufL; p.XcBuf = p.bufA;
  if (keep) {
    p.Ybuf = take(BT * N * C * es);
    p.Qf = take(BT * N * K * 4);
    if (d.io_dtype != DGSCT_BF16) p.Uo32 = take(BT * N * C * 4);
  }
  if (d.io_dtype == DGSCT_BF16) {
    const size_t Kp = (size_t)pad8((int)K);            // bf16 P / Q rows padded to 16 bytes (TMA-describable for any K)
    p.Pb = take(BT * N * Kp * 2);
    p.csm_part = take(BT * (size_t)CSM_MAX_CHUNKS * 2 * K * 4);
    p.fc_s = take(C * 4);
    p.TpB = take(BT * K * C * 2);
    p.Zr = take(BT * N * (size_t)pad8((int)Cd) * 2);
    p.Uo32 = take(BT * N * C * 4);                    // up-sampler output stays fp32 until ln_post
    p.packed = take(make_pack_plan(d).total);
    p.aq = take(BT * (C + H) * 4);
    p.vec_bf = take(BT * (size_t)pad8((int)C) * 2);
    p.vec_bf_o = take(BT * (size_t)pad8((int)C) * 2);
    p.Qb = p.Pb;
    if (keep) { p.Qb = take(BT * N * Kp * 2); p.XcBuf = take(BT * N * C * 2); }
    const fa::TokAttnPlan fp_ = fa::make_tok_attn_plan((int)BT, (int)N, (int)C, (int)K);
    if (fp_.ok) {
      p.fa_part_o = take(fp_.part_o_bytes); p.fa_part_ml = take(fp_.part_ml_bytes);
      p.fa_mref = take(fp_.mref_bytes); p.fa_pfac = take(fp_.mref_bytes); p.fa_stat = take(BT * 2 * fa::KL * 4);
    }
  }
  p.total = off;
  return p;
}

static GemmOperand op(const void* p, long long s_row, long long s_k, long long s_batch, int is_bf16, long long s_group = 0) {
  GemmOperand o; o.p = p; o.s_row = s_row; o.s_k = s_k; o.s_batch = s_batch; o.s_group = s_group; o.is_bf16 = is_bf16;
  return o;
}

static EpiBiasStore epi_store(void* out, int out_bf16, long long ld, long long sb) {
  EpiBiasStore e{}; e.out = out; e.out_bf16 = out_bf16; e.ld = ld; e.sb = sb; e.alpha = 1.f; e.act = ACT_NONE;
  return e;
}

#define DGSCT_GEMM(stage, sh, A, B, ax, epi, name)                            \
  do {                                                                        \
    StageScope _sc(stage, st); count_launch();                                \
    cudaError_t _e = launch_gemm_simt(sh, A, B, ax, epi, st);                 \
    if (_e != cudaSuccess) { set_cuda_error(_e, name); return DGSCT_ERR_CUDA; } \
  } while (0)

static int tap(const void* src, int is_bf16, float* dst, size_t n, cudaStream_t st) {
  if (!dst) return DGSCT_OK;
  StageScope _sc(ST_MISC, st); count_launch();
  tap_copy_kernel<<<(unsigned)((n + 255) / 256 > 2048 ? 2048 : (n + 255) / 256), 256, 0, st>>>(src, is_bf16, dst, n);
  DGSCT_LAUNCH_CHECK("tap_copy");
  return DGSCT_OK;
}

int adapter_fwd_simt(const dgsct_desc& d, const dgsct_params& P, const void* x, const void* other, void* out,
                     void* smap, float* tmap, const dgsct_taps* taps, char* ws, cudaStream_t st, bool keep, const void* resid) {
  const FwdPlan pl = make_fwd_plan(d, keep);
  const int upd = (d.flags & DGSCT_F_NO_STAT_UPDATE) ? 0 : 1;
  const int BT = d.BT, N = d.N, C = d.C, Co = d.Co, No = d.No, K = d.K, Cd = d.C / d.r, H = d.C / 2, G = d.g;
  const int bf = d.io_dtype == DGSCT_BF16;
  const bool train = d.flags & DGSCT_F_TRAIN, use_bn = d.flags & DGSCT_F_USE_BN;
  void* U = ws + pl.bufA; void* V = ws + pl.bufV; void* X1 = ws + pl.bufX1;
  void* Yb = ws + pl.Ybuf; float* Lq = (float*)(ws + pl.Qf);
  void* UoBuf = keep ? (void*)(ws + pl.Uo32) : U; const int uo_bf = keep ? 0 : bf;
  float* qvec = (float*)(ws + pl.qvec);
  float* L = (float*)(ws + pl.bufL); float* Tp = (float*)(ws + pl.Tp); float* Z = (float*)(ws + pl.Z);
  float* osum = (float*)(ws + pl.osum); float* qsum = (float*)(ws + pl.qsum); float* ssum = (float*)(ws + pl.ssum);
  double* bn1_sum = (double*)(ws + pl.bn1_sum); double* bn1_sq = (double*)(ws + pl.bn1_sq);
  double* bn2_sum = (double*)(ws + pl.bn2_sum); double* bn2_sq = (double*)(ws + pl.bn2_sq);
  float* aq1 = (float*)(ws + pl.aq1); float* aq2 = (float*)(ws + pl.aq2); float* hbuf = (float*)(ws + pl.h);
  float* mvec = (float*)(ws + pl.m); float* tau = (float*)(ws + pl.tau); float* sig = (float*)(ws + pl.sig);
  float* bn1_scale = (float*)(ws + pl.bn1_scale); float* bn1_shift = (float*)(ws + pl.bn1_shift);
  float* bn2_scale = (float*)(ws + pl.bn2_scale); float* bn2_shift = (float*)(ws + pl.bn2_shift);
  const float invN = 1.f / (float)N;

  { StageScope _sc(ST_MISC, st);
  DGSCT_CUDA_CHECK(cudaMemsetAsync(ws + pl.acc_begin, 0, pl.acc_end - pl.acc_begin, st)); }

  // (R1) U[b] = Wres (N,No) . other[b] (No,Co) + b_res[n]                       net_trans.py:553
  if (P.rs_rowptr) {      // sparse resampling matrix (bicubic stencil, PVT_AVSModel.py:190-197)
    StageScope _sc(ST_RESAMPLE_CONV, st); count_launch();
    if (!launch_spmm_frames(P.rs_rowptr, P.rs_col, P.rs_val, other, U, bf, BT, N, No, Co, st)) return DGSCT_ERR_UNSUPPORTED;
    DGSCT_LAUNCH_CHECK("resample_spmm");
  } else {
    GemmShape sh{N, Co, No, BT, 1};
    EpiBiasStore e = epi_store(U, bf, Co, (long long)N * Co);
    e.bias_m = (d.flags & DGSCT_F_NO_CONV_BIAS) ? nullptr : P.resample_b;
    DGSCT_GEMM(ST_RESAMPLE_CONV, sh, op(P.resample_w, No, 1, 0, 0), op(other, 1, Co, (long long)No * Co, bf), AxIdentity{}, e, "resample_conv");
  }
  // (R2) V = U . Wfc^T + b_fc ; osum[b,c] = sum_n V                             net_trans.py:554, 592
  {
    GemmShape sh{N, C, Co, BT, 1};
    EpiBiasStore e = epi_store(V, bf, C, (long long)N * C);
    e.bias_n = P.fc_b; e.colsum = osum; e.ncols_total = C;
    DGSCT_GEMM(ST_RESAMPLE_FC, sh, op(U, Co, 1, (long long)N * Co, bf), op(P.fc_w, Co, 1, 0, 0), AxIdentity{}, e, "resample_fc");
  }
  if (taps) { int r = tap(V, bf, taps->V, (size_t)BT * N * C, st); if (r) return r; }
  // (T1) L[b,n,k] = V[b,n,:] . tok[k,:]                                         net_trans.py:574
  {
    GemmShape sh{N, K, C, BT, 1};
    EpiBiasStore e = epi_store(L, 0, K, (long long)N * K);
    DGSCT_GEMM(ST_TOK_LOGITS, sh, op(V, C, 1, (long long)N * C, bf), op(P.tokens, C, 1, 0, 0), AxIdentity{}, e, "tok_logits");
  }
  { StageScope _sc(ST_COL_SOFTMAX, st); count_launch();
  col_softmax_kernel<<<dim3(K, BT), 128, 0, st>>>(L, N, K);                   // net_trans.py:576
  DGSCT_LAUNCH_CHECK("col_softmax"); }
  // (T3) Tp[b,k,c] = tok[k,c] + sum_n P[b,n,k] V[b,n,c]                         net_trans.py:577-580
  {
    GemmShape sh{K, C, N, BT, 1};
    EpiBiasStore e = epi_store(Tp, 0, C, (long long)K * C);
    e.add_mn = P.tokens; e.add_ld = C;
    DGSCT_GEMM(ST_TOK_UPDATE, sh, op(L, 1, K, (long long)N * K, 0), op(V, 1, C, (long long)N * C, bf), AxIdentity{}, e, "tok_update");
  }
  if (taps && taps->Tp) { int r = tap(Tp, 0, taps->Tp, (size_t)BT * K * C, st); if (r) return r; }
  // (T4a) L[b,n,k] = x[b,n,:] . Tp[b,k,:] ; softmax over k                      net_trans.py:583-585
  {
    GemmShape sh{N, K, C, BT, 1};
    EpiBiasStore e = epi_store(Lq, 0, K, (long long)N * K);
    DGSCT_GEMM(ST_X_LOGITS, sh, op(x, C, 1, (long long)N * C, bf), op(Tp, C, 1, (long long)K * C, 0), AxIdentity{}, e, "x_logits");
  }
  {
    const long long rows = (long long)BT * N;
    { StageScope _sc(ST_ROW_SOFTMAX, st); count_launch();
  row_softmax_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(Lq, rows, K);
    DGSCT_LAUNCH_CHECK("row_softmax"); }
  }
  // (T4b) X1 = x + gate_av * Q . Tp                                             net_trans.py:586-589
  {
    GemmShape sh{N, C, K, BT, 1};
    EpiResidual e{X1, x, bf, C, (long long)N * C, P.gate_av};
    DGSCT_GEMM(ST_X_UPDATE, sh, op(Lq, K, 1, (long long)N * K, 0), op(Tp, 1, C, (long long)K * C, 0), AxIdentity{}, e, "x_update");
  }
  if (taps) { int r = tap(X1, bf, taps->X1, (size_t)BT * N * C, st); if (r) return r; }
  // (C1) aq1 = relu(Wa1 obar + b), aq2 = relu(Wa2 obar + b), tau = sigmoid(wt obar + bt)   net_trans.py:593,603
  {
    GemmShape sh{BT, C, C, 1, 1};
    EpiBiasStore e = epi_store(aq1, 0, C, 0); e.bias_n = P.a1_b; e.alpha = invN; e.act = ACT_RELU;
    DGSCT_GEMM(ST_SMALL_LINEAR, sh, op(osum, C, 1, 0, 0), op(P.a1_w, C, 1, 0, 0), AxIdentity{}, e, "aq1");
    GemmShape sh2{BT, H, C, 1, 1};
    EpiBiasStore e2 = epi_store(aq2, 0, H, 0); e2.bias_n = P.a2_b; e2.alpha = invN; e2.act = ACT_RELU;
    DGSCT_GEMM(ST_SMALL_LINEAR, sh2, op(osum, C, 1, 0, 0), op(P.a2_w, C, 1, 0, 0), AxIdentity{}, e2, "aq2");
    if (d.flags & DGSCT_F_TEMPORAL) {
      GemmShape sh3{BT, 1, C, 1, 1};
      EpiBiasStore e3 = epi_store(tau, 0, 1, 0); e3.bias_n = P.t_b; e3.alpha = invN; e3.act = ACT_SIGMOID;
      DGSCT_GEMM(ST_SMALL_LINEAR, sh3, op(osum, C, 1, 0, 0), op(P.t_w, C, 1, 0, 0), AxIdentity{}, e3, "tau");
      if (tmap) DGSCT_CUDA_CHECK(cudaMemcpyAsync(tmap, tau, (size_t)BT * 4, cudaMemcpyDeviceToDevice, st));
    }
    if (taps && taps->obar) {
      // obar = osum / N : reuse the GEMM path with an identity-free scale via tap copy then scale on host side
      // (kept simple: tap the raw sums, the binding divides by N)
      int r = tap(osum, 0, taps->obar, (size_t)BT * C, st); if (r) return r;
    }
  }
  // (C2) qsum[b,c] = sum_n aq1[b,c] * relu(X1[b,n,:] . Wv1[c,:] + b)            net_trans.py:594-595
  {
    GemmShape sh{N, C, C, BT, 1};
    EpiReluMulColsum e{P.v1_b, nullptr, qsum, C, C};      // r1sum[b,c] = sum_n relu(z1), unscaled
    DGSCT_GEMM(ST_CHANNEL_QUERY, sh, op(X1, C, 1, (long long)N * C, bf), op(P.v1_w, C, 1, 0, 0), AxIdentity{}, e, "channel_query");
  }
  // (C3) m = sigmoid(Wc relu(Wb q + bb) + bc)                                   net_trans.py:596-597
  {
    { StageScope _sc(ST_SMALL_LINEAR, st); count_launch();
      channel_q_kernel<<<ceil_div(BT * C, 256), 256, 0, st>>>(aq1, C, qsum, invN, qvec, nullptr, BT, C);
      DGSCT_LAUNCH_CHECK("channel_q"); }
    GemmShape sh{BT, H, C, 1, 1};
    EpiBiasStore e = epi_store(hbuf, 0, H, 0); e.bias_n = P.bott_b; e.act = ACT_RELU;
    DGSCT_GEMM(ST_SMALL_LINEAR, sh, op(qvec, C, 1, 0, 0), op(P.bott_w, C, 1, 0, 0), AxIdentity{}, e, "channel_bottleneck");
    GemmShape sh2{BT, C, H, 1, 1};
    EpiBiasStore e2 = epi_store(mvec, 0, C, 0); e2.bias_n = P.c_b; e2.act = ACT_SIGMOID;
    DGSCT_GEMM(ST_SMALL_LINEAR, sh2, op(hbuf, H, 1, 0, 0), op(P.c_w, H, 1, 0, 0), AxIdentity{}, e2, "channel_gate");
  }
  if (taps && taps->m) { int r = tap(mvec, 0, taps->m, (size_t)BT * C, st); if (r) return r; }
  // (S1) ssum[b,n] = sum_j ws[j] aq2[b,j] relu((X1 (1+m))[b,n,:] . Wv2[j,:] + b)   net_trans.py:598-605
  {
    GemmShape sh{N, H, C, BT, 1};
    EpiSpatialRowsum e{P.v2_b, aq2, P.s_w, ssum, H, N};
    DGSCT_GEMM(ST_SPATIAL_QUERY, sh, op(X1, C, 1, (long long)N * C, bf), op(P.v2_w, C, 1, 0, 0), (AxOnePlusCol{mvec, C}), e, "spatial_query");
  }
  { StageScope _sc(ST_SPATIAL_MAP, st); count_launch();
  spatial_map_kernel<<<BT, 256, 0, st>>>(ssum, P.s_b, N, sig, smap, bf, taps ? taps->s : nullptr);   // :606-608
  DGSCT_LAUNCH_CHECK("spatial_map"); }
  // (B1) Y = LN_before(X1 * (alpha m + beta sig + gamma tau + 1 - alpha))       net_trans.py:611-613, 626-627
  {
    GateLnArgs a{};
    a.x1 = X1; a.in_bf16 = bf; a.y = Yb; a.out_bf16 = bf; a.m = mvec; a.sig = sig;
    a.tau = ((d.flags & DGSCT_F_TEMPORAL) && d.gamma != 0.f) ? tau : nullptr;
    const bool lnb = d.flags & DGSCT_F_LN_BEFORE;
    a.ln_w = lnb ? P.lnb_w : nullptr; a.ln_b = lnb ? P.lnb_b : nullptr;
    a.alpha = d.alpha; a.beta = d.beta; a.gamma = d.gamma; a.eps = d.eps_ln; a.N = N; a.C = C;
    a.rows = (long long)BT * N; a.y_tap = taps ? taps->Y : nullptr;
    const int wpb = 4;
    { StageScope _sc(ST_GATE_LN, st); count_launch();
  gate_ln_kernel<<<(unsigned)((a.rows + wpb - 1) / wpb), wpb * 32, (size_t)wpb * C * 4, st>>>(a);
    DGSCT_LAUNCH_CHECK("gate_ln"); }
  }
  // (B2) Z = down_sampler(Y) (grouped 1x1) ; bn1 batch statistics               net_trans.py:629,636
  {
    GemmShape sh{N, Cd / G, C / G, BT, G};
    EpiBiasStore e = epi_store(Z, 0, Cd, (long long)N * Cd);
    if (train && use_bn) { e.stat_sum = bn1_sum; e.stat_sq = bn1_sq; }
    DGSCT_GEMM(ST_DOWN, sh, op(Yb, C, 1, (long long)N * C, bf, C / G), op(P.down_w, C / G, 1, 0, 0, (long long)(Cd / G) * (C / G)),
               AxIdentity{}, e, "down_sampler");
  }
  if (taps && taps->Z) { int r = tap(Z, 0, taps->Z, (size_t)BT * N * Cd, st); if (r) return r; }
  const double count = (double)BT * (double)N;
  { StageScope _sc(ST_BN_FINALIZE, st); count_launch();
  bn_finalize_kernel<<<ceil_div(Cd, 128), 128, 0, st>>>(Cd, use_bn, train, bn1_sum, bn1_sq, count, P.bn1_w, P.bn1_b,
                                                         P.bn1_rm, P.bn1_rv, d.eps_bn, d.bn_momentum, bn1_scale, bn1_shift,
                                                         (float*)(ws + pl.bn1_mean), (float*)(ws + pl.bn1_rstd), upd);
  DGSCT_LAUNCH_CHECK("bn1_finalize"); }
  // (B3) Uo = up_sampler(relu(bn1(Z))) ; bn2 batch statistics                   net_trans.py:638-643
  {
    GemmShape sh{N, C / G, Cd / G, BT, G};
    EpiBiasStore e = epi_store(UoBuf, uo_bf, C, (long long)N * C);
    if (train && use_bn) { e.stat_sum = bn2_sum; e.stat_sq = bn2_sq; }
    DGSCT_GEMM(ST_UP, sh, op(Z, Cd, 1, (long long)N * Cd, 0, Cd / G), op(P.up_w, Cd / G, 1, 0, 0, (long long)(C / G) * (Cd / G)),
               (AxAffineRelu{bn1_scale, bn1_shift, Cd / G}), e, "up_sampler");
  }
  { StageScope _sc(ST_BN_FINALIZE, st); count_launch();
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(C, use_bn, train, bn2_sum, bn2_sq, count, P.bn2_w, P.bn2_b,
                                                        P.bn2_rm, P.bn2_rv, d.eps_bn, d.bn_momentum, bn2_scale, bn2_shift,
                                                        (float*)(ws + pl.bn2_mean), (float*)(ws + pl.bn2_rstd), upd);
  DGSCT_LAUNCH_CHECK("bn2_finalize"); }
  // (B4) out = gate * LN_post(bn2(Uo))  (or AVS order)                          net_trans.py:667-672
  if (!(d.flags & DGSCT_F_NO_OUTPUT)) {
    OutLnArgs a{};
    a.u = UoBuf; a.in_bf16 = uo_bf; a.out = out; a.out_bf16 = bf; a.scale = bn2_scale; a.shift = bn2_shift;
    const bool lnp = d.flags & DGSCT_F_LN_POST;
    a.ln_w = lnp ? P.lnp_w : nullptr; a.ln_b = lnp ? P.lnp_b : nullptr;
    a.gate = (d.flags & DGSCT_F_USE_GATE) ? P.gate : nullptr;
    a.gate_before_ln = (d.flags & DGSCT_F_GATE_BEFORE_LN) ? 1 : 0;
    a.eps = d.eps_ln; a.C = C; a.rows = (long long)BT * N; a.resid = resid;
    const int wpb = 4;
    { StageScope _sc(ST_OUT_LN, st); count_launch();
  out_ln_kernel<<<(unsigned)((a.rows + wpb - 1) / wpb), wpb * 32, (size_t)wpb * C * 4, st>>>(a);
    DGSCT_LAUNCH_CHECK("out_ln"); }
  }
  return DGSCT_OK;
}

}  // namespace dgsct

// =====================================================================================================
// bf16 tensor-core pipeline
// =====================================================================================================
namespace dgsct {

size_t packed_bytes(const dgsct_desc& d) { return make_pack_plan(d).total; }

int pack_params(const dgsct_desc& d, const dgsct_params& P, char* packed, cudaStream_t st) {
  const PackPlan pp = make_pack_plan(d);
  const int N = d.N, C = d.C, Co = d.Co, No = d.No, K = d.K, Cd = d.C / d.r, H = d.C / 2;
  StageScope _sc(ST_MISC, st);
  PackTable t{};
  size_t biggest = 0;
  auto seg = [&](int kind, const float* src, const float* src2, size_t off, int r, int c, int ld, int g) {
    PackSeg& e = t.seg[t.n++];
    e.src = src; e.src2 = src2; e.dst = packed + off; e.kind = kind; e.rows = r; e.cols = c; e.ld = ld; e.groups = g;
    const size_t n = kind == 2 ? (size_t)(r + c) : (size_t)r * ld;
    if (n > biggest) biggest = n;
  };
  seg(0, P.resample_w, nullptr, pp.res, N, No, pp.ld_res, 1);      // (NULL with the sparse resampler: the segment is skipped)
  seg(0, P.fc_w, nullptr, pp.fc, C, Co, pp.ld_fc, 1);
  seg(0, P.tokens, nullptr, pp.tok, K, C, pp.ld_c, 1);
  seg(0, P.v1_w, nullptr, pp.v1, C, C, pp.ld_c, 1);
  seg(0, P.v2_w, nullptr, pp.v2, H, C, pp.ld_c, 1);
  seg(0, P.a1_w, nullptr, pp.cat1, C, C, pp.ld_c, 1);
  seg(0, P.a2_w, nullptr, pp.cat1 + (size_t)C * pp.ld_c * 2, H, C, pp.ld_c, 1);
  seg(0, P.bott_w, nullptr, pp.bott, H, C, pp.ld_c, 1);
  seg(0, P.c_w, nullptr, pp.catt, C, H, pp.ld_h, 1);
  seg(2, P.a1_b, P.a2_b, pp.cat1_bias, C, H, 0, 1);
  seg(1, P.down_w, nullptr, pp.down, Cd, C, pp.ld_c, d.g);
  seg(1, P.up_w, nullptr, pp.up, C, Cd, pp.ld_up, d.g);
  seg(3, P.fc_w, nullptr, pp.fc_s, C, Co, 32, 1);                  // (ld = 32 only sizes the x extent: a warp per row)
  count_launch();
  // the x extent covers the LARGEST segment with ~8 elements per thread; small segments finish in their first blocks
  size_t bx = (biggest + 256 * 8 - 1) / (256 * 8);
  if (bx > 148 * 4) bx = 148 * 4;
  if (bx < 1) bx = 1;
  pack_all_kernel<<<dim3((unsigned)bx, t.n), 256, 0, st>>>(t);
  DGSCT_LAUNCH_CHECK("pack_params");
  return DGSCT_OK;
}

// one product through the tcgen05 kernel, or through the SIMT kernel when TMA cannot describe the operands
template <typename TE, typename SE>
static int gemm_auto(int stage, const char* name, bool allow_tc, const tc::OperandDesc& A, const tc::OperandDesc& B, int M, int N,
                     int K, int batch, TE te, SE se, cudaStream_t st, int bn_hint = 0) {
  StageScope _sc(stage, st);
  count_launch();
  if (allow_tc) {
    tc::Params p;
    const int rc = tc::make_params(p, A, B, M, N, K, batch, bn_hint, 0, false, tc::epi_row_independent<TE>::value,
                                   tc::epi_stage_bytes<TE>::value, tc::epi_block_stats<TE>::value ? 1 : 0);
    if (rc == DGSCT_OK) {
      cudaError_t e = tc::launch(p, te, st);
      if (e != cudaSuccess) { set_cuda_error(e, name); return DGSCT_ERR_CUDA; }
      return DGSCT_OK;
    }
    if (rc != DGSCT_ERR_UNSUPPORTED) return rc;
  }
  GemmShape sh{M, N, K, batch, 1};
  GemmOperand a{A.ptr, A.mn_major ? 1 : A.ld, A.mn_major ? A.ld : 1, A.batch_stride, 0, 1};
  GemmOperand b{B.ptr, B.mn_major ? 1 : B.ld, B.mn_major ? B.ld : 1, B.batch_stride, 0, 1};
  cudaError_t e = launch_gemm_simt(sh, a, b, AxIdentity{}, se, st);
  if (e != cudaSuccess) { set_cuda_error(e, name); return DGSCT_ERR_CUDA; }
  return DGSCT_OK;
}

// tensor-core kernel or an error (no SIMT twin): for paths whose epilogue has no scalar counterpart
template <typename TE>
static int gemm_tc_only(int stage, const char* name, const tc::OperandDesc& A, const tc::OperandDesc& B, int M, int N, int K,
                        int batch, TE te, cudaStream_t st) {
  StageScope _sc(stage, st);
  count_launch();
  tc::Params p;
  const int rc = tc::make_params(p, A, B, M, N, K, batch, 0, 0, false, tc::epi_row_independent<TE>::value,
                                 tc::epi_stage_bytes<TE>::value);
  if (rc != DGSCT_OK) return rc;
  cudaError_t e = tc::launch(p, te, st);
  if (e != cudaSuccess) { set_cuda_error(e, name); return DGSCT_ERR_CUDA; }
  return DGSCT_OK;
}

static tc::OperandDesc kmaj(const void* p, int rows, int cols, long long ld, long long bs) {
  return tc::OperandDesc{p, 0, ld, bs, rows, cols};
}
static tc::OperandDesc mnmaj(const void* p, int k_rows, int mn_cols, long long ld, long long bs) {
  return tc::OperandDesc{p, 1, ld, bs, k_rows, mn_cols};
}
template <bool OUT_BF16, bool COLSUM = false, bool STATS = false, int ACT = ACT_NONE>
static tc::EpiStore<OUT_BF16, ACT, COLSUM, STATS> tc_store(void* out, long long ld, long long sb) {
  tc::EpiStore<OUT_BF16, ACT, COLSUM, STATS> e{}; e.out = out; e.ld = ld; e.sb = sb; e.alpha = 1.f;
  return e;
}

// file-scope epilogue functors used only by the tensor-core pipeline
struct SimtTokUpdate {
  bf16* o; float* o2; const float* tok; int C, K;
  __device__ void operator()(int b, int, int m, int n, float (&acc)[4][4], int M, int Nn) const {
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
      const int c = m + i, k = n + j;
      if (c < M && k < Nn) {
        const float v = acc[i][j] + tok[(long long)k * C + c];
        o[(long long)b * K * C + (long long)k * C + c] = __float2bfloat16_rn(v);
        o2[(long long)b * K * C + (long long)k * C + c] = v;
      }
    }
  }
};
struct SimtAffineRelu {
  bf16* zr; long long ldz, sbz; const float* scale; const float* shift; float* ztap; int Cd;
  __device__ void operator()(int b, int, int m, int n, float (&acc)[4][4], int M, int Nn) const {
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j)
      if (m + i < M && n + j < Nn) {
        if (ztap) ztap[((long long)b * (sbz / ldz) + m + i) * Cd + n + j] = acc[i][j];
        zr[(long long)b * sbz + (long long)(m + i) * ldz + n + j] =
            __float2bfloat16_rn(fmaxf(fmaf(acc[i][j], scale[n + j], shift[n + j]), 0.f));
      }
  }
};

#define DGSCT_RC(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)


// One launch of the lean per-frame-vector linear (vec_linear.cuh), or DGSCT_ERR_UNSUPPORTED (the caller then takes the GEMM)
static int vec_linear(int stage, const char* name, const void* in, int ld_in, const void* W, int ld_w, int w_trans, const float* bias,
                      int act, float* out32, int ld32, bf16* out16, int ld16, int M, int N, int K, int accumulate, cudaStream_t st) {
  vl::VecLinArgs a{};
  a.in = (const bf16*)in; a.ld_in = ld_in; a.W = (const bf16*)W; a.ld_w = ld_w; a.bias = bias; a.out32 = out32; a.ld32 = ld32;
  a.out16 = out16; a.ld16 = ld16; a.M = M; a.N = N; a.K = K; a.act = act; a.w_trans = w_trans; a.accumulate = accumulate;
  if (!vl::vec_linear_ok(a)) return DGSCT_ERR_UNSUPPORTED;
  StageScope _sc(stage, st); count_launch();
  cudaError_t e = vl::launch_vec_linear(a, st);
  if (e != cudaSuccess) { set_cuda_error(e, name); return DGSCT_ERR_CUDA; }
  return DGSCT_OK;
}

// Fused latent-token attention (attn_fused.cuh): bf16 tensor-core path, shapes the kernels cover, 16-byte aligned rows.
// DGSCT_NO_FUSED_ATTN=1 restores the six-launch GEMM chain (forward and backward), DGSCT_NO_FUSED_ATTN_BWD=1 only the backward's.
static bool fused_attn_enabled(const dgsct_desc& d, const void* x, const PackPlan& pp) {
  if (d.io_dtype != DGSCT_BF16 || (d.flags & DGSCT_F_FORCE_SIMT) || getenv("DGSCT_NO_FUSED_ATTN")) return false;
  if (((uintptr_t)x & 15) != 0 || (pp.ld_c % 8) != 0) return false;
  return fa::make_tok_attn_plan(d.BT, d.N, d.C, d.K).ok && fa::make_x_attn_plan(d.BT, d.N, d.C, d.K).ok;
}
// The backward of the token softmax recomputes P from the statistics only the fused forward leaves behind; when it cannot run
// (very wide rows), the fused forward also writes P for the GEMM chain.
static bool fused_attn_bwd_a_enabled(const dgsct_desc& d, const void* x, const PackPlan& pp) {
  return fused_attn_enabled(d, x, pp) && getenv("DGSCT_NO_FUSED_ATTN_BWD") == nullptr &&
         fa::make_tok_attn_bwd_plan(d.BT, d.N, d.C, d.K).ok;
}

int adapter_fwd_tc(const dgsct_desc& d, const dgsct_params& P, const void* packed_in, const void* x, const void* other,
                   void* out, void* smap, float* tmap, const dgsct_taps* taps, char* ws, cudaStream_t st, bool keep, const void* resid) {
  const FwdPlan pl = make_fwd_plan(d, keep);
  const int upd = (d.flags & DGSCT_F_NO_STAT_UPDATE) ? 0 : 1;
  const PackPlan pp = make_pack_plan(d);
  const int BT = d.BT, N = d.N, C = d.C, Co = d.Co, No = d.No, K = d.K, Cd = d.C / d.r, H = d.C / 2;
  const int ldz = pad8(Cd);
  const bool train = d.flags & DGSCT_F_TRAIN, use_bn = d.flags & DGSCT_F_USE_BN;
  bf16* U = (bf16*)(ws + pl.bufA); bf16* V = (bf16*)(ws + pl.bufV); bf16* X1 = (bf16*)(ws + pl.bufX1);
  float* L = (float*)(ws + pl.bufL); float* Tp = (float*)(ws + pl.Tp); float* Z = (float*)(ws + pl.Z);
  bf16* Pb = (bf16*)(ws + pl.Pb); bf16* TpB = (bf16*)(ws + pl.TpB); bf16* Zr = (bf16*)(ws + pl.Zr);
  bf16* Qb = (bf16*)(ws + pl.Qb); bf16* Yb = (bf16*)(ws + pl.Ybuf); float* qvec = (float*)(ws + pl.qvec);
  float* Uo = (float*)(ws + pl.Uo32);
  float* osum = (float*)(ws + pl.osum); float* qsum = (float*)(ws + pl.qsum); float* ssum = (float*)(ws + pl.ssum);
  double* bn1_sum = (double*)(ws + pl.bn1_sum); double* bn1_sq = (double*)(ws + pl.bn1_sq);
  double* bn2_sum = (double*)(ws + pl.bn2_sum); double* bn2_sq = (double*)(ws + pl.bn2_sq);
  float* hbuf = (float*)(ws + pl.h);
  float* mvec = (float*)(ws + pl.m); float* tau = (float*)(ws + pl.tau); float* sig = (float*)(ws + pl.sig);
  float* bn1_scale = (float*)(ws + pl.bn1_scale); float* bn1_shift = (float*)(ws + pl.bn1_shift);
  float* bn2_scale = (float*)(ws + pl.bn2_scale); float* bn2_shift = (float*)(ws + pl.bn2_shift);
  const float invN = 1.f / (float)N;
  const long long rowsBT = (long long)BT * N;

  const char* pk = (const char*)packed_in;
  if (!pk) {
    DGSCT_RC(pack_params(d, P, ws + pl.packed, st));
    pk = ws + pl.packed;
  }
  const bf16* Wres = (const bf16*)(pk + pp.res); const bf16* Wfc = (const bf16*)(pk + pp.fc);
  const bf16* Wtok = (const bf16*)(pk + pp.tok); const bf16* Wv1 = (const bf16*)(pk + pp.v1);
  const bf16* Wv2 = (const bf16*)(pk + pp.v2); const bf16* Wdn = (const bf16*)(pk + pp.down);
  const bf16* Wup = (const bf16*)(pk + pp.up);

  { StageScope _sc(ST_MISC, st);
    DGSCT_CUDA_CHECK(cudaMemsetAsync(ws + pl.acc_begin, 0, pl.acc_end - pl.acc_begin, st)); }

  const bool fc_first = use_fc_first(d, P);
  if (fc_first) {
    // (R1') O'[b] (No,C) = other[b] (No,Co) . Wfc^T          (frames flattened, Wfc resident)
    // (R2') V[b] (N,C)  = Wres (N,No) . O'[b] + b_res[n] * s[c] + b_fc[c],  s = rowsum(Wfc) ; osum[b,c] = sum_n V
    bf16* Op = U;
    const float* fcs = (const float*)(pk + pp.fc_s);          // row sums of the bf16 fc.weight copy, made by the pack launch
    {
      auto te = tc_store<true>(Op, C, (long long)No * C);
      DGSCT_RC(gemm_tc_only(ST_RESAMPLE_FC, "resample_fc_first", kmaj(other, No, Co, Co, (long long)No * Co),
                            kmaj(Wfc, C, Co, pp.ld_fc, 0), No, C, Co, BT, te, st));
    }
    {
      auto te = tc_store<true, true>(V, C, (long long)N * C);
      te.bias_m = (d.flags & DGSCT_F_NO_CONV_BIAS) ? nullptr : P.resample_b;
      te.bias_m_coln = te.bias_m ? fcs : nullptr;
      te.bias_n = P.fc_b; te.colsum = osum; te.ncols_total = C;
      DGSCT_RC(gemm_tc_only(ST_RESAMPLE_CONV, "resample_conv_reassoc", kmaj(Wres, N, No, pp.ld_res, 0),
                            mnmaj(Op, No, C, C, (long long)No * C), N, C, No, BT, te, st));
    }
  }
  // (R1) U[b] (N,Co) = Wres (N,No) . other[b] (No,Co) + b_res[n]
  if (!fc_first && P.rs_rowptr) {   // sparse resampling matrix (bicubic stencil): gather with fp32 coefficients
    StageScope _sc(ST_RESAMPLE_CONV, st); count_launch();
    if (!launch_spmm_frames(P.rs_rowptr, P.rs_col, P.rs_val, other, U, 1, BT, N, No, Co, st)) return DGSCT_ERR_UNSUPPORTED;
    DGSCT_LAUNCH_CHECK("resample_spmm");
  } else if (!fc_first) {
    auto te = tc_store<true>(U, Co, (long long)N * Co);
    EpiBiasStore se = epi_store(U, 1, Co, (long long)N * Co);
    te.bias_m = se.bias_m = (d.flags & DGSCT_F_NO_CONV_BIAS) ? nullptr : P.resample_b;
    DGSCT_RC(gemm_auto(ST_RESAMPLE_CONV, "resample_conv", true, kmaj(Wres, N, No, pp.ld_res, 0),
                       mnmaj(other, No, Co, Co, (long long)No * Co), N, Co, No, BT, te, se, st));
  }
  // (R2) V = U . Wfc^T + b_fc ; osum[b,c] = sum_n V
  if (!fc_first) {
    auto te = tc_store<true, true>(V, C, (long long)N * C);
    EpiBiasStore se = epi_store(V, 1, C, (long long)N * C);
    te.bias_n = se.bias_n = P.fc_b; te.colsum = se.colsum = osum; te.ncols_total = se.ncols_total = C;
    DGSCT_RC(gemm_auto(ST_RESAMPLE_FC, "resample_fc", true, kmaj(U, N, Co, Co, (long long)N * Co), kmaj(Wfc, C, Co, pp.ld_fc, 0),
                       N, C, Co, BT, te, se, st));
  }
  float* aq = (float*)(ws + pl.aq);
  bf16* vecb = (bf16*)(ws + pl.vec_bf);         // bf16 per-frame vectors of the channel-gate chain (q)
  bf16* vecb_o = (bf16*)(ws + pl.vec_bf_o);     // bf16 obar (own buffer: written on the auxiliary stream)
  const int ldv = pad8(C);
  const float* aq1 = aq; const float* aq2 = aq + C; const int aq_ld = C + H;
  // (C1) per-frame linears on tensor cores: [aq1 | aq2] = relu(obar [Wa1; Wa2]^T + b) ; tau stays SIMT (N = 1).
  // They need only the column sums of V, and nothing needs them before the channel query: they run on an auxiliary stream
  // underneath the latent-token attention.
  SideStream* side = (prof_enabled() || getenv("DGSCT_NO_SIDE")) ? nullptr : side_stream_for(st, 2);
  const cudaStream_t sx = side ? side->s : st;
  auto c1_chain = [&]() -> int {
    { StageScope _sc(ST_SMALL_LINEAR, sx); count_launch();
      scale_to_bf16_kernel<<<ceil_div(BT * C, 256), 256, 0, sx>>>(osum, invN, vecb_o, (size_t)BT * C);
      DGSCT_LAUNCH_CHECK("obar_bf16"); }
    auto te = tc_store<false, false, false, ACT_RELU>(aq, aq_ld, 0);
    te.bias_n = (const float*)(pk + pp.cat1_bias);
    EpiBiasStore se = epi_store(aq, 0, aq_ld, 0); se.bias_n = te.bias_n; se.act = ACT_RELU;
    int rcv = vec_linear(ST_SMALL_LINEAR, "aq12", vecb_o, C, pk + pp.cat1, pp.ld_c, 0, te.bias_n, ACT_RELU, aq, aq_ld, nullptr, 0,
                         BT, C + H, C, 0, sx);
    if (rcv == DGSCT_ERR_UNSUPPORTED)
      rcv = gemm_auto(ST_SMALL_LINEAR, "aq12", true, kmaj(vecb_o, BT, C, C, 0), kmaj(pk + pp.cat1, C + H, C, pp.ld_c, 0),
                      BT, C + H, C, 1, te, se, sx, 64);
    DGSCT_RC(rcv);
    if (d.flags & DGSCT_F_TEMPORAL) {
      cudaStream_t st = sx;                       // (DGSCT_GEMM launches on `st`)
      GemmShape sh3{BT, 1, C, 1, 1};
      EpiBiasStore e3 = epi_store(tau, 0, 1, 0); e3.bias_n = P.t_b; e3.alpha = invN; e3.act = ACT_SIGMOID;
      DGSCT_GEMM(ST_SMALL_LINEAR, sh3, op(osum, C, 1, 0, 0), op(P.t_w, C, 1, 0, 0), AxIdentity{}, e3, "tau");
      if (tmap) DGSCT_CUDA_CHECK(cudaMemcpyAsync(tmap, tau, (size_t)BT * 4, cudaMemcpyDeviceToDevice, sx));
    }
    return DGSCT_OK;
  };
  if (side) {                                     // fork: osum is final
    DGSCT_CUDA_CHECK(cudaEventRecord(side->ev, st));
    DGSCT_CUDA_CHECK(cudaStreamWaitEvent(side->s, side->ev, 0));
    DGSCT_RC(c1_chain());
  }
  if (taps) DGSCT_RC(tap(V, 1, taps->V, (size_t)BT * N * C, st));
  // Latent-token attention: two fused kernels (attn_fused.cuh) when the shape qualifies, else the six-launch GEMM chain.
  const fa::TokAttnPlan fa_a = fa::make_tok_attn_plan(BT, N, C, K);
  const fa::XAttnPlan fa_b = fa::make_x_attn_plan(BT, N, C, K);
  const bool fused_attn = fused_attn_enabled(d, x, pp);
  const bool need_p = keep && !fused_attn_bwd_a_enabled(d, x, pp);     // (P is only read by the GEMM-chain backward)
  const int Kp = pad8(K);
  if (fused_attn) {
    {
      fa::TokAttnArgs a{};
      a.V = V; a.tokb = Wtok; a.tok32 = P.tokens; a.Pb = need_p ? Pb : nullptr;
      a.part_o = (float*)(ws + pl.fa_part_o); a.part_ml = (float*)(ws + pl.fa_part_ml);
      a.mref = (float*)(ws + pl.fa_mref); a.pfac = (float*)(ws + pl.fa_pfac); a.ticket = (unsigned*)(ws + pl.fa_ticket);
      a.stat_ml = (float*)(ws + pl.fa_stat);
      a.Tp = Tp; a.TpB = TpB; a.err_flag = timeout_flag_device_ptr();
      a.N = N; a.C = C; a.K = K; a.Kp = Kp; a.ld_tok = pp.ld_c; a.R = fa_a.R; a.nchunks = fa_a.nchunks; a.splits = fa_a.splits;
      a.zcols = fa_a.zcols; a.nbuf = fa_a.nbuf;
      StageScope _sc(ST_TOK_UPDATE, st); count_launch();
      cudaError_t e = fa::launch_tok_attn(a, fa_a, BT, st);
      if (e != cudaSuccess) { set_cuda_error(e, "tok_attn_fused"); return DGSCT_ERR_CUDA; }
    }
    if (taps && taps->Tp) DGSCT_RC(tap(Tp, 0, taps->Tp, (size_t)BT * K * C, st));
    {
      fa::XAttnArgs a{};
      a.x = (const bf16*)x; a.TpB = TpB; a.X1 = X1; a.Qb = Qb; a.Pb = need_p ? Pb : nullptr; a.pfac = (const float*)(ws + pl.fa_pfac);
      a.gate_av = P.gate_av; a.err_flag = timeout_flag_device_ptr();
      a.N = N; a.C = C; a.K = K; a.Kp = Kp; a.chunk_rows = fa_a.R; a.nchunks = fa_a.nchunks;
      StageScope _sc(ST_X_UPDATE, st); count_launch();
      cudaError_t e = fa::launch_x_attn(a, fa_b, BT, st);
      if (e != cudaSuccess) { set_cuda_error(e, "x_attn_fused"); return DGSCT_ERR_CUDA; }
    }
  } else {
  const bool tok_tc = true;                 // any K: P / Q pitches are padded, ragged N / K edges are TMA zero fill
  // (T1) L[b,n,k] = V[b,n,:] . tok[k,:]  (fp32) ; P = softmax over tokens (bf16)
  {
    auto te = tc_store<false>(L, K, (long long)N * K);
    EpiBiasStore se = epi_store(L, 0, K, (long long)N * K);
    DGSCT_RC(gemm_auto(ST_TOK_LOGITS, "tok_logits", tok_tc, kmaj(V, N, C, C, (long long)N * C), kmaj(Wtok, K, C, pp.ld_c, 0),
                       N, K, C, BT, te, se, st));
    StageScope _sc(ST_COL_SOFTMAX, st); count_launch();
    if (K <= 256) {
      const int groups = 256 / K;
      const dim3 grid(csm_num_chunks(N), BT);
      float* part = (float*)(ws + pl.csm_part);
      count_launch();
      col_softmax_partial_kernel<<<grid, 256, (size_t)2 * groups * K * 4, st>>>(L, part, N, K);
      col_softmax_apply_kernel<<<grid, 256, 0, st>>>(L, part, Pb, N, K, Kp);
    } else {
      return DGSCT_ERR_UNSUPPORTED;
    }
    DGSCT_LAUNCH_CHECK("col_softmax_bf16");
  }
  // (T3) Tp[b,k,c] = tok[k,c] + sum_n P[b,n,k] V[b,n,c]     computed as D(c,k): both operands token-major
  {
    tc::EpiTokUpdate te{TpB, Tp, P.tokens, C, K};
    // SIMT twin computes the same D(c,k) orientation: rows m = c, columns n = k
    SimtTokUpdate se{TpB, Tp, P.tokens, C, K};
    DGSCT_RC(gemm_auto(ST_TOK_UPDATE, "tok_update", tok_tc, mnmaj(V, N, C, C, (long long)N * C), mnmaj(Pb, N, K, Kp, (long long)N * Kp),
                       C, K, N, BT, te, se, st));
  }
  if (taps && taps->Tp) DGSCT_RC(tap(Tp, 0, taps->Tp, (size_t)BT * K * C, st));
  // (T4a) Q[b,n,:] = softmax_k(x[b,n,:] . Tp[b,k,:])  (bf16, reuses the P buffer)
  if (K == 32) {
    tc::EpiRowSoftmax32 te{}; te.out = Qb; te.ld = Kp; te.sb = (long long)N * Kp;
    EpiBiasStore se = epi_store(L, 0, K, (long long)N * K);   // unused twin (operands are always TMA-describable here)
    DGSCT_RC(gemm_auto(ST_X_LOGITS, "x_logits", true, kmaj(x, N, C, C, (long long)N * C), kmaj(TpB, K, C, C, (long long)K * C),
                       N, K, C, BT, te, se, st));
  } else {
    auto te = tc_store<false>(L, K, (long long)N * K);
    EpiBiasStore se = epi_store(L, 0, K, (long long)N * K);
    DGSCT_RC(gemm_auto(ST_X_LOGITS, "x_logits", true, kmaj(x, N, C, C, (long long)N * C), kmaj(TpB, K, C, C, (long long)K * C),
                       N, K, C, BT, te, se, st));
    StageScope _sc(ST_ROW_SOFTMAX, st); count_launch(2);
    row_softmax_kernel<<<(unsigned)((rowsBT + 7) / 8), 256, 0, st>>>(L, rowsBT, K);
    // Q as bf16 for the next product
    pack_rows_kernel<<<(unsigned)(((size_t)rowsBT * K + 255) / 256 > 4096 ? 4096 : ((size_t)rowsBT * K + 255) / 256), 256, 0, st>>>(
        L, (int)rowsBT, K, Qb, Kp);
    DGSCT_LAUNCH_CHECK("row_softmax");
  }
  // (T4b) X1 = x + gate_av * Q . Tp
  {
    tc::EpiResidual te{X1, (const bf16*)x, C, (long long)N * C, P.gate_av, 0.f};
    EpiResidual se{X1, x, 1, C, (long long)N * C, P.gate_av};
    DGSCT_RC(gemm_auto(ST_X_UPDATE, "x_update", tok_tc, kmaj(Qb, N, K, Kp, (long long)N * Kp), mnmaj(TpB, K, C, C, (long long)K * C),
                       N, C, K, BT, te, se, st));
  }
  }
  if (taps) DGSCT_RC(tap(X1, 1, taps->X1, (size_t)BT * N * C, st));
  // (C1) joins here: [aq1 | aq2] (and tau) were computed on the auxiliary stream underneath the token attention
  if (side) {
    DGSCT_CUDA_CHECK(cudaEventRecord(side->ev, side->s));
    DGSCT_CUDA_CHECK(cudaStreamWaitEvent(st, side->ev, 0));
  } else {
    DGSCT_RC(c1_chain());
  }
  if (taps && taps->obar) DGSCT_RC(tap(osum, 0, taps->obar, (size_t)BT * C, st));
  // (C2) qsum[b,c] = sum_n aq1[b,c] relu(X1[b,n,:] . Wv1[c,:] + b)
  {
    tc::EpiReluMulColsum te{P.v1_b, nullptr, qsum, C, aq_ld};     // r1sum[b,c] = sum_n relu(z1), unscaled
    EpiReluMulColsum se{P.v1_b, nullptr, qsum, C, aq_ld};
    DGSCT_RC(gemm_auto(ST_CHANNEL_QUERY, "channel_query", true, kmaj(X1, N, C, C, (long long)N * C), kmaj(Wv1, C, C, pp.ld_c, 0),
                       N, C, C, BT, te, se, st));
  }
  // (C3) m = sigmoid(Wc relu(Wb q + bb) + bc)   (two small tensor-core linears, h kept in bf16)
  {
    bf16* hb = (bf16*)hbuf;                 // (BT, ld_h) bf16 in the fp32-sized h buffer
    { StageScope _sc(ST_SMALL_LINEAR, st); count_launch();
      channel_q_kernel<<<ceil_div(BT * C, 256), 256, 0, st>>>(aq1, aq_ld, qsum, invN, qvec, vecb, BT, C);
      DGSCT_LAUNCH_CHECK("channel_q"); }
    auto te = tc_store<true, false, false, ACT_RELU>(hb, pp.ld_h, 0);
    te.bias_n = P.bott_b;
    EpiBiasStore se = epi_store(hb, 1, pp.ld_h, 0); se.bias_n = P.bott_b; se.act = ACT_RELU;
    int rcv = vec_linear(ST_SMALL_LINEAR, "channel_bottleneck", vecb, C, pk + pp.bott, pp.ld_c, 0, P.bott_b, ACT_RELU, nullptr, 0, hb,
                         pp.ld_h, BT, H, C, 0, st);
    if (rcv == DGSCT_ERR_UNSUPPORTED)
      rcv = gemm_auto(ST_SMALL_LINEAR, "channel_bottleneck", true, kmaj(vecb, BT, C, C, 0), kmaj(pk + pp.bott, H, C, pp.ld_c, 0),
                      BT, H, C, 1, te, se, st, 64);
    DGSCT_RC(rcv);
    auto te2 = tc_store<false, false, false, ACT_SIGMOID>(mvec, C, 0);
    te2.bias_n = P.c_b;
    EpiBiasStore se2 = epi_store(mvec, 0, C, 0); se2.bias_n = P.c_b; se2.act = ACT_SIGMOID;
    rcv = vec_linear(ST_SMALL_LINEAR, "channel_gate", hb, pp.ld_h, pk + pp.catt, pp.ld_h, 0, P.c_b, ACT_SIGMOID, mvec, C, nullptr, 0,
                     BT, C, H, 0, st);
    if (rcv == DGSCT_ERR_UNSUPPORTED)
      rcv = gemm_auto(ST_SMALL_LINEAR, "channel_gate", true, kmaj(hb, BT, H, pp.ld_h, 0), kmaj(pk + pp.catt, C, H, pp.ld_h, 0),
                      BT, C, H, 1, te2, se2, st, 64);
    DGSCT_RC(rcv);
  }
  if (taps && taps->m) DGSCT_RC(tap(mvec, 0, taps->m, (size_t)BT * C, st));
  // (S1) Xc = X1 (1+m) (bf16, in the U buffer) ; ssum[b,n] = sum_j ws[j] aq2[b,j] relu(Xc[b,n,:] . Wv2[j,:] + b)
  {
    bf16* Xc = (bf16*)(ws + pl.XcBuf);
    { StageScope _sc(ST_SPATIAL_QUERY, st); count_launch();
      launch_scale_cols(X1, mvec, Xc, rowsBT, N, C, st);
      DGSCT_LAUNCH_CHECK("scale_cols"); }
    tc::EpiSpatialRowsum te{P.v2_b, aq2, P.s_w, ssum, aq_ld, N, 0.f};
    EpiSpatialRowsum se{P.v2_b, aq2, P.s_w, ssum, aq_ld, N};
    DGSCT_RC(gemm_auto(ST_SPATIAL_QUERY, "spatial_query", true, kmaj(Xc, N, C, C, (long long)N * C), kmaj(Wv2, H, C, pp.ld_c, 0),
                       N, H, C, BT, te, se, st));
  }
  // The spatial map (softmax_n tanh s: an output, and the sigmoid gate the backward reads) is off the critical path: gate_ln
  // derives the gate from the logits itself, so the map kernel runs on the auxiliary stream beside gate_ln / the bottleneck.
  const bool map_aside = side != nullptr;
  {
    const cudaStream_t sm_st = map_aside ? side->s : st;
    if (map_aside) {
      DGSCT_CUDA_CHECK(cudaEventRecord(side->ev, st));
      DGSCT_CUDA_CHECK(cudaStreamWaitEvent(side->s, side->ev, 0));
    }
    StageScope _sc(ST_SPATIAL_MAP, sm_st); count_launch();
    spatial_map_kernel<<<BT, 256, 0, sm_st>>>(ssum, P.s_b, N, sig, smap, 1, taps ? taps->s : nullptr);
    DGSCT_LAUNCH_CHECK("spatial_map");
    if (map_aside) DGSCT_CUDA_CHECK(cudaEventRecord(side->ev, side->s));
  }
  // (B1) Y = LN_before(X1 * (alpha m + beta sig + gamma tau + 1 - alpha))  -> V buffer
  {
    GateLnArgs a{};
    a.x1 = X1; a.in_bf16 = 1; a.y = Yb; a.out_bf16 = 1; a.m = mvec; a.sig = sig;
    if (map_aside) { a.ssum = ssum; a.s_bias = P.s_b; }
    a.tau = ((d.flags & DGSCT_F_TEMPORAL) && d.gamma != 0.f) ? tau : nullptr;
    const bool lnb = d.flags & DGSCT_F_LN_BEFORE;
    a.ln_w = lnb ? P.lnb_w : nullptr; a.ln_b = lnb ? P.lnb_b : nullptr;
    a.alpha = d.alpha; a.beta = d.beta; a.gamma = d.gamma; a.eps = d.eps_ln; a.N = N; a.C = C;
    a.rows = rowsBT; a.y_tap = taps ? taps->Y : nullptr;
    const int wpb = 4;
    StageScope _sc(ST_GATE_LN, st); count_launch();
    if (!launch_gate_ln_vec(a, st))
      gate_ln_kernel<<<(unsigned)((a.rows + wpb - 1) / wpb), wpb * 32, (size_t)wpb * C * 4, st>>>(a);
    DGSCT_LAUNCH_CHECK("gate_ln");
  }
  const double count = (double)BT * (double)N;
  const bool fuse_bn1 = !(train && use_bn) && !keep;     // bn1 affine known before the down projection
  if (fuse_bn1) {
    StageScope _sc(ST_BN_FINALIZE, st); count_launch();
    bn_finalize_kernel<<<ceil_div(Cd, 128), 128, 0, st>>>(Cd, use_bn, 0, bn1_sum, bn1_sq, count, P.bn1_w, P.bn1_b, P.bn1_rm,
                                                           P.bn1_rv, d.eps_bn, d.bn_momentum, bn1_scale, bn1_shift,
                                                           (float*)(ws + pl.bn1_mean), (float*)(ws + pl.bn1_rstd), upd);
    DGSCT_LAUNCH_CHECK("bn1_finalize");
  }
  // (B2) Z = Y . Wdown^T (dense block-diagonal) ; eval: Zr = relu(bn1(Z)) fused into the epilogue
  {
    if (fuse_bn1) {
      tc::EpiAffineRelu te{Zr, ldz, N, bn1_scale, bn1_shift, taps ? taps->Z : nullptr, Cd};
      SimtAffineRelu se{Zr, ldz, (long long)N * ldz, bn1_scale, bn1_shift, taps ? taps->Z : nullptr, Cd};
      DGSCT_RC(gemm_auto(ST_DOWN, "down_sampler", true, kmaj(Yb, N, C, C, (long long)N * C), kmaj(Wdn, Cd, C, pp.ld_c, 0),
                         N, Cd, C, BT, te, se, st));
    } else {
      auto te = tc_store<false, false, true>(Z, Cd, (long long)N * Cd);
      EpiBiasStore se = epi_store(Z, 0, Cd, (long long)N * Cd);
      te.stat_sum = se.stat_sum = bn1_sum; te.stat_sq = se.stat_sq = bn1_sq;   // (harmless extra moments when BN is frozen)
      DGSCT_RC(gemm_auto(ST_DOWN, "down_sampler", true, kmaj(Yb, N, C, C, (long long)N * C), kmaj(Wdn, Cd, C, pp.ld_c, 0),
                         N, Cd, C, BT, te, se, st));
      if (taps && taps->Z) DGSCT_RC(tap(Z, 0, taps->Z, (size_t)BT * N * Cd, st));
      StageScope _sc(ST_BN_FINALIZE, st); count_launch();
      {   // finalize + bn1 + ReLU in one launch (the running-stat race: only block 0 reads-modifies-writes rm / rv, nobody reads them
          // in train mode; in eval mode they are read-only)
        const long long tot = rowsBT * Cd;
        long long blocks = (tot + 255) / 256;
        if (blocks > 2048) blocks = 2048;
        if (blocks < 1) blocks = 1;
        bn_finalize_relu_kernel<<<(unsigned)blocks, 256, (size_t)2 * Cd * 4, st>>>(Z, Zr, rowsBT, Cd, ldz, use_bn, (train && use_bn) ? 1 : 0, bn1_sum,
                                                                                 bn1_sq, count, P.bn1_w, P.bn1_b, P.bn1_rm, P.bn1_rv, d.eps_bn,
                                                                                 d.bn_momentum, bn1_scale, bn1_shift, (float*)(ws + pl.bn1_mean),
                                                                                 (float*)(ws + pl.bn1_rstd), upd);
      }
      DGSCT_LAUNCH_CHECK("bn1_finalize");
    }
  }
  // (B3) Uo = Zr . Wup^T (dense block-diagonal) ; bn2 batch statistics in train mode
  {
    EpiBiasStore se = epi_store(Uo, 0, C, (long long)N * C);
    if (train && use_bn) {
      auto te = tc_store<false, false, true>(Uo, C, (long long)N * C);
      te.stat_sum = se.stat_sum = bn2_sum; te.stat_sq = se.stat_sq = bn2_sq;
      DGSCT_RC(gemm_auto(ST_UP, "up_sampler", true, kmaj(Zr, N, Cd, ldz, (long long)N * ldz), kmaj(Wup, C, Cd, pp.ld_up, 0),
                         N, C, Cd, BT, te, se, st));
    } else {
      auto te = tc_store<false>(Uo, C, (long long)N * C);
      DGSCT_RC(gemm_auto(ST_UP, "up_sampler", true, kmaj(Zr, N, Cd, ldz, (long long)N * ldz), kmaj(Wup, C, Cd, pp.ld_up, 0),
                         N, C, Cd, BT, te, se, st));
    }
  }
  { StageScope _sc(ST_BN_FINALIZE, st); count_launch();
    bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(C, use_bn, train, bn2_sum, bn2_sq, count, P.bn2_w, P.bn2_b, P.bn2_rm,
                                                          P.bn2_rv, d.eps_bn, d.bn_momentum, bn2_scale, bn2_shift,
                                                          (float*)(ws + pl.bn2_mean), (float*)(ws + pl.bn2_rstd), upd);
    DGSCT_LAUNCH_CHECK("bn2_finalize"); }
  if (!(d.flags & DGSCT_F_NO_OUTPUT)) {
    OutLnArgs a{};
    a.u = Uo; a.in_bf16 = 0; a.out = out; a.out_bf16 = 1; a.scale = bn2_scale; a.shift = bn2_shift;
    const bool lnp = d.flags & DGSCT_F_LN_POST;
    a.ln_w = lnp ? P.lnp_w : nullptr; a.ln_b = lnp ? P.lnp_b : nullptr;
    a.gate = (d.flags & DGSCT_F_USE_GATE) ? P.gate : nullptr;
    a.gate_before_ln = (d.flags & DGSCT_F_GATE_BEFORE_LN) ? 1 : 0;
    a.eps = d.eps_ln; a.C = C; a.rows = rowsBT; a.resid = resid;
    const int wpb = 4;
    StageScope _sc(ST_OUT_LN, st); count_launch();
    if (!launch_out_ln_vec(a, st))
      out_ln_kernel<<<(unsigned)((a.rows + wpb - 1) / wpb), wpb * 32, (size_t)wpb * C * 4, st>>>(a);
    DGSCT_LAUNCH_CHECK("out_ln");
  }
  if (map_aside) DGSCT_CUDA_CHECK(cudaStreamWaitEvent(st, side->ev, 0));   // the spatial map is an output of the call
  return DGSCT_OK;
}

}  // namespace dgsct

#include "adapter_bwd.inl"

using namespace dgsct;
extern "C" {
// launch geometry of the fused attention kernels (host only; see include/dgsct.h)
int dgsct_debug_attn_plan(int kernel, int BT, int N, int C, int K, int* out8) {
  if (!out8) return DGSCT_ERR_NULL;
  if (BT < 1 || N < 1 || C < 1 || K < 1) return DGSCT_ERR_BAD_DESC;
  for (int i = 0; i < 8; ++i) out8[i] = 0;
  if (kernel == 0) {
    const fa::TokAttnPlan p = fa::make_tok_attn_plan(BT, N, C, K);
    if (p.ok) { out8[0] = 1; out8[1] = p.R; out8[2] = p.nbuf; out8[3] = p.nchunks; out8[4] = p.splits; out8[5] = p.Z; out8[6] = p.zcols; out8[7] = (int)p.smem; }
  } else if (kernel == 1) {
    const fa::XAttnPlan p = fa::make_x_attn_plan(BT, N, C, K);
    if (p.ok) { out8[0] = 1; out8[1] = p.WB; out8[2] = p.coop ? 1 : 0; out8[4] = p.gx; out8[5] = 1; out8[6] = C; out8[7] = (int)p.smem; }
  } else if (kernel == 2) {
    const fa::XAttnBwdPlan p = fa::make_x_attn_bwd_plan(BT, N, C, K);
    if (p.ok) { out8[0] = 1; out8[1] = p.R; out8[2] = p.nbuf; out8[3] = p.nchunks; out8[4] = p.splits; out8[5] = p.Z; out8[6] = p.zcols; out8[7] = (int)p.smem; }
  } else if (kernel == 3) {
    const fa::TokAttnBwdPlan p = fa::make_tok_attn_bwd_plan(BT, N, C, K);
    if (p.ok) { out8[0] = 1; out8[1] = p.R; out8[2] = p.nbuf; out8[3] = p.nchunks; out8[4] = p.splits; out8[5] = p.Z; out8[6] = p.zcols; out8[7] = (int)p.smem; }
  } else {
    return DGSCT_ERR_BAD_DESC;
  }
  return DGSCT_OK;
}
}  // extern "C"
