// VisualAdapter.forward as a sequence of CUDA kernels (reference: DG-SCT/AVE/nets/net_trans.py:552-674).
// This file is the orchestration + the exact-fp32 SIMT path; the bf16 tensor-core kernels (gemm_tc.cuh) are
// substituted for the large products when io_dtype is bf16.
#include "adapter_fwd.h"

#include "epilogues.cuh"
#include "profile.h"
#include "rowwise.cuh"

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

FwdPlan make_fwd_plan(const dgsct_desc& d) {
  FwdPlan p{};
  const size_t BT = d.BT, N = d.N, C = d.C, Co = d.Co, K = d.K, Cd = d.C / d.r, H = d.C / 2;
  const size_t es = d.io_dtype == DGSCT_BF16 ? 2 : 4;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  p.bufA = take(BT * N * (Co > C ? Co : C) * es);   // U (BT,N,Co) then up-sampler output (BT,N,C)
  p.bufV = take(BT * N * C * es);                   // V then Y
  p.bufX1 = take(BT * N * C * es);
  p.bufL = take(BT * N * K * 4);                    // token logits -> probabilities (P then Q), fp32
  p.Tp = take(BT * K * C * 4);
  p.Z = take(BT * N * Cd * 4);
  p.acc_begin = off;                                // ---- zero-initialised accumulators
  p.osum = take(BT * C * 4);
  p.qsum = take(BT * C * 4);
  p.ssum = take(BT * N * 4);
  p.bn1_sum = take(Cd * 8); p.bn1_sq = take(Cd * 8);
  p.bn2_sum = take(C * 8); p.bn2_sq = take(C * 8);
  p.acc_end = off;
  p.aq1 = take(BT * C * 4); p.aq2 = take(BT * H * 4); p.h = take(BT * H * 4); p.m = take(BT * C * 4);
  p.tau = take(BT * 4); p.sig = take(BT * N * 4);
  p.bn1_scale = take(Cd * 4); p.bn1_shift = take(Cd * 4);
  p.bn2_scale = take(C * 4); p.bn2_shift = take(C * 4);
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
                     void* smap, float* tmap, const dgsct_taps* taps, char* ws, cudaStream_t st) {
  const FwdPlan pl = make_fwd_plan(d);
  const int BT = d.BT, N = d.N, C = d.C, Co = d.Co, No = d.No, K = d.K, Cd = d.C / d.r, H = d.C / 2, G = d.g;
  const int bf = d.io_dtype == DGSCT_BF16;
  const bool train = d.flags & DGSCT_F_TRAIN, use_bn = d.flags & DGSCT_F_USE_BN;
  void* U = ws + pl.bufA; void* V = ws + pl.bufV; void* X1 = ws + pl.bufX1;
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
  {
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
    EpiBiasStore e = epi_store(L, 0, K, (long long)N * K);
    DGSCT_GEMM(ST_X_LOGITS, sh, op(x, C, 1, (long long)N * C, bf), op(Tp, C, 1, (long long)K * C, 0), AxIdentity{}, e, "x_logits");
  }
  {
    const long long rows = (long long)BT * N;
    { StageScope _sc(ST_ROW_SOFTMAX, st); count_launch();
  row_softmax_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(L, rows, K);
    DGSCT_LAUNCH_CHECK("row_softmax"); }
  }
  // (T4b) X1 = x + gate_av * Q . Tp                                             net_trans.py:586-589
  {
    GemmShape sh{N, C, K, BT, 1};
    EpiResidual e{X1, x, bf, C, (long long)N * C, P.gate_av};
    DGSCT_GEMM(ST_X_UPDATE, sh, op(L, K, 1, (long long)N * K, 0), op(Tp, 1, C, (long long)K * C, 0), AxIdentity{}, e, "x_update");
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
    EpiReluMulColsum e{P.v1_b, aq1, qsum, C};
    DGSCT_GEMM(ST_CHANNEL_QUERY, sh, op(X1, C, 1, (long long)N * C, bf), op(P.v1_w, C, 1, 0, 0), AxIdentity{}, e, "channel_query");
  }
  // (C3) m = sigmoid(Wc relu(Wb q + bb) + bc)                                   net_trans.py:596-597
  {
    GemmShape sh{BT, H, C, 1, 1};
    EpiBiasStore e = epi_store(hbuf, 0, H, 0); e.bias_n = P.bott_b; e.alpha = invN; e.act = ACT_RELU;
    DGSCT_GEMM(ST_SMALL_LINEAR, sh, op(qsum, C, 1, 0, 0), op(P.bott_w, C, 1, 0, 0), AxIdentity{}, e, "channel_bottleneck");
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
    a.x1 = X1; a.in_bf16 = bf; a.y = V; a.out_bf16 = bf; a.m = mvec; a.sig = sig;
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
    DGSCT_GEMM(ST_DOWN, sh, op(V, C, 1, (long long)N * C, bf, C / G), op(P.down_w, C / G, 1, 0, 0, (long long)(Cd / G) * (C / G)),
               AxIdentity{}, e, "down_sampler");
  }
  if (taps && taps->Z) { int r = tap(Z, 0, taps->Z, (size_t)BT * N * Cd, st); if (r) return r; }
  const double count = (double)BT * (double)N;
  { StageScope _sc(ST_BN_FINALIZE, st); count_launch();
  bn_finalize_kernel<<<ceil_div(Cd, 128), 128, 0, st>>>(Cd, use_bn, train, bn1_sum, bn1_sq, count, P.bn1_w, P.bn1_b,
                                                         P.bn1_rm, P.bn1_rv, d.eps_bn, d.bn_momentum, bn1_scale, bn1_shift);
  DGSCT_LAUNCH_CHECK("bn1_finalize"); }
  // (B3) Uo = up_sampler(relu(bn1(Z))) ; bn2 batch statistics                   net_trans.py:638-643
  {
    GemmShape sh{N, C / G, Cd / G, BT, G};
    EpiBiasStore e = epi_store(U, bf, C, (long long)N * C);
    if (train && use_bn) { e.stat_sum = bn2_sum; e.stat_sq = bn2_sq; }
    DGSCT_GEMM(ST_UP, sh, op(Z, Cd, 1, (long long)N * Cd, 0, Cd / G), op(P.up_w, Cd / G, 1, 0, 0, (long long)(C / G) * (Cd / G)),
               (AxAffineRelu{bn1_scale, bn1_shift, Cd / G}), e, "up_sampler");
  }
  { StageScope _sc(ST_BN_FINALIZE, st); count_launch();
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(C, use_bn, train, bn2_sum, bn2_sq, count, P.bn2_w, P.bn2_b,
                                                        P.bn2_rm, P.bn2_rv, d.eps_bn, d.bn_momentum, bn2_scale, bn2_shift);
  DGSCT_LAUNCH_CHECK("bn2_finalize"); }
  // (B4) out = gate * LN_post(bn2(Uo))  (or AVS order)                          net_trans.py:667-672
  {
    OutLnArgs a{};
    a.u = U; a.in_bf16 = bf; a.out = out; a.out_bf16 = bf; a.scale = bn2_scale; a.shift = bn2_shift;
    const bool lnp = d.flags & DGSCT_F_LN_POST;
    a.ln_w = lnp ? P.lnp_w : nullptr; a.ln_b = lnp ? P.lnp_b : nullptr;
    a.gate = (d.flags & DGSCT_F_USE_GATE) ? P.gate : nullptr;
    a.gate_before_ln = (d.flags & DGSCT_F_GATE_BEFORE_LN) ? 1 : 0;
    a.eps = d.eps_ln; a.C = C; a.rows = (long long)BT * N;
    const int wpb = 4;
    { StageScope _sc(ST_OUT_LN, st); count_launch();
  out_ln_kernel<<<(unsigned)((a.rows + wpb - 1) / wpb), wpb * 32, (size_t)wpb * C * 4, st>>>(a);
    DGSCT_LAUNCH_CHECK("out_ln"); }
  }
  return DGSCT_OK;
}

}  // namespace dgsct
