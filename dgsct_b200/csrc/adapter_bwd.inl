// Backward of VisualAdapter.forward (gradients w.r.t. x, vis_token and every parameter; reference = autograd of
// DG-SCT/AVE/nets/net_trans.py:552-674).  The call first RE-RUNS the forward into its own workspace with a
// non-aliasing plan (activation recomputation: the training forward keeps nothing and costs the same memory as
// inference), then walks the stages in reverse.  Large products go through gemm_auto (tcgen05 when the operands are
// bf16 and TMA-describable, SIMT otherwise); everything else is in bwd_kernels.cuh.
// (textually included at the end of adapter_fwd.cu: one translation unit owns every header-defined kernel)
#include "bwd_kernels.cuh"

#include <stdlib.h>

#include <vector>

namespace dgsct {

// ---- generic output epilogues for the backward products ------------------------------------------------------------
enum { OUT_STORE = 0, OUT_ADD = 1, OUT_ATOMIC = 2 };

// SIMT: value = alpha * acc * colscale[b, n] (+ resid[b,m,n]) (+ bcast[b, n]); store / add / atomicAdd; optionally
// written transposed (out[b, n, m]).
struct EpiOut {
  void* out; int out_bf16; long long ld, sb; int mode; int transposed;
  float alpha; const float* alpha_ptr;                  // alpha_ptr (device scalar) multiplies alpha when non-NULL
  const float* colscale; int cs_ld;                     // per-frame per-column scale (1+m), or NULL
  const void* resid; int resid_bf16;                    // same indexing as out
  const float* bcast; int bc_ld; float bc_scale;        // + bc_scale * bcast[b, n]
  __device__ __forceinline__ void operator()(int b, int gi, int m, int n, float (&acc)[4][4], int M, int N) const {
    const float al = alpha * (alpha_ptr ? alpha_ptr[0] : 1.f);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int gm = m + i, gn = n + j;
        if (gm < M && gn < N) {
          float v = al * acc[i][j];
          if (colscale) v *= colscale[(long long)b * cs_ld + gn];
          if (bcast) v += bc_scale * bcast[(long long)b * bc_ld + gn];
          const size_t idx = transposed ? (size_t)((long long)b * sb + (long long)gn * ld + gm)
                                        : (size_t)((long long)b * sb + (long long)gm * ld + gn);
          if (resid) v += ld_act(resid, idx, resid_bf16);
          if (mode == OUT_ATOMIC) atomicAdd(reinterpret_cast<float*>(out) + idx, v);
          else if (mode == OUT_ADD) st_act(out, idx, ld_act(out, idx, out_bf16) + v, out_bf16);
          else st_act(out, idx, v, out_bf16);
        }
      }
  }
};

namespace tc {
// tensor-core twin (row-major output only; transposed outputs use EpiOutT)
template <bool OUT_BF16, int MODE, bool RESID, bool BCAST, bool CSCALE>
struct EpiOut {
  // staged (row-contiguous) I/O only for pure stores: staging a residual / accumulate LOAD measured slower than the direct
  // per-row access (see EpiResidual)
  static constexpr bool kHasChunk32 = (MODE == OUT_STORE) && !RESID;
  static constexpr bool kHasPair = OUT_BF16 && (RESID || MODE == OUT_ADD);
  static constexpr int kStageBytes = OUT_BF16 ? 2048 : 4096;
  static constexpr bool kRowIndependent = true;
  void* out; long long ld, sb; float alpha; const float* alpha_ptr; float al;
  const float* colscale; int cs_ld; const bf16* resid; const float* bcast; int bc_ld; float bc_scale;
  __device__ __forceinline__ void begin(int, int, bool) { al = alpha * (alpha_ptr ? alpha_ptr[0] : 1.f); }
  __device__ __forceinline__ void end(int, int, bool) {}
  __device__ __forceinline__ void transform(int b, int n0, float (&v)[16], int nvalid) const {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] *= al;
    if (CSCALE) {
      float cs[16];
      load16_guard(colscale + (long long)b * cs_ld + n0, cs, nvalid);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] *= cs[j];
    }
    if (BCAST) {
      float bc[16];
      load16_guard(bcast + (long long)b * bc_ld + n0, bc, nvalid);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaf(bc_scale, bc[j], v[j]);
    }
  }
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
    const int nvalid = nvalid16(n0, N);
    if (!row_ok || nvalid <= 0) return;
    const size_t idx = (size_t)((long long)b * sb + (long long)m * ld + n0);
    transform(b, n0, v, nvalid);
    if (RESID) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nvalid) v[j] += __bfloat162float(resid[idx + j]);
    }
    if (MODE == OUT_ATOMIC) {
      red_add16(reinterpret_cast<float*>(out) + idx, v, nvalid);
    } else {
      if (MODE == OUT_ADD) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < nvalid) v[j] += OUT_BF16 ? __bfloat162float(reinterpret_cast<const bf16*>(out)[idx + j])
                                           : reinterpret_cast<const float*>(out)[idx + j];
      }
      store_row16<OUT_BF16>(out, idx, v, nvalid);
    }
  }
  // read-modify-write blocks (bf16 residual and / or accumulate into a bf16 output): both 16-column halves at once, every
  // 128-bit load of the block issued before the first dependent use
  __device__ __forceinline__ void pair(int b, int m, int n0, float (&v0)[16], float (&v1)[16], bool two, bool row_ok, int N) {
    const int nv0 = nvalid16(n0, N), nv1 = two ? nvalid16(n0 + 16, N) : 0;
    const size_t idx = (size_t)((long long)b * sb + (long long)m * ld + n0);
    const bool fast = row_ok && nv0 == 16 && nv1 == 16 && ((idx * 2) & 15) == 0;
    if (!__all_sync(FULL, fast || !row_ok) || !__any_sync(FULL, fast)) {      // ragged / misaligned block somewhere in the warp
      chunk(b, m, n0, v0, row_ok, N);
      if (two) chunk(b, m, n0 + 16, v1, row_ok, N);
      return;
    }
    Raw16 r0{}, r1{}, a0{}, a1{};
    if (fast) {
      if (RESID) { r0 = ldg_raw16(resid + idx); r1 = ldg_raw16(resid + idx + 16); }
      if (MODE == OUT_ADD) { a0 = ldg_raw16(reinterpret_cast<const bf16*>(out) + idx); a1 = ldg_raw16(reinterpret_cast<const bf16*>(out) + idx + 16); }
    }
    const int bs = row_ok ? b : 0;
    transform(bs, n0, v0, 16);
    transform(bs, n0 + 16, v1, 16);
    if (!fast) return;
    if (RESID) { add_raw16(r0, v0); add_raw16(r1, v1); }
    if (MODE == OUT_ADD) { add_raw16(a0, v0); add_raw16(a1, v1); }
    store_row16<true>(out, idx, v0, 16);
    store_row16<true>(out, idx + 16, v1, 16);
  }
  // 32 columns through the warp's staging buffer: residual / accumulate loads and the store are row-contiguous
  __device__ __forceinline__ void chunk32(int b, int m, int n0, float (&v0)[16], float (&v1)[16], bool two, bool row_ok,
                                          int N, uint8_t* stage) {
    const int nv0 = nvalid16(n0, N), nv1 = two ? nvalid16(n0 + 16, N) : 0;
    const long long off = (long long)b * sb + (long long)m * ld;
    const int bs = row_ok ? b : 0;                       // rows past M still execute (warp-cooperative I/O): keep loads in range
    transform(bs, n0, v0, nv0);
    if (two) transform(bs, n0 + 16, v1, nv1);
    if (RESID) {
      float r0[16], r1[16];
      warp_load_block32_bf16(resid, off, n0, nv0 + nv1, row_ok, stage, r0, r1);
#pragma unroll
      for (int j = 0; j < 16; ++j) { v0[j] += r0[j]; v1[j] += r1[j]; }
    }
    if (MODE == OUT_ADD && OUT_BF16) {
      float r0[16], r1[16];
      warp_load_block32_bf16(reinterpret_cast<const bf16*>(out), off, n0, nv0 + nv1, row_ok, stage, r0, r1);
#pragma unroll
      for (int j = 0; j < 16; ++j) { v0[j] += r0[j]; v1[j] += r1[j]; }
    }
    warp_store_block32<OUT_BF16>(out, off, n0, v0, v1, nv0 + nv1, row_ok, stage);
  }
};
struct EpiNone {   // placeholder for products that never take the tensor-core kernel
  __device__ __forceinline__ void begin(int, int, bool) {}
  __device__ __forceinline__ void end(int, int, bool) {}
  __device__ __forceinline__ void chunk(int, int, int, float (&)[16], bool, int) {}
};
// transposed: accumulator is D(m, n); writes out[b, n, m] (fp32) -- lanes are consecutive m: coalesced
template <int MODE>
struct EpiOutT {
  float* out; long long ld, sb; float alpha; const float* alpha_ptr; float al;
  __device__ __forceinline__ void begin(int, int, bool) { al = alpha * (alpha_ptr ? alpha_ptr[0] : 1.f); }
  __device__ __forceinline__ void end(int, int, bool) {}
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
    const int nvalid = nvalid16(n0, N);
    if (!row_ok) return;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < nvalid) {
        float* o = out + (long long)b * sb + (long long)(n0 + j) * ld + m;
        if (MODE == OUT_ATOMIC) atomicAdd(o, al * v[j]);
        else if (MODE == OUT_ADD) *o += al * v[j];
        else *o = al * v[j];
      }
  }
};
}  // namespace tc

// ---- dispatcher (same policy as the forward's gemm_auto) ------------------------------------------------------------
struct Gemm {
  int stage; const char* name; cudaStream_t st; bool allow_tc;
};

static GemmOperand simt_operand(const tc::OperandDesc& o, int is_bf16) {
  return GemmOperand{o.ptr, o.mn_major ? 1 : o.ld, o.mn_major ? o.ld : 1, o.batch_stride, 0, is_bf16};
}

// sum_b: the product is a weight gradient sum_b A_b . B_b^T -- the tensor-core kernel accumulates groups of frames inside
// the tile (its functor is called with the frame-group index and must add atomically); the SIMT twin runs one frame per
// batch entry and adds atomically.
template <typename TE, typename AX>
static int run_gemm(const Gemm& g, const tc::OperandDesc& A, int a_bf16, const tc::OperandDesc& B, int b_bf16, int M, int N, int K,
                    int batch, TE te, const EpiOut& se, AX ax, bool ax_identity, int bn_hint = 0, bool sum_b = false,
                    bool split_k_ok = false, bool* used_tc = nullptr) {
  StageScope _sc(g.stage, g.st);
  count_launch();
  if (used_tc) *used_tc = false;
  if (g.allow_tc && a_bf16 && b_bf16 && ax_identity) {
    tc::Params p;
    const int rc = tc::make_params(p, A, B, M, N, K, batch, bn_hint, 0, sum_b, tc::epi_row_independent<TE>::value,
                                   tc::epi_stage_bytes<TE>::value);
    if (rc == DGSCT_OK) {
      cudaError_t e = tc::launch(p, te, g.st);
      if (e != cudaSuccess) { set_cuda_error(e, g.name); return DGSCT_ERR_CUDA; }
      if (used_tc) *used_tc = true;
      return DGSCT_OK;
    }
    if (rc != DGSCT_ERR_UNSUPPORTED) return rc;
  }
  GemmShape sh{M, N, K, batch, 1, 1};
  EpiOut se2 = se;
  // skinny per-frame-vector products (M = BT rows): split the contraction over CTAs and add atomically.  The output is
  // fp32 and either accumulating (OUT_ADD) or zero-initialised by the caller (OUT_STORE targets live in the zero region).
  // (additive epilogue terms -- residual, broadcast row -- would be applied once per split: such products never split)
  if (split_k_ok && M <= 128 && K >= 256 && batch == 1 && !se.out_bf16 && !se.resid && !se.bcast && se.mode != OUT_ATOMIC) {
    const int ctas = ceil_div(N, SG_BN) * ceil_div(M, SG_BM);
    int ks = (2 * 148 + ctas - 1) / ctas;
    if (ks > K / 64) ks = K / 64;
    if (ks > 1) { sh.ksplit = ks; se2.mode = OUT_ATOMIC; }
  }
  cudaError_t e = launch_gemm_simt(sh, simt_operand(A, a_bf16), simt_operand(B, b_bf16), ax, se2, g.st);
  if (e != cudaSuccess) { set_cuda_error(e, g.name); return DGSCT_ERR_CUDA; }
  return DGSCT_OK;
}

// tensor-core kernel or nothing: DGSCT_ERR_UNSUPPORTED tells the caller to take its unfused path
template <typename TE>
static int run_tc_only(const Gemm& g, const tc::OperandDesc& A, const tc::OperandDesc& B, int M, int N, int K, int batch, TE te) {
  tc::Params p;
  const int rc = tc::make_params(p, A, B, M, N, K, batch, 0, 0, false, tc::epi_row_independent<TE>::value,
                                 tc::epi_stage_bytes<TE>::value);
  if (rc != DGSCT_OK) return rc;
  StageScope _sc(g.stage, g.st);
  count_launch();
  cudaError_t e = tc::launch(p, te, g.st);
  if (e != cudaSuccess) { set_cuda_error(e, g.name); return DGSCT_ERR_CUDA; }
  return DGSCT_OK;
}

// same, for functors that keep per-CTA column partials in shared memory (kBlockStats)
template <typename TE>
static int run_tc_only_stats(const Gemm& g, const tc::OperandDesc& A, const tc::OperandDesc& B, int M, int N, int K, int batch, TE te) {
  tc::Params p;
  const int rc = tc::make_params(p, A, B, M, N, K, batch, 0, 0, false, tc::epi_row_independent<TE>::value,
                                 tc::epi_stage_bytes<TE>::value, 1);
  if (rc != DGSCT_OK) return rc;
  StageScope _sc(g.stage, g.st);
  count_launch();
  cudaError_t e = tc::launch(p, te, g.st);
  if (e != cudaSuccess) { set_cuda_error(e, g.name); return DGSCT_ERR_CUDA; }
  return DGSCT_OK;
}

static tc::OperandDesc K_(const void* p, int rows, int cols, long long ld, long long bs) { return tc::OperandDesc{p, 0, ld, bs, rows, cols}; }
static tc::OperandDesc MN_(const void* p, int k_rows, int mn_cols, long long ld, long long bs) { return tc::OperandDesc{p, 1, ld, bs, k_rows, mn_cols}; }

static EpiOut eo(void* out, int out_bf16, long long ld, long long sb, int mode) {
  EpiOut e{}; e.out = out; e.out_bf16 = out_bf16; e.ld = ld; e.sb = sb; e.mode = mode; e.alpha = 1.f;
  return e;
}
template <typename T>
static T to_tc(const EpiOut& e) {   // copy the shared fields into a tensor-core functor
  T t{}; t.out = e.out; t.ld = e.ld; t.sb = e.sb; t.alpha = e.alpha; t.alpha_ptr = e.alpha_ptr;
  return t;
}
template <typename T>
static T to_tc_full(const EpiOut& e) {
  T t = to_tc<T>(e);
  t.colscale = e.colscale; t.cs_ld = e.cs_ld; t.resid = (const bf16*)e.resid; t.bcast = e.bcast; t.bc_ld = e.bc_ld; t.bc_scale = e.bc_scale;
  return t;
}
template <typename T>
static T to_tcT(const EpiOut& e) {
  T t{}; t.out = (float*)e.out; t.ld = e.ld; t.sb = e.sb; t.alpha = e.alpha; t.alpha_ptr = e.alpha_ptr;
  return t;
}

#define RC(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)
#define KCHECK(name) DGSCT_LAUNCH_CHECK(name)

struct BwdPlan {
  size_t fwd;                        // keep-mode forward workspace
  size_t gUo, gZ, gY, gX1, S1, S2, gV, gU, gQ, gP, dTp, dTpB;
  size_t zero_begin, dm, dsig, ds, dtau, daq, dobar, dr1, colA, colB, colA1, colB1, dWup_d, dWd_d, zero_end;
  size_t vA, vB, vC, vD, wA, wB, wC, vAb, vBb, wcatb, coefA, coefB, coefC, Wd_f32, Wup_f32, opm, out_fwd, smap_fwd, tmap_fwd, csm_part, fc_t, fc_s;
  size_t total;
};

static BwdPlan make_bwd_plan(const dgsct_desc& d, bool have_saved = false) {
  BwdPlan p{};
  const size_t BT = d.BT, N = d.N, C = d.C, Co = d.Co, K = d.K, Cd = d.C / d.r, H = d.C / 2;
  const size_t es = d.io_dtype == DGSCT_BF16 ? 2 : 4;
  const size_t ldz = pad8((int)Cd);
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  p.fwd = take(have_saved ? 256 : make_fwd_plan(d, true).total);      // recompute region (unused with saved activations)
  p.gUo = take(BT * N * C * es);
  p.gZ = take(BT * N * ldz * es);
  p.gY = take(BT * N * C * es);
  p.gX1 = take(BT * N * C * es);
  p.S1 = take(BT * N * C * es);          // z1 -> dz1 ; later dXc
  p.S2 = take(BT * N * H * es);          // z2 -> dz2
  p.gV = take(BT * N * C * es);
  { size_t gu = BT * N * Co;                           // dU (BT,N,Co), or dO' (BT,No,C) of the re-associated resampler
    if (BT * (size_t)d.No * C > gu) gu = BT * (size_t)d.No * C;
    p.gU = take(gu * es); }
  p.gQ = take(BT * N * (size_t)pad8((int)K) * es);   // dQraw -> dL2 ; later dL1 (bf16 rows padded to 16 bytes)
  p.gP = take(BT * N * K * 4);           // dP fp32
  p.dTpB = take(BT * K * C * 2);
  p.zero_begin = off;
  p.dTp = take(BT * K * C * 4);          // (zeroed: the fused attention backward accumulates into it)
  p.dm = take(BT * C * 4); p.dsig = take(BT * N * 4); p.ds = take(BT * N * 4); p.dtau = take(BT * 4);
  p.daq = take(BT * (C + H) * 4); p.dobar = take(BT * C * 4); p.dr1 = take(BT * C * 4);
  p.colA = take(C * 4); p.colB = take(C * 4); p.colA1 = take(C * 4); p.colB1 = take(C * 4);
  p.dWup_d = take(C * ldz * 4); p.dWd_d = take(Cd * C * 4);
  p.vB = take(BT * C * 4); p.vC = take(BT * C * 4);     // split-K targets of the skinny products (added atomically)
  p.fc_t = take(C * 4);                                 // sum_{b,n} dV b_res (re-associated resampler)
  p.zero_end = off;
  p.fc_s = take(C * 4);
  p.vA = take(BT * C * 4); p.vD = take(BT * C * 4);
  p.wA = take(BT * C * 4); p.wB = take(BT * (C + H) * 4); p.wC = take(BT * C * 4);   // second vector chain (own stream: own scratch)
  p.vAb = take(BT * (size_t)pad8((int)C) * 2); p.vBb = take(BT * (size_t)pad8((int)H) * 2);      // bf16 operand copies of the
  p.wcatb = take(BT * (size_t)pad8((int)(C + H)) * 2);                                           // per-frame gradient vectors
  p.coefA = take(C * 4); p.coefB = take(C * 4); p.coefC = take(C * 4);
  p.Wd_f32 = take(Cd * C * 4); p.Wup_f32 = take(C * ldz * 4);
  p.opm = take(BT * C * 4);
  p.csm_part = take(BT * (size_t)CSM_MAX_CHUNKS * K * 4);
  p.out_fwd = take(BT * N * C * es); p.smap_fwd = take(BT * N * es); p.tmap_fwd = take(BT * 4);
  p.total = off;
  return p;
}

size_t bwd_workspace_bytes(const dgsct_desc& d, bool have_saved) { return make_bwd_plan(d, have_saved).total; }

int adapter_bwd(const dgsct_desc& d_in, const dgsct_params& P, const void* packed_in, const void* x, const void* other,
                const void* dout, const void* dsmap, const float* dtmap, void* dx, void* dother, const dgsct_grads& G,
                char* ws, cudaStream_t st, const char* saved) {
  dgsct_desc d = d_in;
  d.flags |= DGSCT_F_NO_STAT_UPDATE | DGSCT_F_NO_OUTPUT;   // the recompute pass must not touch the running buffers
  const BwdPlan bp = make_bwd_plan(d, saved != nullptr);
  const FwdPlan fp = make_fwd_plan(d, true);
  const PackPlan pp = make_pack_plan(d);
  const int BT = d.BT, N = d.N, C = d.C, Co = d.Co, No = d.No, K = d.K, Cd = d.C / d.r, H = d.C / 2, ldz = pad8(Cd);
  const int bf = d.io_dtype == DGSCT_BF16;
  const bool tcp = bf && !(d.flags & DGSCT_F_FORCE_SIMT);   // forward ran the tensor-core pipeline
  if (bf && !tcp) return DGSCT_ERR_UNSUPPORTED;
  const bool train = d.flags & DGSCT_F_TRAIN, use_bn = d.flags & DGSCT_F_USE_BN;
  const long long R = (long long)BT * N;
  const float invN = 1.f / (float)N;
  // ---- 0. the forward's intermediates: saved by a DGSCT_F_KEEP forward, or recomputed here with the keep plan ----------
  char* fw = saved ? const_cast<char*>(saved) : ws + bp.fwd;      // (read-only when saved, except the packed-weights slot)
  const char* pk = (const char*)packed_in;
  if (saved) {
    if (tcp && !pk) pk = fw + fp.packed;                          // packed inside the keep forward's workspace
  } else if (tcp) {
    if (!pk) { RC(pack_params(d, P, fw + fp.packed, st)); pk = fw + fp.packed; }
    RC(adapter_fwd_tc(d, P, pk, x, other, ws + bp.out_fwd, ws + bp.smap_fwd, (float*)(ws + bp.tmap_fwd), nullptr, fw, st, true));
  } else {
    RC(adapter_fwd_simt(d, P, x, other, ws + bp.out_fwd, ws + bp.smap_fwd, (float*)(ws + bp.tmap_fwd), nullptr, fw, st, true));
  }
  // saved tensors (activation dtype unless noted)
  const void* U = fw + fp.bufA; const void* V = fw + fp.bufV; const void* X1 = fw + fp.bufX1; const void* Y = fw + fp.Ybuf;
  const void* Pm = tcp ? (const void*)(fw + fp.Pb) : (const void*)(fw + fp.bufL);
  const void* Qm = tcp ? (const void*)(fw + fp.Qb) : (const void*)(fw + fp.Qf);
  const int ldq = tcp ? pad8(K) : K;                  // row pitch of P, Q, dQ / dL1 (bf16 rows are padded to 16 bytes)
  const long long sbq = (long long)N * ldq;
  const float* Tp = (const float*)(fw + fp.Tp); const void* TpOp = tcp ? (const void*)(fw + fp.TpB) : (const void*)Tp;
  const float* Z = (const float*)(fw + fp.Z); const float* Uo = (const float*)(fw + fp.Uo32);
  const float* osum = (const float*)(fw + fp.osum); const float* r1sum = (const float*)(fw + fp.r1sum);
  const float* aq1 = tcp ? (const float*)(fw + fp.aq) : (const float*)(fw + fp.aq1);
  const float* aq2 = tcp ? aq1 + C : (const float*)(fw + fp.aq2);
  const int aq_ld = tcp ? C + H : C, aq2_ld = tcp ? C + H : H;
  const float* qvec = (const float*)(fw + fp.qvec); const float* mvec = (const float*)(fw + fp.m);
  const float* ssum = (const float*)(fw + fp.ssum); const float* sig = (const float*)(fw + fp.sig); const float* tau = (const float*)(fw + fp.tau);
  const float* sc1 = (const float*)(fw + fp.bn1_scale); const float* sh1 = (const float*)(fw + fp.bn1_shift);
  const float* sc2 = (const float*)(fw + fp.bn2_scale); const float* sh2 = (const float*)(fw + fp.bn2_shift);
  // h: fp32 (BT,H) in the SIMT path, bf16 (BT, ld_h) in the tensor-core path -> fp32 copy for the small products
  float* hf = (float*)(ws + bp.vD);
  bool hf_ready = !tcp;
  auto ensure_h_f32 = [&]() -> int {      // only the SIMT fallbacks of the channel-gate chain read h as fp32: converted on demand
    if (hf_ready) return DGSCT_OK;
    StageScope _sc(ST_MISC, st); count_launch();
    tap_copy_kernel<<<64, 256, 0, st>>>(fw + fp.h, 1, hf, (size_t)BT * pp.ld_h);   // (BT, ld_h) bf16 -> fp32, same pitch
    KCHECK("h_f32");
    hf_ready = true;
    return DGSCT_OK;
  };
  const float* hsrc = tcp ? hf : (const float*)(fw + fp.h);
  const int h_ld = tcp ? pp.ld_h : H;

  void* gUo = ws + bp.gUo; void* gZ = ws + bp.gZ; void* gY = ws + bp.gY; void* gX1 = ws + bp.gX1;
  void* S1 = ws + bp.S1; void* S2 = ws + bp.S2; void* gV = ws + bp.gV; void* gU = ws + bp.gU; void* gQ = ws + bp.gQ;
  float* gP = (float*)(ws + bp.gP); float* dTp = (float*)(ws + bp.dTp); bf16* dTpB = (bf16*)(ws + bp.dTpB);
  float* dm = (float*)(ws + bp.dm); float* dsig = (float*)(ws + bp.dsig); float* ds = (float*)(ws + bp.ds);
  float* dtau = (float*)(ws + bp.dtau); float* daq = (float*)(ws + bp.daq); float* dobar = (float*)(ws + bp.dobar);
  float* dr1 = (float*)(ws + bp.dr1); float* colA = (float*)(ws + bp.colA); float* colB = (float*)(ws + bp.colB);
  float* colA1 = (float*)(ws + bp.colA1); float* colB1 = (float*)(ws + bp.colB1);       // bn1-backward sums (own zeroed slots)
  float* dWup_d = (float*)(ws + bp.dWup_d); float* dWd_d = (float*)(ws + bp.dWd_d);
  float* vA = (float*)(ws + bp.vA); float* vB = (float*)(ws + bp.vB); float* vC = (float*)(ws + bp.vC);
  float* wA = (float*)(ws + bp.wA); float* wB = (float*)(ws + bp.wB); float* wC = (float*)(ws + bp.wC);
  bf16* vAb = (bf16*)(ws + bp.vAb); bf16* vBb = (bf16*)(ws + bp.vBb); bf16* wcatb = (bf16*)(ws + bp.wcatb);
  const int ldcb = pad8(C), ldhb = pad8(H), ldcat = pad8(C + H);
  // The per-frame-vector products (M or K = BT rows) run on the tcgen05 kernel with bf16 copies of the gradient vectors and
  // the forward's packed weights (round 1: eight launches of the SIMT kernel, ~15 us each); DGSCT_VEC_SIMT=1 keeps the fp32 path.
  const bool vec_tc = tcp && C % 8 == 0 && H % 8 == 0 && getenv("DGSCT_VEC_SIMT") == nullptr;
  // D = A . B^T through the tensor-core kernel, or DGSCT_ERR_UNSUPPORTED (the caller then runs its SIMT twin)
  auto vec_gemm = [&](const Gemm& g, const tc::OperandDesc& A, const tc::OperandDesc& B, int M_, int N_, int K_d, float* out, int ld,
                      int mode) -> int {
    EpiOut e = eo(out, 0, ld, 0, mode);
    if (mode == OUT_ATOMIC) return run_tc_only(g, A, B, M_, N_, K_d, 1, to_tc_full<tc::EpiOut<false, OUT_ATOMIC, false, false, false>>(e));
    return run_tc_only(g, A, B, M_, N_, K_d, 1, to_tc_full<tc::EpiOut<false, OUT_STORE, false, false, false>>(e));
  };
  float* coefA = (float*)(ws + bp.coefA); float* coefB = (float*)(ws + bp.coefB); float* coefC = (float*)(ws + bp.coefC);
  float* Wd_f = (float*)(ws + bp.Wd_f32); float* Wup_f = (float*)(ws + bp.Wup_f32); float* opm = (float*)(ws + bp.opm);
  float* daq1 = daq; float* daq2 = daq + C; const int daq_ld = C + H;

  { StageScope _sc(ST_MISC, st);
    DGSCT_CUDA_CHECK(cudaMemsetAsync(ws + bp.zero_begin, 0, bp.zero_end - bp.zero_begin, st)); }
  // weight operands: bf16 packed copies (tensor-core path) or fp32 masters / densified fp32 (SIMT path)
  const void* Wup_op; const void* Wd_op; const void* Wv2_op; const void* Wv1_op; const void* Wfc_op; const void* Wres_op; const void* Tok_op;
  int ld_up, ld_dn, ld_v, ld_fc, ld_res, ld_tok;
  if (tcp) {
    Wup_op = pk + pp.up; Wd_op = pk + pp.down; Wv2_op = pk + pp.v2; Wv1_op = pk + pp.v1; Wfc_op = pk + pp.fc; Wres_op = pk + pp.res;
    Tok_op = pk + pp.tok; ld_up = pp.ld_up; ld_dn = pp.ld_c; ld_v = pp.ld_c; ld_fc = pp.ld_fc; ld_res = pp.ld_res; ld_tok = pp.ld_c;
  } else {
    StageScope _sc(ST_MISC, st); count_launch(2);
    densify_grouped_kernel<<<ceil_div(Cd * C, 256), 256, 0, st>>>(P.down_w, Cd, C, d.g, Wd_f, C);
    densify_grouped_kernel<<<ceil_div(C * ldz, 256), 256, 0, st>>>(P.up_w, C, Cd, d.g, Wup_f, ldz);
    KCHECK("densify");
    Wup_op = Wup_f; Wd_op = Wd_f; Wv2_op = P.v2_w; Wv1_op = P.v1_w; Wfc_op = P.fc_w; Wres_op = P.resample_w; Tok_op = P.tokens;
    ld_up = ldz; ld_dn = C; ld_v = C; ld_fc = Co; ld_res = No; ld_tok = C;
  }
  const int wbf = tcp ? 1 : 0;
  const Gemm GB{ST_MISC, "bwd_gemm", st, tcp};
  const Gemm G_UD{ST_BWD_UPDOWN, "bwd_updown", st, tcp}, G_SP{ST_BWD_SPATIAL, "bwd_spatial", st, tcp}, G_CH{ST_BWD_CHANNEL, "bwd_channel", st, tcp},
      G_TK{ST_BWD_TOKEN, "bwd_token", st, tcp}, G_FC{ST_BWD_FC, "bwd_fc", st, tcp}, G_DW{ST_BWD_DWRES, "bwd_dwres", st, tcp},
      G_DO{ST_BWD_DOTHER, "bwd_dother", st, tcp};
  (void)GB;
  // weight-gradient products go to the side stream (same stream when unavailable, profiling, or DGSCT_NO_SIDE is set)
  // (tensor-core path only: the exact-fp32 path rebuilds relu(bn1(Z)) in a buffer that is reused later)
  SideStream* side = (!tcp || prof_enabled() || getenv("DGSCT_NO_SIDE")) ? nullptr : side_stream_for(st);
  const cudaStream_t sst = side ? side->s : st;
  bool side_used = false;
  auto fork_side = [&]() -> int {        // the side stream waits for everything enqueued on `st` so far
    if (!side) return DGSCT_OK;
    DGSCT_CUDA_CHECK(cudaEventRecord(side->ev, st));
    DGSCT_CUDA_CHECK(cudaStreamWaitEvent(side->s, side->ev, 0));
    side_used = true;
    return DGSCT_OK;
  };
  auto join_side = [&]() -> int {        // `st` waits for the side stream
    if (!side || !side_used) return DGSCT_OK;
    DGSCT_CUDA_CHECK(cudaEventRecord(side->ev, side->s));
    DGSCT_CUDA_CHECK(cudaStreamWaitEvent(st, side->ev, 0));
    side_used = false;
    return DGSCT_OK;
  };
  // (stage "bwd_dw" = the weight-gradient products sum_b A_b^T B_b of the small matrices; "bwd_dwres" = the resampler's)
  SideStream* side2 = side ? side_stream_for(st, 1) : nullptr;
  const cudaStream_t s2 = side2 ? side2->s : st;
  bool side2_used = false;
  auto fork_side2 = [&]() -> int {
    if (!side2) return DGSCT_OK;
    DGSCT_CUDA_CHECK(cudaEventRecord(side2->ev, st));
    DGSCT_CUDA_CHECK(cudaStreamWaitEvent(side2->s, side2->ev, 0));
    side2_used = true;
    return DGSCT_OK;
  };
  auto join_side2 = [&]() -> int {
    if (!side2 || !side2_used) return DGSCT_OK;
    DGSCT_CUDA_CHECK(cudaEventRecord(side2->ev, side2->s));
    DGSCT_CUDA_CHECK(cudaStreamWaitEvent(st, side2->ev, 0));
    side2_used = false;
    return DGSCT_OK;
  };
  const Gemm S_UD{ST_BWD_DW, "bwd_updown_dw", sst, tcp}, S_SP{ST_BWD_DW, "bwd_spatial_dw", sst, tcp},
      S_CH{ST_BWD_DW, "bwd_channel_dw", sst, tcp}, S_TK{ST_BWD_DW, "bwd_token_dw", sst, tcp},
      S_FC{ST_BWD_DW, "bwd_fc_dw", sst, tcp}, S_DW{ST_BWD_DWRES, "bwd_dwres", sst, tcp};
  (void)G_DW;
  const int rgrid = (int)((R + 7) / 8 < 148 * 4 ? (R + 7) / 8 : 148 * 4);
  const int ew_grid = 148 * 8;

  // ---- 1. output stage: gate, ln_post, bn2 ----------------------------------------------------------------------
  bool bn2_sums_done = false;
  {
    BwdOutArgs a{};
    a.dout = dout; a.dout_bf16 = bf; a.uo = Uo; a.scale = sc2; a.shift = sh2;
    const bool lnp = d.flags & DGSCT_F_LN_POST;
    a.ln_w = lnp ? P.lnp_w : nullptr; a.ln_b = lnp ? P.lnp_b : nullptr;
    a.gate = (d.flags & DGSCT_F_USE_GATE) ? P.gate : nullptr;
    a.gate_before_ln = (d.flags & DGSCT_F_GATE_BEFORE_LN) ? 1 : 0;
    a.eps = d.eps_ln; a.C = C; a.rows = R; a.db2 = gUo; a.db2_bf16 = bf;
    a.d_lnw = lnp ? G.lnp_w : nullptr; a.d_lnb = lnp ? G.lnp_b : nullptr; a.d_gate = a.gate ? G.gate : nullptr;
    const int wpb = C <= 768 ? 8 : 4;
    static PerDeviceOnce attr1;
    if (attr1.first()) DGSCT_CUDA_CHECK(cudaFuncSetAttribute(bwd_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    StageScope _sc(ST_BWD_OUT, st); count_launch();
    { cudaError_t pre = cudaGetLastError(); if (pre != cudaSuccess) { set_cuda_error(pre, "before bwd_out"); return DGSCT_ERR_CUDA; } }
    // vectorised kernel: also produces the bn2-backward column sums (colA = sum dB2, colB = sum dB2*Uo)
    bn2_sums_done = launch_bwd_out_vec(a, use_bn ? colA : nullptr, use_bn ? colB : nullptr, st);
    if (!bn2_sums_done) bwd_out_kernel<<<rgrid, wpb * 32, (size_t)(2 + 2 * wpb) * C * 4, st>>>(a);
    KCHECK("bwd_out");
  }
  // bn2 backward: dUo = A*dB2 + B*Uo + Cc
  if (use_bn) {
    StageScope _sc(ST_BN_FINALIZE, st); count_launch(3);
    if (!bn2_sums_done) col_reduce_kernel<<<dim3(ceil_div(C, 32), 64), 256, 0, st>>>(gUo, bf, C, Uo, 0, C, R, C, colA, colB);
    // coefficients + combine in one pass when the vector kernel applies (P.bn2_w may sit at any 4-byte offset: checked inside)
    if (!launch_bn_bwd_apply(gUo, bf, C, Uo, C, colA, colB, (const float*)(fw + fp.bn2_mean), (const float*)(fw + fp.bn2_rstd), P.bn2_w,
                             (double)R, train, G.bn2_w, G.bn2_b, R, C, st)) {
      bn_bwd_coeffs_kernel<<<ceil_div(C, 128), 128, 0, st>>>(C, 1, train, colA, colB, (const float*)(fw + fp.bn2_mean),
                                                             (const float*)(fw + fp.bn2_rstd), P.bn2_w, (double)R, coefA, coefB, coefC,
                                                             G.bn2_w, G.bn2_b);
      launch_affine_combine(gUo, bf, C, Uo, C, coefA, coefB, coefC, R, C, ew_grid, st);
    }
    KCHECK("bn2_bwd");
  }
  // ---- 2. up projection: dZr = dUo . Wup ; dWup += dUo^T . Zr ---------------------------------------------------
  bool relu_bn1_sums_done = false;
  {
    // Zr = relu(Z*sc1+sh1) as a GEMM operand: the tensor-core forward kept it (bf16, pitch ldz); the fp32 path
    // rebuilds it into S2 (free until step 5)
    const void* Zr = fw + fp.Zr;
    if (!tcp) {
      StageScope _sc(ST_MISC, st); count_launch();
      bn_relu_f32_kernel<<<ew_grid, 256, 0, st>>>(Z, sc1, sh1, (float*)S2, R, Cd, ldz);
      KCHECK("zr_f32");
      Zr = S2;
    }
    // tensor-core path: ReLU mask of bn1 and the bn1-backward column sums are this product's epilogue
    int rcz = DGSCT_ERR_UNSUPPORTED;
    if (tcp && getenv("DGSCT_NO_FUSED_BWD_EPI") == nullptr) {
      tc::EpiReluMaskStats te{};
      te.out = (bf16*)gZ; te.ld = ldz; te.sb = (long long)N * ldz; te.z = Z; te.zld = Cd; te.zsb = (long long)N * Cd; te.sc = sc1; te.sh = sh1;
      te.colA = use_bn ? colA1 : nullptr; te.colB = use_bn ? colB1 : nullptr;
      rcz = run_tc_only_stats(G_UD, K_(gUo, N, C, C, (long long)N * C), MN_(Wup_op, C, Cd, ld_up, 0), N, Cd, C, BT, te);
      if (rcz != DGSCT_OK && rcz != DGSCT_ERR_UNSUPPORTED) return rcz;
      relu_bn1_sums_done = (rcz == DGSCT_OK);
    }
    if (rcz == DGSCT_ERR_UNSUPPORTED) {
      EpiOut e1 = eo(gZ, bf, ldz, (long long)N * ldz, OUT_STORE);
      RC(run_gemm(G_UD, K_(gUo, N, C, C, (long long)N * C), bf, MN_(Wup_op, C, Cd, ld_up, 0), wbf, N, Cd, C, BT,
                  to_tc_full<tc::EpiOut<true, OUT_STORE, false, false, false>>(e1), e1, AxIdentity{}, true));
    }
    RC(fork_side());                                                  // gUo is final; Zr is a saved/recomputed activation
    EpiOut e2 = eo(dWup_d, 0, ldz, 0, OUT_ATOMIC);
    RC(run_gemm(S_UD, MN_(gUo, N, C, C, (long long)N * C), bf, MN_(Zr, N, Cd, ldz, (long long)N * ldz), bf, C, Cd, N, BT,
                to_tc_full<tc::EpiOut<false, OUT_ATOMIC, false, false, false>>(e2), e2, AxIdentity{}, true, 0, true));
    { StageScope _sc(ST_MISC, sst); count_launch();
      ungroup_grad_kernel<<<ceil_div(C * (Cd / d.g), 256), 256, 0, sst>>>(dWup_d, C, Cd, d.g, ldz, G.up_w);
      KCHECK("ungroup_up"); }
  }
  // relu + bn1 backward on dZr -> dZ (in place in gZ)
  {
    StageScope _sc(ST_BN_FINALIZE, st); count_launch(relu_bn1_sums_done ? 1 : 4);
    if (!relu_bn1_sums_done) relu_mask_kernel<<<ew_grid, 256, 0, st>>>(gZ, bf, ldz, Z, Cd, sc1, sh1, R, Cd);
    if (use_bn) {
      if (!relu_bn1_sums_done)
      { const int gx = ceil_div(Cd, 32);                      // narrow (C/r columns): spread the rows over ~4 blocks per SM
        long long gy = (R + 255) / 256;
        if (gy > (148 * 4) / gx) gy = (148 * 4) / gx;
        if (gy < 1) gy = 1;
        col_reduce_kernel<<<dim3(gx, (unsigned)gy), 256, 0, st>>>(gZ, bf, ldz, Z, 0, Cd, R, Cd, colA1, colB1); }
      if (!launch_bn_bwd_apply(gZ, bf, ldz, Z, Cd, colA1, colB1, (const float*)(fw + fp.bn1_mean), (const float*)(fw + fp.bn1_rstd), P.bn1_w,
                               (double)R, train, G.bn1_w, G.bn1_b, R, Cd, st)) {
        bn_bwd_coeffs_kernel<<<ceil_div(Cd, 128), 128, 0, st>>>(Cd, 1, train, colA1, colB1, (const float*)(fw + fp.bn1_mean),
                                                                (const float*)(fw + fp.bn1_rstd), P.bn1_w, (double)R, coefA, coefB, coefC,
                                                                G.bn1_w, G.bn1_b);
        launch_affine_combine(gZ, bf, ldz, Z, Cd, coefA, coefB, coefC, R, Cd, ew_grid, st);
      }
    }
    KCHECK("bn1_bwd");
  }
  // ---- 3. down projection: dY = dZ . Wd ; dWd += dZ^T . Y --------------------------------------------------------
  {
    EpiOut e1 = eo(gY, bf, C, (long long)N * C, OUT_STORE);
    RC(run_gemm(G_UD, K_(gZ, N, Cd, ldz, (long long)N * ldz), bf, MN_(Wd_op, Cd, C, ld_dn, 0), wbf, N, C, Cd, BT,
                to_tc_full<tc::EpiOut<true, OUT_STORE, false, false, false>>(e1), e1, AxIdentity{}, true));
    RC(fork_side());                                                  // gZ is final (bn1 backward applied)
    EpiOut e2 = eo(dWd_d, 0, C, 0, OUT_ATOMIC); e2.transposed = 1;      // D(c, cd) -> dWd_d[cd, c]
    RC(run_gemm(S_UD, MN_(Y, N, C, C, (long long)N * C), bf, MN_(gZ, N, Cd, ldz, (long long)N * ldz), bf, C, Cd, N, BT,
                to_tcT<tc::EpiOutT<OUT_ATOMIC>>(e2), e2, AxIdentity{}, true, 0, true));
    { StageScope _sc(ST_MISC, sst); count_launch();
      ungroup_grad_kernel<<<ceil_div(Cd * (C / d.g), 256), 256, 0, sst>>>(dWd_d, Cd, C, d.g, C, G.down_w);
      KCHECK("ungroup_down"); }
  }
  // ---- 4. ln_before + gate combine ------------------------------------------------------------------------------
  const bool use_tau = (d.flags & DGSCT_F_TEMPORAL) && d.gamma != 0.f;
  {
    BwdGateLnArgs a{};
    a.dy = gY; a.dy_bf16 = bf; a.x1 = X1; a.x1_bf16 = bf; a.m = mvec; a.sig = sig; a.tau = use_tau ? tau : nullptr;
    const bool lnb = d.flags & DGSCT_F_LN_BEFORE;
    a.ln_w = lnb ? P.lnb_w : nullptr;
    a.alpha = d.alpha; a.beta = d.beta; a.gamma = d.gamma; a.eps = d.eps_ln; a.N = N; a.C = C; a.rows = R;
    a.dx1 = gX1; a.dx1_bf16 = bf; a.dm_acc = dm; a.dsig = dsig; a.dtau = use_tau ? dtau : nullptr;
    a.d_lnw = lnb ? G.lnb_w : nullptr; a.d_lnb = lnb ? G.lnb_b : nullptr;
    const int wpb = C <= 768 ? 8 : 4;
    static PerDeviceOnce attr2;
    if (attr2.first()) DGSCT_CUDA_CHECK(cudaFuncSetAttribute(bwd_gate_ln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    StageScope _sc(ST_BWD_GATE_LN, st); count_launch();
    if (!launch_bwd_gate_ln_vec(a, BT, st))
      bwd_gate_ln_kernel<<<rgrid, wpb * 32, (size_t)(2 + 2 * wpb) * C * 4, st>>>(a);
    KCHECK("bwd_gate_ln");
  }
  // ---- 5. spatial gate --------------------------------------------------------------------------------------------
  {
    { StageScope _sc(ST_SPATIAL_MAP, st); count_launch();
      bwd_spatial_map_kernel<<<BT, 256, 0, st>>>(ssum, P.s_b, N, dsmap, bf, dsig, ds, G.s_b);
      KCHECK("bwd_spatial_map"); }
    // z2 = Xc . Wv2^T (pre-bias) -> S2
    EpiOut e1 = eo(S2, bf, H, (long long)N * H, OUT_STORE);
    if (tcp) {
      RC(run_gemm(G_SP, K_(fw + fp.XcBuf, N, C, C, (long long)N * C), 1, K_(Wv2_op, H, C, ld_v, 0), 1, N, H, C, BT,
                  to_tc_full<tc::EpiOut<true, OUT_STORE, false, false, false>>(e1), e1, AxIdentity{}, true));
    } else {
      RC(run_gemm(G_SP, K_(X1, N, C, C, (long long)N * C), 0, K_(Wv2_op, H, C, ld_v, 0), 0, N, H, C, BT,
                  to_tc_full<tc::EpiOut<true, OUT_STORE, false, false, false>>(e1), e1, (AxOnePlusCol{mvec, C}), false));
    }
    { StageScope _sc(ST_SPATIAL_QUERY, st); count_launch();
      bwd_spatial_mask_kernel<<<dim3(ceil_div(H, 32), BT, N >= 512 ? 8 : 1), 256, 0, st>>>(S2, bf, H, N, P.v2_b, ds, P.s_w, aq2, aq2_ld,
                                                                                        G.s_w, daq2, daq_ld, G.v2_b);
      KCHECK("bwd_spatial_mask"); }
    // dXc = dz2 . Wv2 -> S1 ; dX1 += dXc (1+m) ; dm += sum_n dXc X1
    {
      EpiOut e2 = eo(S1, bf, C, (long long)N * C, OUT_STORE);
      RC(run_gemm(G_SP, K_(S2, N, H, H, (long long)N * H), bf, MN_(Wv2_op, H, C, ld_v, 0), wbf, N, C, H, BT,
                  to_tc_full<tc::EpiOut<true, OUT_STORE, false, false, false>>(e2), e2, AxIdentity{}, true));
      StageScope _sc(ST_SPATIAL_QUERY, st); count_launch();
      bwd_xc_kernel<<<dim3(ceil_div(C, 32), BT, N >= 512 ? 8 : 1), 256, 0, st>>>(S1, gX1, X1, bf, C, N, mvec, dm);
      KCHECK("bwd_xc");
    }
    // opm = 1 + m   (operand scale of the exact-fp32 path's dWv2 product; the tensor-core path kept Xc = X1 (1+m))
    if (!tcp) { StageScope _sc(ST_MISC, st); count_launch();
      one_plus_kernel<<<ceil_div(BT * C, 256), 256, 0, st>>>(mvec, opm, BT * C);
      KCHECK("opm"); }
    // dWv2[j,c] += sum_b (1+m[b,c]) sum_n dz2[b,n,j] X1[b,n,c]
    if (tcp) {      // Xc = X1 (1+m) was kept by the forward: plain sum over frames
      RC(fork_side());                                                // S2 = dz2 is not written again in this call
      EpiOut e3 = eo(G.v2_w, 0, C, 0, OUT_ATOMIC);
      RC(run_gemm(S_SP, MN_(S2, N, H, H, (long long)N * H), 1, MN_(fw + fp.XcBuf, N, C, C, (long long)N * C), 1, H, C, N, BT,
                  to_tc_full<tc::EpiOut<false, OUT_ATOMIC, false, false, false>>(e3), e3, AxIdentity{}, true, 0, true));
    } else {
      EpiOut e3 = eo(G.v2_w, 0, C, 0, OUT_ATOMIC); e3.colscale = opm; e3.cs_ld = C;
      RC(run_gemm(G_SP, MN_(S2, N, H, H, (long long)N * H), bf, MN_(X1, N, C, C, (long long)N * C), bf, H, C, N, BT,
                  to_tc_full<tc::EpiOut<false, OUT_ATOMIC, false, false, true>>(e3), e3, AxIdentity{}, true));
    }
  }
  // ---- 6. channel gate ---------------------------------------------------------------------------------------------
  {
    // dpm = dm * m (1-m) -> vA (BT,C)
    // (the parameter gradients of this chain -- two outer products over the BT rows and two bias sums -- go to the side stream)
    const Gemm GS{ST_BWD_SMALL, "bwd_small", st, false}, GSW{ST_BWD_SMALL, "bwd_small_dw", sst, false};
    const bf16* hb = (const bf16*)(fw + fp.h);                 // forward's h (BT, ld_h) bf16 and q (BT, pad8(C)) bf16
    const bf16* qb = (const bf16*)(fw + fp.vec_bf);
    count_launch();
    vec_op_kernel<<<ceil_div(BT * C, 256), 256, 0, st>>>(VOP_SIGMOID_BWD, dm, C, mvec, C, 0.f, vA, C, BT, C, vec_tc ? vAb : nullptr, ldcb);
    KCHECK("dpm");
    // d_cb += colsum dpm ; dWc (C,H) += dpm^T h ; dh = dpm . Wc -> vB (BT,H)
    RC(fork_side());                                                  // vA = dpm is final
    col_reduce_kernel<<<dim3(ceil_div(C, 32), 1), 256, 0, sst>>>(vA, 0, C, nullptr, 0, 0, BT, C, G.c_b, nullptr);
    int rcv = vec_tc ? vec_gemm(GSW, MN_(vAb, BT, C, ldcb, 0), MN_(hb, BT, H, pp.ld_h, 0), C, H, BT, G.c_w, H, OUT_ATOMIC) : DGSCT_ERR_UNSUPPORTED;
    if (rcv == DGSCT_ERR_UNSUPPORTED) {
      RC(ensure_h_f32());
      if (side) { RC(fork_side()); }                                  // (the conversion ran on the caller's stream)
      EpiOut e1 = eo(G.c_w, 0, H, 0, OUT_ADD);
      RC(run_gemm(GSW, MN_(vA, BT, C, C, 0), 0, MN_(hsrc, BT, H, h_ld, 0), 0, C, H, BT, 1, tc::EpiNone{}, e1, AxIdentity{}, true));
    } else if (rcv) return rcv;
    rcv = vec_tc ? vec_gemm(GS, K_(vAb, BT, C, ldcb, 0), MN_(pk + pp.catt, C, H, pp.ld_h, 0), BT, H, C, vB, H, OUT_STORE) : DGSCT_ERR_UNSUPPORTED;
    if (rcv == DGSCT_ERR_UNSUPPORTED) {
      EpiOut e2 = eo(vB, 0, H, 0, OUT_STORE);
      RC(run_gemm(GS, K_(vA, BT, C, C, 0), 0, MN_(P.c_w, C, H, H, 0), 0, BT, H, C, 1, tc::EpiNone{}, e2, AxIdentity{}, true, 0, false, true));
    } else if (rcv) return rcv;
    // dph = dh * 1[h>0] -> vB in place ; d_bb += colsum ; dWb (H,C) += dph^T q ; dq = dph . Wb -> vC (BT,C)
    count_launch(2);
    if (!tcp || !vec_tc) { RC(ensure_h_f32()); }
    vec_op_kernel<<<ceil_div(BT * H, 256), 256, 0, st>>>(VOP_RELU_BWD, vB, H, hsrc, h_ld, 0.f, vB, H, BT, H, vec_tc ? vBb : nullptr, ldhb,
                                                         (tcp && vec_tc) ? hb : nullptr);
    RC(fork_side());                                                  // vB = dph is final
    col_reduce_kernel<<<dim3(ceil_div(H, 32), 1), 256, 0, sst>>>(vB, 0, H, nullptr, 0, 0, BT, H, G.bott_b, nullptr);
    KCHECK("dph");
    rcv = vec_tc ? vec_gemm(GSW, MN_(vBb, BT, H, ldhb, 0), MN_(qb, BT, C, pad8(C), 0), H, C, BT, G.bott_w, C, OUT_ATOMIC) : DGSCT_ERR_UNSUPPORTED;
    if (rcv == DGSCT_ERR_UNSUPPORTED) {
      EpiOut e3 = eo(G.bott_w, 0, C, 0, OUT_ADD);
      RC(run_gemm(GSW, MN_(vB, BT, H, H, 0), 0, MN_(qvec, BT, C, C, 0), 0, H, C, BT, 1, tc::EpiNone{}, e3, AxIdentity{}, true));
    } else if (rcv) return rcv;
    rcv = vec_tc ? vec_gemm(GS, K_(vBb, BT, H, ldhb, 0), MN_(pk + pp.bott, H, C, pp.ld_c, 0), BT, C, H, vC, C, OUT_STORE) : DGSCT_ERR_UNSUPPORTED;
    if (rcv == DGSCT_ERR_UNSUPPORTED) {
      EpiOut e4 = eo(vC, 0, C, 0, OUT_STORE);
      RC(run_gemm(GS, K_(vB, BT, H, H, 0), 0, MN_(P.bott_w, H, C, C, 0), 0, BT, C, H, 1, tc::EpiNone{}, e4, AxIdentity{}, true, 0, false, true));
    } else if (rcv) return rcv;
    // q = aq1 * r1sum / N : daq1 = dq * r1sum / N ; dr1 = dq * aq1 / N
    count_launch();
    dq_split_kernel<<<ceil_div(BT * C, 256), 256, 0, st>>>(vC, r1sum, aq1, aq_ld, invN, daq1, daq_ld, dr1, BT, C);
    KCHECK("dq_split");
  }
  {
    // z1 = X1 . Wv1^T (pre-bias) -> S1 ; dz1 = dr1 * 1[z1+b>0] ; dX1 += dz1 . Wv1 ; dWv1 += dz1^T X1
    {
      // tensor-core path: the ReLU mask, dr1 broadcast and d_b1 column sums are the GEMM's epilogue (no pass over (BT,N,C))
      int rcm = DGSCT_ERR_UNSUPPORTED;
      if (tcp && getenv("DGSCT_NO_FUSED_BWD_EPI") == nullptr) {
        tc::EpiMaskBcast te{};
        te.out = (bf16*)S1; te.ld = C; te.sb = (long long)N * C; te.bias = P.v1_b; te.bc = dr1; te.bc_ld = C; te.dbias = G.v1_b;
        rcm = run_tc_only_stats(G_CH, K_(X1, N, C, C, (long long)N * C), K_(Wv1_op, C, C, ld_v, 0), N, C, C, BT, te);
        if (rcm != DGSCT_OK && rcm != DGSCT_ERR_UNSUPPORTED) return rcm;
      }
      if (rcm == DGSCT_ERR_UNSUPPORTED) {
        EpiOut e1 = eo(S1, bf, C, (long long)N * C, OUT_STORE);
        RC(run_gemm(G_CH, K_(X1, N, C, C, (long long)N * C), bf, K_(Wv1_op, C, C, ld_v, 0), wbf, N, C, C, BT,
                    to_tc_full<tc::EpiOut<true, OUT_STORE, false, false, false>>(e1), e1, AxIdentity{}, true));
        StageScope _sc(ST_CHANNEL_QUERY, st); count_launch();
        bwd_channel_mask_kernel<<<dim3(ceil_div(C, 32), BT, N >= 512 ? 8 : 1), 256, 0, st>>>(S1, bf, C, N, P.v1_b, dr1, G.v1_b);
        KCHECK("bwd_channel_mask");
      }
    }
    EpiOut e2 = eo(gX1, bf, C, (long long)N * C, OUT_ADD);
    RC(run_gemm(G_CH, K_(S1, N, C, C, (long long)N * C), bf, MN_(Wv1_op, C, C, ld_v, 0), wbf, N, C, C, BT,
                to_tc_full<tc::EpiOut<true, OUT_ADD, false, false, false>>(e2), e2, AxIdentity{}, true));
    RC(fork_side());                                                  // S1 = dz1 is not written again in this call
    EpiOut e3 = eo(G.v1_w, 0, C, 0, OUT_ATOMIC);
    RC(run_gemm(S_CH, MN_(S1, N, C, C, (long long)N * C), bf, MN_(X1, N, C, C, (long long)N * C), bf, C, C, N, BT,
                to_tc_full<tc::EpiOut<false, OUT_ATOMIC, false, false, false>>(e3), e3, AxIdentity{}, true, 0, true));
  }
  {
    // aq1 = relu(Wa1 obar + b), aq2 = relu(Wa2 obar + b), tau = sigmoid(wt obar + bt)   [obar = osum / N]
    // Runs on the second auxiliary stream with scratch of its own: d obar is first read by sum_frames_kernel / the dV products.
    RC(fork_side2());                                                 // daq1 / daq2 / dtau are final
    const Gemm GS{ST_BWD_SMALL, "bwd_small", s2, false};
    const bf16* obar_b = (const bf16*)(fw + fp.vec_bf_o);         // forward's bf16 obar (BT, pad8(C))
    // [dpa1 | dpa2] = [daq1 | daq2] * 1[[aq1 | aq2] > 0] in ONE pass over the concatenated (BT, C+H) vectors -> wB (+ bf16 copy)
    count_launch((d.flags & DGSCT_F_TEMPORAL) ? 2 : 1);
    if (d.flags & DGSCT_F_TEMPORAL)
      cast_scale_kernel<<<ceil_div(BT * C, 256), 256, 0, s2>>>(osum, invN, wA, 0, (size_t)BT * C);                   // obar -> wA (temporal gate)
    bool cat_ok = vec_tc && aq_ld == C + H && aq2_ld == C + H && daq_ld == C + H;
    if (cat_ok) {
      vec_op_kernel<<<ceil_div(BT * (C + H), 256), 256, 0, s2>>>(VOP_RELU_BWD, daq1, daq_ld, aq1, aq_ld, 0.f, wB, C + H, BT, C + H, wcatb, ldcat);
      KCHECK("dpa");
      count_launch(2);
      col_reduce_kernel<<<dim3(ceil_div(C, 32), 1), 256, 0, s2>>>(wB, 0, C + H, nullptr, 0, 0, BT, C, G.a1_b, nullptr);
      col_reduce_kernel<<<dim3(ceil_div(H, 32), 1), 256, 0, s2>>>(wB + C, 0, C + H, nullptr, 0, 0, BT, H, G.a2_b, nullptr);
      int r1 = vec_gemm(GS, MN_(wcatb, BT, C, ldcat, 0), MN_(obar_b, BT, C, pad8(C), 0), C, C, BT, G.a1_w, C, OUT_ATOMIC);
      int r2 = r1 ? r1 : vec_gemm(GS, MN_(wcatb + C, BT, H, ldcat, 0), MN_(obar_b, BT, C, pad8(C), 0), H, C, BT, G.a2_w, C, OUT_ATOMIC);
      // d obar = [dpa1 | dpa2] . [Wa1 ; Wa2]  : one product over K = C + H (the forward's packed [a1_w ; a2_w])
      int r3 = r2 ? r2 : vec_gemm(GS, K_(wcatb, BT, C + H, ldcat, 0), MN_(pk + pp.cat1, C + H, C, pp.ld_c, 0), BT, C, C + H, dobar, C, OUT_ATOMIC);
      if (r3 == DGSCT_ERR_UNSUPPORTED && r1 == DGSCT_ERR_UNSUPPORTED) cat_ok = false;       // nothing was launched: SIMT chain below
      else if (r3) return r3;
    }
    if (!cat_ok) {
      if (!(d.flags & DGSCT_F_TEMPORAL)) { count_launch(); cast_scale_kernel<<<ceil_div(BT * C, 256), 256, 0, s2>>>(osum, invN, wA, 0, (size_t)BT * C); }
      count_launch(2);
      vec_op_kernel<<<ceil_div(BT * C, 256), 256, 0, s2>>>(VOP_RELU_BWD, daq1, daq_ld, aq1, aq_ld, 0.f, wB, C, BT, C);  // dpa1 -> wB
      vec_op_kernel<<<ceil_div(BT * H, 256), 256, 0, s2>>>(VOP_RELU_BWD, daq2, daq_ld, aq2, aq2_ld, 0.f, wC, H, BT, H); // dpa2 -> wC
      KCHECK("dpa");
      count_launch(2);
      col_reduce_kernel<<<dim3(ceil_div(C, 32), 1), 256, 0, s2>>>(wB, 0, C, nullptr, 0, 0, BT, C, G.a1_b, nullptr);
      col_reduce_kernel<<<dim3(ceil_div(H, 32), 1), 256, 0, s2>>>(wC, 0, H, nullptr, 0, 0, BT, H, G.a2_b, nullptr);
      EpiOut e1 = eo(G.a1_w, 0, C, 0, OUT_ADD);
      RC(run_gemm(GS, MN_(wB, BT, C, C, 0), 0, MN_(wA, BT, C, C, 0), 0, C, C, BT, 1, tc::EpiNone{}, e1, AxIdentity{}, true));
      EpiOut e2 = eo(G.a2_w, 0, C, 0, OUT_ADD);
      RC(run_gemm(GS, MN_(wC, BT, H, H, 0), 0, MN_(wA, BT, C, C, 0), 0, H, C, BT, 1, tc::EpiNone{}, e2, AxIdentity{}, true));
      EpiOut e3 = eo(dobar, 0, C, 0, OUT_ADD);
      RC(run_gemm(GS, K_(wB, BT, C, C, 0), 0, MN_(P.a1_w, C, C, C, 0), 0, BT, C, C, 1, tc::EpiNone{}, e3, AxIdentity{}, true, 0, false, true));
      RC(run_gemm(GS, K_(wC, BT, H, H, 0), 0, MN_(P.a2_w, H, C, C, 0), 0, BT, C, H, 1, tc::EpiNone{}, e3, AxIdentity{}, true, 0, false, true));
    }
    if (d.flags & DGSCT_F_TEMPORAL) {
      count_launch();
      temporal_bwd_kernel<<<BT, 128, 0, s2>>>(use_tau ? dtau : nullptr, dtmap, tau, wA, P.t_w, C, dobar, G.t_w, G.t_b);
      KCHECK("temporal_bwd");
    }
  }
  // ---- 7. latent-token attention -----------------------------------------------------------------------------------
  {
    // Fused kernels (attn_fused.cuh) when the shape qualifies: x_attn_bwd replaces five launches (dQraw, row-softmax backward,
    // two dT' products, dx), tok_attn_bwd six (dP, the split column-softmax backward, two dV products, the d-token product).
    const fa::XAttnBwdPlan fb_b = fa::make_x_attn_bwd_plan(BT, N, C, K);
    const bool fused_b = tcp && fb_b.ok && getenv("DGSCT_NO_FUSED_ATTN_BWD") == nullptr && getenv("DGSCT_NO_FUSED_ATTN") == nullptr &&
                         (((uintptr_t)x | (uintptr_t)dx) & 15) == 0;
    const fa::TokAttnBwdPlan fb_a = fa::make_tok_attn_bwd_plan(BT, N, C, K);
    const bool fused_a = tcp && fb_a.ok && fused_attn_bwd_a_enabled(d, x, pp);
    if (fused_b) {
      fa::XAttnBwdArgs a{};
      a.G = (const bf16*)gX1; a.x = (const bf16*)x; a.TpB = (const bf16*)TpOp; a.Qb = (const bf16*)Qm;
      a.dx = (bf16*)dx; a.dTp = dTp; a.gate_av = P.gate_av; a.d_gate_av = G.gate_av; a.err_flag = timeout_flag_device_ptr();
      a.N = N; a.C = C; a.K = K; a.Kp = ldq; a.R = fb_b.R; a.nchunks = fb_b.nchunks; a.splits = fb_b.splits; a.zcols = fb_b.zcols; a.nbuf = fb_b.nbuf;
      StageScope _sc(ST_BWD_TOKEN, st); count_launch();
      cudaError_t e = fa::launch_x_attn_bwd(a, fb_b, BT, st);
      if (e != cudaSuccess) { set_cuda_error(e, "x_attn_bwd_fused"); return DGSCT_ERR_CUDA; }
    } else {
      // dQraw = G . Tp^T -> gQ ; dL2 (in place) ; d_gate_av
      EpiOut e1 = eo(gQ, bf, ldq, sbq, OUT_STORE);
      RC(run_gemm(G_TK, K_(gX1, N, C, C, (long long)N * C), bf, K_(TpOp, K, C, C, (long long)K * C), wbf, N, K, C, BT,
                  to_tc_full<tc::EpiOut<true, OUT_STORE, false, false, false>>(e1), e1, AxIdentity{}, true));
      { StageScope _sc(ST_ROW_SOFTMAX, st); count_launch();
        row_softmax_bwd_kernel<<<rgrid, 256, 0, st>>>(gQ, Qm, bf, R, K, ldq, P.gate_av, G.gate_av);
        KCHECK("row_softmax_bwd"); }
      // dTp[b,k,c] = g sum_n Q G + sum_n dL2 x     (computed as D(c,k), written transposed)
      EpiOut e2 = eo(dTp, 0, C, (long long)K * C, OUT_STORE); e2.transposed = 1; e2.alpha_ptr = P.gate_av;
      RC(run_gemm(G_TK, MN_(gX1, N, C, C, (long long)N * C), bf, MN_(Qm, N, K, ldq, sbq), bf, C, K, N, BT,
                  to_tcT<tc::EpiOutT<OUT_STORE>>(e2), e2, AxIdentity{}, true));
      EpiOut e3 = eo(dTp, 0, C, (long long)K * C, OUT_ADD); e3.transposed = 1;
      RC(run_gemm(G_TK, MN_(x, N, C, C, (long long)N * C), bf, MN_(gQ, N, K, ldq, sbq), bf, C, K, N, BT,
                  to_tcT<tc::EpiOutT<OUT_ADD>>(e3), e3, AxIdentity{}, true));
      // dx = G + dL2 . Tp
      EpiOut e4 = eo(dx, bf, C, (long long)N * C, OUT_STORE); e4.resid = gX1; e4.resid_bf16 = bf;
      RC(run_gemm(G_TK, K_(gQ, N, K, ldq, sbq), bf, MN_(TpOp, K, C, C, (long long)K * C), wbf, N, C, K, BT,
                  to_tc_full<tc::EpiOut<true, OUT_STORE, true, false, false>>(e4), e4, AxIdentity{}, true));
    }
    // d tokens += sum_b dTp[b]
    RC(join_side2());                                                 // d obar (second vector chain) is needed from here on
    const void* dTpOp = dTp;
    long long dTp_bs = (long long)K * C;
    RC(fork_side());                                                  // dT' and d obar are final: the frame sums are parameter gradients
    { StageScope _sc(ST_MISC, sst); count_launch();
      sum_frames_kernel<<<ceil_div(K * C, 256), 256, 0, sst>>>(dTp, BT, (size_t)K * C, G.tokens, G.fc_b, C, dobar);  // + d fc_b (exact)
      KCHECK("dtok_sum"); }
    if (tcp && !fused_a) {
      StageScope _sc(ST_MISC, st); count_launch();
      cast_scale_kernel<<<ew_grid, 256, 0, st>>>(dTp, 1.f, dTpB, 1, (size_t)BT * K * C); dTpOp = dTpB;
      KCHECK("dtp_cast");
    }
    if (fused_a) {
      float* Dv = (float*)(ws + bp.csm_part);               // D[b,k] = sum_n dP P  (BT x K)
      { StageScope _sc(ST_COL_SOFTMAX, st); count_launch();
        fa::dtp_prep_kernel<<<ceil_div(BT * K * 32, 256), 256, 0, st>>>(dTp, Tp, P.tokens, dTpB, Dv, BT, K, C);
        KCHECK("dtp_prep"); }
      fa::TokAttnBwdArgs a{};
      a.V = (const bf16*)V; a.tokb = (const bf16*)Tok_op; a.dTpB = dTpB; a.stat_ml = (const float*)(fw + fp.fa_stat); a.D = Dv;
      a.dobar = dobar; a.gV = (bf16*)gV; a.d_tok = G.tokens; a.err_flag = timeout_flag_device_ptr(); a.invN = invN;
      a.N = N; a.C = C; a.K = K; a.Kp = ldq; a.ld_tok = ld_tok; a.R = fb_a.R; a.nchunks = fb_a.nchunks; a.splits = fb_a.splits;
      a.zcols = fb_a.zcols; a.nbuf = fb_a.nbuf;
      StageScope _sc(ST_BWD_TOKEN, st); count_launch();
      cudaError_t e = fa::launch_tok_attn_bwd(a, fb_a, BT, st);
      if (e != cudaSuccess) { set_cuda_error(e, "tok_attn_bwd_fused"); return DGSCT_ERR_CUDA; }
    } else {
      // dP = V . dTp^T (fp32) ; dL1 = P (dP - sum_n dP P) -> gQ
      EpiOut e5 = eo(gP, 0, K, (long long)N * K, OUT_STORE);
      RC(run_gemm(G_TK, K_(V, N, C, C, (long long)N * C), bf, K_(dTpOp, K, C, C, dTp_bs), wbf, N, K, C, BT,
                  to_tc_full<tc::EpiOut<false, OUT_STORE, false, false, false>>(e5), e5, AxIdentity{}, true));
      { StageScope _sc(ST_COL_SOFTMAX, st); count_launch();
        if (K > 256) return DGSCT_ERR_UNSUPPORTED;
        { const dim3 grid(csm_num_chunks(N), BT);
          float* part = (float*)(ws + bp.csm_part);
          count_launch();
          col_softmax_bwd_partial_kernel<<<grid, 256, (size_t)(256 / K) * K * 4, st>>>(gP, Pm, bf, part, N, K, ldq);
          col_softmax_bwd_apply_kernel<<<grid, 256, 0, st>>>(gP, Pm, bf, part, gQ, bf, N, K, ldq); }
        KCHECK("col_softmax_bwd"); }
      // dV = P . dTp + dL1 . T + dobar / N
      EpiOut e6 = eo(gV, bf, C, (long long)N * C, OUT_STORE); e6.bcast = dobar; e6.bc_ld = C; e6.bc_scale = invN;
      const void* dL1 = gQ; long long dL1_ld = ldq, dL1_bs = sbq;
      {
        RC(run_gemm(G_TK, K_(Pm, N, K, ldq, sbq), bf, MN_(dTpOp, K, C, C, (long long)K * C), wbf, N, C, K, BT,
                    to_tc_full<tc::EpiOut<true, OUT_STORE, false, true, false>>(e6), e6, AxIdentity{}, true));
        EpiOut e7 = eo(gV, bf, C, (long long)N * C, OUT_ADD);
        RC(run_gemm(G_TK, K_(gQ, N, K, ldq, sbq), bf, MN_(Tok_op, K, C, ld_tok, 0), wbf, N, C, K, BT,
                    to_tc_full<tc::EpiOut<true, OUT_ADD, false, false, false>>(e7), e7, AxIdentity{}, true));
      }
      // d tokens[k,c] += sum_b sum_n dL1[b,n,k] V[b,n,c]   (D(c,k) transposed, atomics over frames)
      RC(fork_side());                                                  // dL1 is final; sum_frames already added its share
      EpiOut e8 = eo(G.tokens, 0, C, 0, OUT_ATOMIC); e8.transposed = 1;
      RC(run_gemm(S_TK, MN_(V, N, C, C, (long long)N * C), bf, MN_(dL1, N, K, dL1_ld, dL1_bs), bf, C, K, N, BT,
                  to_tcT<tc::EpiOutT<OUT_ATOMIC>>(e8), e8, AxIdentity{}, true, 0, true));
  
    }
  }
  // ---- 8. token resampler ---------------------------------------------------------------------------------------------
  const bool fc_first = tcp && use_fc_first(d, P);
  if (fc_first) {
    // forward was V = Wres . O' + b_res (x) s + b_fc with O' = other . Wfc^T kept in the U buffer, s = rowsum(Wfc)
    const void* Op = U;
    const float* fcs = (const float*)(pk + pp.fc_s); float* fct = (float*)(ws + bp.fc_t);   // (row sums of fc.weight: from the pack launch)
    const bool has_bias = !(d.flags & DGSCT_F_NO_CONV_BIAS) && P.resample_b;
    RC(fork_side());                                                  // gV is final (d fc_b: see sum_frames_kernel)
    // dO'[b] (No,C) = Wres^T . dV[b]  -> gU
    EpiOut e1 = eo(gU, bf, C, (long long)No * C, OUT_STORE);
    RC(run_gemm(G_DO, MN_(Wres_op, N, No, ld_res, 0), wbf, MN_(gV, N, C, C, (long long)N * C), bf, No, C, N, BT,
                to_tc_full<tc::EpiOut<true, OUT_STORE, false, false, false>>(e1), e1, AxIdentity{}, true));
    // side: dWres += sum_b dV_b . O'_b^T ; d_bres[n] += sum dV s ; t[c] = sum dV b_res
    if (G.resample_w) {
      EpiOut e3 = eo(G.resample_w, 0, No, 0, OUT_ATOMIC);
      RC(run_gemm(S_DW, K_(gV, N, C, C, (long long)N * C), bf, K_(Op, No, C, C, (long long)No * C), bf, N, No, C, BT,
                  to_tc_full<tc::EpiOut<false, OUT_ATOMIC, false, false, false>>(e3), e3, AxIdentity{}, true, 0, true));
    }
    if (has_bias && G.resample_b) {
      StageScope _sc(ST_MISC, sst); count_launch();
      if (!launch_dv_bias_sums(gV, bf, R, N, C, fcs, P.resample_b, G.resample_b, fct, sst)) {
        count_launch();
        rowsum_frames_weighted_kernel<<<N, 256, 0, sst>>>(gV, bf, BT, N, C, fcs, G.resample_b);
        col_reduce_rowweight_kernel<<<dim3(ceil_div(C, 32), 64), 256, 0, sst>>>(gV, bf, R, N, C, P.resample_b, fct);
      }
      KCHECK("dres_b_reassoc");
    }
    // dother[b] (No,Co) = dO'[b] . Wfc ; side: dWfc += sum_b dO'_b^T . other_b (+ t broadcast over co)
    RC(fork_side());                                                  // gU = dO' is final
    EpiOut e4 = eo(dother, bf, Co, (long long)No * Co, OUT_STORE);
    RC(run_gemm(G_FC, K_(gU, No, C, C, (long long)No * C), bf, MN_(Wfc_op, C, Co, ld_fc, 0), wbf, No, Co, C, BT,
                to_tc_full<tc::EpiOut<true, OUT_STORE, false, false, false>>(e4), e4, AxIdentity{}, true));
    EpiOut e2 = eo(G.fc_w, 0, Co, 0, OUT_ATOMIC);
    RC(run_gemm(S_FC, MN_(gU, No, C, C, (long long)No * C), bf, MN_(other, No, Co, Co, (long long)No * Co), bf, C, Co, No, BT,
                to_tc_full<tc::EpiOut<false, OUT_ATOMIC, false, false, false>>(e2), e2, AxIdentity{}, true, 0, true));
    if (has_bias) {
      StageScope _sc(ST_MISC, sst); count_launch();
      add_col_broadcast_kernel<<<ceil_div(C * Co, 256), 256, 0, sst>>>(G.fc_w, fct, C, Co);
      KCHECK("dfc_w_rank1");
    }
  } else {
    RC(fork_side());                                                  // gV is final (d fc_b: see sum_frames_kernel)
    // dU = dV . Wfc ; dWfc += dV^T . U
    EpiOut e1 = eo(gU, bf, Co, (long long)N * Co, OUT_STORE);
    RC(run_gemm(G_FC, K_(gV, N, C, C, (long long)N * C), bf, MN_(Wfc_op, C, Co, ld_fc, 0), wbf, N, Co, C, BT,
                to_tc_full<tc::EpiOut<true, OUT_STORE, false, false, false>>(e1), e1, AxIdentity{}, true));
    EpiOut e2 = eo(G.fc_w, 0, Co, 0, OUT_ATOMIC);
    RC(run_gemm(S_FC, MN_(gV, N, C, C, (long long)N * C), bf, MN_(U, N, Co, Co, (long long)N * Co), bf, C, Co, N, BT,
                to_tc_full<tc::EpiOut<false, OUT_ATOMIC, false, false, false>>(e2), e2, AxIdentity{}, true, 0, true));
    // dWres += sum_b dU_b . other_b^T ; d_bres ; dother_b = Wres^T . dU_b
    RC(fork_side());                                                  // gU is final
    if (G.resample_w) {
      EpiOut e3 = eo(G.resample_w, 0, No, 0, OUT_ATOMIC);
      RC(run_gemm(S_DW, K_(gU, N, Co, Co, (long long)N * Co), bf, K_(other, No, Co, Co, (long long)No * Co), bf, N, No, Co, BT,
                  to_tc_full<tc::EpiOut<false, OUT_ATOMIC, false, false, false>>(e3), e3, AxIdentity{}, true, 0, true));
    }
    if (G.resample_b && !(d.flags & DGSCT_F_NO_CONV_BIAS)) {
      StageScope _sc(ST_MISC, sst); count_launch();
      rowsum_frames_kernel<<<N, 256, 0, sst>>>(gU, bf, BT, N, Co, G.resample_b);
      KCHECK("dres_b");
    }
    if (P.rs_rowptr) {      // dother[b] = M^T . dU[b] from the transposed stencil
      StageScope _sc(ST_BWD_DOTHER, st); count_launch();
      if (!launch_spmm_frames(P.rst_rowptr, P.rst_col, P.rst_val, gU, dother, bf, BT, No, N, Co, st)) return DGSCT_ERR_UNSUPPORTED;
      KCHECK("dother_spmm");
    } else {
      EpiOut e4 = eo(dother, bf, Co, (long long)No * Co, OUT_STORE);
      RC(run_gemm(G_DO, MN_(Wres_op, N, No, ld_res, 0), wbf, MN_(gU, N, Co, Co, (long long)N * Co), bf, No, Co, N, BT,
                  to_tc_full<tc::EpiOut<true, OUT_STORE, false, false, false>>(e4), e4, AxIdentity{}, true));
    }
  }
  RC(join_side2());
  RC(join_side());
  return DGSCT_OK;
}

}  // namespace dgsct
