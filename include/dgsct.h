/* dgsct.h -- C ABI of the B200-native DG-SCT adapter engine (libdgsct.so).
 *
 * The reference (haoyi-duan/DG-SCT) has no native code and no FFI: its boundary for this path is the
 * Python call  VisualAdapter.forward(x, vis_token)  (DG-SCT/AVE/nets/net_trans.py:552-674 and the per-task
 * copies listed in SURVEY.md section 2.1).  This header is what a ctypes binding of that call binds to;
 * `dgsct_b200/adapter.py` is that binding and INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless noted; no allocation and
 *     no synchronisation happen inside any call; work is enqueued on `stream` (a cudaStream_t).
 *   - activations are token-major: x (BT,N,C), other (BT,No,Co), out (BT,N,C), spatial_map (BT,N),
 *     temporal_map (BT/T, T).  The reference's (BT,C,N,1) tensors are permuted views of these buffers
 *     (net_trans.py:891), so the binding passes them through without a copy.
 *   - parameters are the fp32 master tensors of the module's state_dict, contiguous.
 *   - return value: 0 on success, a negative dgsct_status otherwise; nothing throws across the ABI.
 */
#ifndef DGSCT_H_
#define DGSCT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGSCT_VERSION 100  /* 0.1.0 */

typedef enum dgsct_status {
  DGSCT_OK = 0,
  DGSCT_ERR_BAD_DESC = -1,      /* inconsistent dimensions / unsupported flag combination          */
  DGSCT_ERR_NULL = -2,          /* a required pointer is NULL                                       */
  DGSCT_ERR_WORKSPACE = -3,     /* workspace too small or misaligned (256 B)                        */
  DGSCT_ERR_CUDA = -4,          /* a CUDA runtime/driver call failed (see dgsct_last_cuda_error)    */
  DGSCT_ERR_UNSUPPORTED = -5,   /* valid request the engine does not implement (yet)                */
  DGSCT_ERR_ARCH = -6,          /* device is not sm_100 (checked on every compute entry point)      */
  DGSCT_ERR_TIMEOUT = -7        /* a bounded tcgen05 pipeline wait expired in an EARLIER call: sticky until
                                   dgsct_tc_error_flag(1); everything computed since is invalid      */
} dgsct_status;

typedef enum dgsct_dtype { DGSCT_F32 = 0, DGSCT_BF16 = 1 } dgsct_dtype;

/* flag bits of dgsct_desc.flags */
#define DGSCT_F_TRAIN          (1u << 0)  /* BatchNorm uses batch statistics and updates running stats */
#define DGSCT_F_USE_BN         (1u << 1)  /* use_bn   (net_trans.py:507,632,640)                        */
#define DGSCT_F_USE_GATE       (1u << 2)  /* use_gate (net_trans.py:460,670)                            */
#define DGSCT_F_LN_BEFORE      (1u << 3)  /* opt.is_before_layernorm (net_trans.py:626)                 */
#define DGSCT_F_LN_POST        (1u << 4)  /* opt.is_post_layernorm   (net_trans.py:667)                 */
#define DGSCT_F_GATE_BEFORE_LN (1u << 5)  /* AVS order: gate, then ln_post (PVT_AVSModel.py:308-313)    */
#define DGSCT_F_TEMPORAL       (1u << 6)  /* compute temporal_att_maps (few-shot/nets/net_trans.py:534) */
#define DGSCT_F_NO_CONV_BIAS   (1u << 7)  /* resampling matrix has no bias (bicubic variant)            */
#define DGSCT_F_FORCE_SIMT     (1u << 8)  /* debugging: never take the tcgen05 kernels                  */
#define DGSCT_F_NO_STAT_UPDATE (1u << 9)  /* TRAIN statistics without touching the running buffers       */
#define DGSCT_F_NO_OUTPUT      (1u << 10) /* internal (backward recompute): skip the final ln_post/gate pass */
#define DGSCT_F_KEEP           (1u << 11) /* training forward: every intermediate the backward needs survives in `workspace`
                                             (non-aliasing plan, larger dgsct_workspace_bytes) for dgsct_adapter_bwd_saved */

typedef struct dgsct_desc {
  int32_t BT;       /* frames in the batch (clips x T)                        */
  int32_t T;        /* frames per clip (temporal gate only; BT % T == 0)      */
  int32_t C, N;     /* width / tokens of the adapted stream                   */
  int32_t Co, No;   /* width / tokens of the other (prompt) stream            */
  int32_t K;        /* latent tokens (rows of my_tokens)                      */
  int32_t r, g;     /* bottleneck reduction factor, conv groups               */
  int32_t io_dtype; /* dgsct_dtype of x, other, out, spatial_map              */
  uint32_t flags;
  float alpha, beta, gamma;   /* gate mixing weights (net_trans.py:611; few-shot :537-549) */
  float eps_ln, eps_bn, bn_momentum;
} dgsct_desc;

/* fp32 master parameters, same names as the reference state_dict.
 * `resample_w` is conv_adapter.weight viewed (N,No) -- or, for the bicubic variant, the dense
 * interpolation matrix built by the binding; `resample_b` may be NULL with DGSCT_F_NO_CONV_BIAS. */
typedef struct dgsct_params {
  const float *resample_w, *resample_b;        /* (N,No), (N)            conv_adapter            */
  const float *fc_w, *fc_b;                    /* (C,Co), (C)            fc                      */
  const float *a1_w, *a1_b;                    /* (C,C), (C)             fc_affine_audio_1       */
  const float *v1_w, *v1_b;                    /* (C,C), (C)             fc_affine_video_1       */
  const float *bott_w, *bott_b;                /* (C/2,C), (C/2)         fc_affine_bottleneck    */
  const float *v2_w, *v2_b;                    /* (C/2,C), (C/2)         fc_affine_video_2       */
  const float *a2_w, *a2_b;                    /* (C/2,C), (C/2)         fc_affine_audio_2       */
  const float *s_w, *s_b;                      /* (1,C/2), (1)           fc_affine_v_s_att       */
  const float *c_w, *c_b;                      /* (C,C/2), (C)           fc_affine_v_c_att       */
  const float *tokens;                         /* (K,C)                  my_tokens               */
  const float *gate_av;                        /* (1)                                            */
  const float *gate;                           /* (1) or NULL                                    */
  const float *down_w, *up_w;                  /* (C/r, C/g), (C, C/(r g))  grouped 1x1 convs    */
  const float *bn1_w, *bn1_b, *bn2_w, *bn2_b;  /* (C/r) x2, (C) x2  or NULL                      */
  float *bn1_rm, *bn1_rv, *bn2_rm, *bn2_rv;    /* running stats, updated in place when TRAIN     */
  const float *lnb_w, *lnb_b, *lnp_w, *lnp_b;  /* (C) each or NULL                               */
  const float *t_w, *t_b;                      /* (1,C), (1)  temporal_gated.0 or NULL           */
  /* Optional SPARSE resampling matrix (CSR, fp32 values): the AVS-S4 bicubic interpolation is a 16-tap stencil
   * (PVT_AVSModel.py:190-197), not a dense (N,No) matrix.  When rs_rowptr != NULL the engine computes
   * U[b] = M . other[b] with one gather kernel (fp32 coefficients, no bf16 rounding of the stencil) and
   * dother[b] = M^T . dU[b] from the transposed CSR; resample_w may then be NULL.  Build both with dgsct_bicubic_csr. */
  const int32_t *rs_rowptr, *rs_col;           /* (N+1), (nnz)   CSR of M   (N x No)             */
  const float *rs_val;                         /* (nnz)                                          */
  const int32_t *rst_rowptr, *rst_col;         /* (No+1), (nnz)  CSR of M^T (No x N)             */
  const float *rst_val;                        /* (nnz)                                          */
} dgsct_params;

/* fp32 gradient buffers, one per entry of dgsct_params (same shapes).  dgsct_adapter_bwd ADDS into them, so the
 * caller zero-fills (or accumulates across micro-batches, main_trans.py:135-136).  Entries whose parameter is absent
 * for this descriptor (no gate / BN / LN / temporal gate; the bicubic resampler has no trainable matrix) may be NULL. */
typedef struct dgsct_grads {
  float *resample_w, *resample_b, *fc_w, *fc_b, *a1_w, *a1_b, *v1_w, *v1_b, *bott_w, *bott_b, *v2_w, *v2_b, *a2_w, *a2_b,
      *s_w, *s_b, *c_w, *c_b, *tokens, *gate_av, *gate, *down_w, *up_w, *bn1_w, *bn1_b, *bn2_w, *bn2_b, *lnb_w, *lnb_b,
      *lnp_w, *lnp_b, *t_w, *t_b;
} dgsct_grads;

/* Optional taps of intermediate tensors (fp32, token-major) for stage-by-stage parity tests.
 * Any pointer may be NULL. */
typedef struct dgsct_taps {
  float *V;      /* (BT,N,C)  resampled prompt tokens          */
  float *Tp;     /* (BT,K,C)  updated latent tokens            */
  float *X1;     /* (BT,N,C)  x after latent-token attention   */
  float *obar;   /* (BT,C)    mean_n V                         */
  float *m;      /* (BT,C)    channel gate                     */
  float *s;      /* (BT,N)    spatial logits incl. bias        */
  float *Y;      /* (BT,N,C)  gated (+ln_before) activations   */
  float *Z;      /* (BT,N,C/r) bottleneck pre-BN               */
} dgsct_taps;

int dgsct_version(void);
const char *dgsct_status_string(int status);
/* text of the last CUDA error seen by the calling thread (host pointer, never NULL) */
const char *dgsct_last_cuda_error(void);

/* bytes of scratch the forward needs for `desc` (0 if desc is invalid) */
size_t dgsct_workspace_bytes(const dgsct_desc *desc);

/* bf16 operand copies of the weights for the tensor-core path ("packed weights"): size of the buffer and the
 * conversion itself (device buffer, 256-B aligned).  Re-pack whenever the fp32 masters change. */
size_t dgsct_packed_bytes(const dgsct_desc *desc);
int dgsct_pack_params(const dgsct_desc *desc, const dgsct_params *params, void *packed, void *stream);

/* VisualAdapter.forward.  x,other,out,spatial_map are `io_dtype`; temporal_map is fp32 or NULL; `taps` may be
 * NULL.  io_dtype fp32 runs exact fp32 SIMT kernels; io_dtype bf16 runs the tcgen05 tensor-core kernels
 * with fp32 accumulation and fp32 softmax/LayerNorm/BatchNorm statistics.  `packed` is the buffer filled by
 * dgsct_pack_params for these parameters, or NULL to convert the weights inside the call (bf16 only). */
int dgsct_adapter_fwd(const dgsct_desc *desc, const dgsct_params *params, const void *packed, const void *x,
                      const void *other, void *out, void *spatial_map, float *temporal_map, const dgsct_taps *taps,
                      void *workspace, size_t workspace_bytes, void *stream);

/* Backward of VisualAdapter.forward.  The forward keeps no state: this call recomputes the forward into its own
 * workspace (activation recomputation; BatchNorm batch statistics are recomputed, running buffers are NOT updated
 * again) and then back-propagates.  d_out / d_spatial_map / dx / dother are `io_dtype`; d_spatial_map and
 * d_temporal_map (fp32, (BT)) may be NULL.  `desc` must equal the forward's (same TRAIN flag). */
size_t dgsct_bwd_workspace_bytes(const dgsct_desc *desc);
int dgsct_adapter_bwd(const dgsct_desc *desc, const dgsct_params *params, const void *packed, const void *x,
                      const void *other, const void *d_out, const void *d_spatial_map, const float *d_temporal_map,
                      void *dx, void *dother, const dgsct_grads *grads, void *workspace, size_t workspace_bytes,
                      void *stream);

/* Backward from SAVED activations: `saved_forward_workspace` is the workspace of the dgsct_adapter_fwd call made with
 * DGSCT_F_KEEP for the same desc / parameters / inputs (untouched since).  No recomputation happens, so the call costs
 * the backward proper; the price is the memory of one kept workspace per adapter call in flight (0.1-0.6 GB at the AVE
 * shapes, ~12 GB for the 48 calls of a step: small against 180 GB of HBM3e).  `desc` is the forward's, with or without
 * DGSCT_F_KEEP. */
size_t dgsct_bwd_saved_workspace_bytes(const dgsct_desc *desc);
int dgsct_adapter_bwd_saved(const dgsct_desc *desc, const dgsct_params *params, const void *packed,
                            const void *saved_forward_workspace, const void *x, const void *other, const void *d_out,
                            const void *d_spatial_map, const float *d_temporal_map, void *dx, void *dother,
                            const dgsct_grads *grads, void *workspace, size_t workspace_bytes, void *stream);

/* BOTH directions of one position in one call.  The interleave loop calls
 * audio_adapter_blocks_pX[i](f_a, f_v) and vis_adapter_blocks_pX[i](f_v, f_a) back to back on the same two tensors
 * (net_trans.py:891-892, 900-901) and later adds the results to the token streams (:895-906).  desc_a / params_a describe
 * the audio adapter (x = f_a (BT,Na,Ca), other = f_v), desc_v / params_v the visual one (x = f_v, other = f_a); the two run
 * concurrently on forked streams and are joined before the call returns.  base_a / base_v (io_dtype, shapes of f_a / f_v, may be
 * NULL): out_a = base_a + adapter_a(...), the caller's residual add folded into the last pass (pass the output of the frozen
 * block, which the reference adds the adapter result to).  Workspaces as in dgsct_adapter_fwd (one per adapter).
 * The backward (saved activations) returns d_f_a = dx(audio) + dother(visual) and d_f_v likewise; tmp_f_a / tmp_f_v are scratch
 * tensors of the shapes of f_a / f_v; the gradient w.r.t. base_* is d_out_* itself. */
int dgsct_adapter_pair_fwd(const dgsct_desc *desc_a, const dgsct_params *params_a, const void *packed_a, const dgsct_desc *desc_v,
                           const dgsct_params *params_v, const void *packed_v, const void *f_a, const void *f_v, const void *base_a,
                           const void *base_v, void *out_a, void *out_v, void *smap_a, void *smap_v, float *tmap_a, float *tmap_v,
                           void *ws_a, size_t ws_a_bytes, void *ws_v, size_t ws_v_bytes, void *stream);
int dgsct_adapter_pair_bwd_saved(const dgsct_desc *desc_a, const dgsct_params *params_a, const void *packed_a, const void *saved_a,
                                 const dgsct_desc *desc_v, const dgsct_params *params_v, const void *packed_v, const void *saved_v,
                                 const void *f_a, const void *f_v, const void *d_out_a, const void *d_out_v, const void *d_smap_a,
                                 const void *d_smap_v, void *d_f_a, void *d_f_v, void *tmp_f_a, void *tmp_f_v,
                                 const dgsct_grads *grads_a, const dgsct_grads *grads_v, void *ws_a, size_t ws_a_bytes, void *ws_v,
                                 size_t ws_v_bytes, void *stream);

/* the two ops directly behind the adapters in the task networks.
 *   dgsct_pool_tokens_*     out (BT,C) = sum_n spatial_map[b,n] f[b,n,c]  -- torch.bmm(spatial_att_maps, f) of
 *                           DG-SCT/AVE/nets/net_trans.py:922,924 (the adapter's second return value pools the tokens);
 *                           bwd writes d_map (BT,N) and / or d_f (BT,N,C) (either may be NULL)
 *   dgsct_temporal_gates_*  the sigmoid gates closing TemporalAttention.forward (net_trans.py:243-249): rows = T*B rows of
 *                           D channels; a_kv / v_kv = encoder outputs, vq / aq = decoder outputs; wa, ba, wv, bv = the
 *                           audio_gated.0 / video_gated.0 Linear(D,1) parameters (fp32); `gates` (rows,2) fp32 keeps
 *                           (audio_gate, video_gate) for the backward; parameter gradients are ADDED.
 * activations are io_dtype (dgsct_dtype). */
int dgsct_pool_tokens_fwd(const void *spatial_map, const void *f, void *out, int BT, int N, int C, int io_dtype, void *stream);
int dgsct_pool_tokens_bwd(const void *spatial_map, const void *f, const void *d_out, void *d_map, void *d_f, int BT, int N, int C,
                          int io_dtype, void *stream);
int dgsct_temporal_gates_fwd(const void *a_kv, const void *v_kv, const void *vq, const void *aq, const float *wa, const float *ba,
                             const float *wv, const float *bv, float gamma, void *vq_out, void *aq_out, void *av_gate, float *gates,
                             int rows, int D, int io_dtype, void *stream);
int dgsct_temporal_gates_bwd(const void *a_kv, const void *v_kv, const void *vq, const void *aq, const float *wa, const float *wv,
                             float gamma, const float *gates, const void *d_vq_out, const void *d_aq_out, const void *d_av_gate,
                             void *d_a_kv, void *d_v_kv, void *d_vq, void *d_aq, float *d_wa, float *d_ba, float *d_wv, float *d_bv,
                             int rows, int D, int io_dtype, void *stream);

/* Optimizer over FLAT buffers (the gradient buckets of the data-parallel wrapper): torch.optim.Adam semantics
 * (DG-SCT/AVE/main_trans.py:276: Adam over the trainable adapter parameters), one fused kernel per bucket.
 *   dgsct_adam_tick   increments the device-resident step counter (once per optimizer step; graph-capturable)
 *   dgsct_adam_flat   p, m, v fp32 (n elements, 16-B aligned); g fp32 or bf16 (g_is_bf16); grad_scale multiplies g first
 *                     (1/world after a SUM all-reduce); bias corrections use *step_dev
 *   dgsct_cast_bucket_bf16   dst = bf16(scale * src): communication copy of an fp32 bucket for a bf16 all-reduce */
int dgsct_adam_tick(int *step_dev, void *stream);
int dgsct_adam_flat(float *p, const void *g, int g_is_bf16, float *m, float *v, size_t n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, float grad_scale, const int *step_dev, void *stream);
int dgsct_cast_bucket_bf16(const float *src, void *dst_bf16, size_t n, float scale, void *stream);

/* Stage profiler (used by bench.py for the per-kernel roofline).  When enabled every pipeline stage is
 * bracketed by CUDA events on the launching stream; collect() synchronises on them, writes accumulated
 * milliseconds and call counts per stage (arrays of >= dgsct_num_stages() entries) and clears the log. */
int dgsct_num_stages(void);
const char *dgsct_stage_name(int stage);
int dgsct_profile_enable(int on);
int dgsct_profile_collect(double *ms_per_stage, long long *calls_per_stage, int n);
/* kernels launched by the engine since the library was loaded (bench.py's gpu_launches) */
long long dgsct_launch_count(void);

/* Bring-up entry (tests): D[b] (M,N) fp32 = A[b] . B[b]^T with bf16 operands through the tcgen05 kernel
 * (use_tc=1) or the SIMT kernel (use_tc=0).  K-major operand: element (row,k) at row*ld + k; MN-major operand:
 * element (row,k) at k*ld + row; batch strides in elements, 0 = shared.  bn_hint: low 16 bits tile N (0 = auto). */
int dgsct_debug_gemm_bf16(const void *A, const void *B, float *D, int M, int N, int K, int batch, int a_mn_major,
                          int b_mn_major, long long lda, long long ldb, long long a_batch_stride,
                          long long b_batch_stride, int use_tc, int bn_hint, int stages_hint, void *stream);
/* value of the sticky flag (host-mapped memory) raised when a bounded pipeline wait expired (0 = healthy); synchronises
 * the current device first; reset != 0 clears it.  The compute entry points test the same flag WITHOUT synchronising and
 * return DGSCT_ERR_TIMEOUT once it is set. */
int dgsct_tc_error_flag(int reset);

/* Host helper (no device work): launch geometry of the fused latent-token attention kernels (attn_fused.cuh) for one adapter
 * call of BT frames, N tokens x C channels, K latent tokens.  kernel: 0 = tokens over V (forward), 1 = stream over tokens
 * (forward), 2 = backward of 1, 3 = backward of 0.  out[8] = {supported, rows_per_chunk, buffers, chunks, splits,
 * column_ranges, columns_per_range, dynamic_smem_bytes} (kernel 1: {supported, warps_per_cta, cooperative, 0, ctas_per_frame,
 * 1, C, smem}).  Tests pin the properties the bit-for-bit batch == clips guarantee rests on (chunk rows and splits of the
 * forward depend on the frame's shape only). */
int dgsct_debug_attn_plan(int kernel, int BT, int N, int C, int K, int *out8);

/* Host helper: dense (n_out^2, n_in^2) matrix of bicubic (A=-0.75, align_corners=False) interpolation
 * on a square token grid, written to HOST memory `w` (row-major).  Used for the AVS-S4 resampler
 * (PVT_AVSModel.py:190-197). */
int dgsct_bicubic_matrix(int n_in_side, int n_out_side, float *w_host);
/* Host helper: the same interpolation as CSR (transpose = 0: rows = n_out^2 outputs, <= 16 taps each; transpose = 1: rows =
 * n_in^2 inputs).  rowptr_host has rows+1 entries; col_host / val_host hold up to dgsct_bicubic_csr_capacity() entries
 * (taps that clamp onto the same border pixel are merged).  Returns the number of stored entries, or a negative status. */
int dgsct_bicubic_csr_capacity(int n_in_side, int n_out_side);
int dgsct_bicubic_csr(int n_in_side, int n_out_side, int transpose, int32_t *rowptr_host, int32_t *col_host, float *val_host);

#ifdef __cplusplus
}
#endif
#endif /* DGSCT_H_ */
