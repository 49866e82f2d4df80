// C ABI of libdgsct.so (see include/dgsct.h).
#include <math.h>
#include <string.h>

#include <mutex>
#include <utility>
#include <vector>

#include "adapter_fwd.h"
#include "profile.h"
#include "side_stream.h"

namespace dgsct {
static thread_local char g_err[512] = "";
void set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorName(e), cudaGetErrorString(e));
}

int check_arch() {
  static int cached[kMaxDevices] = {};      // 0 unknown, 1 ok, 2 wrong architecture
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return DGSCT_ERR_CUDA;
  if (dev >= 0 && dev < kMaxDevices && cached[dev]) return cached[dev] == 1 ? DGSCT_OK : DGSCT_ERR_ARCH;
  int major = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return DGSCT_ERR_CUDA;
  const int ok = (major == 10);             // the library holds sm_100a code only
  if (dev >= 0 && dev < kMaxDevices) cached[dev] = ok ? 1 : 2;
  return ok ? DGSCT_OK : DGSCT_ERR_ARCH;
}

static int* g_flag_host = nullptr;
static std::once_flag g_flag_once;
int* timeout_flag_device_ptr() {
  std::call_once(g_flag_once, [] {
    void* h = nullptr;
    if (cudaHostAlloc(&h, 64, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess) {
      g_flag_host = reinterpret_cast<int*>(h);
      *g_flag_host = 0;
    }
  });
  return g_flag_host;      // unified addressing: portable mapped host memory has the same address on every device
}
int timeout_flag_host_value(int reset) {
  if (!g_flag_host) return 0;
  const int v = *reinterpret_cast<volatile int*>(g_flag_host);
  if (reset) *reinterpret_cast<volatile int*>(g_flag_host) = 0;
  return v;
}
// common prologue of the compute entry points
static int entry_checks() {
  const int rc = check_arch();
  if (rc != DGSCT_OK) return rc;
  if (timeout_flag_host_value(0)) return DGSCT_ERR_TIMEOUT;
  return DGSCT_OK;
}
}  // namespace dgsct

using namespace dgsct;

// ---- both directions of one position in one call ------------------------------------------------
__global__ void __launch_bounds__(256) pair_add_kernel(void* __restrict__ dst, const void* __restrict__ src, size_t n, int bf) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    st_act(dst, i, ld_act(dst, i, bf) + ld_act(src, i, bf), bf);
}
static int pair_consistent(const dgsct_desc* a, const dgsct_desc* v) {
  if (!a || !v) return DGSCT_ERR_NULL;
  if (a->BT != v->BT || a->C != v->Co || a->N != v->No || v->C != a->Co || v->N != a->No || a->io_dtype != v->io_dtype) return DGSCT_ERR_BAD_DESC;
  return DGSCT_OK;
}

extern "C" {

int dgsct_version(void) { return DGSCT_VERSION; }

const char* dgsct_status_string(int s) {
  switch (s) {
    case DGSCT_OK: return "ok";
    case DGSCT_ERR_BAD_DESC: return "bad descriptor";
    case DGSCT_ERR_NULL: return "null pointer";
    case DGSCT_ERR_WORKSPACE: return "workspace too small or misaligned";
    case DGSCT_ERR_CUDA: return "CUDA error";
    case DGSCT_ERR_UNSUPPORTED: return "unsupported";
    case DGSCT_ERR_ARCH: return "device is not sm_100";
    case DGSCT_ERR_TIMEOUT: return "a tcgen05 pipeline wait timed out in an earlier call (results since then are invalid)";
    default: return "unknown status";
  }
}

const char* dgsct_last_cuda_error(void) { return g_err; }

size_t dgsct_workspace_bytes(const dgsct_desc* desc) {
  if (validate_desc(desc) != DGSCT_OK) return 0;
  return make_fwd_plan(*desc, (desc->flags & DGSCT_F_KEEP) != 0).total;
}

size_t dgsct_packed_bytes(const dgsct_desc* desc) {
  if (validate_desc(desc) != DGSCT_OK) return 0;
  return packed_bytes(*desc);
}

int dgsct_pack_params(const dgsct_desc* desc, const dgsct_params* params, void* packed, void* stream) {
  int rc = validate_desc(desc);
  if (rc != DGSCT_OK) return rc;
  if (!params || !packed) return DGSCT_ERR_NULL;
  const dgsct_params& P = *params;
  if ((!P.resample_w && !P.rs_rowptr) || !P.fc_w || !P.tokens || !P.v1_w || !P.v2_w || !P.down_w || !P.up_w) return DGSCT_ERR_NULL;
  if ((uintptr_t)packed & 255) return DGSCT_ERR_WORKSPACE;
  if ((rc = entry_checks()) != DGSCT_OK) return rc;
  return pack_params(*desc, P, (char*)packed, (cudaStream_t)stream);
}

int dgsct_adapter_fwd(const dgsct_desc* desc, const dgsct_params* params, const void* packed, const void* x,
                      const void* other, void* out, void* spatial_map, float* temporal_map, const dgsct_taps* taps,
                      void* workspace, size_t workspace_bytes, void* stream) {
  int rc = validate_desc(desc);
  if (rc != DGSCT_OK) return rc;
  if (!params || !x || !other || !out || !spatial_map || !workspace) return DGSCT_ERR_NULL;
  if ((rc = entry_checks()) != DGSCT_OK) return rc;
  const dgsct_desc& d = *desc;
  const dgsct_params& P = *params;
  if ((!P.resample_w && !P.rs_rowptr) || !P.fc_w || !P.fc_b || !P.a1_w || !P.a1_b || !P.v1_w || !P.v1_b || !P.bott_w || !P.bott_b ||
      !P.v2_w || !P.v2_b || !P.a2_w || !P.a2_b || !P.s_w || !P.s_b || !P.c_w || !P.c_b || !P.tokens || !P.gate_av ||
      !P.down_w || !P.up_w)
    return DGSCT_ERR_NULL;
  if (!(d.flags & DGSCT_F_NO_CONV_BIAS) && !P.resample_b) return DGSCT_ERR_NULL;
  if (P.rs_rowptr && (!P.rs_col || !P.rs_val || !(d.flags & DGSCT_F_NO_CONV_BIAS))) return DGSCT_ERR_NULL;
  if ((d.flags & DGSCT_F_USE_GATE) && !P.gate) return DGSCT_ERR_NULL;
  if ((d.flags & DGSCT_F_USE_BN) &&
      (!P.bn1_w || !P.bn1_b || !P.bn2_w || !P.bn2_b || !P.bn1_rm || !P.bn1_rv || !P.bn2_rm || !P.bn2_rv))
    return DGSCT_ERR_NULL;
  if ((d.flags & DGSCT_F_LN_BEFORE) && (!P.lnb_w || !P.lnb_b)) return DGSCT_ERR_NULL;
  if ((d.flags & DGSCT_F_LN_POST) && (!P.lnp_w || !P.lnp_b)) return DGSCT_ERR_NULL;
  if ((d.flags & DGSCT_F_TEMPORAL) && (!P.t_w || !P.t_b)) return DGSCT_ERR_NULL;
  const bool keep = (d.flags & DGSCT_F_KEEP) != 0;
  if (workspace_bytes < make_fwd_plan(d, keep).total || ((uintptr_t)workspace & 255)) return DGSCT_ERR_WORKSPACE;
  if (packed && ((uintptr_t)packed & 255)) return DGSCT_ERR_WORKSPACE;
  if (d.io_dtype == DGSCT_BF16 && !(d.flags & DGSCT_F_FORCE_SIMT))
    return adapter_fwd_tc(d, P, packed, x, other, out, spatial_map, temporal_map, taps, (char*)workspace, (cudaStream_t)stream, keep);
  return adapter_fwd_simt(d, P, x, other, out, spatial_map, temporal_map, taps, (char*)workspace, (cudaStream_t)stream, keep);
}

int dgsct_adapter_pair_fwd(const dgsct_desc* desc_a, const dgsct_params* params_a, const void* packed_a, const dgsct_desc* desc_v,
                           const dgsct_params* params_v, const void* packed_v, const void* f_a, const void* f_v, const void* base_a,
                           const void* base_v, void* out_a, void* out_v, void* smap_a, void* smap_v, float* tmap_a, float* tmap_v,
                           void* ws_a, size_t ws_a_bytes, void* ws_v, size_t ws_v_bytes, void* stream) {
  int rc = pair_consistent(desc_a, desc_v);
  if (rc != DGSCT_OK) return rc;
  if ((rc = validate_desc(desc_a)) != DGSCT_OK || (rc = validate_desc(desc_v)) != DGSCT_OK) return rc;
  if (!params_a || !params_v || !f_a || !f_v || !out_a || !out_v || !smap_a || !smap_v || !ws_a || !ws_v) return DGSCT_ERR_NULL;
  if ((rc = entry_checks()) != DGSCT_OK) return rc;
  const bool keep_a = (desc_a->flags & DGSCT_F_KEEP) != 0, keep_v = (desc_v->flags & DGSCT_F_KEEP) != 0;
  if (ws_a_bytes < make_fwd_plan(*desc_a, keep_a).total || ws_v_bytes < make_fwd_plan(*desc_v, keep_v).total ||
      ((uintptr_t)ws_a & 255) || ((uintptr_t)ws_v & 255))
    return DGSCT_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  // the two adapters read the same two tensors and are independent (net_trans.py:891-892): the audio one runs on an
  // auxiliary stream forked from the caller's, the visual one on the caller's stream
  SideStream* side = (prof_enabled() || getenv("DGSCT_NO_SIDE")) ? nullptr : side_stream_for(st, 3);
  cudaStream_t sa = side ? side->s : st;
  if (side) {
    DGSCT_CUDA_CHECK(cudaEventRecord(side->ev, st));
    DGSCT_CUDA_CHECK(cudaStreamWaitEvent(side->s, side->ev, 0));
  }
  auto run = [&](const dgsct_desc& d, const dgsct_params& P, const void* pk, const void* x, const void* o, void* out, void* smap,
                 float* tmap, void* ws, bool keep, const void* base, cudaStream_t s) -> int {
    if (d.io_dtype == DGSCT_BF16 && !(d.flags & DGSCT_F_FORCE_SIMT))
      return adapter_fwd_tc(d, P, pk, x, o, out, smap, tmap, nullptr, (char*)ws, s, keep, base);
    return adapter_fwd_simt(d, P, x, o, out, smap, tmap, nullptr, (char*)ws, s, keep, base);
  };
  rc = run(*desc_a, *params_a, packed_a, f_a, f_v, out_a, smap_a, tmap_a, ws_a, keep_a, base_a, sa);
  const int rc2 = run(*desc_v, *params_v, packed_v, f_v, f_a, out_v, smap_v, tmap_v, ws_v, keep_v, base_v, st);
  if (side) {
    DGSCT_CUDA_CHECK(cudaEventRecord(side->ev, side->s));
    DGSCT_CUDA_CHECK(cudaStreamWaitEvent(st, side->ev, 0));
  }
  return rc != DGSCT_OK ? rc : rc2;
}

size_t dgsct_bwd_workspace_bytes(const dgsct_desc* desc) {
  if (validate_desc(desc) != DGSCT_OK) return 0;
  return bwd_workspace_bytes(*desc);
}

static int bwd_entry(const dgsct_desc* desc, const dgsct_params* params, const void* packed, const void* saved, const void* x,
                     const void* other, const void* d_out, const void* d_spatial_map, const float* d_temporal_map, void* dx,
                     void* dother, const dgsct_grads* grads, void* workspace, size_t workspace_bytes, void* stream, bool need_saved) {
  int rc = validate_desc(desc);
  if (rc != DGSCT_OK) return rc;
  if (!params || !x || !other || !d_out || !dx || !dother || !grads || !workspace) return DGSCT_ERR_NULL;
  if (need_saved && !saved) return DGSCT_ERR_NULL;
  if ((rc = entry_checks()) != DGSCT_OK) return rc;
  const dgsct_desc& d = *desc;
  const dgsct_grads& G = *grads;
  if (params->rs_rowptr && (!params->rst_rowptr || !params->rst_col || !params->rst_val)) return DGSCT_ERR_NULL;
  if (!G.fc_w || !G.fc_b || !G.a1_w || !G.a1_b || !G.v1_w || !G.v1_b || !G.bott_w || !G.bott_b || !G.v2_w || !G.v2_b ||
      !G.a2_w || !G.a2_b || !G.s_w || !G.s_b || !G.c_w || !G.c_b || !G.tokens || !G.gate_av || !G.down_w || !G.up_w)
    return DGSCT_ERR_NULL;
  if ((d.flags & DGSCT_F_USE_GATE) && !G.gate) return DGSCT_ERR_NULL;
  if ((d.flags & DGSCT_F_USE_BN) && (!G.bn1_w || !G.bn1_b || !G.bn2_w || !G.bn2_b)) return DGSCT_ERR_NULL;
  if ((d.flags & DGSCT_F_LN_BEFORE) && (!G.lnb_w || !G.lnb_b)) return DGSCT_ERR_NULL;
  if ((d.flags & DGSCT_F_LN_POST) && (!G.lnp_w || !G.lnp_b)) return DGSCT_ERR_NULL;
  if ((d.flags & DGSCT_F_TEMPORAL) && (!G.t_w || !G.t_b)) return DGSCT_ERR_NULL;
  if (workspace_bytes < bwd_workspace_bytes(d, saved != nullptr) || ((uintptr_t)workspace & 255)) return DGSCT_ERR_WORKSPACE;
  if (packed && ((uintptr_t)packed & 255)) return DGSCT_ERR_WORKSPACE;
  if (saved && ((uintptr_t)saved & 255)) return DGSCT_ERR_WORKSPACE;
  return adapter_bwd(d, *params, packed, x, other, d_out, d_spatial_map, d_temporal_map, dx, dother, G, (char*)workspace,
                     (cudaStream_t)stream, (const char*)saved);
}

int dgsct_adapter_bwd(const dgsct_desc* desc, const dgsct_params* params, const void* packed, const void* x, const void* other,
                      const void* d_out, const void* d_spatial_map, const float* d_temporal_map, void* dx, void* dother,
                      const dgsct_grads* grads, void* workspace, size_t workspace_bytes, void* stream) {
  return bwd_entry(desc, params, packed, nullptr, x, other, d_out, d_spatial_map, d_temporal_map, dx, dother, grads, workspace,
                   workspace_bytes, stream, false);
}

size_t dgsct_bwd_saved_workspace_bytes(const dgsct_desc* desc) {
  if (validate_desc(desc) != DGSCT_OK) return 0;
  return bwd_workspace_bytes(*desc, true);
}

int dgsct_adapter_bwd_saved(const dgsct_desc* desc, const dgsct_params* params, const void* packed,
                            const void* saved_forward_workspace, const void* x, const void* other, const void* d_out,
                            const void* d_spatial_map, const float* d_temporal_map, void* dx, void* dother,
                            const dgsct_grads* grads, void* workspace, size_t workspace_bytes, void* stream) {
  return bwd_entry(desc, params, packed, saved_forward_workspace, x, other, d_out, d_spatial_map, d_temporal_map, dx, dother,
                   grads, workspace, workspace_bytes, stream, true);
}

int dgsct_adapter_pair_bwd_saved(const dgsct_desc* desc_a, const dgsct_params* params_a, const void* packed_a, const void* saved_a,
                                 const dgsct_desc* desc_v, const dgsct_params* params_v, const void* packed_v, const void* saved_v,
                                 const void* f_a, const void* f_v, const void* d_out_a, const void* d_out_v, const void* d_smap_a,
                                 const void* d_smap_v, void* d_f_a, void* d_f_v, void* tmp_f_a, void* tmp_f_v,
                                 const dgsct_grads* grads_a, const dgsct_grads* grads_v, void* ws_a, size_t ws_a_bytes, void* ws_v,
                                 size_t ws_v_bytes, void* stream) {
  int rc = pair_consistent(desc_a, desc_v);
  if (rc != DGSCT_OK) return rc;
  if (!d_f_a || !d_f_v || !tmp_f_a || !tmp_f_v) return DGSCT_ERR_NULL;
  cudaStream_t st = (cudaStream_t)stream;
  SideStream* side = (prof_enabled() || getenv("DGSCT_NO_SIDE")) ? nullptr : side_stream_for(st, 3);
  cudaStream_t sa = side ? side->s : st;
  if (side) {
    DGSCT_CUDA_CHECK(cudaEventRecord(side->ev, st));
    DGSCT_CUDA_CHECK(cudaStreamWaitEvent(side->s, side->ev, 0));
  }
  // audio adapter: x = f_a (dx -> d_f_a), other = f_v (dother -> tmp_f_v); visual adapter: x = f_v (dx -> d_f_v), other = f_a
  rc = bwd_entry(desc_a, params_a, packed_a, saved_a, f_a, f_v, d_out_a, d_smap_a, nullptr, d_f_a, tmp_f_v, grads_a, ws_a, ws_a_bytes,
                 (void*)sa, true);
  const int rc2 = bwd_entry(desc_v, params_v, packed_v, saved_v, f_v, f_a, d_out_v, d_smap_v, nullptr, d_f_v, tmp_f_a, grads_v, ws_v,
                            ws_v_bytes, stream, true);
  if (side) {
    DGSCT_CUDA_CHECK(cudaEventRecord(side->ev, side->s));
    DGSCT_CUDA_CHECK(cudaStreamWaitEvent(st, side->ev, 0));
  }
  if (rc != DGSCT_OK || rc2 != DGSCT_OK) return rc != DGSCT_OK ? rc : rc2;
  const int bf = desc_a->io_dtype == DGSCT_BF16;
  const size_t na = (size_t)desc_a->BT * desc_a->N * desc_a->C, nv = (size_t)desc_v->BT * desc_v->N * desc_v->C;
  StageScope _sc(ST_MISC, st); count_launch(2);
  pair_add_kernel<<<148 * 4, 256, 0, st>>>(d_f_a, tmp_f_a, na, bf);     // d f_a = dx(audio adapter) + dother(visual adapter)
  pair_add_kernel<<<148 * 4, 256, 0, st>>>(d_f_v, tmp_f_v, nv, bf);
  DGSCT_LAUNCH_CHECK("pair_add");
  return DGSCT_OK;
}

// Cubic convolution coefficients (A = -0.75), the kernel torch uses for mode='bicubic'.
static void cubic_coeffs(double t, double w[4]) {
  const double A = -0.75;
  auto c1 = [&](double x) { return ((A + 2) * x - (A + 3)) * x * x + 1; };          // |x| <= 1
  auto c2 = [&](double x) { return ((A * x - 5 * A) * x + 8 * A) * x - 4 * A; };    // 1 < |x| < 2
  w[0] = c2(t + 1.0); w[1] = c1(t); w[2] = c1(1.0 - t); w[3] = c2(2.0 - t);
}

int dgsct_bicubic_matrix(int n_in, int n_out, float* w) {
  if (n_in <= 0 || n_out <= 0 || !w) return DGSCT_ERR_BAD_DESC;
  const size_t cols = (size_t)n_in * n_in;
  memset(w, 0, sizeof(float) * cols * (size_t)n_out * n_out);
  const double scale = (double)n_in / (double)n_out;   // align_corners = False
  for (int oy = 0; oy < n_out; ++oy) {
    const double sy = scale * (oy + 0.5) - 0.5;
    const int iy = (int)floor(sy);
    double wy[4]; cubic_coeffs(sy - iy, wy);
    for (int ox = 0; ox < n_out; ++ox) {
      const double sx = scale * (ox + 0.5) - 0.5;
      const int ix = (int)floor(sx);
      double wx[4]; cubic_coeffs(sx - ix, wx);
      float* row = w + ((size_t)oy * n_out + ox) * cols;
      for (int a = 0; a < 4; ++a) {
        int yy = iy - 1 + a; yy = yy < 0 ? 0 : (yy >= n_in ? n_in - 1 : yy);   // border clamp
        for (int b = 0; b < 4; ++b) {
          int xx = ix - 1 + b; xx = xx < 0 ? 0 : (xx >= n_in ? n_in - 1 : xx);
          row[(size_t)yy * n_in + xx] += (float)(wy[a] * wx[b]);
        }
      }
    }
  }
  return DGSCT_OK;
}

int dgsct_bicubic_csr_capacity(int n_in, int n_out) {
  if (n_in <= 0 || n_out <= 0) return DGSCT_ERR_BAD_DESC;
  return 16 * n_out * n_out;
}

int dgsct_bicubic_csr(int n_in, int n_out, int transpose, int32_t* rowptr, int32_t* col, float* val) {
  if (n_in <= 0 || n_out <= 0 || !rowptr || !col || !val) return DGSCT_ERR_BAD_DESC;
  const int rows_out = n_out * n_out, rows_in = n_in * n_in;
  // taps of every output pixel (border-clamped duplicates merged), in fp64
  std::vector<std::vector<std::pair<int, double>>> taps(rows_out);
  const double scale = (double)n_in / (double)n_out;   // align_corners = False
  for (int oy = 0; oy < n_out; ++oy) {
    const double sy = scale * (oy + 0.5) - 0.5;
    const int iy = (int)floor(sy);
    double wy[4]; cubic_coeffs(sy - iy, wy);
    for (int ox = 0; ox < n_out; ++ox) {
      const double sx = scale * (ox + 0.5) - 0.5;
      const int ix = (int)floor(sx);
      double wx[4]; cubic_coeffs(sx - ix, wx);
      auto& t = taps[(size_t)oy * n_out + ox];
      for (int a = 0; a < 4; ++a) {
        int yy = iy - 1 + a; yy = yy < 0 ? 0 : (yy >= n_in ? n_in - 1 : yy);
        for (int b = 0; b < 4; ++b) {
          int xx = ix - 1 + b; xx = xx < 0 ? 0 : (xx >= n_in ? n_in - 1 : xx);
          const int c = yy * n_in + xx;
          bool merged = false;
          for (auto& e : t) if (e.first == c) { e.second += wy[a] * wx[b]; merged = true; break; }
          if (!merged) t.emplace_back(c, wy[a] * wx[b]);
        }
      }
    }
  }
  int nnz = 0;
  if (!transpose) {
    for (int r = 0; r < rows_out; ++r) {
      rowptr[r] = nnz;
      for (auto& e : taps[r]) { col[nnz] = e.first; val[nnz] = (float)e.second; ++nnz; }
    }
    rowptr[rows_out] = nnz;
  } else {
    std::vector<int> cnt(rows_in + 1, 0);
    for (int r = 0; r < rows_out; ++r) for (auto& e : taps[r]) ++cnt[e.first + 1];
    for (int i = 0; i < rows_in; ++i) cnt[i + 1] += cnt[i];
    for (int i = 0; i <= rows_in; ++i) rowptr[i] = cnt[i];
    std::vector<int> fill(cnt.begin(), cnt.end() - 1);
    for (int r = 0; r < rows_out; ++r)
      for (auto& e : taps[r]) { const int q = fill[e.first]++; col[q] = r; val[q] = (float)e.second; }
    nnz = cnt[rows_in];
  }
  return nnz;
}

}  // extern "C"
