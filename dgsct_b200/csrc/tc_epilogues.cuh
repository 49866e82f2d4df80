// Epilogue functors for gemm_tc_kernel.  A thread owns one accumulator row (TMEM lane) and receives it in
// chunks of 16 consecutive columns; `chunk` is always called by all 32 lanes of a warp (shuffles are legal).
//
// Keep these bodies SMALL: the first version used one run-time-generic functor (activation switch, optional
// bias/transposed/tap paths per element) whose 16x-unrolled body was ~32 KB of SASS; the epilogue warps then ran
// at ~0.1 IPC on instruction-cache misses and every product was epilogue-bound (profiles/r01_gemm_ablation.txt).
// Everything optional is a template parameter or hoisted out of the per-element loop.  Register arrays are
// only ever indexed with compile-time constants (a run-time index silently moves them to local memory).
#pragma once
#include "epilogues.cuh"
#include "gemm_tc.cuh"

namespace dgsct {
namespace tc {

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

__device__ __forceinline__ void load16(const float* __restrict__ p, float (&d)[16]) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p) + j);
    d[4 * j] = t.x; d[4 * j + 1] = t.y; d[4 * j + 2] = t.z; d[4 * j + 3] = t.w;
  }
}
// 16 floats starting at p[0]; elements >= nvalid read as 0 (ragged right edge or unaligned base)
__device__ __forceinline__ void load16_guard(const float* __restrict__ p, float (&d)[16], int nvalid) {
  if (nvalid == 16 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) { load16(p, d); return; }
#pragma unroll
  for (int j = 0; j < 16; ++j) d[j] = j < nvalid ? p[j] : 0.f;   // static indices: keep d[] in registers
}

template <bool OUT_BF16>
__device__ __forceinline__ void store_row16(void* out, size_t idx, const float (&v)[16], int nvalid) {
  if (OUT_BF16) {
    bf16* o = reinterpret_cast<bf16*>(out) + idx;
    if (nvalid == 16 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
      uint4 a, b;
      a.x = pack_bf16x2(v[0], v[1]); a.y = pack_bf16x2(v[2], v[3]); a.z = pack_bf16x2(v[4], v[5]); a.w = pack_bf16x2(v[6], v[7]);
      b.x = pack_bf16x2(v[8], v[9]); b.y = pack_bf16x2(v[10], v[11]); b.z = pack_bf16x2(v[12], v[13]); b.w = pack_bf16x2(v[14], v[15]);
      reinterpret_cast<uint4*>(o)[0] = a;
      reinterpret_cast<uint4*>(o)[1] = b;
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nvalid) o[j] = __float2bfloat16_rn(v[j]);
    }
  } else {
    float* o = reinterpret_cast<float*>(out) + idx;
    if (nvalid == 16 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) reinterpret_cast<float4*>(o)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nvalid) o[j] = v[j];
    }
  }
}

// 16 bf16 starting at p (16-byte aligned, all valid) as two 128-bit loads; unpack after ALL loads of a block were issued
struct Raw16 { uint4 a, b; };
__device__ __forceinline__ Raw16 ldg_raw16(const bf16* __restrict__ p) {
  Raw16 r; r.a = reinterpret_cast<const uint4*>(p)[0]; r.b = reinterpret_cast<const uint4*>(p)[1];
  return r;
}
__device__ __forceinline__ void add_raw16(const Raw16& r, float (&v)[16]) {
  const uint32_t w[8] = {r.a.x, r.a.y, r.a.z, r.a.w, r.b.x, r.b.y, r.b.z, r.b.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
    v[2 * j] += __bfloat162float(t.x);
    v[2 * j + 1] += __bfloat162float(t.y);
  }
}
// fp32 atomic accumulation of 16 consecutive values: four vector reductions when the address is 16-byte aligned
__device__ __forceinline__ void red_add16(float* o, const float (&v)[16], int nvalid) {
  if (nvalid == 16 && (reinterpret_cast<uintptr_t>(o) & 15) == 0) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + 4 * j), "f"(v[4 * j]), "f"(v[4 * j + 1]), "f"(v[4 * j + 2]),
                   "f"(v[4 * j + 3])
                   : "memory");
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < nvalid) atomicAdd(o + j, v[j]);
  }
}

__device__ __forceinline__ int nvalid16(int n0, int N) { const int r = N - n0; return r < 0 ? 0 : (r > 16 ? 16 : r); }

// ---- warp-cooperative block I/O through the per-warp staging buffer (Params::stage_out) -------------------------------
// A warp owns a 32-row x 32-column block of the output, lane = row.  Written straight from registers every store
// instruction touches 32 different rows (16 bytes each: 32 L2 requests, half-filled sectors).  Staged through shared
// memory (XOR-swizzled 16-byte chunks, conflict free both ways) each instruction instead covers whole 64-byte (bf16) or
// 128-byte (fp32) row segments.  `my_row_off` = element offset of this lane's row in the destination (any per-row
// mapping: frames, flattened frames, padded pitches); `ncols` = valid columns of the block (0..32).
constexpr unsigned FULL = 0xffffffffu;

template <bool BF16>
__device__ __forceinline__ void warp_store_block32(void* out, long long my_row_off, int n0, const float (&v0)[16],
                                                   const float (&v1)[16], int ncols, bool row_ok, uint8_t* stage) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
  if (BF16) {
    uint4 c[4];
    c[0] = make_uint4(pack_bf16x2(v0[0], v0[1]), pack_bf16x2(v0[2], v0[3]), pack_bf16x2(v0[4], v0[5]), pack_bf16x2(v0[6], v0[7]));
    c[1] = make_uint4(pack_bf16x2(v0[8], v0[9]), pack_bf16x2(v0[10], v0[11]), pack_bf16x2(v0[12], v0[13]), pack_bf16x2(v0[14], v0[15]));
    c[2] = make_uint4(pack_bf16x2(v1[0], v1[1]), pack_bf16x2(v1[2], v1[3]), pack_bf16x2(v1[4], v1[5]), pack_bf16x2(v1[6], v1[7]));
    c[3] = make_uint4(pack_bf16x2(v1[8], v1[9]), pack_bf16x2(v1[10], v1[11]), pack_bf16x2(v1[12], v1[13]), pack_bf16x2(v1[14], v1[15]));
    const int f = (lane >> 1) & 3;
#pragma unroll
    for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(stage + lane * 64 + ((j ^ f) << 4)) = c[j];
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = 8 * i + (lane >> 2), ch = lane & 3;
      const uint4 val = *reinterpret_cast<const uint4*>(stage + r * 64 + ((ch ^ ((r >> 1) & 3)) << 4));
      const long long roff = __shfl_sync(FULL, my_row_off, r);
      const int rok = __shfl_sync(FULL, (int)row_ok, r);
      const int col = ch * 8;
      if (rok && col < ncols) {
        bf16* dst = reinterpret_cast<bf16*>(out) + roff + n0 + col;
        if (col + 8 <= ncols && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
          *reinterpret_cast<uint4*>(dst) = val;
        } else {
          const uint32_t w[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (col + e < ncols) {
              const unsigned short h = (unsigned short)((w[e >> 1] >> ((e & 1) * 16)) & 0xffffu);
              reinterpret_cast<unsigned short*>(dst)[e] = h;
            }
        }
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      *reinterpret_cast<float4*>(stage + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(v0[4 * j], v0[4 * j + 1], v0[4 * j + 2], v0[4 * j + 3]);
      *reinterpret_cast<float4*>(stage + lane * 128 + (((j + 4) ^ (lane & 7)) << 4)) = make_float4(v1[4 * j], v1[4 * j + 1], v1[4 * j + 2], v1[4 * j + 3]);
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = 4 * i + (lane >> 3), ch = lane & 7;
      const float4 val = *reinterpret_cast<const float4*>(stage + r * 128 + ((ch ^ (r & 7)) << 4));
      const long long roff = __shfl_sync(FULL, my_row_off, r);
      const int rok = __shfl_sync(FULL, (int)row_ok, r);
      const int col = ch * 4;
      if (rok && col < ncols) {
        float* dst = reinterpret_cast<float*>(out) + roff + n0 + col;
        if (col + 4 <= ncols && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
          *reinterpret_cast<float4*>(dst) = val;
        } else {
          const float w[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (col + e < ncols) dst[e] = w[e];
        }
      }
    }
  }
}

// the reverse for a bf16 source: r0|r1 receive this lane's 32 columns (0 where invalid)
__device__ __forceinline__ void warp_load_block32_bf16(const bf16* src, long long my_row_off, int n0, int ncols, bool row_ok,
                                                       uint8_t* stage, float (&r0)[16], float (&r1)[16]) {
  const int lane = threadIdx.x & 31;
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = 8 * i + (lane >> 2), ch = lane & 3;
    const long long roff = __shfl_sync(FULL, my_row_off, r);
    const int rok = __shfl_sync(FULL, (int)row_ok, r);
    const int col = ch * 8;
    uint4 val = make_uint4(0u, 0u, 0u, 0u);
    if (rok && col < ncols) {
      const bf16* p = src + roff + n0 + col;
      if (col + 8 <= ncols && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        val = *reinterpret_cast<const uint4*>(p);
      } else {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int e = 0; e < 8; ++e)
          if (col + e < ncols) w[e >> 1] |= (uint32_t)reinterpret_cast<const unsigned short*>(p)[e] << ((e & 1) * 16);
        val = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    *reinterpret_cast<uint4*>(stage + r * 64 + ((ch ^ ((r >> 1) & 3)) << 4)) = val;
  }
  __syncwarp();
  const int f = (lane >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint4 u = *reinterpret_cast<const uint4*>(stage + lane * 64 + ((j ^ f) << 4));
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&w[q]);
      if (j < 2) { r0[8 * j + 2 * q] = __bfloat162float(t.x); r0[8 * j + 2 * q + 1] = __bfloat162float(t.y); }
      else { r1[8 * (j - 2) + 2 * q] = __bfloat162float(t.x); r1[8 * (j - 2) + 2 * q + 1] = __bfloat162float(t.y); }
    }
  }
}

// Per-frame column sums of a warp's 16-column chunk when the lanes' rows may belong to different frames (flattened M
// extent): one masked 16-shuffle reduction per frame present in the warp (one when the slab lies inside a frame).
// f(bb, col_of_lane, sum, owner) is called for every frame bb with the owning lanes' totals.
template <typename F>
__device__ __forceinline__ void segmented_colsum16(const float (&v)[16], int b, bool row_ok, int n0, int N, F&& f) {
  const int lane = threadIdx.x & 31;
  const int b_lo = __reduce_min_sync(FULL, row_ok ? b : 0x7fffffff);
  const int b_hi = __reduce_max_sync(FULL, row_ok ? b : -1);
  const int col = n0 + colsum16_owner_col(lane);
  const bool owner = ((lane & 1) == 0) && col < N;
  for (int bb = b_lo; bb <= b_hi; ++bb) {
    float w[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) w[j] = (row_ok && b == bb) ? v[j] : 0.f;
    const float s = warp_colsum16(w);
    f(bb, col, s, owner);
  }
}

// out[b, m, n] = act(alpha*acc + bias_m[m] + bias_n[n])  (+ per-frame column sums, + global fp64 column moments)
template <bool OUT_BF16, int ACT, bool COLSUM, bool STATS>
struct EpiStore {
  static constexpr bool kHasChunk32 = true;
  static constexpr int kStageBytes = OUT_BF16 ? 2048 : 4096;
  static constexpr bool kRowIndependent = true;
  static constexpr bool kBlockStats = STATS;
  void* out; long long ld, sb;
  const float* bias_m; const float* bias_n;
  float alpha;
  float* colsum; int ncols_total;
  double* stat_sum; double* stat_sq;
  float* blk;                 // CTA-level partial moments in shared memory ([2][N] floats), set by the kernel (may be NULL)
  __device__ __forceinline__ void flush_block(const float* b, int N, int tid, int nthreads) const {
    for (int c = tid; c < N; c += nthreads) {
      atomicAdd(&stat_sum[c], (double)b[c]);
      atomicAdd(&stat_sq[c], (double)b[N + c]);
    }
  }
  const float* bias_m_coln;   // non-NULL: the row bias enters as the rank-1 term bias_m[m] * bias_m_coln[n] (DGSCT_FC_FIRST)
  __device__ __forceinline__ void begin(int, int, bool) {}
  __device__ __forceinline__ void end(int, int, bool) {}
  __device__ __forceinline__ void transform(int m, int n0, float (&v)[16], bool row_ok, int nvalid) const {
    float bm = (bias_m && row_ok) ? bias_m[m] : 0.f;
    if (bias_m_coln) {
      float sn[16];
      load16_guard(bias_m_coln + n0, sn, nvalid);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaf(bm, sn[j], alpha * v[j]);
      bm = 0.f;
      if (bias_n) {
        float bn[16];
        load16_guard(bias_n + n0, bn, nvalid);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] += bn[j];
      }
    } else
    if (bias_n) {
      float bn[16];
      load16_guard(bias_n + n0, bn, nvalid);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaf(alpha, v[j], bn[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] *= alpha;
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float x = v[j] + bm;
      if (ACT == ACT_RELU) x = fmaxf(x, 0.f);
      if (ACT == ACT_SIGMOID) x = __frcp_rn(1.f + __expf(-x));
      v[j] = x;
    }
  }
  __device__ __forceinline__ void reduce(int b, int n0, float (&v)[16], bool row_ok, int nvalid, int N) const {
    if (COLSUM || STATS) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = (row_ok && j < nvalid) ? v[j] : 0.f;
      if (COLSUM) {
        float* cs = colsum; const int nt = ncols_total;
        segmented_colsum16(v, b, row_ok, n0, N, [cs, nt](int bb, int col, float s, bool owner) {
          if (owner) atomicAdd(&cs[(long long)bb * nt + col], s);
        });
      }
      if (STATS) {
        float q[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) q[j] = v[j] * v[j];
        const float s = warp_colsum16(v);
        const float s2 = warp_colsum16(q);
        const int lane = threadIdx.x & 31;
        const int col = n0 + colsum16_owner_col(lane);
        if (((lane & 1) == 0) && col < N) {
          if (blk) { atomicAdd(&blk[col], s); atomicAdd(&blk[N + col], s2); }
          else { atomicAdd(&stat_sum[col], (double)s); atomicAdd(&stat_sq[col], (double)s2); }
        }
      }
    }
  }
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
    const int nvalid = nvalid16(n0, N);
    transform(m, n0, v, row_ok, nvalid);
    if (out && row_ok && nvalid > 0) store_row16<OUT_BF16>(out, (size_t)((long long)b * sb + (long long)m * ld + n0), v, nvalid);
    reduce(b, n0, v, row_ok, nvalid, N);
  }
  // 32 columns at once through the warp's staging buffer (v1 holds columns n0+16.. when `two`)
  __device__ __forceinline__ void chunk32(int b, int m, int n0, float (&v0)[16], float (&v1)[16], bool two, bool row_ok,
                                          int N, uint8_t* stage) {
    const int nv0 = nvalid16(n0, N), nv1 = two ? nvalid16(n0 + 16, N) : 0;
    transform(m, n0, v0, row_ok, nv0);
    if (two) transform(m, n0 + 16, v1, row_ok, nv1);
    if (out) warp_store_block32<OUT_BF16>(out, (long long)b * sb + (long long)m * ld, n0, v0, v1, nv0 + nv1, row_ok, stage);
    reduce(b, n0, v0, row_ok, nv0, N);
    if (two) reduce(b, n0 + 16, v1, row_ok, nv1, N);
  }
};

// latent-token update written token-major: accumulator is D(c, k); out[b, k, c] = acc + tok[k, c]  (bf16 + fp32)
struct EpiTokUpdate {
  bf16* out; float* out32; const float* tok; int C, K;
  __device__ __forceinline__ void begin(int, int, bool) {}
  __device__ __forceinline__ void end(int, int, bool) {}
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
    if (!row_ok) return;
    const int nvalid = nvalid16(n0, N);
    const long long base = (long long)b * K * C + m;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      if (j < nvalid) {
        const long long idx = base + (long long)(n0 + j) * C;      // lanes = consecutive c: coalesced
        const float x = v[j] + tok[(long long)(n0 + j) * C + m];
        out[idx] = __float2bfloat16_rn(x);
        out32[idx] = x;
      }
    }
  }
};

// out[b,m,n] = resid[b,m,n] + gate[0]*acc      (bf16 in / bf16 out)
struct EpiResidual {
  // measured (bench.py stage table, DGSCT_NO_STAGE_OUT A/B): staging the residual LOAD costs more than the row-contiguous
  // accesses save (x_update 2.33 ms staged vs 1.79 ms direct over the 48 calls), so this functor keeps the direct path;
  // pure stores keep the staged one (up_sampler fp32: 1.27 vs 1.74 ms)
  static constexpr bool kHasChunk32 = false;
  static constexpr bool kHasPair = true;
  static constexpr int kStageBytes = 2048;
  static constexpr bool kRowIndependent = true;
  bf16* out; const bf16* resid; long long ld, sb; const float* gate; float g;
  // both halves of a 32-column block: the four 128-bit residual loads are in flight together
  __device__ __forceinline__ void pair(int b, int m, int n0, float (&v0)[16], float (&v1)[16], bool two, bool row_ok, int N) {
    const int nv0 = nvalid16(n0, N), nv1 = two ? nvalid16(n0 + 16, N) : 0;
    const size_t idx = (size_t)((long long)b * sb + (long long)m * ld + n0);
    if (row_ok && nv0 == 16 && nv1 == 16 && ((idx * 2) & 15) == 0) {
      const Raw16 r0 = ldg_raw16(resid + idx), r1 = ldg_raw16(resid + idx + 16);
#pragma unroll
      for (int j = 0; j < 16; ++j) { v0[j] *= g; v1[j] *= g; }
      add_raw16(r0, v0); add_raw16(r1, v1);
      store_row16<true>(out, idx, v0, 16);
      store_row16<true>(out, idx + 16, v1, 16);
      return;
    }
    chunk(b, m, n0, v0, row_ok, N);
    if (two) chunk(b, m, n0 + 16, v1, row_ok, N);
  }
  __device__ __forceinline__ void begin(int, int, bool) { g = gate[0]; }
  __device__ __forceinline__ void end(int, int, bool) {}
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
    const int nvalid = nvalid16(n0, N);
    if (!row_ok || nvalid <= 0) return;
    const size_t idx = (size_t)((long long)b * sb + (long long)m * ld + n0);
    if (nvalid == 16 && ((idx * 2) & 15) == 0) {
      const uint4* r = reinterpret_cast<const uint4*>(resid + idx);
      const uint4 ra = r[0], rb = r[1];
      const uint32_t w[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
        v[2 * j] = fmaf(g, v[2 * j], __bfloat162float(t.x));
        v[2 * j + 1] = fmaf(g, v[2 * j + 1], __bfloat162float(t.y));
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nvalid) v[j] = fmaf(g, v[j], __bfloat162float(resid[idx + j]));
    }
    store_row16<true>(out, idx, v, nvalid);
  }
  __device__ __forceinline__ void chunk32(int b, int m, int n0, float (&v0)[16], float (&v1)[16], bool two, bool row_ok,
                                          int N, uint8_t* stage) {
    const int ncols = nvalid16(n0, N) + (two ? nvalid16(n0 + 16, N) : 0);
    const long long off = (long long)b * sb + (long long)m * ld;
    float r0[16], r1[16];
    warp_load_block32_bf16(resid, off, n0, ncols, row_ok, stage, r0, r1);
#pragma unroll
    for (int j = 0; j < 16; ++j) { v0[j] = fmaf(g, v0[j], r0[j]); v1[j] = fmaf(g, v1[j], r1[j]); }
    warp_store_block32<true>(out, off, n0, v0, v1, ncols, row_ok, stage);
  }
};

// colsum[b,n] += mul[b,n] * sum_m relu(acc + bias[n])
struct EpiReluMulColsum {
  static constexpr bool kRowIndependent = true;
  const float* bias; const float* mul; float* colsum; int ncols; int mul_ld;
  __device__ __forceinline__ void begin(int, int, bool) {}
  __device__ __forceinline__ void end(int, int, bool) {}
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
    const int nvalid = nvalid16(n0, N);
    float bn[16];
    load16_guard(bias + n0, bn, nvalid);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = (row_ok && j < nvalid) ? fmaxf(v[j] + bn[j], 0.f) : 0.f;
    float* cs = colsum; const float* mu = mul; const int nc = ncols, mld = mul_ld;
    segmented_colsum16(v, b, row_ok, n0, N, [cs, mu, nc, mld](int bb, int col, float s, bool owner) {
      if (owner) atomicAdd(&cs[(long long)bb * nc + col], mu ? s * mu[(long long)bb * mld + col] : s);
    });
  }
};

// rowsum[b,m] += sum_n relu(acc + bias[n]) * mul[b,n] * w[n]
struct EpiSpatialRowsum {
  static constexpr bool kRowIndependent = true;
  const float* bias; const float* mul; const float* w; float* rowsum; int mul_ld; int nrows; float acc;
  __device__ __forceinline__ void begin(int, int, bool) { acc = 0.f; }
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
    const int nvalid = nvalid16(n0, N);
    float bn[16], f[16], ww[16];
    load16_guard(bias + n0, bn, nvalid);
    load16_guard(mul + (long long)b * mul_ld + n0, f, nvalid);
    load16_guard(w + n0, ww, nvalid);
#pragma unroll
    for (int j = 0; j < 16; ++j) acc = fmaf(fmaxf(v[j] + bn[j], 0.f), f[j] * ww[j], acc);
  }
  __device__ __forceinline__ void end(int b, int m, bool row_ok) {
    if (row_ok) atomicAdd(&rowsum[(long long)b * nrows + m], acc);
  }
};

// ---- backward epilogue that absorbs the row kernel which used to follow its GEMM (round 2) -----------------------------------
// (Measured on B200, scripts/gpu_r2w.sh: this one saves 13-16 us per call.  The same treatment of the spatial gate's dXc product
//  -- read-modify-write of dX1, a load of X1 and per-frame column sums in the epilogue -- cost 4-29 us MORE per call than the
//  separate bwd_xc_kernel and was removed again: a TMEM-drain epilogue is a poor place for three extra memory streams.)
// Channel gate: z1 = X1 . Wv1^T is only needed as a ReLU mask.  out[b,m,n] = (acc + bias[n] > 0) ? bc[b,n] : 0 (bf16) is dz1
// directly (bwd_channel_mask_kernel read and rewrote the whole (BT,N,C) tensor); d bias[n] += column sums through the CTA's
// shared-memory partials (one global atomic per column per CTA).
struct EpiMaskBcast {
  static constexpr bool kRowIndependent = true;
  static constexpr bool kBlockStats = true;
  bf16* out; long long ld, sb; const float* bias; const float* bc; int bc_ld; float* dbias; float* blk;
  __device__ __forceinline__ void flush_block(const float* b, int N, int tid, int nthreads) const {
    if (!dbias) return;
    for (int c = tid; c < N; c += nthreads) {
      const float s = b[c];
      if (s != 0.f) atomicAdd(&dbias[c], s);
    }
  }
  __device__ __forceinline__ void begin(int, int, bool) {}
  __device__ __forceinline__ void end(int, int, bool) {}
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
    const int nvalid = nvalid16(n0, N);
    float bn[16], d[16];
    load16_guard(bias + n0, bn, nvalid);
    load16_guard(bc + (long long)(row_ok ? b : 0) * bc_ld + n0, d, nvalid);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = (row_ok && j < nvalid && v[j] + bn[j] > 0.f) ? d[j] : 0.f;
    if (row_ok && nvalid > 0) store_row16<true>(out, (size_t)((long long)b * sb + (long long)m * ld + n0), v, nvalid);
    if (dbias) {
      const float s = warp_colsum16(v);
      const int lane = threadIdx.x & 31;
      const int col = n0 + colsum16_owner_col(lane);
      if (((lane & 1) == 0) && col < N) {
        if (blk) atomicAdd(&blk[col], s);
        else atomicAdd(&dbias[col], s);
      }
    }
  }
};

// Bottleneck: dZr = dUo . Wup must pass the ReLU of bn1 and feed bn1's backward sums.  out[b,m,n] = (z sc[n] + sh[n] > 0) ?
// acc : 0 (bf16), colA[n] += out, colB[n] += out z  (z = the forward's fp32 pre-BN activation) -- relu_mask_kernel and
// col_reduce_kernel each made a pass over the (BT,N,C/r) tensors for this.
struct EpiReluMaskStats {
  static constexpr bool kRowIndependent = true;
  static constexpr bool kBlockStats = true;
  bf16* out; long long ld, sb; const float* z; long long zld, zsb; const float* sc; const float* sh; float* colA; float* colB; float* blk;
  __device__ __forceinline__ void flush_block(const float* b, int N, int tid, int nthreads) const {
    if (!colA) return;
    for (int c = tid; c < N; c += nthreads) {
      if (b[c] != 0.f) atomicAdd(&colA[c], b[c]);
      if (b[N + c] != 0.f) atomicAdd(&colB[c], b[N + c]);
    }
  }
  __device__ __forceinline__ void begin(int, int, bool) {}
  __device__ __forceinline__ void end(int, int, bool) {}
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
    const int nvalid = nvalid16(n0, N);
    float zz[16];
    {
      float s1[16], s2[16];
      load16_guard(sc + n0, s1, nvalid);
      load16_guard(sh + n0, s2, nvalid);
      load16_guard(z + (long long)(row_ok ? b : 0) * zsb + (long long)(row_ok ? m : 0) * zld + n0, zz, nvalid);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = (row_ok && j < nvalid && fmaf(zz[j], s1[j], s2[j]) > 0.f) ? v[j] : 0.f;
    }
    if (row_ok && nvalid > 0) store_row16<true>(out, (size_t)((long long)b * sb + (long long)m * ld + n0), v, nvalid);
    if (colA) {
#pragma unroll
      for (int j = 0; j < 16; ++j) zz[j] *= v[j];
      const float sa = warp_colsum16(v);
      const float sbz = warp_colsum16(zz);
      const int lane = threadIdx.x & 31;
      const int col = n0 + colsum16_owner_col(lane);
      if (((lane & 1) == 0) && col < N) {
        if (blk) { atomicAdd(&blk[col], sa); atomicAdd(&blk[N + col], sbz); }
        else { atomicAdd(&colA[col], sa); atomicAdd(&colB[col], sbz); }
      }
    }
  }
};

// Q[b,m,0:32] = softmax_k(acc[0:32])    (N == 32, one N-block), bf16
struct EpiRowSoftmax32 {
  static constexpr bool kWholeRow = true;
  static constexpr bool kRowIndependent = true;
  bf16* out; long long ld, sb; float lo[16];
  __device__ __forceinline__ void begin(int, int, bool) {}
  __device__ __forceinline__ void end(int, int, bool) {}
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
    if (n0 == 0) {
#pragma unroll
      for (int j = 0; j < 16; ++j) lo[j] = v[j];
      return;
    }
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 16; ++j) mx = fmaxf(mx, fmaxf(lo[j], v[j]));
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) { lo[j] = __expf(lo[j] - mx); v[j] = __expf(v[j] - mx); s += lo[j] + v[j]; }
    const float inv = __frcp_rn(s);
#pragma unroll
    for (int j = 0; j < 16; ++j) { lo[j] *= inv; v[j] *= inv; }
    if (row_ok) {
      const size_t idx = (size_t)((long long)b * sb + (long long)m * ld);
      store_row16<true>(out, idx, lo, 16);
      store_row16<true>(out, idx + 16, v, 16);
    }
  }
};

// zr[b,m,n] = relu(acc*scale[n] + shift[n]) (bf16) ; optional fp32 tap of acc        (bn1 + ReLU folded)
struct EpiAffineRelu {
  static constexpr bool kHasChunk32 = true;
  static constexpr int kStageBytes = 2048;
  static constexpr bool kRowIndependent = true;
  bf16* zr; long long ldz; int rows_per_batch; const float* scale; const float* shift; float* ztap; int Cd;
  __device__ __forceinline__ void begin(int, int, bool) {}
  __device__ __forceinline__ void end(int, int, bool) {}
  __device__ __forceinline__ void transform(long long row, int n0, float (&v)[16], bool row_ok, int nvalid) const {
    if (ztap && row_ok) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j < nvalid) ztap[row * Cd + n0 + j] = v[j];
    }
    float sc[16], sh[16];
    load16_guard(scale + n0, sc, nvalid);
    load16_guard(shift + n0, sh, nvalid);
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = fmaxf(fmaf(v[j], sc[j], sh[j]), 0.f);
  }
  __device__ __forceinline__ void chunk(int b, int m, int n0, float (&v)[16], bool row_ok, int N) {
    const int nvalid = nvalid16(n0, N);
    if (!row_ok || nvalid <= 0) return;
    const long long row = (long long)b * rows_per_batch + m;
    transform(row, n0, v, row_ok, nvalid);
    store_row16<true>(zr, (size_t)(row * ldz + n0), v, nvalid);
  }
  __device__ __forceinline__ void chunk32(int b, int m, int n0, float (&v0)[16], float (&v1)[16], bool two, bool row_ok,
                                          int N, uint8_t* stage) {
    const int nv0 = nvalid16(n0, N), nv1 = two ? nvalid16(n0 + 16, N) : 0;
    const long long row = (long long)b * rows_per_batch + m;
    transform(row, n0, v0, row_ok, nv0);
    if (two) transform(row, n0 + 16, v1, row_ok, nv1);
    warp_store_block32<true>(zr, row * ldz, n0, v0, v1, nv0 + nv1, row_ok, stage);
  }
};

}  // namespace tc
}  // namespace dgsct
