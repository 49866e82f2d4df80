// Fused latent-token attention (reference: DG-SCT/AVE/nets/net_trans.py:572-589), bf16 I/O, fp32 statistics.
//
//   P  = softmax_n(T V^T)        T' = T + P V              (tokens attend over the resampled other-modality tokens V)
//   Q  = softmax_k(X T'^T)       X1 = X + gate_av * Q T'    (the adapted stream attends over the updated latent tokens)
//
// Round 1/2 ran this as six launches (two products + a split column softmax, two products + a row softmax) with the logits,
// P and Q making HBM round trips.  Here it is two kernels, each reading its token tile ONCE:
//
//   tok_attn_fwd_kernel   grid (splits, BT, column halves).  A CTA walks chunks of R rows of V[b]: the chunk is staged in shared
//                         memory by bulk-async copies (one per row, padded pitch), S = V_chunk T^T on warp-level tensor-core
//                         MMAs, flash-style running max / sum over n, O += P~^T V_chunk with the accumulator in registers.
//                         Partials (m, l, O) go to global memory; the LAST CTA of a frame (ticket) combines them into T'.
//   x_attn_fwd_kernel     grid (row-tile groups, BT).  A warp owns 16 rows of X[b]: S = X_tile T'^T, softmax over the K latent
//                         tokens in registers (quad shuffles), Q stays in registers as the A operand of Q T', the residual is
//                         added in place in the staged tile, which leaves through one bulk-async store per row.
//
// The tensor-core work per byte is tiny (64 flop/B): warp-level mma.sync is enough to stay under the HBM roofline, so the
// kernels are organised around the memory system (bulk copies, one pass over every tile), not around tcgen05.
#pragma once
#include "gemm_tc.cuh"

namespace dgsct {
namespace fa {

constexpr int KL = 32;          // latent tokens handled per tile (K <= 32; rows >= K are zero / masked)
constexpr int PS_PITCH = 80;    // bytes per row of the P~ staging tile: 32 bf16 + 16 B pad (odd multiple of 16: conflict-free ldmatrix)

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
// D(16x8, fp32) += A(16x16, bf16, row) . B(16x8, bf16, col)
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}
// bulk-async copies (TMA unit, 1-D): global -> shared with mbarrier completion; shared -> global in a bulk group
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(reinterpret_cast<uint64_t>(dst)), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// unbounded-looking waits are bounded like the GEMM's (sticky error flag, no hang)
__device__ __forceinline__ bool wait_bar(uint32_t bar, uint32_t parity, int* err_flag) { return tc::mbar_wait(bar, parity, err_flag); }

// S(16 x 32) = A_tile(16 rows x C, row-major, `pitch` bytes) . T^T over channels [k_begin, k_end) (multiples of 16), T = (32 x C,
// row-major, same pitch).  acc[nt] covers latent
// tokens nt*8 .. nt*8+7; thread (g = lane>>2, t = lane&3) holds rows g / g+8, columns nt*8 + 2t, +1.
__device__ __forceinline__ void rows_times_tokT(float (&acc)[4][4], uint32_t a_tile, uint32_t t_base, int pitch, int k_begin, int k_end, int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const uint32_t a_addr = a_tile + (uint32_t)((lane & 15) * pitch + (lane >> 4) * 16);
  const uint32_t b_addr = t_base + (uint32_t)(((lane & 7) + ((lane >> 4) << 3)) * pitch + ((lane >> 3) & 1) * 16);
  const uint32_t b_addr2 = b_addr + (uint32_t)(16 * pitch);
#pragma unroll 2
  for (int kk = k_begin; kk < k_end; kk += 16) {
    uint32_t a[4], b[4], c[4];
    ldsm_x4(a_addr + kk * 2, a[0], a[1], a[2], a[3]);
    ldsm_x4(b_addr + kk * 2, b[0], b[1], b[2], b[3]);
    ldsm_x4(b_addr2 + kk * 2, c[0], c[1], c[2], c[3]);
    mma16816(acc[0], a, b[0], b[1]);
    mma16816(acc[1], a, b[2], b[3]);
    mma16816(acc[2], a, c[0], c[1]);
    mma16816(acc[3], a, c[2], c[3]);
  }
}

// A fragments (bf16) of a 16 x 32 fp32 tile held in the accumulator layout above: k-step ks covers columns ks*16 .. +15
__device__ __forceinline__ void acc_to_afrag(const float (&acc)[4][4], uint32_t (&a)[2][4]) {
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) {
    a[ks][0] = pack_bf16x2(acc[2 * ks][0], acc[2 * ks][1]);
    a[ks][1] = pack_bf16x2(acc[2 * ks][2], acc[2 * ks][3]);
    a[ks][2] = pack_bf16x2(acc[2 * ks + 1][0], acc[2 * ks + 1][1]);
    a[ks][3] = pack_bf16x2(acc[2 * ks + 1][2], acc[2 * ks + 1][3]);
  }
}

// zero `rows` rows of `row_bytes` bytes (16-B multiple) starting at shared address base (pitch bytes apart); nthr threads cooperate
__device__ __forceinline__ void zero_rows(uint8_t* base, int rows, int pitch, int row_bytes, int tid, int nthr) {
  const int v = row_bytes >> 4;
  for (int i = tid; i < rows * v; i += nthr) {
    const int r = i / v, c = i - r * v;
    *reinterpret_cast<uint4*>(base + (size_t)r * pitch + c * 16) = make_uint4(0, 0, 0, 0);
  }
}

// =====================================================================================================================
// Kernel A: tokens attend over V
// =====================================================================================================================
struct TokAttnArgs {
  const bf16* V;            // (BT, N, C)
  const bf16* tokb;         // (K, ld_tok) bf16 latent tokens (packed weights)
  const float* tok32;       // (K, C) fp32 master
  bf16* Pb;                 // (BT, N, Kp) out: P~ relative to the chunk's running max (x_attn rescales rows to the final P); may be NULL
  float* part_o;            // (BT, splits, 32, C) fp32 partial outputs
  float* part_ml;           // (BT, splits, 2, 32) running max / sum of each split
  float* mref;              // (BT, nchunks, 32)   running max each chunk's P~ refers to
  float* pfac;              // (BT, nchunks, 32)   out: exp(mref - M) / L
  float* stat_ml;           // (BT, 2, 32)         out: final max M and sum L of every latent token's softmax over n (backward)
  unsigned* ticket;         // (BT) zero on entry
  float* Tp;                // (BT, K, C) fp32 out
  bf16* TpB;                // (BT, K, C) bf16 out
  int* err_flag;
  int N, C, K, Kp, ld_tok, R, nchunks, splits, zcols, nbuf;
};

template <int NTW>
__global__ void __launch_bounds__(256) tok_attn_fwd_kernel(const TokAttnArgs a) {
  extern __shared__ __align__(128) uint8_t fa_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int C = a.C, N = a.N, K = a.K, R = a.R;
  const int pitch = C * 2 + 16;
  const int b = blockIdx.y, split = blockIdx.x, z = blockIdx.z;
  uint8_t* Tsm = fa_smem;
  uint8_t* Vsm = Tsm + (size_t)KL * pitch;
  uint8_t* Psm = Vsm + (size_t)a.nbuf * R * pitch;
  float* wmax = reinterpret_cast<float*>(Psm + (size_t)R * PS_PITCH);   // [8][32]
  float* wsum = wmax + 8 * KL;                                          // [8][32]
  float* m_run = wsum + 8 * KL;                                         // [32]
  float* l_run = m_run + KL;
  float* scl = l_run + KL;
  float* mnew = scl + KL;
  float* fch = mnew + KL;                                               // exp(chunk max - new running max)
  float* coef = fch + KL;                                               // combine: [splits <= 32][32]
  float* sred = coef + 32 * KL;                                         // [16][8 warps][32 lanes] partial logits (K-split pass 1)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sred + 16 * 8 * 32);     // [0] tokens, [1..2] V buffers
  int* s_last = reinterpret_cast<int*>(bars + 4);

  const int c_begin = (int)((long long)split * a.nchunks / a.splits), c_end = (int)((long long)(split + 1) * a.nchunks / a.splits);
  const bf16* Vb = a.V + (size_t)b * N * C;
  const uint32_t row_bytes = (uint32_t)C * 2;

  if (tid == 0) {
    for (int i = 0; i < 3; ++i) tc::mbar_init(tc::smem_u32(&bars[i]), 1);
    tc::fence_barrier_init();
  }
  if (tid < KL) { m_run[tid] = -INFINITY; l_run[tid] = 0.f; }
  __syncthreads();
  // A bulk copy is issued by ONE lane at a time (the compiler serialises the per-lane copies of a warp through the uniform
  // datapath, ~80 cycles each: ncu, round 2) -- the rows of a chunk are therefore dealt out over all eight warps (row r to warp
  // r % 8), and one thread posts the byte count; copies that land before it only drive the transaction count negative.
  const int rlane = lane * 8 + warp;                    // the row this thread copies (if < rows)
  auto issue_chunk = [&](int chunk, int buf) {          // called by EVERY thread of the CTA
    const int n0 = chunk * R;
    const int rows = min(R, N - n0);
    const uint32_t bar = tc::smem_u32(&bars[1 + buf]);
    if (tid == 0) tc::mbar_expect_tx(bar, (uint32_t)rows * row_bytes);
    const uint32_t dst = tc::smem_u32(Vsm + (size_t)buf * R * pitch);
    if (rlane < rows) bulk_g2s(dst + (uint32_t)(rlane * pitch), Vb + (size_t)(n0 + rlane) * C, row_bytes, bar);
  };
  {
    const uint32_t bar = tc::smem_u32(&bars[0]);
    if (tid == 0) tc::mbar_expect_tx(bar, (uint32_t)K * row_bytes);
    if (lane < 4 && lane * 8 + warp < K)
      bulk_g2s(tc::smem_u32(Tsm) + (uint32_t)((lane * 8 + warp) * pitch), a.tokb + (size_t)(lane * 8 + warp) * a.ld_tok, row_bytes, bar);
    if (c_begin < c_end) issue_chunk(c_begin, 0);
  }
  if (K < KL) zero_rows(Tsm + (size_t)K * pitch, KL - K, pitch, (int)row_bytes, tid, 256);

  // this CTA's output columns [zc0, zc0 + zcols) and the warp's column pairs (16 columns each)
  const int zc0 = z * a.zcols;
  const int zcols = min(a.zcols, C - zc0);
  const int npairs = zcols >> 4;
  const int p_begin = warp * (NTW / 2);
  float o[2][NTW][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) o[mt][nt][j] = 0.f;

  bool ok = true;
  uint32_t ph[2] = {0, 0};
  // pass 1: R / 16 row tiles; with fewer than 8 of them the channel range of a tile is split over KS warps (partials summed
  // through shared memory by the tile's first warp, its "owner") so that a 1536-channel dot product is not one warp's chain
  const int pass1_warps = R >> 4;
  const int KS = 8 / pass1_warps;
  const int mtile = warp / KS, kpart = warp - mtile * KS;
  const bool owner = kpart == 0;
  const int ksteps = ((C >> 4) + KS - 1) / KS;
  const int k_begin = min(C, kpart * ksteps * 16), k_end = min(C, k_begin + ksteps * 16);
  for (int chunk = c_begin; chunk < c_end && ok; ++chunk) {
    const int it = chunk - c_begin;
    const int buf = a.nbuf == 2 ? (it & 1) : 0;
    const int n0 = chunk * R;
    const int rows = min(R, N - n0);
    uint8_t* Vc = Vsm + (size_t)buf * R * pitch;
    if (a.nbuf == 2 && chunk + 1 < c_end) issue_chunk(chunk + 1, buf ^ 1);   // (buffer released by the trailing sync)
    if (rows < R) zero_rows(Vc + (size_t)rows * pitch, R - rows, pitch, (int)row_bytes, tid, 256);
    if (it == 0) ok = wait_bar(tc::smem_u32(&bars[0]), 0, a.err_flag);
    if (ok) ok = wait_bar(tc::smem_u32(&bars[1 + buf]), ph[buf], a.err_flag);
    ph[buf] ^= 1;
    ok = __syncthreads_and(ok);
    if (!ok) break;
    // ---- pass 1: S = V_chunk . T^T, 16 rows per warp ----
    float s[4][4];
    rows_times_tokT(s, tc::smem_u32(Vc + (size_t)mtile * 16 * pitch), tc::smem_u32(Tsm), pitch, k_begin, k_end, lane);
    if (KS > 1) {
      if (!owner) {
#pragma unroll
        for (int i = 0; i < 16; ++i) sred[(i * 8 + warp) * 32 + lane] = s[i >> 2][i & 3];
      }
      __syncthreads();
      if (owner) {
        for (int j = 1; j < KS; ++j)
#pragma unroll
          for (int i = 0; i < 16; ++i) s[i >> 2][i & 3] += sred[(i * 8 + warp + j) * 32 + lane];
      }
    }
    if (owner) {
      const bool v0 = (mtile * 16 + g) < rows, v1 = (mtile * 16 + g + 8) < rows;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        if (!v0) { s[nt][0] = -INFINITY; s[nt][1] = -INFINITY; }
        if (!v1) { s[nt][2] = -INFINITY; s[nt][3] = -INFINITY; }
        float m0 = fmaxf(s[nt][0], s[nt][2]), m1 = fmaxf(s[nt][1], s[nt][3]);
#pragma unroll
        for (int off = 4; off < 32; off <<= 1) {
          m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, off));
          m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, off));
        }
        if (g == 0) { wmax[warp * KL + nt * 8 + 2 * t] = m0; wmax[warp * KL + nt * 8 + 2 * t + 1] = m1; }
      }
    } else {
      wmax[warp * KL + lane] = -INFINITY;
    }
    __syncthreads();
    if (tid < KL) {
      float mc = wmax[tid];
#pragma unroll
      for (int w = 1; w < 8; ++w) mc = fmaxf(mc, wmax[w * KL + tid]);
      // P~ of a chunk refers to the chunk's OWN maximum (its bf16 rounding is then independent of how the frame is split over
      // CTAs: a batch and its clips give the same bits); the running statistics carry the two rescale factors
      const float mo = m_run[tid];
      const float mn = fmaxf(mo, mc);
      scl[tid] = (mo == -INFINITY) ? 0.f : __expf(mo - mn);
      fch[tid] = __expf(mc - mn);
      m_run[tid] = mn;
      mnew[tid] = mc;
      if (z == 0) a.mref[((size_t)b * a.nchunks + chunk) * KL + tid] = mc;
    }
    __syncthreads();
    if (owner) {
      const int r0 = mtile * 16 + g;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = nt * 8 + 2 * t;
        const float mA = mnew[col], mB = mnew[col + 1];
        float p0 = __expf(s[nt][0] - mA), p1 = __expf(s[nt][1] - mB), p2 = __expf(s[nt][2] - mA), p3 = __expf(s[nt][3] - mB);
        if (col >= K) { p0 = 0.f; p2 = 0.f; }
        if (col + 1 >= K) { p1 = 0.f; p3 = 0.f; }
        const uint32_t u0 = pack_bf16x2(p0, p1), u1 = pack_bf16x2(p2, p3);
        *reinterpret_cast<uint32_t*>(Psm + (size_t)r0 * PS_PITCH + col * 2) = u0;
        *reinterpret_cast<uint32_t*>(Psm + (size_t)(r0 + 8) * PS_PITCH + col * 2) = u1;
        if (a.Pb && z == 0 && col < a.Kp) {
          if (r0 < rows) *reinterpret_cast<uint32_t*>(a.Pb + ((size_t)b * N + n0 + r0) * a.Kp + col) = u0;
          if (r0 + 8 < rows) *reinterpret_cast<uint32_t*>(a.Pb + ((size_t)b * N + n0 + r0 + 8) * a.Kp + col) = u1;
        }
        float c0 = p0 + p2, c1 = p1 + p3;
#pragma unroll
        for (int off = 4; off < 32; off <<= 1) {
          c0 += __shfl_xor_sync(0xffffffffu, c0, off);
          c1 += __shfl_xor_sync(0xffffffffu, c1, off);
        }
        if (g == 0) { wsum[warp * KL + col] = c0; wsum[warp * KL + col + 1] = c1; }
      }
    } else {
      wsum[warp * KL + lane] = 0.f;
    }
    __syncthreads();
    if (tid < KL) {
      float ls = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) ls += wsum[w * KL + tid];
      l_run[tid] = l_run[tid] * scl[tid] + ls * fch[tid];
    }
    // ---- pass 2: O = O * scale + f_chunk * (P~^T . V_chunk)  (M = 32 latent tokens, N = this warp's columns, K-dim = rows) ----
    if (p_begin < npairs) {
      const float sc0 = scl[g], sc1 = scl[g + 8], sc2 = scl[16 + g], sc3 = scl[24 + g];
      const float f0 = fch[g], f1 = fch[g + 8], f2 = fch[16 + g], f3 = fch[24 + g];
      const uint32_t ps = tc::smem_u32(Psm);
      const uint32_t vs = tc::smem_u32(Vc);
      const int j = lane >> 3, r = lane & 7;
      const uint32_t a_off = (uint32_t)((r + (j >> 1) * 8) * PS_PITCH + (j & 1) * 16);
      const uint32_t b_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * pitch + (lane >> 4) * 16);
#pragma unroll
      for (int pi = 0; pi < NTW / 2; ++pi) {
        if (p_begin + pi < npairs) {
          float e[2][2][4];
#pragma unroll
          for (int i = 0; i < 16; ++i) e[i >> 3][(i >> 2) & 1][i & 3] = 0.f;
          const uint32_t vcol = vs + b_off + (uint32_t)((zc0 + (p_begin + pi) * 16) * 2);
          for (int ks = 0; ks < (R >> 4); ++ks) {
            uint32_t a0[4], a1[4], b0, b1, b2, b3;
            ldsm_x4_t(ps + a_off + (uint32_t)(ks * 16 * PS_PITCH), a0[0], a0[1], a0[2], a0[3]);
            ldsm_x4_t(ps + a_off + (uint32_t)(ks * 16 * PS_PITCH) + 32, a1[0], a1[1], a1[2], a1[3]);
            ldsm_x4_t(vcol + (uint32_t)(ks * 16 * pitch), b0, b1, b2, b3);
            mma16816(e[0][0], a0, b0, b1);
            mma16816(e[0][1], a0, b2, b3);
            mma16816(e[1][0], a1, b0, b1);
            mma16816(e[1][1], a1, b2, b3);
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            float (&o0)[4] = o[0][2 * pi + h];
            float (&o1)[4] = o[1][2 * pi + h];
            o0[0] = fmaf(f0, e[0][h][0], o0[0] * sc0); o0[1] = fmaf(f0, e[0][h][1], o0[1] * sc0);
            o0[2] = fmaf(f1, e[0][h][2], o0[2] * sc1); o0[3] = fmaf(f1, e[0][h][3], o0[3] * sc1);
            o1[0] = fmaf(f2, e[1][h][0], o1[0] * sc2); o1[1] = fmaf(f2, e[1][h][1], o1[1] * sc2);
            o1[2] = fmaf(f3, e[1][h][2], o1[2] * sc3); o1[3] = fmaf(f3, e[1][h][3], o1[3] * sc3);
          }
        }
      }
    }
    __syncthreads();                                   // V buffer, P~ tile and the per-warp statistics are free again
    if (a.nbuf == 1 && chunk + 1 < c_end) issue_chunk(chunk + 1, 0);
  }
  ok = __syncthreads_and(ok);
  if (!ok) return;
  if (a.splits == 1) {
    // ---- the whole frame went through this CTA: T' = T + O / l straight from the accumulators ----
#pragma unroll
    for (int pi = 0; pi < NTW / 2; ++pi) {
      if (p_begin + pi < npairs) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = zc0 + (p_begin + pi) * 16 + h * 8 + 2 * t;
          const int nt = 2 * pi + h;
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              const int k = mt * 16 + g + hh * 8;
              if (k < K) {
                const float inv = 1.f / l_run[k];
                const float* tk = a.tok32 + (size_t)k * C + c;          // (scalar loads: a parameter may sit at any 4-byte offset)
                const float v0 = fmaf(o[mt][nt][2 * hh], inv, tk[0]), v1 = fmaf(o[mt][nt][2 * hh + 1], inv, tk[1]);
                const size_t idx = ((size_t)b * K + k) * C + c;
                *reinterpret_cast<float2*>(a.Tp + idx) = make_float2(v0, v1);
                *reinterpret_cast<uint32_t*>(a.TpB + idx) = pack_bf16x2(v0, v1);
              }
            }
          }
        }
      }
    }
    if (z == 0) {
      if (tid < KL && a.stat_ml) { a.stat_ml[(size_t)b * 2 * KL + tid] = m_run[tid]; a.stat_ml[(size_t)b * 2 * KL + KL + tid] = l_run[tid]; }
      for (int i = tid; i < a.nchunks * KL; i += 256) {
        const int k = i & (KL - 1);
        const float l = l_run[k];
        a.pfac[(size_t)b * a.nchunks * KL + i] = l > 0.f ? __expf(a.mref[(size_t)b * a.nchunks * KL + i] - m_run[k]) / l : 0.f;
      }
    }
    return;
  }
  // ---- partial results of this (split, column half) ----
  {
    float* po = a.part_o + ((size_t)b * a.splits + split) * KL * C;
#pragma unroll
    for (int pi = 0; pi < NTW / 2; ++pi) {
      if (p_begin + pi < npairs) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int c = zc0 + (p_begin + pi) * 16 + h * 8 + 2 * t;
          const int nt = 2 * pi + h;
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            *reinterpret_cast<float2*>(po + (size_t)(mt * 16 + g) * C + c) = make_float2(o[mt][nt][0], o[mt][nt][1]);
            *reinterpret_cast<float2*>(po + (size_t)(mt * 16 + g + 8) * C + c) = make_float2(o[mt][nt][2], o[mt][nt][3]);
          }
        }
      }
    }
    if (z == 0 && tid < KL) {
      float* pm = a.part_ml + ((size_t)b * a.splits + split) * 2 * KL;
      pm[tid] = m_run[tid];
      pm[KL + tid] = l_run[tid];
    }
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned total = gridDim.x * gridDim.z;
    *s_last = (atomicAdd(&a.ticket[b], 1u) == total - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!*s_last) return;
  __threadfence();
  // ---- combine (last CTA of the frame): M = max m_s, L = sum l_s e^(m_s - M), T' = T + sum_s e^(m_s - M) O_s / L ----
  const float* pml = a.part_ml + (size_t)b * a.splits * 2 * KL;
  if (tid < KL) {
    float M = -INFINITY;
    for (int sidx = 0; sidx < a.splits; ++sidx) M = fmaxf(M, __ldcg(pml + sidx * 2 * KL + tid));
    float L = 0.f;
    for (int sidx = 0; sidx < a.splits; ++sidx) {
      const float ms = __ldcg(pml + sidx * 2 * KL + tid);
      L += (ms == -INFINITY) ? 0.f : __ldcg(pml + sidx * 2 * KL + KL + tid) * __expf(ms - M);
    }
    const float invL = L > 0.f ? 1.f / L : 0.f;
    for (int sidx = 0; sidx < a.splits; ++sidx) {
      const float ms = __ldcg(pml + sidx * 2 * KL + tid);
      coef[sidx * KL + tid] = (ms == -INFINITY) ? 0.f : __expf(ms - M) * invL;
    }
    mnew[tid] = M;
    scl[tid] = invL;
    if (a.stat_ml) { a.stat_ml[(size_t)b * 2 * KL + tid] = M; a.stat_ml[(size_t)b * 2 * KL + KL + tid] = L; }
  }
  __syncthreads();
  for (int i = tid; i < a.nchunks * KL; i += 256) {
    const int k = i & (KL - 1);
    const float mr = __ldcg(a.mref + (size_t)b * a.nchunks * KL + i);
    a.pfac[(size_t)b * a.nchunks * KL + i] = __expf(mr - mnew[k]) * scl[k];
  }
  const float* pob = a.part_o + (size_t)b * a.splits * KL * C;
  const int c4 = C >> 2;
  const int items = K * c4;
  for (int i0 = tid; i0 < items; i0 += 4 * 256) {          // four independent 16-byte items per thread: loads of all of them in flight
    float4 acc[4];
    int kk[4], cc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * 256;
      kk[u] = i < items ? i / c4 : 0;
      cc[u] = i < items ? (i - kk[u] * c4) * 4 : 0;
      const float* tk = a.tok32 + (size_t)kk[u] * C + cc[u];      // (scalar loads: a parameter may sit at any 4-byte offset)
      acc[u] = make_float4(tk[0], tk[1], tk[2], tk[3]);
    }
    for (int sidx = 0; sidx < a.splits; ++sidx) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldcg(reinterpret_cast<const float4*>(pob + ((size_t)sidx * KL + kk[u]) * C + cc[u]));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float w = coef[sidx * KL + kk[u]];
        acc[u].x = fmaf(w, v[u].x, acc[u].x); acc[u].y = fmaf(w, v[u].y, acc[u].y);
        acc[u].z = fmaf(w, v[u].z, acc[u].z); acc[u].w = fmaf(w, v[u].w, acc[u].w);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (i0 + u * 256 < items) {
        const size_t idx = ((size_t)b * K + kk[u]) * C + cc[u];
        *reinterpret_cast<float4*>(a.Tp + idx) = acc[u];
        *reinterpret_cast<uint2*>(a.TpB + idx) = make_uint2(pack_bf16x2(acc[u].x, acc[u].y), pack_bf16x2(acc[u].z, acc[u].w));
      }
    }
  }
}

// opt in to the largest dynamic shared-memory size the kernel can have (227 KB minus its static shared memory)
template <typename KernelT>
inline cudaError_t set_max_dyn_smem(KernelT kernel) {
  cudaFuncAttributes fa_{};
  cudaError_t e = cudaFuncGetAttributes(&fa_, kernel);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa_.sharedSizeBytes);
}

// Column ranges per frame (grid.z).  Every CTA of a frame runs the full-width logits product, but the products that are cut by
// output columns (P~^T V, dT', dx, dV, d tokens) split cleanly, and so do their accumulators: when a call has fewer CTAs than the
// machine has SMs (deep stages: one split per frame), the frame is dealt out over up to four column ranges.  The arithmetic of a
// column does not depend on the range it falls in, so this may follow the number of frames without touching a single bit.
// Chosen by a two-term cost model: CTAs needed / CTAs resident (whole waves) x work per CTA, where `full` parts of the work are
// repeated by every range and `sep` parts are divided among them.
inline void fa_widen_z(int BT, int splits, int C, size_t smem, int full, int sep, int& Z, int& zcols, int& ntw) {
  int best = Z;
  double best_cost = 0.0;
  for (int z = Z; z <= 4; ++z) {
    if (z > Z && (C / 16) / z < 4) break;                     // at least 64 columns per range
    const int zc = ((C / 16 + z - 1) / z) * 16;
    const int nt = 2 * ((zc / 16 + 7) / 8);
    int cps = (int)((227 * 1024) / (smem + 1024));
    if (cps > 2) cps = 2;
    if (nt >= 10 || cps < 1) cps = 1;                          // register budget: one CTA per SM for the widest accumulators
    const long long ctas = (long long)BT * splits * z;
    const long long waves = (ctas + 148LL * cps - 1) / (148LL * cps);
    const double cost = (double)waves * ((double)full + (double)sep / z);
    if (z == Z || cost < best_cost * 0.95) { best = z; best_cost = cost; }
  }
  Z = best;
  zcols = ((C / 16 + Z - 1) / Z) * 16;
  ntw = 2 * ((zcols / 16 + 7) / 8);
}

struct TokAttnPlan {
  bool ok; int R, nbuf, nchunks, splits, Z, zcols, ntw; size_t smem;
  size_t part_o_bytes, part_ml_bytes, mref_bytes, ticket_bytes;
};
inline size_t tok_attn_smem(int C, int R, int nbuf) {
  const size_t pitch = (size_t)C * 2 + 16;
  return (size_t)KL * pitch + (size_t)nbuf * R * pitch + (size_t)R * PS_PITCH + (size_t)(16 * KL + 5 * KL + 32 * KL + 16 * 8 * 32) * 4 + 64;
}
inline TokAttnPlan make_tok_attn_plan(int BT, int N, int C, int K) {
  TokAttnPlan p{};
  if (C % 32 || C < 32 || K > KL || K < 1 || N < 1) return p;
  p.Z = (C + 767) / 768;
  p.zcols = ((C / 16 + p.Z - 1) / p.Z) * 16;
  const int pairs = p.zcols / 16;
  const int ppw = (pairs + 7) / 8;
  p.ntw = 2 * ppw;
  if (p.ntw > 12) return p;
  // rows per chunk / V buffers: two CTAs per SM (<= 104 KB each) when the register budget allows two, else one big CTA
  // (order from a sweep on B200, profiles/r02_fused_attention_chunk_sweep.txt: the phases of a chunk are latency-bound, so more
  //  rows per chunk beat a second buffer -- C = 192: (128,1) 53.6 us vs (64,2) 63.7; C = 768: (64,1) 44.7 vs (32,2) 47.8)
  static const int cand_small[][2] = {{128, 2}, {128, 1}, {64, 2}, {64, 1}, {32, 2}, {32, 1}, {16, 1}};
  static const int cand_big[][2] = {{128, 1}, {64, 1}, {32, 1}, {16, 1}};
  const size_t small = 104 * 1024, big = 225 * 1024;
  int R = 0, nbuf = 0;
  // (the chunk size must not depend on the number of frames: it fixes which rows share a reference maximum, i.e. the bits of P~)
  if (p.ntw < 10)
    for (auto& cnd : cand_small) if (tok_attn_smem(C, cnd[0], cnd[1]) <= small) { R = cnd[0]; nbuf = cnd[1]; break; }
  if (!R) for (auto& cnd : cand_big) if (tok_attn_smem(C, cnd[0], cnd[1]) <= big) { R = cnd[0]; nbuf = cnd[1]; break; }
  if (!R) return p;
  while (R > 16 && R / 2 >= N) R /= 2;                       // short frames: no point staging empty rows
  p.R = R;
  p.nchunks = (N + R - 1) / R;
  // Splits of a frame: every split costs a 32 x C fp32 partial (written, then read by the combining CTA), i.e. 128 * splits / N
  // of the frame's own bytes -- one split (no partials, T' straight from the accumulators) up to ~570 rows, then one per 384
  // rows.  The count depends on the FRAME's shape only, never on how many frames a call holds: the association of the fp32
  // sums, hence every bit of T', is the same for a batch and for its clips (tests/test_gpu_zfullsize.py).
  (void)BT;
  {
    int sp = (N + 192) / 384;        // (sweep 256 .. 2304 rows per split on B200: flat around 384, profiles/r02_fused_attention_chunk_sweep.txt)
    if (sp < 1) sp = 1;
    if (sp > p.nchunks) sp = p.nchunks;
    if (sp > 32) sp = 32;
    p.splits = sp;
  }
  fa_widen_z(BT, p.splits, C, tok_attn_smem(C, p.R, nbuf), 1, 1, p.Z, p.zcols, p.ntw);
  p.nbuf = (nbuf == 2 && (p.nchunks + p.splits - 1) / p.splits > 1) ? 2 : 1;
  p.smem = tok_attn_smem(C, p.R, p.nbuf);
  p.part_o_bytes = (size_t)BT * p.splits * KL * C * 4;
  p.part_ml_bytes = (size_t)BT * p.splits * 2 * KL * 4;
  p.mref_bytes = (size_t)BT * p.nchunks * KL * 4;
  p.ticket_bytes = (size_t)BT * 4;
  p.ok = true;
  return p;
}

template <int NTW>
inline cudaError_t launch_tok_attn_t(const TokAttnArgs& a, const TokAttnPlan& pl, int BT, cudaStream_t st) {
  static PerDeviceOnce attr;
  if (attr.first()) {
    cudaError_t e = set_max_dyn_smem(tok_attn_fwd_kernel<NTW>);
    if (e != cudaSuccess) return e;
  }
  tok_attn_fwd_kernel<NTW><<<dim3(pl.splits, BT, pl.Z), 256, pl.smem, st>>>(a);
  return cudaGetLastError();
}
inline cudaError_t launch_tok_attn(const TokAttnArgs& a, const TokAttnPlan& pl, int BT, cudaStream_t st) {
  switch (pl.ntw) {
    case 2: return launch_tok_attn_t<2>(a, pl, BT, st);
    case 4: return launch_tok_attn_t<4>(a, pl, BT, st);
    case 6: return launch_tok_attn_t<6>(a, pl, BT, st);
    case 8: return launch_tok_attn_t<8>(a, pl, BT, st);
    case 10: return launch_tok_attn_t<10>(a, pl, BT, st);
    default: return launch_tok_attn_t<12>(a, pl, BT, st);
  }
}

// softmax over the 32 columns (latent tokens, columns >= K masked) of a 16 x 32 tile in the accumulator layout: rows g / g+8 of
// the warp's tile live in the four lanes of a quad
__device__ __forceinline__ void row_softmax_frag(float (&s)[4][4], int K, int t) {
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int col = nt * 8 + 2 * t;
    if (col >= K) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
    if (col + 1 >= K) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
    mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
    mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  float sm0 = 0.f, sm1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    s[nt][0] = __expf(s[nt][0] - mx0); s[nt][1] = __expf(s[nt][1] - mx0);
    s[nt][2] = __expf(s[nt][2] - mx1); s[nt][3] = __expf(s[nt][3] - mx1);
    sm0 += s[nt][0] + s[nt][1]; sm1 += s[nt][2] + s[nt][3];
  }
  sm0 += __shfl_xor_sync(0xffffffffu, sm0, 1); sm0 += __shfl_xor_sync(0xffffffffu, sm0, 2);
  sm1 += __shfl_xor_sync(0xffffffffu, sm1, 1); sm1 += __shfl_xor_sync(0xffffffffu, sm1, 2);
  const float i0 = 1.f / sm0, i1 = 1.f / sm1;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) { s[nt][0] *= i0; s[nt][1] *= i0; s[nt][2] *= i1; s[nt][3] *= i1; }
}

// rescale 16 rows of P~ (bf16, pitch Kp) to the final P: one warp, lane -> (row lane/2, 16 columns)
__device__ __forceinline__ void rescale_p_rows(bf16* Prows, const float* fac, int rows, int K, int Kp, int lane) {
  const int r = lane >> 1, kb = (lane & 1) * 16;
  if (r >= rows) return;
  bf16* prow = Prows + (size_t)r * Kp;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int k0 = kb + u * 8;
    if (k0 < Kp) {
      uint4 v = *reinterpret_cast<const uint4*>(prow + k0);
      uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float2 f = unpack_bf16x2(w[q]);
        const int k = k0 + 2 * q;
        f.x = k < K ? f.x * fac[k] : 0.f;
        f.y = k + 1 < K ? f.y * fac[k + 1] : 0.f;
        w[q] = pack_bf16x2(f.x, f.y);
      }
      *reinterpret_cast<uint4*>(prow + k0) = v;
    }
  }
}

// X1 columns [c_begin, c_end) of a 16-row tile: X + gate * Q . T', written over the staged tile (xw = the tile's shared memory)
__device__ __forceinline__ void q_times_tok_residual(const uint32_t (&aq)[2][4], uint32_t ts, uint8_t* xw, int pitch, int c_begin,
                                                     int c_end, float gate, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const uint32_t b_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * pitch + (lane >> 4) * 16);
  uint8_t* xr0 = xw + (size_t)g * pitch + t * 4;
  uint8_t* xr1 = xr0 + (size_t)8 * pitch;
#pragma unroll 2
  for (int c0 = c_begin; c0 < c_end; c0 += 16) {
    float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t b0, b1, b2, b3;
      ldsm_x4_t(ts + b_off + (uint32_t)(ks * 16 * pitch + c0 * 2), b0, b1, b2, b3);
      mma16816(d0, aq[ks], b0, b1);
      mma16816(d1, aq[ks], b2, b3);
    }
    uint32_t* p00 = reinterpret_cast<uint32_t*>(xr0 + c0 * 2);
    uint32_t* p01 = reinterpret_cast<uint32_t*>(xr0 + c0 * 2 + 16);
    uint32_t* p10 = reinterpret_cast<uint32_t*>(xr1 + c0 * 2);
    uint32_t* p11 = reinterpret_cast<uint32_t*>(xr1 + c0 * 2 + 16);
    float2 f;
    f = unpack_bf16x2(*p00); *p00 = pack_bf16x2(fmaf(gate, d0[0], f.x), fmaf(gate, d0[1], f.y));
    f = unpack_bf16x2(*p10); *p10 = pack_bf16x2(fmaf(gate, d0[2], f.x), fmaf(gate, d0[3], f.y));
    f = unpack_bf16x2(*p01); *p01 = pack_bf16x2(fmaf(gate, d1[0], f.x), fmaf(gate, d1[1], f.y));
    f = unpack_bf16x2(*p11); *p11 = pack_bf16x2(fmaf(gate, d1[2], f.x), fmaf(gate, d1[3], f.y));
  }
}

// =====================================================================================================================
// Kernel B: the adapted stream attends over the updated latent tokens
// =====================================================================================================================
struct XAttnArgs {
  const bf16* x;            // (BT, N, C)
  const bf16* TpB;          // (BT, K, C)
  bf16* X1;                 // (BT, N, C) out
  bf16* Qb;                 // (BT, N, Kp) out
  bf16* Pb;                 // (BT, N, Kp) in/out: P~ -> P (NULL: nothing to rescale)
  const float* pfac;        // (BT, nchunks, 32)
  const float* gate_av;     // device scalar
  int* err_flag;
  int N, C, K, Kp, chunk_rows, nchunks;
};

__global__ void __launch_bounds__(128) x_attn_fwd_kernel(const XAttnArgs a) {
  extern __shared__ __align__(128) uint8_t fa_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int WB = blockDim.x >> 5;
  const int C = a.C, N = a.N, K = a.K;
  const int pitch = C * 2 + 16;
  const int b = blockIdx.y;
  uint8_t* Tsm = fa_smem;
  uint8_t* xw = Tsm + (size_t)KL * pitch + (size_t)warp * 16 * pitch;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Tsm + (size_t)(KL + 16 * WB) * pitch);   // [0] T', [1 + warp] the warp's tile
  const uint32_t row_bytes = (uint32_t)C * 2;
  if (tid == 0) {
    for (int i = 0; i < 1 + WB; ++i) tc::mbar_init(tc::smem_u32(&bars[i]), 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  if (warp == 0) {
    const uint32_t bar = tc::smem_u32(&bars[0]);
    if (lane == 0) tc::mbar_expect_tx(bar, (uint32_t)K * row_bytes);
    __syncwarp();
    if (lane < K) bulk_g2s(tc::smem_u32(Tsm) + (uint32_t)(lane * pitch), a.TpB + ((size_t)b * K + lane) * C, row_bytes, bar);
  }
  if (K < KL) zero_rows(Tsm + (size_t)K * pitch, KL - K, pitch, (int)row_bytes, tid, blockDim.x);
  __syncthreads();
  const float gate = a.gate_av[0];
  const int ntiles = (N + 15) >> 4;
  const uint32_t barw = tc::smem_u32(&bars[1 + warp]);
  const uint32_t xs = tc::smem_u32(xw), ts = tc::smem_u32(Tsm);
  uint32_t ph = 0;
  bool first = true;
  for (int tile = blockIdx.x * WB + warp; tile < ntiles; tile += gridDim.x * WB) {
    const int n0 = tile << 4;
    const int rows = min(16, N - n0);
    const size_t grow = (size_t)b * N + n0;
    if (lane == 0) tc::mbar_expect_tx(barw, (uint32_t)rows * row_bytes);
    __syncwarp();
    if (lane < rows) bulk_g2s(xs + (uint32_t)(lane * pitch), a.x + (grow + lane) * C, row_bytes, barw);
    if (rows < 16) zero_rows(xw + (size_t)rows * pitch, 16 - rows, pitch, (int)row_bytes, lane, 32);
    // rescale this tile's rows of P~ to the final P (kernel A wrote them relative to the running max of their chunk)
    if (a.Pb) rescale_p_rows(a.Pb + grow * a.Kp, a.pfac + ((size_t)b * a.nchunks + n0 / a.chunk_rows) * KL, rows, K, a.Kp, lane);
    __syncwarp();
    bool ok = true;
    if (first) { ok = wait_bar(tc::smem_u32(&bars[0]), 0, a.err_flag); first = false; }
    if (ok) ok = wait_bar(barw, ph, a.err_flag);
    ph ^= 1;
    if (!__all_sync(0xffffffffu, ok)) return;
    // ---- S = X_tile . T'^T ; Q = softmax over the K latent tokens ----
    float s[4][4];
    rows_times_tokT(s, xs, ts, pitch, 0, C, lane);
    row_softmax_frag(s, K, t);
    uint32_t aq[2][4];
    acc_to_afrag(s, aq);
    if (a.Qb) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = nt * 8 + 2 * t;
        if (col < a.Kp) {
          if (g < rows) *reinterpret_cast<uint32_t*>(a.Qb + (grow + g) * a.Kp + col) = aq[nt >> 1][(nt & 1) * 2];
          if (g + 8 < rows) *reinterpret_cast<uint32_t*>(a.Qb + (grow + g + 8) * a.Kp + col) = aq[nt >> 1][(nt & 1) * 2 + 1];
        }
      }
    }
    // ---- X1 = X + gate * Q . T'  (16 output columns per step, written over the staged tile) ----
    q_times_tok_residual(aq, ts, xw, pitch, 0, C, gate, lane);
    fence_proxy_async();
    __syncwarp();
    if (lane < rows) {
      bulk_s2g(a.X1 + (grow + lane) * C, xs + (uint32_t)(lane * pitch), row_bytes);
      bulk_commit();
      bulk_wait_read0();
    }
    __syncwarp();
  }
}

// Wide rows (C > 384): the four warps of a CTA share ONE 16-row tile -- each takes a quarter of the channels of the logits
// (partials summed through shared memory; every warp then holds the full 16 x 32 tile and its softmax) and a quarter of the output
// columns -- so that a 1536-channel row is not a single warp's dependent chain and T' (32 x C) is staged once per CTA, not per warp.
__global__ void __launch_bounds__(128) x_attn_fwd_coop_kernel(const XAttnArgs a) {
  extern __shared__ __align__(128) uint8_t fa_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int C = a.C, N = a.N, K = a.K;
  const int pitch = C * 2 + 16;
  const int b = blockIdx.y;
  uint8_t* Tsm = fa_smem;
  uint8_t* xw = Tsm + (size_t)KL * pitch;
  float* sred = reinterpret_cast<float*>(xw + (size_t)16 * pitch);                  // [16][4 warps][32 lanes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sred + 16 * 4 * 32);                 // [0] T', [1] tile
  const uint32_t row_bytes = (uint32_t)C * 2;
  if (tid == 0) {
    tc::mbar_init(tc::smem_u32(&bars[0]), 1);
    tc::mbar_init(tc::smem_u32(&bars[1]), 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  const uint32_t xs = tc::smem_u32(xw), ts = tc::smem_u32(Tsm), bar1 = tc::smem_u32(&bars[1]);
  const int ntiles = (N + 15) >> 4;
  const int rlane = lane * 4 + warp;                  // bulk copies are dealt out over the four warps (see tok_attn_fwd_kernel)
  auto issue_tile = [&](int tile) {                   // EVERY thread of the CTA
    const int n0 = tile << 4;
    const int rows = min(16, N - n0);
    if (tid == 0) tc::mbar_expect_tx(bar1, (uint32_t)rows * row_bytes);
    if (rlane < rows) bulk_g2s(xs + (uint32_t)(rlane * pitch), a.x + ((size_t)b * N + n0 + rlane) * C, row_bytes, bar1);
  };
  {
    const uint32_t bar = tc::smem_u32(&bars[0]);
    if (tid == 0) tc::mbar_expect_tx(bar, (uint32_t)K * row_bytes);
    if (rlane < K) bulk_g2s(ts + (uint32_t)(rlane * pitch), a.TpB + ((size_t)b * K + rlane) * C, row_bytes, bar);
    if ((int)blockIdx.x < ntiles) issue_tile(blockIdx.x);
  }
  if (K < KL) zero_rows(Tsm + (size_t)K * pitch, KL - K, pitch, (int)row_bytes, tid, 128);
  const float gate = a.gate_av[0];
  const int cq = ((C >> 4) + 3) / 4 * 16;                 // channels / output columns per warp (multiple of 16)
  const int c_begin = min(C, warp * cq), c_end = min(C, c_begin + cq);
  uint32_t ph = 0;
  bool first = true;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int n0 = tile << 4;
    const int rows = min(16, N - n0);
    const size_t grow = (size_t)b * N + n0;
    if (warp == 2 && a.Pb) rescale_p_rows(a.Pb + grow * a.Kp, a.pfac + ((size_t)b * a.nchunks + n0 / a.chunk_rows) * KL, rows, K, a.Kp, lane);
    bool ok = true;
    if (first) { ok = wait_bar(tc::smem_u32(&bars[0]), 0, a.err_flag); first = false; }
    if (ok) ok = wait_bar(bar1, ph, a.err_flag);
    ph ^= 1;
    // (after the wait: the load was issued behind the previous tile's bulk store, which reads all 16 rows of the buffer)
    if (rows < 16) zero_rows(xw + (size_t)rows * pitch, 16 - rows, pitch, (int)row_bytes, tid, 128);
    if (!__syncthreads_and(ok)) return;
    float s[4][4];
    rows_times_tokT(s, xs, ts, pitch, c_begin, c_end, lane);
#pragma unroll
    for (int i = 0; i < 16; ++i) sred[(i * 4 + warp) * 32 + lane] = s[i >> 2][i & 3];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 16; ++i)
      s[i >> 2][i & 3] = sred[(i * 4 + 0) * 32 + lane] + sred[(i * 4 + 1) * 32 + lane] + sred[(i * 4 + 2) * 32 + lane] + sred[(i * 4 + 3) * 32 + lane];
    row_softmax_frag(s, K, t);
    uint32_t aq[2][4];
    acc_to_afrag(s, aq);
    if (warp == 3 && a.Qb) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = nt * 8 + 2 * t;
        if (col < a.Kp) {
          if (g < rows) *reinterpret_cast<uint32_t*>(a.Qb + (grow + g) * a.Kp + col) = aq[nt >> 1][(nt & 1) * 2];
          if (g + 8 < rows) *reinterpret_cast<uint32_t*>(a.Qb + (grow + g + 8) * a.Kp + col) = aq[nt >> 1][(nt & 1) * 2 + 1];
        }
      }
    }
    q_times_tok_residual(aq, ts, xw, pitch, c_begin, c_end, gate, lane);
    fence_proxy_async();
    __syncthreads();                                   // every column range of the tile is final; sred is free again
    if (rlane < rows) {
      bulk_s2g(a.X1 + (grow + rlane) * C, xs + (uint32_t)(rlane * pitch), row_bytes);
      bulk_commit();
      bulk_wait_read0();
    }
    if (tile + (int)gridDim.x < ntiles) {
      __syncthreads();                                 // every row of the tile has left: the buffer may be overwritten
      issue_tile(tile + gridDim.x);
    }
  }
}

struct XAttnPlan { bool ok; bool coop; int WB, gx; size_t smem; };
inline XAttnPlan make_x_attn_plan(int BT, int N, int C, int K) {
  XAttnPlan p{};
  if (C % 32 || C < 32 || K > KL || K < 1) return p;
  const size_t pitch = (size_t)C * 2 + 16;
  const int ntiles = (N + 15) / 16;
  if (C > 384) {                                           // one tile per CTA iteration, four warps share it
    p.coop = true;
    p.WB = 4;
    p.smem = (size_t)(KL + 16) * pitch + 16 * 4 * 32 * 4 + 64;
    if (p.smem > 225 * 1024) return p;
    int want = (3 * 148 + BT - 1) / BT;                     // T' is staged per CTA: ~3 tiles per CTA when the frame has them
    if (want < 1) want = 1;
    int gx = (ntiles + 2) / 3;
    if (gx > want) gx = want;
    if (gx < 1) gx = 1;
    if ((long long)gx * BT < 148) { gx = (148 + BT - 1) / BT; if (gx > ntiles) gx = ntiles; }
    p.gx = gx;
    p.ok = true;
    return p;
  }
  int WB = 0;
  if ((KL + 64) * pitch + 64 <= 104 * 1024) WB = 4;
  else if ((KL + 32) * pitch + 64 <= 220 * 1024) WB = 2;
  else if ((KL + 16) * pitch + 64 <= 220 * 1024) WB = 1;
  else return p;
  p.WB = WB;
  p.smem = (size_t)(KL + 16 * WB) * pitch + 64;
  const int groups = (ntiles + WB - 1) / WB;
  int want = (6 * 148 + BT - 1) / BT;
  if (want < 1) want = 1;
  p.gx = groups < want ? groups : want;
  p.ok = true;
  return p;
}
inline cudaError_t launch_x_attn(const XAttnArgs& a, const XAttnPlan& pl, int BT, cudaStream_t st) {
  static PerDeviceOnce attr;
  if (attr.first()) {
    cudaError_t e = set_max_dyn_smem(x_attn_fwd_kernel);
    if (e == cudaSuccess) e = set_max_dyn_smem(x_attn_fwd_coop_kernel);
    if (e != cudaSuccess) return e;
  }
  if (pl.coop) x_attn_fwd_coop_kernel<<<dim3(pl.gx, BT), 128, pl.smem, st>>>(a);
  else x_attn_fwd_kernel<<<dim3(pl.gx, BT), 32 * pl.WB, pl.smem, st>>>(a);
  return cudaGetLastError();
}


// =====================================================================================================================
// Backward
// =====================================================================================================================
// A fragments of a 16 x 32 bf16 tile staged row-major in shared memory ([row][latent token], PS_PITCH bytes per row)
__device__ __forceinline__ void stage_to_afrag(uint32_t tile, uint32_t (&a)[2][4], int lane) {
  const uint32_t addr = tile + (uint32_t)((lane & 15) * PS_PITCH + (lane >> 4) * 16);
  ldsm_x4(addr, a[0][0], a[0][1], a[0][2], a[0][3]);
  ldsm_x4(addr + 32, a[1][0], a[1][1], a[1][2], a[1][3]);
}
// 16 x 32 tile in the accumulator layout -> bf16 staging tile ([row][latent token]); rows r0 = g and g + 8 of the tile
__device__ __forceinline__ void frag_to_stage(const float (&v)[4][4], uint8_t* tile, int g, int t) {
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int col = nt * 8 + 2 * t;
    *reinterpret_cast<uint32_t*>(tile + (size_t)g * PS_PITCH + col * 2) = pack_bf16x2(v[nt][0], v[nt][1]);
    *reinterpret_cast<uint32_t*>(tile + (size_t)(g + 8) * PS_PITCH + col * 2) = pack_bf16x2(v[nt][2], v[nt][3]);
  }
}
// 16 x 32 tile of a (rows, Kp) bf16 matrix in global memory -> accumulator layout (rows >= rows_valid and columns >= K read as 0)
__device__ __forceinline__ void load_frag_rows(const bf16* base, int Kp, int K, int rows_valid, int g, int t, float (&v)[4][4]) {
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int col = nt * 8 + 2 * t;
    float2 lo = make_float2(0.f, 0.f), hi = make_float2(0.f, 0.f);
    if (col < Kp) {
      if (g < rows_valid) lo = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(base + (size_t)g * Kp + col));
      if (g + 8 < rows_valid) hi = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(base + (size_t)(g + 8) * Kp + col));
    }
    v[nt][0] = col < K ? lo.x : 0.f; v[nt][1] = col + 1 < K ? lo.y : 0.f;
    v[nt][2] = col < K ? hi.x : 0.f; v[nt][3] = col + 1 < K ? hi.y : 0.f;
  }
}
// acc(32 x this warp's columns) += A^T . B over the R rows of a chunk: A staged [row][latent token] (PS_PITCH), B a row tile
// [row][channel] (pitch); warp-private column pairs p_begin .. of the CTA's column range starting at zc0
template <int NTW>
__device__ __forceinline__ void accum_tokens_by_cols(float (&o)[2][NTW][4], uint32_t a_stage, uint32_t b_tile, int pitch, int R, int zc0,
                                                     int p_begin, int npairs, int lane) {
  const int j = lane >> 3, r = lane & 7;
  const uint32_t a_off = (uint32_t)((r + (j >> 1) * 8) * PS_PITCH + (j & 1) * 16);
  const uint32_t b_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * pitch + (lane >> 4) * 16);
  for (int ks = 0; ks < (R >> 4); ++ks) {
    uint32_t a0[4], a1[4];
    ldsm_x4_t(a_stage + a_off + (uint32_t)(ks * 16 * PS_PITCH), a0[0], a0[1], a0[2], a0[3]);
    ldsm_x4_t(a_stage + a_off + (uint32_t)(ks * 16 * PS_PITCH) + 32, a1[0], a1[1], a1[2], a1[3]);
    const uint32_t vrow = b_tile + b_off + (uint32_t)(ks * 16 * pitch);
#pragma unroll
    for (int pi = 0; pi < NTW / 2; ++pi) {
      if (p_begin + pi < npairs) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(vrow + (uint32_t)((zc0 + (p_begin + pi) * 16) * 2), b0, b1, b2, b3);
        mma16816(o[0][2 * pi], a0, b0, b1);
        mma16816(o[0][2 * pi + 1], a0, b2, b3);
        mma16816(o[1][2 * pi], a1, b0, b1);
        mma16816(o[1][2 * pi + 1], a1, b2, b3);
      }
    }
  }
}
// add the warp's (32 x columns) accumulator into a (K, C) fp32 matrix in global memory
template <int NTW>
__device__ __forceinline__ void atomic_flush_tokens(const float (&o)[2][NTW][4], float* dst, int C, int K, int zc0, int p_begin, int npairs,
                                                    int g, int t) {
#pragma unroll
  for (int pi = 0; pi < NTW / 2; ++pi) {
    if (p_begin + pi < npairs) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = zc0 + (p_begin + pi) * 16 + h * 8 + 2 * t;
        const int nt = 2 * pi + h;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const int k0 = mt * 16 + g, k1 = k0 + 8;
          if (k0 < K) { atomicAdd(dst + (size_t)k0 * C + c, o[mt][nt][0]); atomicAdd(dst + (size_t)k0 * C + c + 1, o[mt][nt][1]); }
          if (k1 < K) { atomicAdd(dst + (size_t)k1 * C + c, o[mt][nt][2]); atomicAdd(dst + (size_t)k1 * C + c + 1, o[mt][nt][3]); }
        }
      }
    }
  }
}

// ---- backward of  Q = softmax_k(X T'^T),  X1 = X + gate * Q T'  ---------------------------------------------------------
//   dQraw = G T'^T ; d_gate += sum Q dQraw ; dL2 = gate * Q (dQraw - rowsum(Q dQraw))
//   dx = G + dL2 T' ;  dT'[b] += (gate Q)^T G + dL2^T X                      (G = dL/dX1)
struct XAttnBwdArgs {
  const bf16* G; const bf16* x; const bf16* TpB; const bf16* Qb;
  bf16* dx; float* dTp; const float* gate_av; float* d_gate_av;
  int* err_flag;
  int N, C, K, Kp, R, nchunks, splits, zcols, nbuf;
};

template <int NTW>
__global__ void __launch_bounds__(256, (NTW <= 8 ? 2 : 1)) x_attn_bwd_kernel(const XAttnBwdArgs a) {
  extern __shared__ __align__(128) uint8_t fa_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int C = a.C, N = a.N, K = a.K, R = a.R;
  const int pitch = C * 2 + 16;
  const int b = blockIdx.y, split = blockIdx.x, z = blockIdx.z;
  uint8_t* Tsm = fa_smem;
  uint8_t* Gs0 = Tsm + (size_t)KL * pitch;              // nbuf x (G tile, X tile): the next chunk lands while this one is worked on
  uint8_t* Qs = Gs0 + (size_t)a.nbuf * 2 * R * pitch;   // [R][32] bf16 gate * Q
  uint8_t* Ls = Qs + (size_t)R * PS_PITCH;              // [R][32] bf16 dL2
  float* sred = reinterpret_cast<float*>(Ls + (size_t)R * PS_PITCH);
  float* red = sred + 16 * 8 * 32;                      // [32]
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 32);
  const int c_begin = (int)((long long)split * a.nchunks / a.splits), c_end = (int)((long long)(split + 1) * a.nchunks / a.splits);
  const uint32_t row_bytes = (uint32_t)C * 2;
  if (tid == 0) {
    for (int i = 0; i < 3; ++i) tc::mbar_init(tc::smem_u32(&bars[i]), 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  const uint32_t ts = tc::smem_u32(Tsm);
  const uint32_t buf_bytes = 2u * (uint32_t)R * (uint32_t)pitch;
  const int rlane = lane * 8 + warp;                    // the row this thread copies (see tok_attn_fwd_kernel)
  auto issue_chunk = [&](int chunk, int buf) {          // EVERY thread of the CTA
    const int n0 = chunk * R;
    const int rows = min(R, N - n0);
    const uint32_t bar = tc::smem_u32(&bars[1 + buf]);
    const uint32_t gsb = tc::smem_u32(Gs0) + (uint32_t)buf * buf_bytes, xsb = gsb + (uint32_t)R * (uint32_t)pitch;
    if (tid == 0) tc::mbar_expect_tx(bar, 2u * (uint32_t)rows * row_bytes);
    if (rlane < rows) {
      bulk_g2s(gsb + (uint32_t)(rlane * pitch), a.G + ((size_t)b * N + n0 + rlane) * C, row_bytes, bar);
      bulk_g2s(xsb + (uint32_t)(rlane * pitch), a.x + ((size_t)b * N + n0 + rlane) * C, row_bytes, bar);
    }
  };
  {
    const uint32_t bar = tc::smem_u32(&bars[0]);
    if (tid == 0) tc::mbar_expect_tx(bar, (uint32_t)K * row_bytes);
    if (lane < 4 && lane * 8 + warp < K)
      bulk_g2s(ts + (uint32_t)((lane * 8 + warp) * pitch), a.TpB + ((size_t)b * K + lane * 8 + warp) * C, row_bytes, bar);
    if (c_begin < c_end) issue_chunk(c_begin, 0);
    if (a.nbuf == 2 && c_begin + 1 < c_end) issue_chunk(c_begin + 1, 1);
  }
  if (K < KL) zero_rows(Tsm + (size_t)K * pitch, KL - K, pitch, (int)row_bytes, tid, 256);
  const float gate = a.gate_av[0];
  const int zc0 = z * a.zcols;
  const int zcols = min(a.zcols, C - zc0);
  const int npairs = zcols >> 4;
  const int p_begin = warp * (NTW / 2);
  float o[2][NTW][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) o[mt][nt][jj] = 0.f;
  const int pass1_warps = R >> 4;
  const int KS = 8 / pass1_warps;
  const int mtile = warp / KS, kpart = warp - mtile * KS;
  const bool owner = kpart == 0;
  const int ksteps = ((C >> 4) + KS - 1) / KS;
  const int k_begin = min(C, kpart * ksteps * 16), k_end = min(C, k_begin + ksteps * 16);
  // dx columns of this warp: the CTA's column range cut into KS parts of whole 16-column steps
  const int dsteps = ((zcols >> 4) + KS - 1) / KS;
  const int d_begin = zc0 + min(zcols, kpart * dsteps * 16), d_end = zc0 + min(zcols, (kpart + 1) * dsteps * 16);
  float gacc = 0.f;
  uint32_t ph0 = 0, ph1 = 0;
  bool ok = true;
  for (int chunk = c_begin; chunk < c_end; ++chunk) {
    const int n0 = chunk * R;
    const int rows = min(R, N - n0);
    const int buf = a.nbuf == 2 ? ((chunk - c_begin) & 1) : 0;
    uint8_t* Gs = Gs0 + (size_t)buf * buf_bytes;
    uint8_t* Xs = Gs + (size_t)R * pitch;
    const uint32_t gs = tc::smem_u32(Gs), xs = tc::smem_u32(Xs);
    if (chunk == c_begin) ok = wait_bar(tc::smem_u32(&bars[0]), 0, a.err_flag);
    if (ok) ok = wait_bar(tc::smem_u32(&bars[1 + buf]), buf ? ph1 : ph0, a.err_flag);
    if (buf) ph1 ^= 1; else ph0 ^= 1;
    if (rows < R) {                                    // (after the wait: see x_attn_fwd_coop_kernel)
      zero_rows(Gs + (size_t)rows * pitch, R - rows, pitch, (int)row_bytes, tid, 256);
      zero_rows(Xs + (size_t)rows * pitch, R - rows, pitch, (int)row_bytes, tid, 256);
    }
    ok = __syncthreads_and(ok);
    if (!ok) break;
    // ---- dQraw = G_tile . T'^T ----
    float s[4][4];
    rows_times_tokT(s, gs + (uint32_t)(mtile * 16 * pitch), ts, pitch, k_begin, k_end, lane);
    if (KS > 1) {
      if (!owner) {
#pragma unroll
        for (int i = 0; i < 16; ++i) sred[(i * 8 + warp) * 32 + lane] = s[i >> 2][i & 3];
      }
      __syncthreads();
      if (owner) {
        for (int jj = 1; jj < KS; ++jj)
#pragma unroll
          for (int i = 0; i < 16; ++i) s[i >> 2][i & 3] += sred[(i * 8 + warp + jj) * 32 + lane];
      }
    }
    if (owner) {
      const int rv = rows - mtile * 16;
      float q[4][4];
      load_frag_rows(a.Qb + ((size_t)b * N + n0 + mtile * 16) * a.Kp, a.Kp, K, rv, g, t, q);
      float d0 = 0.f, d1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) { d0 += q[nt][0] * s[nt][0] + q[nt][1] * s[nt][1]; d1 += q[nt][2] * s[nt][2] + q[nt][3] * s[nt][3]; }
      d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
      d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
      if (t == 0) gacc += d0 + d1;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        s[nt][0] = gate * q[nt][0] * (s[nt][0] - d0); s[nt][1] = gate * q[nt][1] * (s[nt][1] - d0);
        s[nt][2] = gate * q[nt][2] * (s[nt][2] - d1); s[nt][3] = gate * q[nt][3] * (s[nt][3] - d1);
        q[nt][0] *= gate; q[nt][1] *= gate; q[nt][2] *= gate; q[nt][3] *= gate;
      }
      frag_to_stage(q, Qs + (size_t)mtile * 16 * PS_PITCH, g, t);
      frag_to_stage(s, Ls + (size_t)mtile * 16 * PS_PITCH, g, t);
    }
    __syncthreads();
    // ---- dT' += (gate Q)^T G + dL2^T X ----
    if (p_begin < npairs) {
      accum_tokens_by_cols<NTW>(o, tc::smem_u32(Qs), gs, pitch, R, zc0, p_begin, npairs, lane);
      accum_tokens_by_cols<NTW>(o, tc::smem_u32(Ls), xs, pitch, R, zc0, p_begin, npairs, lane);
    }
    __syncthreads();
    // ---- dx = G + dL2 . T'  (in place over the staged G tile; the KS warps of a row tile split the columns) ----
    if (d_begin < d_end) {
      uint32_t al[2][4];
      stage_to_afrag(tc::smem_u32(Ls + (size_t)mtile * 16 * PS_PITCH), al, lane);
      q_times_tok_residual(al, ts, Gs + (size_t)mtile * 16 * pitch, pitch, d_begin, d_end, 1.f, lane);
    }
    fence_proxy_async();
    __syncthreads();
    if (rlane < rows) {
      bulk_s2g(a.dx + ((size_t)b * N + n0 + rlane) * C + zc0, gs + (uint32_t)(rlane * pitch + zc0 * 2), (uint32_t)zcols * 2);
      bulk_commit();
      bulk_wait_read0();
    }
    if (chunk + a.nbuf < c_end) {
      __syncthreads();                                 // every row of the tile has left: this buffer may be overwritten
      issue_chunk(chunk + a.nbuf, buf);
    }
  }
  if (!__syncthreads_and(ok)) return;
  if (p_begin < npairs) atomic_flush_tokens<NTW>(o, a.dTp + (size_t)b * K * C, C, K, zc0, p_begin, npairs, g, t);
  if (z == 0 && a.d_gate_av) {
    gacc = warp_sum(gacc);
    if (lane == 0) red[warp] = gacc;
    __syncthreads();
    if (tid == 0) {
      float tot = 0.f;
      for (int w = 0; w < 8; ++w) tot += red[w];
      if (tot != 0.f) atomicAdd(a.d_gate_av, tot);
    }
  }
}

struct XAttnBwdPlan { bool ok; int R, nbuf, nchunks, splits, Z, zcols, ntw; size_t smem; };
inline size_t x_attn_bwd_smem(int C, int R, int nbuf) {
  const size_t pitch = (size_t)C * 2 + 16;
  return (size_t)KL * pitch + (size_t)nbuf * 2 * R * pitch + 2 * (size_t)R * PS_PITCH + (size_t)(16 * 8 * 32 + 32) * 4 + 64;
}
inline int fa_splits(int BT, int Z, int N, int nchunks, size_t smem, int ntw) {
  int cps = (int)((227 * 1024) / (smem + 1024));
  if (cps > 2) cps = 2;
  if (ntw >= 10 || cps < 1) cps = 1;
  const int slots = 148 * cps;
  const int frames = BT * Z;
  const int s_min = (74 + frames - 1) / frames;
  int sp = (N + 192) / 384;
  if (sp < s_min) sp = s_min;
  if (sp < 1) sp = 1;
  if ((long long)frames * sp > slots && frames <= slots) sp = slots / frames;
  if (sp > nchunks) sp = nchunks;
  if (sp > 32) sp = 32;
  return sp;
}
// (rows per chunk, buffers) of the backward kernels: two CTAs per SM (<= 104 KB) when the call has the CTAs to fill them, else
// one CTA per SM with what the 227 KB allow.  The kernels can double-buffer their chunks (nbuf = 2), but the sweep on B200
// (profiles/r02_fused_attention_chunk_sweep.txt) never preferred it to twice the rows per chunk: single buffers, largest R.
template <typename SmemFn>
inline bool fa_bwd_tile(int BT, int N, int C, int ntw, SmemFn smem_of, int& R, int& nbuf) {
  static const int small_c[][2] = {{128, 1}, {64, 1}, {32, 1}, {16, 1}};
  static const int big_c[][2] = {{64, 1}, {32, 1}, {16, 1}};
  const long long ctas_est = (long long)BT * ((N + 192) / 384 > 0 ? (N + 192) / 384 : 1);
  R = 0;
  if (ntw < 10 && ctas_est > 148)
    for (auto& c : small_c) if (smem_of(C, c[0], c[1]) <= (size_t)104 * 1024) { R = c[0]; nbuf = c[1]; break; }
  if (!R) for (auto& c : big_c) if (smem_of(C, c[0], c[1]) <= (size_t)225 * 1024) { R = c[0]; nbuf = c[1]; break; }
  if (!R) return false;
  while (R > 16 && R / 2 >= N) R /= 2;
  return true;
}
inline XAttnBwdPlan make_x_attn_bwd_plan(int BT, int N, int C, int K) {
  XAttnBwdPlan p{};
  if (C % 32 || C < 32 || K > KL || K < 1 || N < 1) return p;
  p.Z = (C + 767) / 768;
  p.zcols = ((C / 16 + p.Z - 1) / p.Z) * 16;
  p.ntw = 2 * ((p.zcols / 16 + 7) / 8);
  if (p.ntw > 12) return p;
  if (!fa_bwd_tile(BT, N, C, p.ntw, x_attn_bwd_smem, p.R, p.nbuf)) return p;
  p.nchunks = (N + p.R - 1) / p.R;
  p.splits = fa_splits(BT, p.Z, N, p.nchunks, x_attn_bwd_smem(C, p.R, p.nbuf), p.ntw);
  if ((p.nchunks + p.splits - 1) / p.splits < 2) p.nbuf = 1;
  p.smem = x_attn_bwd_smem(C, p.R, p.nbuf);
  fa_widen_z(BT, p.splits, C, p.smem, 1, 3, p.Z, p.zcols, p.ntw);
  p.ok = true;
  return p;
}
template <int NTW>
inline cudaError_t launch_x_attn_bwd_t(const XAttnBwdArgs& a, const XAttnBwdPlan& pl, int BT, cudaStream_t st) {
  static PerDeviceOnce attr;
  if (attr.first()) {
    cudaError_t e = set_max_dyn_smem(x_attn_bwd_kernel<NTW>);
    if (e != cudaSuccess) return e;
  }
  x_attn_bwd_kernel<NTW><<<dim3(pl.splits, BT, pl.Z), 256, pl.smem, st>>>(a);
  return cudaGetLastError();
}
inline cudaError_t launch_x_attn_bwd(const XAttnBwdArgs& a, const XAttnBwdPlan& pl, int BT, cudaStream_t st) {
  switch (pl.ntw) {
    case 2: return launch_x_attn_bwd_t<2>(a, pl, BT, st);
    case 4: return launch_x_attn_bwd_t<4>(a, pl, BT, st);
    case 6: return launch_x_attn_bwd_t<6>(a, pl, BT, st);
    case 8: return launch_x_attn_bwd_t<8>(a, pl, BT, st);
    case 10: return launch_x_attn_bwd_t<10>(a, pl, BT, st);
    default: return launch_x_attn_bwd_t<12>(a, pl, BT, st);
  }
}

// ---- per (frame, latent token): bf16 copy of dT' and D[b,k] = sum_c dT'[b,k,c] (T'[b,k,c] - T[k,c])  (= sum_n dP P, the
// softmax-backward inner product, without a pass over the N rows: O = T' - T = P V) ----------------------------------------
__global__ void __launch_bounds__(256) dtp_prep_kernel(const float* __restrict__ dTp, const float* __restrict__ Tp, const float* __restrict__ tok,
                                                       bf16* __restrict__ dTpB, float* __restrict__ D, int BT, int K, int C) {
  const int warp = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= BT * K) return;
  const int k = warp % K;
  const float* d = dTp + (size_t)warp * C;
  const float* tp = Tp + (size_t)warp * C;
  const float* tk = tok + (size_t)k * C;
  bf16* o = dTpB + (size_t)warp * C;
  float acc = 0.f;
  for (int c = lane * 2; c < C; c += 64) {
    const float2 dv = *reinterpret_cast<const float2*>(d + c);
    const float2 tv = *reinterpret_cast<const float2*>(tp + c);
    acc += dv.x * (tv.x - tk[c]) + dv.y * (tv.y - tk[c + 1]);
    *reinterpret_cast<uint32_t*>(o + c) = pack_bf16x2(dv.x, dv.y);
  }
  acc = warp_sum(acc);
  if (lane == 0) D[warp] = acc;
}

// ---- backward of  P = softmax_n(V T^T),  T' = T + P^T V  ------------------------------------------------------------------
//   dP = V dT'^T ; dL1 = P (dP - D) ; dV = P dT' + dL1 T + dobar / N ; d tokens += dL1^T V
// P is recomputed in fp32 from the logits V T^T and the forward's final (max, sum) of every latent token: it is never stored.
struct TokAttnBwdArgs {
  const bf16* V; const bf16* tokb; const bf16* dTpB; const float* stat_ml; const float* D; const float* dobar;
  bf16* gV; float* d_tok;
  int* err_flag;
  float invN;
  int N, C, K, Kp, ld_tok, R, nchunks, splits, zcols, nbuf;
};

template <int NTW>
__global__ void __launch_bounds__(256, (NTW <= 8 ? 2 : 1)) tok_attn_bwd_kernel(const TokAttnBwdArgs a) {
  extern __shared__ __align__(128) uint8_t fa_smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int C = a.C, N = a.N, K = a.K, R = a.R;
  const int pitch = C * 2 + 16;
  const int b = blockIdx.y, split = blockIdx.x, z = blockIdx.z;
  uint8_t* Tsm = fa_smem;                                // latent tokens T
  uint8_t* Dsm = Tsm + (size_t)KL * pitch;               // dT'[b]
  uint8_t* Vsm0 = Dsm + (size_t)KL * pitch;              // nbuf V tiles: the next chunk lands while this one is worked on
  uint8_t* Ps = Vsm0 + (size_t)a.nbuf * R * pitch;       // [R][32] bf16 P
  uint8_t* Ls = Ps + (size_t)R * PS_PITCH;               // [R][32] bf16 dL1
  float* sred = reinterpret_cast<float*>(Ls + (size_t)R * PS_PITCH);
  float* dvec = sred + 16 * 8 * 32;                      // [32] D, [32] M, [32] 1 / L
  uint64_t* bars = reinterpret_cast<uint64_t*>(dvec + 96);
  const int c_begin = (int)((long long)split * a.nchunks / a.splits), c_end = (int)((long long)(split + 1) * a.nchunks / a.splits);
  const uint32_t row_bytes = (uint32_t)C * 2;
  if (tid == 0) {
    for (int i = 0; i < 3; ++i) tc::mbar_init(tc::smem_u32(&bars[i]), 1);
    tc::fence_barrier_init();
  }
  if (tid < KL) {
    dvec[tid] = tid < K ? a.D[(size_t)b * K + tid] : 0.f;
    const float L = a.stat_ml[(size_t)b * 2 * KL + KL + tid];
    dvec[KL + tid] = a.stat_ml[(size_t)b * 2 * KL + tid];
    dvec[2 * KL + tid] = (tid < K && L > 0.f) ? 1.f / L : 0.f;
  }
  __syncthreads();
  const uint32_t ts = tc::smem_u32(Tsm), ds = tc::smem_u32(Dsm);
  const uint32_t buf_bytes = (uint32_t)R * (uint32_t)pitch;
  const int rlane = lane * 8 + warp;                    // the row this thread copies (see tok_attn_fwd_kernel)
  auto issue_chunk = [&](int chunk, int buf) {          // EVERY thread of the CTA
    const int n0 = chunk * R;
    const int rows = min(R, N - n0);
    const uint32_t bar = tc::smem_u32(&bars[1 + buf]);
    const uint32_t vsb = tc::smem_u32(Vsm0) + (uint32_t)buf * buf_bytes;
    if (tid == 0) tc::mbar_expect_tx(bar, (uint32_t)rows * row_bytes);
    if (rlane < rows) bulk_g2s(vsb + (uint32_t)(rlane * pitch), a.V + ((size_t)b * N + n0 + rlane) * C, row_bytes, bar);
  };
  {
    const uint32_t bar = tc::smem_u32(&bars[0]);
    if (tid == 0) tc::mbar_expect_tx(bar, 2u * (uint32_t)K * row_bytes);
    if (lane < 4 && lane * 8 + warp < K) {
      const int k = lane * 8 + warp;
      bulk_g2s(ts + (uint32_t)(k * pitch), a.tokb + (size_t)k * a.ld_tok, row_bytes, bar);
      bulk_g2s(ds + (uint32_t)(k * pitch), a.dTpB + ((size_t)b * K + k) * C, row_bytes, bar);
    }
    if (c_begin < c_end) issue_chunk(c_begin, 0);
    if (a.nbuf == 2 && c_begin + 1 < c_end) issue_chunk(c_begin + 1, 1);
  }
  if (K < KL) {
    zero_rows(Tsm + (size_t)K * pitch, KL - K, pitch, (int)row_bytes, tid, 256);
    zero_rows(Dsm + (size_t)K * pitch, KL - K, pitch, (int)row_bytes, tid, 256);
  }
  const int zc0 = z * a.zcols;
  const int zcols = min(a.zcols, C - zc0);
  const int npairs = zcols >> 4;
  const int p_begin = warp * (NTW / 2);
  float o[2][NTW][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) o[mt][nt][jj] = 0.f;
  const int pass1_warps = R >> 4;
  const int KS = 8 / pass1_warps;
  const int mtile = warp / KS, kpart = warp - mtile * KS;
  const bool owner = kpart == 0;
  const int ksteps = ((C >> 4) + KS - 1) / KS;
  const int k_begin = min(C, kpart * ksteps * 16), k_end = min(C, k_begin + ksteps * 16);
  const int dsteps = ((zcols >> 4) + KS - 1) / KS;
  const int d_begin = zc0 + min(zcols, kpart * dsteps * 16), d_end = zc0 + min(zcols, (kpart + 1) * dsteps * 16);
  uint32_t ph0 = 0, ph1 = 0;
  bool ok = true;
  for (int chunk = c_begin; chunk < c_end; ++chunk) {
    const int n0 = chunk * R;
    const int rows = min(R, N - n0);
    const int buf = a.nbuf == 2 ? ((chunk - c_begin) & 1) : 0;
    uint8_t* Vsm = Vsm0 + (size_t)buf * buf_bytes;
    const uint32_t vs = tc::smem_u32(Vsm);
    if (chunk == c_begin) ok = wait_bar(tc::smem_u32(&bars[0]), 0, a.err_flag);
    if (ok) ok = wait_bar(tc::smem_u32(&bars[1 + buf]), buf ? ph1 : ph0, a.err_flag);
    if (buf) ph1 ^= 1; else ph0 ^= 1;
    if (rows < R) zero_rows(Vsm + (size_t)rows * pitch, R - rows, pitch, (int)row_bytes, tid, 256);
    ok = __syncthreads_and(ok);
    if (!ok) break;
    // ---- logits V_tile . T^T (-> P) and dP = V_tile . dT'^T ----
    float s[4][4], p[4][4];
    rows_times_tokT(p, vs + (uint32_t)(mtile * 16 * pitch), ts, pitch, k_begin, k_end, lane);
    rows_times_tokT(s, vs + (uint32_t)(mtile * 16 * pitch), ds, pitch, k_begin, k_end, lane);
    if (KS > 1) {
      if (!owner) {
#pragma unroll
        for (int i = 0; i < 16; ++i) sred[(i * 8 + warp) * 32 + lane] = p[i >> 2][i & 3];
      }
      __syncthreads();
      if (owner) {
        for (int jj = 1; jj < KS; ++jj)
#pragma unroll
          for (int i = 0; i < 16; ++i) p[i >> 2][i & 3] += sred[(i * 8 + warp + jj) * 32 + lane];
      }
      __syncthreads();
      if (!owner) {
#pragma unroll
        for (int i = 0; i < 16; ++i) sred[(i * 8 + warp) * 32 + lane] = s[i >> 2][i & 3];
      }
      __syncthreads();
      if (owner) {
        for (int jj = 1; jj < KS; ++jj)
#pragma unroll
          for (int i = 0; i < 16; ++i) s[i >> 2][i & 3] += sred[(i * 8 + warp + jj) * 32 + lane];
      }
    }
    if (owner) {
      const int rv = rows - mtile * 16;
      const bool v0 = g < rv, v1 = g + 8 < rv;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int col = nt * 8 + 2 * t;
        const float dA = dvec[col], dB = dvec[col + 1];
        const float mA = dvec[KL + col], mB = dvec[KL + col + 1], iA = dvec[2 * KL + col], iB = dvec[2 * KL + col + 1];
        p[nt][0] = v0 ? __expf(p[nt][0] - mA) * iA : 0.f; p[nt][1] = v0 ? __expf(p[nt][1] - mB) * iB : 0.f;
        p[nt][2] = v1 ? __expf(p[nt][2] - mA) * iA : 0.f; p[nt][3] = v1 ? __expf(p[nt][3] - mB) * iB : 0.f;
        s[nt][0] = p[nt][0] * (s[nt][0] - dA); s[nt][1] = p[nt][1] * (s[nt][1] - dB);
        s[nt][2] = p[nt][2] * (s[nt][2] - dA); s[nt][3] = p[nt][3] * (s[nt][3] - dB);
      }
      frag_to_stage(p, Ps + (size_t)mtile * 16 * PS_PITCH, g, t);
      frag_to_stage(s, Ls + (size_t)mtile * 16 * PS_PITCH, g, t);
    }
    __syncthreads();
    // ---- d tokens += dL1^T V ----
    if (p_begin < npairs) accum_tokens_by_cols<NTW>(o, tc::smem_u32(Ls), vs, pitch, R, zc0, p_begin, npairs, lane);
    __syncthreads();
    // ---- dV = P . dT' + dL1 . T + dobar / N  (written over the staged V tile) ----
    if (d_begin < d_end) {
      uint32_t ap[2][4], al[2][4];
      stage_to_afrag(tc::smem_u32(Ps + (size_t)mtile * 16 * PS_PITCH), ap, lane);
      stage_to_afrag(tc::smem_u32(Ls + (size_t)mtile * 16 * PS_PITCH), al, lane);
      const uint32_t b_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * pitch + (lane >> 4) * 16);
      uint8_t* vr0 = Vsm + (size_t)(mtile * 16 + g) * pitch + t * 4;
      uint8_t* vr1 = vr0 + (size_t)8 * pitch;
      const float* dob = a.dobar ? a.dobar + (size_t)b * C : nullptr;
#pragma unroll 2
      for (int c0 = d_begin; c0 < d_end; c0 += 16) {
        float e0[4] = {0.f, 0.f, 0.f, 0.f}, e1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(ds + b_off + (uint32_t)(ks * 16 * pitch + c0 * 2), b0, b1, b2, b3);
          mma16816(e0, ap[ks], b0, b1);
          mma16816(e1, ap[ks], b2, b3);
          ldsm_x4_t(ts + b_off + (uint32_t)(ks * 16 * pitch + c0 * 2), b0, b1, b2, b3);
          mma16816(e0, al[ks], b0, b1);
          mma16816(e1, al[ks], b2, b3);
        }
        const int ca = c0 + 2 * t, cb = ca + 8;
        const float oa0 = dob ? dob[ca] * a.invN : 0.f, oa1 = dob ? dob[ca + 1] * a.invN : 0.f;
        const float ob0 = dob ? dob[cb] * a.invN : 0.f, ob1 = dob ? dob[cb + 1] * a.invN : 0.f;
        *reinterpret_cast<uint32_t*>(vr0 + c0 * 2) = pack_bf16x2(e0[0] + oa0, e0[1] + oa1);
        *reinterpret_cast<uint32_t*>(vr1 + c0 * 2) = pack_bf16x2(e0[2] + oa0, e0[3] + oa1);
        *reinterpret_cast<uint32_t*>(vr0 + c0 * 2 + 16) = pack_bf16x2(e1[0] + ob0, e1[1] + ob1);
        *reinterpret_cast<uint32_t*>(vr1 + c0 * 2 + 16) = pack_bf16x2(e1[2] + ob0, e1[3] + ob1);
      }
    }
    fence_proxy_async();
    __syncthreads();
    if (rlane < rows) {
      bulk_s2g(a.gV + ((size_t)b * N + n0 + rlane) * C + zc0, vs + (uint32_t)(rlane * pitch + zc0 * 2), (uint32_t)zcols * 2);
      bulk_commit();
      bulk_wait_read0();
    }
    if (chunk + a.nbuf < c_end) {
      __syncthreads();                                 // every row of the tile has left: this buffer may be overwritten
      issue_chunk(chunk + a.nbuf, buf);
    }
  }
  if (!__syncthreads_and(ok)) return;
  if (p_begin < npairs && a.d_tok) atomic_flush_tokens<NTW>(o, a.d_tok, C, K, zc0, p_begin, npairs, g, t);
}

struct TokAttnBwdPlan { bool ok; int R, nbuf, nchunks, splits, Z, zcols, ntw; size_t smem; };
inline size_t tok_attn_bwd_smem(int C, int R, int nbuf) {
  const size_t pitch = (size_t)C * 2 + 16;
  return 2 * (size_t)KL * pitch + (size_t)nbuf * R * pitch + 2 * (size_t)R * PS_PITCH + (size_t)(16 * 8 * 32 + 96) * 4 + 64;
}
inline TokAttnBwdPlan make_tok_attn_bwd_plan(int BT, int N, int C, int K) {
  TokAttnBwdPlan p{};
  if (C % 32 || C < 32 || K > KL || K < 1 || N < 1) return p;
  p.Z = (C + 767) / 768;
  p.zcols = ((C / 16 + p.Z - 1) / p.Z) * 16;
  p.ntw = 2 * ((p.zcols / 16 + 7) / 8);
  if (p.ntw > 12) return p;
  if (!fa_bwd_tile(BT, N, C, p.ntw, tok_attn_bwd_smem, p.R, p.nbuf)) return p;
  p.nchunks = (N + p.R - 1) / p.R;
  p.splits = fa_splits(BT, p.Z, N, p.nchunks, tok_attn_bwd_smem(C, p.R, p.nbuf), p.ntw);
  if ((p.nchunks + p.splits - 1) / p.splits < 2) p.nbuf = 1;
  p.smem = tok_attn_bwd_smem(C, p.R, p.nbuf);
  fa_widen_z(BT, p.splits, C, p.smem, 2, 3, p.Z, p.zcols, p.ntw);
  p.ok = true;
  return p;
}
template <int NTW>
inline cudaError_t launch_tok_attn_bwd_t(const TokAttnBwdArgs& a, const TokAttnBwdPlan& pl, int BT, cudaStream_t st) {
  static PerDeviceOnce attr;
  if (attr.first()) {
    cudaError_t e = set_max_dyn_smem(tok_attn_bwd_kernel<NTW>);
    if (e != cudaSuccess) return e;
  }
  tok_attn_bwd_kernel<NTW><<<dim3(pl.splits, BT, pl.Z), 256, pl.smem, st>>>(a);
  return cudaGetLastError();
}
inline cudaError_t launch_tok_attn_bwd(const TokAttnBwdArgs& a, const TokAttnBwdPlan& pl, int BT, cudaStream_t st) {
  switch (pl.ntw) {
    case 2: return launch_tok_attn_bwd_t<2>(a, pl, BT, st);
    case 4: return launch_tok_attn_bwd_t<4>(a, pl, BT, st);
    case 6: return launch_tok_attn_bwd_t<6>(a, pl, BT, st);
    case 8: return launch_tok_attn_bwd_t<8>(a, pl, BT, st);
    case 10: return launch_tok_attn_bwd_t<10>(a, pl, BT, st);
    default: return launch_tok_attn_bwd_t<12>(a, pl, BT, st);
  }
}

}  // namespace fa
}  // namespace dgsct
