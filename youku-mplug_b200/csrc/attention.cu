// Fused attention (softmax(scale * Q K^T [+causal]) V) forward and backward.
//
// Attention is ~3 % of the path's FLOPs, so these kernels use warp-level
// mma.sync (m16n8k16 bf16, fp32 accumulate) flash-style tiles; the score matrix never touches
// HBM.  The tcgen05 budget is spent on the GEMMs.  Sequences are addressed through ymp_seqmap
// so the kernels read Q/K/V straight out of packed QKV GEMM outputs in any of the path's
// layouts (ViT [3,heads,hd], GPT per-head [q|k|v], TimeSformer per-frame sequences with a
// shared cls row, abstractor cross attention).
#include <math_constants.h>

#include "common.h"
#include "ptx.cuh"

namespace ymp {

struct SeqMap {
  int seq_div, n_prefix, prefix_per_seq;
  long outer_stride, inner_stride, pos_stride, prefix_base, prefix_stride;
};
static SeqMap to_map(const ymp_seqmap& m) {
  SeqMap r;
  r.seq_div = m.seq_div > 0 ? m.seq_div : 1;
  r.n_prefix = m.n_prefix; r.prefix_per_seq = m.prefix_per_seq;
  r.outer_stride = m.outer_stride; r.inner_stride = m.inner_stride; r.pos_stride = m.pos_stride;
  r.prefix_base = m.prefix_base; r.prefix_stride = m.prefix_stride;
  return r;
}
__device__ __forceinline__ long map_row(const SeqMap& m, int s, int i) {
  const int outer = s / m.seq_div, inner = s - outer * m.seq_div;
  if (i < m.n_prefix) return m.prefix_base + (long)(m.prefix_per_seq ? s : outer) * m.prefix_stride + i;
  return (long)outer * m.outer_stride + (long)inner * m.inner_stride + (long)(i - m.n_prefix) * m.pos_stride;
}

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct AttnKParams {
  const __nv_bfloat16 *q, *k, *v;
  __nv_bfloat16* o;
  float* lse;
  int ldq, ldk, ldv, ldo, hsq, hsk, hsv, hso;
  SeqMap mq, mkv, mo;
  int n_seq, n_heads, s_q, s_kv, causal;
  float scale_log2, scale;
  // backward only
  const __nv_bfloat16* dout;
  __nv_bfloat16 *dq, *dk, *dv;
  int lddo, hsdo, lddq, lddk, lddv, hsdq, hsdk, hsdv;
  SeqMap mdo, mdq, mdkv;
};

// Load a [64 x D] bf16 tile (rows r0..r0+63 of sequence s through `m`) into padded smem.
template <int D>
__device__ __forceinline__ void load_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, const SeqMap& m,
                                          int s, int r0, int n_valid, int ld, int col_off) {
  constexpr int CH = D / 8, LDS = D + 8;
  for (int idx = threadIdx.x; idx < 64 * CH; idx += blockDim.x) {
    const int r = idx / CH, c = idx - r * CH;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (r0 + r < n_valid)
      val = __ldg(reinterpret_cast<const uint4*>(src + map_row(m, s, r0 + r) * ld + col_off + c * 8));
    *reinterpret_cast<uint4*>(dst + r * LDS + c * 8) = val;
  }
}

// ------------------------------------------------------------------------------ forward
template <int D>
__global__ void __launch_bounds__(128) attn_fwd_kernel(const AttnKParams p) {
  constexpr int LDS = D + 8, KS = D / 16;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* Ks = Qs + 64 * LDS;
  __nv_bfloat16* Vs = Ks + 64 * LDS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 64, h = blockIdx.y, s = blockIdx.z;
  const int g = lane >> 2, t4 = lane & 3;

  load_tile<D>(Qs, p.q, p.mq, s, q0, p.s_q, p.ldq, h * p.hsq);
  __syncthreads();
  uint32_t qf[KS][4];
#pragma unroll
  for (int kk = 0; kk < KS; ++kk)
    ldsm_x4(qf[kk], smem_u32(Qs + (warp * 16 + (lane & 15)) * LDS + kk * 16 + (lane >> 4) * 8));

  float o_acc[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) { o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f; }
  float m_i[2] = {-CUDART_INF_F, -CUDART_INF_F}, l_i[2] = {0.f, 0.f};
  const int kv_end = p.causal ? min(p.s_kv, q0 + 64) : p.s_kv;

  for (int kv0 = 0; kv0 < kv_end; kv0 += 64) {
    __syncthreads();
    load_tile<D>(Ks, p.k, p.mkv, s, kv0, p.s_kv, p.ldk, h * p.hsk);
    load_tile<D>(Vs, p.v, p.mkv, s, kv0, p.s_kv, p.ldv, h * p.hsv);
    __syncthreads();
    float sc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < KS; ++kk) {
#pragma unroll
      for (int nbp = 0; nbp < 4; ++nbp) {
        uint32_t b[4];
        ldsm_x4(b, smem_u32(Ks + (nbp * 16 + (lane & 7) + (lane >> 4) * 8) * LDS + kk * 16 + ((lane >> 3) & 1) * 8));
        mma16816(sc[2 * nbp], qf[kk], b[0], b[1]);
        mma16816(sc[2 * nbp + 1], qf[kk], b[2], b[3]);
      }
    }
    float mx[2] = {-CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = kv0 + nb * 8 + t4 * 2 + (e & 1);
        const int row = q0 + warp * 16 + g + (e >> 1) * 8;
        float v = sc[nb][e] * p.scale_log2;
        if (col >= p.s_kv || (p.causal && col > row)) v = -CUDART_INF_F;
        sc[nb][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
    float alpha[2], msafe[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float mnew = fmaxf(m_i[r], mx[r]);
      msafe[r] = (mnew == -CUDART_INF_F) ? 0.f : mnew;
      alpha[r] = exp2f(m_i[r] - msafe[r]);
      m_i[r] = mnew;
      l_i[r] *= alpha[r];
    }
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = exp2f(sc[nb][e] - msafe[e >> 1]);
        sc[nb][e] = pv;
        l_i[e >> 1] += pv;
      }
    }
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      o_acc[i][0] *= alpha[0]; o_acc[i][1] *= alpha[0];
      o_acc[i][2] *= alpha[1]; o_acc[i][3] *= alpha[1];
    }
#pragma unroll
    for (int kk2 = 0; kk2 < 4; ++kk2) {
      uint32_t pa[4];
      pa[0] = pack_bf16(sc[2 * kk2][0], sc[2 * kk2][1]);
      pa[1] = pack_bf16(sc[2 * kk2][2], sc[2 * kk2][3]);
      pa[2] = pack_bf16(sc[2 * kk2 + 1][0], sc[2 * kk2 + 1][1]);
      pa[3] = pack_bf16(sc[2 * kk2 + 1][2], sc[2 * kk2 + 1][3]);
#pragma unroll
      for (int dbp = 0; dbp < KS; ++dbp) {
        uint32_t b[4];
        ldsm_x4_t(b, smem_u32(Vs + (kk2 * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + dbp * 16 + (lane >> 4) * 8));
        mma16816(o_acc[2 * dbp], pa, b[0], b[1]);
        mma16816(o_acc[2 * dbp + 1], pa, b[2], b[3]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 1);
    l_i[r] += __shfl_xor_sync(0xffffffffu, l_i[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int qi = q0 + warp * 16 + g + r * 8;
    if (qi >= p.s_q) continue;
    const float inv = l_i[r] > 0.f ? 1.f / l_i[r] : 0.f;
    __nv_bfloat16* orow = p.o + map_row(p.mo, s, qi) * p.ldo + h * p.hso;
#pragma unroll
    for (int nb = 0; nb < D / 8; ++nb)
      *reinterpret_cast<uint32_t*>(orow + nb * 8 + t4 * 2) =
          pack_bf16(o_acc[nb][2 * r] * inv, o_acc[nb][2 * r + 1] * inv);
    if (p.lse && t4 == 0)
      p.lse[((size_t)s * p.n_heads + h) * p.s_q + qi] = m_i[r] * 0.6931471805599453f + logf(l_i[r]);
  }
}

// ------------------------------------------------------------------------------ backward
// One CTA per (sequence, head).  kv tiles outer (dK/dV of a warp's 16 kv rows live in
// registers), q tiles inner; dQ accumulates in shared memory (fp32) - no atomics, deterministic.
template <int D>
__global__ void __launch_bounds__(128) attn_bwd_kernel(const AttnKParams p, const int sq_pad) {
  constexpr int LDS = D + 8, KS = D / 16, LDD = 72;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* Ks = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* Vs = Ks + 64 * LDS;
  __nv_bfloat16* Qs = Vs + 64 * LDS;
  __nv_bfloat16* dOs = Qs + 64 * LDS;
  __nv_bfloat16* dSs = dOs + 64 * LDS;            // [64 kv][LDD] holds dS^T
  float* lse_s = reinterpret_cast<float*>(dSs + 64 * LDD);
  float* delta_s = lse_s + sq_pad;
  float* dQacc = delta_s + sq_pad;                // [sq_pad][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, s = blockIdx.y;
  const int g = lane >> 2, t4 = lane & 3;

  for (int i = threadIdx.x; i < sq_pad * D; i += blockDim.x) dQacc[i] = 0.f;
  // delta[i] = sum_d dO[i,d] * O[i,d];   lse in log2 units
  for (int i = warp; i < sq_pad; i += 4) {
    float acc = 0.f;
    if (i < p.s_q) {
      const __nv_bfloat16* orow = p.o + map_row(p.mo, s, i) * p.ldo + h * p.hso;
      const __nv_bfloat16* drow = p.dout + map_row(p.mdo, s, i) * p.lddo + h * p.hsdo;
      for (int d = lane * 2; d < D; d += 64) {
        const uint32_t a = *reinterpret_cast<const uint32_t*>(orow + d);
        const uint32_t b = *reinterpret_cast<const uint32_t*>(drow + d);
        acc += bf16_lo(a) * bf16_lo(b) + bf16_hi(a) * bf16_hi(b);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      delta_s[i] = acc;
      lse_s[i] = (i < p.s_q) ? p.lse[((size_t)s * p.n_heads + h) * p.s_q + i] * 1.4426950408889634f
                             : CUDART_INF_F;
    }
  }

  for (int kv0 = 0; kv0 < p.s_kv; kv0 += 64) {
    __syncthreads();
    load_tile<D>(Ks, p.k, p.mkv, s, kv0, p.s_kv, p.ldk, h * p.hsk);
    load_tile<D>(Vs, p.v, p.mkv, s, kv0, p.s_kv, p.ldv, h * p.hsv);
    float dk_acc[D / 8][4], dv_acc[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      dk_acc[i][0] = dk_acc[i][1] = dk_acc[i][2] = dk_acc[i][3] = 0.f;
      dv_acc[i][0] = dv_acc[i][1] = dv_acc[i][2] = dv_acc[i][3] = 0.f;
    }
    const int qt0 = p.causal ? (kv0 / 64) * 64 : 0;
    for (int qi0 = qt0; qi0 < p.s_q; qi0 += 64) {
      __syncthreads();
      load_tile<D>(Qs, p.q, p.mq, s, qi0, p.s_q, p.ldq, h * p.hsq);
      load_tile<D>(dOs, p.dout, p.mdo, s, qi0, p.s_q, p.lddo, h * p.hsdo);
      __syncthreads();
      // S^T (16 kv x 64 q) = K_w Q^T ; dP^T = V_w dO^T
      float st[8][4], dpt[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
        dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
      }
#pragma unroll
      for (int kk = 0; kk < KS; ++kk) {
        uint32_t kf[4], vf[4];
        ldsm_x4(kf, smem_u32(Ks + (warp * 16 + (lane & 15)) * LDS + kk * 16 + (lane >> 4) * 8));
        ldsm_x4(vf, smem_u32(Vs + (warp * 16 + (lane & 15)) * LDS + kk * 16 + (lane >> 4) * 8));
#pragma unroll
        for (int nbp = 0; nbp < 4; ++nbp) {
          uint32_t b[4];
          const int off = (nbp * 16 + (lane & 7) + (lane >> 4) * 8) * LDS + kk * 16 + ((lane >> 3) & 1) * 8;
          ldsm_x4(b, smem_u32(Qs + off));
          mma16816(st[2 * nbp], kf, b[0], b[1]);
          mma16816(st[2 * nbp + 1], kf, b[2], b[3]);
          ldsm_x4(b, smem_u32(dOs + off));
          mma16816(dpt[2 * nbp], vf, b[0], b[1]);
          mma16816(dpt[2 * nbp + 1], vf, b[2], b[3]);
        }
      }
      // P^T and dS^T
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int kvr = kv0 + warp * 16 + g + (e >> 1) * 8;
          const int qc = qi0 + nb * 8 + t4 * 2 + (e & 1);
          float pv = exp2f(st[nb][e] * p.scale_log2 - lse_s[qc]);
          if (kvr >= p.s_kv || (p.causal && kvr > qc)) pv = 0.f;
          st[nb][e] = pv;
          dpt[nb][e] = pv * (dpt[nb][e] - delta_s[qc]);
        }
      }
      // dV += P^T dO ; dK += dS^T Q   (k dimension = the 64 q of this tile)
#pragma unroll
      for (int kk2 = 0; kk2 < 4; ++kk2) {
        uint32_t pa[4], da[4];
        pa[0] = pack_bf16(st[2 * kk2][0], st[2 * kk2][1]);
        pa[1] = pack_bf16(st[2 * kk2][2], st[2 * kk2][3]);
        pa[2] = pack_bf16(st[2 * kk2 + 1][0], st[2 * kk2 + 1][1]);
        pa[3] = pack_bf16(st[2 * kk2 + 1][2], st[2 * kk2 + 1][3]);
        da[0] = pack_bf16(dpt[2 * kk2][0], dpt[2 * kk2][1]);
        da[1] = pack_bf16(dpt[2 * kk2][2], dpt[2 * kk2][3]);
        da[2] = pack_bf16(dpt[2 * kk2 + 1][0], dpt[2 * kk2 + 1][1]);
        da[3] = pack_bf16(dpt[2 * kk2 + 1][2], dpt[2 * kk2 + 1][3]);
        // stash dS^T (bf16) for the dQ product
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          *reinterpret_cast<uint32_t*>(dSs + (warp * 16 + g) * LDD + (2 * kk2 + hb) * 8 + t4 * 2) = da[2 * hb];
          *reinterpret_cast<uint32_t*>(dSs + (warp * 16 + g + 8) * LDD + (2 * kk2 + hb) * 8 + t4 * 2) = da[2 * hb + 1];
        }
#pragma unroll
        for (int dbp = 0; dbp < KS; ++dbp) {
          uint32_t b[4];
          const int off = (kk2 * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + dbp * 16 + (lane >> 4) * 8;
          ldsm_x4_t(b, smem_u32(dOs + off));
          mma16816(dv_acc[2 * dbp], pa, b[0], b[1]);
          mma16816(dv_acc[2 * dbp + 1], pa, b[2], b[3]);
          ldsm_x4_t(b, smem_u32(Qs + off));
          mma16816(dk_acc[2 * dbp], da, b[0], b[1]);
          mma16816(dk_acc[2 * dbp + 1], da, b[2], b[3]);
        }
      }
      __syncthreads();
      // dQ[qi0 + 16*warp .. +15][:] += dS(16 q x 64 kv) K(64 kv x D)
      {
        float* dq_rows = dQacc + (size_t)(qi0 + warp * 16) * D;
        float acc[D / 8][4];
#pragma unroll
        for (int nb = 0; nb < D / 8; ++nb) {
          const float2 lo = *reinterpret_cast<const float2*>(dq_rows + g * D + nb * 8 + t4 * 2);
          const float2 hi = *reinterpret_cast<const float2*>(dq_rows + (g + 8) * D + nb * 8 + t4 * 2);
          acc[nb][0] = lo.x; acc[nb][1] = lo.y; acc[nb][2] = hi.x; acc[nb][3] = hi.y;
        }
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          uint32_t a[4];
          const int mi = lane >> 3, r = lane & 7;
          ldsm_x4_t(a, smem_u32(dSs + (kb * 16 + ((mi >> 1) & 1) * 8 + r) * LDD + warp * 16 + (mi & 1) * 8));
#pragma unroll
          for (int dbp = 0; dbp < KS; ++dbp) {
            uint32_t b[4];
            ldsm_x4_t(b, smem_u32(Ks + (kb * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + dbp * 16 + (lane >> 4) * 8));
            mma16816(acc[2 * dbp], a, b[0], b[1]);
            mma16816(acc[2 * dbp + 1], a, b[2], b[3]);
          }
        }
#pragma unroll
        for (int nb = 0; nb < D / 8; ++nb) {
          *reinterpret_cast<float2*>(dq_rows + g * D + nb * 8 + t4 * 2) = make_float2(acc[nb][0], acc[nb][1]);
          *reinterpret_cast<float2*>(dq_rows + (g + 8) * D + nb * 8 + t4 * 2) = make_float2(acc[nb][2], acc[nb][3]);
        }
      }
    }
    // write dK (scaled) and dV for this warp's 16 kv rows
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int kvr = kv0 + warp * 16 + g + r * 8;
      if (kvr >= p.s_kv) continue;
      const long row = map_row(p.mdkv, s, kvr);
      __nv_bfloat16* dkrow = p.dk + row * p.lddk + h * p.hsdk;
      __nv_bfloat16* dvrow = p.dv + row * p.lddv + h * p.hsdv;
#pragma unroll
      for (int nb = 0; nb < D / 8; ++nb) {
        *reinterpret_cast<uint32_t*>(dkrow + nb * 8 + t4 * 2) =
            pack_bf16(dk_acc[nb][2 * r] * p.scale, dk_acc[nb][2 * r + 1] * p.scale);
        *reinterpret_cast<uint32_t*>(dvrow + nb * 8 + t4 * 2) = pack_bf16(dv_acc[nb][2 * r], dv_acc[nb][2 * r + 1]);
      }
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < p.s_q * (D / 2); idx += blockDim.x) {
    const int i = idx / (D / 2), d = (idx - i * (D / 2)) * 2;
    __nv_bfloat16* dqrow = p.dq + map_row(p.mdq, s, i) * p.lddq + h * p.hsdq;
    *reinterpret_cast<uint32_t*>(dqrow + d) = pack_bf16(dQacc[i * D + d] * p.scale, dQacc[i * D + d + 1] * p.scale);
  }
}

// ------------------------------------------------------------------------------ tiny sequences
// TimeSformer temporal attention: S = num_frames (<= 16).  One warp per (sequence, head); SIMT,
// HBM-bound (reads q,k,v once, writes o once).
constexpr int SMALL_MAX_S = 16;

struct SmallParams {
  const __nv_bfloat16 *q, *k, *v, *dout;
  __nv_bfloat16 *o, *dq, *dk, *dv;
  int ld, hs, ldo, hso, ldd, hsd;  // qkv share ld/head stride; o/dout share; dq/dk/dv share
  int n_seq, n_heads, S, D;
  float scale;
};

// rows of sequence s are s*S .. s*S+S-1 (dense); per-warp smem: q,k,v(,do) as fp32 [S][D] + p[S][S]
template <bool BWD>
__global__ void __launch_bounds__(128) attn_small_dense_kernel(const SmallParams p) {
  extern __shared__ __align__(16) uint8_t smem_attn[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S, D = p.D, R = p.D + 1;  // R: padded smem row (bank-conflict free)
  const int per_warp = (BWD ? 4 : 3) * S * R + 2 * S * S;
  float* base = reinterpret_cast<float*>(smem_attn) + (size_t)warp * per_warp;
  float* qs = base;
  float* ks = qs + S * R;
  float* vs = ks + S * R;
  float* dos = vs + S * R;  // only BWD
  float* ps = base + (BWD ? 4 : 3) * S * R;
  float* dss = ps + S * S;
  const long total = (long)p.n_seq * p.n_heads;
  for (long w = (long)blockIdx.x * 4 + warp; w < total; w += (long)gridDim.x * 4) {
    const int s = (int)(w / p.n_heads), h = (int)(w % p.n_heads);
    const long row0 = (long)s * S;
    __syncwarp();
    for (int idx = lane; idx < S * (D / 2); idx += 32) {
      const int i = idx / (D / 2), d = (idx - i * (D / 2)) * 2;
      const size_t off = (size_t)(row0 + i) * p.ld + (size_t)h * p.hs + d;
      uint32_t a = *reinterpret_cast<const uint32_t*>(p.q + off);
      qs[i * R + d] = bf16_lo(a); qs[i * R + d + 1] = bf16_hi(a);
      a = *reinterpret_cast<const uint32_t*>(p.k + off);
      ks[i * R + d] = bf16_lo(a); ks[i * R + d + 1] = bf16_hi(a);
      a = *reinterpret_cast<const uint32_t*>(p.v + off);
      vs[i * R + d] = bf16_lo(a); vs[i * R + d + 1] = bf16_hi(a);
      if (BWD) {
        a = *reinterpret_cast<const uint32_t*>(p.dout + (size_t)(row0 + i) * p.ldo + (size_t)h * p.hso + d);
        dos[i * R + d] = bf16_lo(a); dos[i * R + d + 1] = bf16_hi(a);
      }
    }
    __syncwarp();
    // scores + softmax: lane handles (i, j) pairs
    for (int ij = lane; ij < S * S; ij += 32) {
      const int i = ij / S, j = ij - i * S;
      float acc = 0.f;
      for (int d = 0; d < D; ++d) acc = fmaf(qs[i * R + d], ks[j * R + d], acc);
      ps[ij] = acc * p.scale;
    }
    __syncwarp();
    for (int i = lane; i < S; i += 32) {
      float mx = -CUDART_INF_F;
      for (int j = 0; j < S; ++j) mx = fmaxf(mx, ps[i * S + j]);
      float sum = 0.f;
      for (int j = 0; j < S; ++j) { const float e = __expf(ps[i * S + j] - mx); ps[i * S + j] = e; sum += e; }
      const float inv = 1.f / sum;
      for (int j = 0; j < S; ++j) ps[i * S + j] = __bfloat162float(__float2bfloat16(ps[i * S + j] * inv));
    }
    __syncwarp();
    if (!BWD) {
      for (int idx = lane; idx < S * (D / 2); idx += 32) {
        const int i = idx / (D / 2), d = (idx - i * (D / 2)) * 2;
        float a0 = 0.f, a1 = 0.f;
        for (int j = 0; j < S; ++j) {
          a0 = fmaf(ps[i * S + j], vs[j * R + d], a0);
          a1 = fmaf(ps[i * S + j], vs[j * R + d + 1], a1);
        }
        *reinterpret_cast<uint32_t*>(p.o + (size_t)(row0 + i) * p.ldo + (size_t)h * p.hso + d) = pack_bf16(a0, a1);
      }
    } else {
      // dP = dO V^T ; dS = P * (dP - rowsum(P*dP))
      for (int ij = lane; ij < S * S; ij += 32) {
        const int i = ij / S, j = ij - i * S;
        float acc = 0.f;
        for (int d = 0; d < D; ++d) acc = fmaf(dos[i * R + d], vs[j * R + d], acc);
        dss[ij] = acc;
      }
      __syncwarp();
      for (int i = lane; i < S; i += 32) {
        float dot = 0.f;
        for (int j = 0; j < S; ++j) dot = fmaf(ps[i * S + j], dss[i * S + j], dot);
        for (int j = 0; j < S; ++j) dss[i * S + j] = ps[i * S + j] * (dss[i * S + j] - dot);
      }
      __syncwarp();
      for (int idx = lane; idx < S * (D / 2); idx += 32) {
        const int i = idx / (D / 2), d = (idx - i * (D / 2)) * 2;
        float q0 = 0.f, q1 = 0.f, k0 = 0.f, k1 = 0.f, v0 = 0.f, v1 = 0.f;
        for (int j = 0; j < S; ++j) {
          const float ds_ij = dss[i * S + j], ds_ji = dss[j * S + i], p_ji = ps[j * S + i];
          q0 = fmaf(ds_ij, ks[j * R + d], q0); q1 = fmaf(ds_ij, ks[j * R + d + 1], q1);
          k0 = fmaf(ds_ji, qs[j * R + d], k0); k1 = fmaf(ds_ji, qs[j * R + d + 1], k1);
          v0 = fmaf(p_ji, dos[j * R + d], v0); v1 = fmaf(p_ji, dos[j * R + d + 1], v1);
        }
        const size_t off = (size_t)(row0 + i) * p.ldd + (size_t)h * p.hsd + d;
        *reinterpret_cast<uint32_t*>(p.dq + off) = pack_bf16(q0 * p.scale, q1 * p.scale);
        *reinterpret_cast<uint32_t*>(p.dk + off) = pack_bf16(k0 * p.scale, k1 * p.scale);
        *reinterpret_cast<uint32_t*>(p.dv + off) = pack_bf16(v0, v1);
      }
    }
  }
}

static int fill_params(const ymp_attn_args* a, AttnKParams& p, const char* who) {
  YMP_CHECK_ARG(a && a->q && a->k && a->v, "%s: null q/k/v", who);
  YMP_CHECK_ARG(a->head_dim == 64 || a->head_dim == 80 || a->head_dim == 96 || a->head_dim == 128,
                "%s: head_dim %d not in {64,80,96,128}", who, a->head_dim);
  YMP_CHECK_ARG(a->n_seq > 0 && a->n_heads > 0 && a->s_q > 0 && a->s_kv > 0, "%s: bad sizes", who);
  YMP_CHECK_ARG(a->ldq % 8 == 0 && a->ldk % 8 == 0 && a->ldv % 8 == 0 && a->ldo % 8 == 0, "%s: row strides must be multiples of 8", who);
  YMP_CHECK_ARG(a->q_head_stride % 8 == 0 && a->k_head_stride % 8 == 0 && a->v_head_stride % 8 == 0 && a->o_head_stride % 8 == 0, "%s: head strides must be multiples of 8", who);
  YMP_CHECK_ARG(aligned16(a->q) && aligned16(a->k) && aligned16(a->v), "%s: q/k/v must be 16-byte aligned", who);
  YMP_CHECK_ARG(!a->causal || a->s_q == a->s_kv, "%s: causal needs s_q == s_kv", who);
  p.q = (const __nv_bfloat16*)a->q; p.k = (const __nv_bfloat16*)a->k; p.v = (const __nv_bfloat16*)a->v;
  p.o = (__nv_bfloat16*)a->o; p.lse = a->lse;
  p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv; p.ldo = a->ldo;
  p.hsq = a->q_head_stride; p.hsk = a->k_head_stride; p.hsv = a->v_head_stride; p.hso = a->o_head_stride;
  p.mq = to_map(a->map_q); p.mkv = to_map(a->map_kv); p.mo = to_map(a->map_o);
  p.n_seq = a->n_seq; p.n_heads = a->n_heads; p.s_q = a->s_q; p.s_kv = a->s_kv; p.causal = a->causal;
  p.scale = a->scale; p.scale_log2 = a->scale * 1.4426950408889634f;
  return YMP_OK;
}

template <int D>
static int launch_fwd(const AttnKParams& p, cudaStream_t st) {
  const int smem = 3 * 64 * (D + 8) * 2;
  static bool set = false;
  if (!set) { YMP_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); set = true; }
  dim3 grid((p.s_q + 63) / 64, p.n_heads, p.n_seq);
  attn_fwd_kernel<D><<<grid, 128, smem, st>>>(p);
  YMP_LAUNCH_CHECK();
  return YMP_OK;
}
template <int D>
static int launch_bwd(const AttnKParams& p, cudaStream_t st) {
  const int sq_pad = (p.s_q + 63) / 64 * 64;
  const size_t smem = (size_t)4 * 64 * (D + 8) * 2 + 64 * 72 * 2 + (size_t)2 * sq_pad * 4 + (size_t)sq_pad * D * 4;
  if (smem > 227 * 1024) return set_error(YMP_ENOSUP, "ymp_attn_bwd: s_q=%d too long for the smem-resident dQ (head_dim %d)", p.s_q, D);
  static size_t cur = 0;
  if (smem > cur) { YMP_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cur = smem; }
  dim3 grid(p.n_heads, p.n_seq);
  attn_bwd_kernel<D><<<grid, 128, smem, st>>>(p, sq_pad);
  YMP_LAUNCH_CHECK();
  return YMP_OK;
}

}  // namespace ymp

extern "C" int ymp_attn_fwd(const ymp_attn_args* a, void* stream) {
  using namespace ymp;
  AttnKParams p = {};
  int rc = fill_params(a, p, "ymp_attn_fwd");
  if (rc) return rc;
  YMP_CHECK_ARG(a->o && aligned16(a->o), "ymp_attn_fwd: bad o");
  cudaStream_t st = (cudaStream_t)stream;
  switch (a->head_dim) {
    case 64: return launch_fwd<64>(p, st);
    case 80: return launch_fwd<80>(p, st);
    case 96: return launch_fwd<96>(p, st);
    default: return launch_fwd<128>(p, st);
  }
}

extern "C" int ymp_attn_bwd(const ymp_attn_bwd_args* b, void* stream) {
  using namespace ymp;
  YMP_CHECK_ARG(b != nullptr, "ymp_attn_bwd: null args");
  const ymp_attn_args* a = &b->fwd;
  AttnKParams p = {};
  int rc = fill_params(a, p, "ymp_attn_bwd");
  if (rc) return rc;
  YMP_CHECK_ARG(a->o && a->lse && b->dout && b->dq && b->dk && b->dv, "ymp_attn_bwd: null o/lse/dout/dq/dk/dv");
  YMP_CHECK_ARG(b->lddo % 8 == 0 && b->lddq % 8 == 0 && b->lddk % 8 == 0 && b->lddv % 8 == 0, "ymp_attn_bwd: grad row strides must be multiples of 8");
  p.dout = (const __nv_bfloat16*)b->dout; p.dq = (__nv_bfloat16*)b->dq; p.dk = (__nv_bfloat16*)b->dk; p.dv = (__nv_bfloat16*)b->dv;
  p.lddo = b->lddo; p.hsdo = b->do_head_stride;
  p.lddq = b->lddq; p.lddk = b->lddk; p.lddv = b->lddv;
  p.hsdq = b->dq_head_stride; p.hsdk = b->dk_head_stride; p.hsdv = b->dv_head_stride;
  p.mdo = to_map(b->map_do); p.mdq = to_map(b->map_dq); p.mdkv = to_map(b->map_dkv);
  cudaStream_t st = (cudaStream_t)stream;
  switch (a->head_dim) {
    case 64: return launch_bwd<64>(p, st);
    case 80: return launch_bwd<80>(p, st);
    case 96: return launch_bwd<96>(p, st);
    default: return launch_bwd<128>(p, st);
  }
}

static int small_common(const ymp_attn_small_args* a, ymp::SmallParams& p, bool bwd, const char* who) {
  using namespace ymp;
  YMP_CHECK_ARG(a && a->q && a->k && a->v, "%s: null q/k/v", who);
  YMP_CHECK_ARG(a->S > 0 && a->S <= SMALL_MAX_S, "%s: S=%d must be in [1,%d]", who, a->S, SMALL_MAX_S);
  YMP_CHECK_ARG(a->D > 0 && a->D % 2 == 0 && a->D <= 128, "%s: bad head_dim %d", who, a->D);
  YMP_CHECK_ARG(a->n_seq > 0 && a->n_heads > 0, "%s: bad sizes", who);
  p.q = (const __nv_bfloat16*)a->q; p.k = (const __nv_bfloat16*)a->k; p.v = (const __nv_bfloat16*)a->v;
  p.o = (__nv_bfloat16*)a->o; p.dout = (const __nv_bfloat16*)a->dout;
  p.dq = (__nv_bfloat16*)a->dq; p.dk = (__nv_bfloat16*)a->dk; p.dv = (__nv_bfloat16*)a->dv;
  p.ld = a->ld; p.hs = a->head_stride; p.ldo = a->ldo; p.hso = a->o_head_stride; p.ldd = a->ldd; p.hsd = a->d_head_stride;
  p.n_seq = a->n_seq; p.n_heads = a->n_heads; p.S = a->S; p.D = a->D; p.scale = a->scale;
  if (bwd) YMP_CHECK_ARG(a->dout && a->dq && a->dk && a->dv, "%s: null grads", who);
  else YMP_CHECK_ARG(a->o != nullptr, "%s: null o", who);
  return YMP_OK;
}

extern "C" int ymp_attn_small_fwd(const ymp_attn_small_args* a, void* stream) {
  using namespace ymp;
  SmallParams p = {};
  int rc = small_common(a, p, false, "ymp_attn_small_fwd");
  if (rc) return rc;
  const size_t smem = (size_t)4 * (3 * p.S * (p.D + 1) + 2 * p.S * p.S) * 4;
  static size_t cur = 0;
  if (smem > cur) { YMP_CUDA(cudaFuncSetAttribute(attn_small_dense_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cur = smem; }
  const long total = (long)p.n_seq * p.n_heads;
  const int blocks = (int)min((total + 3) / 4, (long)num_sms() * 8);
  attn_small_dense_kernel<false><<<blocks, 128, smem, (cudaStream_t)stream>>>(p);
  YMP_LAUNCH_CHECK();
  return YMP_OK;
}

extern "C" int ymp_attn_small_bwd(const ymp_attn_small_args* a, void* stream) {
  using namespace ymp;
  SmallParams p = {};
  int rc = small_common(a, p, true, "ymp_attn_small_bwd");
  if (rc) return rc;
  const size_t smem = (size_t)4 * (4 * p.S * (p.D + 1) + 2 * p.S * p.S) * 4;
  static size_t cur = 0;
  if (smem > cur) { YMP_CUDA(cudaFuncSetAttribute(attn_small_dense_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cur = smem; }
  const long total = (long)p.n_seq * p.n_heads;
  const int blocks = (int)min((total + 3) / 4, (long)num_sms() * 8);
  attn_small_dense_kernel<true><<<blocks, 128, smem, (cudaStream_t)stream>>>(p);
  YMP_LAUNCH_CHECK();
  return YMP_OK;
}
