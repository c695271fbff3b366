// Skinny GEMM for single-token decoding: y[M, N] = epilogue(x[M, K] . w[N, K]^T) with M <= 8 rows
// (beam width / decode batch).  Replaces the GPT-3 layer's linears at the KV-cache step of
// models/modeling_distributed_gpt3.py:868-938,580-595,1348-1350.  With a handful of rows the work is one pass over
// the weight matrix: HBM-bound, so no tensor cores - a 128-row tcgen05 tile would occupy N/128 CTAs (16 for N = 2048)
// and stream the weights at a fraction of the memory bandwidth.
//   * a warp owns FOUR output columns and one K slice; its lanes stride the slice with 128-bit loads (four independent
//     weight streams per lane per iteration), x is re-read through L1/L2 (M*K*2 bytes, shared by every CTA);
//   * the 8 warps of a CTA are ksplit K-slices x (8 / ksplit) column groups: the host picks ksplit so that even the
//     N = hidden GEMMs put >= 3 CTAs on every SM (enough loads in flight to cover the HBM latency);
//   * partial sums are combined with warp shuffles (+ shared memory across K slices) and lane (m, col) applies
//       v = acc + bias[n] ; v = gelu(v) (optional) ; v += residual[m, n] (bf16 or fp32) ; store bf16 or fp32.
// Algorithmic bytes per call: N*K*2 (weights) + M*(K + N)*2..4.
#include "common.h"
#include "ptx.cuh"

namespace ymp {

constexpr int SK_MAXM = 8, SK_COLS = 4, SK_WARPS = 8;

struct SkinnyParams {
  const __nv_bfloat16* x;
  const __nv_bfloat16* w;
  const __nv_bfloat16* bias;
  const void* residual;
  void* y;
  int M, N, K, ldx, ldw, ldr, ldy;
  int act, res_f32, out_f32, ksplit;
};

__device__ __forceinline__ void fma8(float& acc, const uint4& x, const uint4& w) {
  acc = fmaf(bf16_lo(x.x), bf16_lo(w.x), acc); acc = fmaf(bf16_hi(x.x), bf16_hi(w.x), acc);
  acc = fmaf(bf16_lo(x.y), bf16_lo(w.y), acc); acc = fmaf(bf16_hi(x.y), bf16_hi(w.y), acc);
  acc = fmaf(bf16_lo(x.z), bf16_lo(w.z), acc); acc = fmaf(bf16_hi(x.z), bf16_hi(w.z), acc);
  acc = fmaf(bf16_lo(x.w), bf16_lo(w.w), acc); acc = fmaf(bf16_hi(x.w), bf16_hi(w.w), acc);
}

__global__ void __launch_bounds__(SK_WARPS * 32) gemm_skinny_kernel(const SkinnyParams p) {
  __shared__ float part[SK_WARPS][SK_MAXM * SK_COLS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ks = p.ksplit, kpart = warp % ks, cgrp = warp / ks, groups = SK_WARPS / ks;
  const int n0 = (blockIdx.x * groups + cgrp) * SK_COLS;
  const bool live = n0 < p.N;
  float acc[SK_MAXM][SK_COLS];
#pragma unroll
  for (int m = 0; m < SK_MAXM; ++m)
#pragma unroll
    for (int c = 0; c < SK_COLS; ++c) acc[m][c] = 0.f;
  if (live) {
    const uint4* wr[SK_COLS];
#pragma unroll
    for (int c = 0; c < SK_COLS; ++c) wr[c] = reinterpret_cast<const uint4*>(p.w + (size_t)min(n0 + c, p.N - 1) * p.ldw);
    const int nvec = p.K >> 3;
    const int v_lo = (int)((long)nvec * kpart / ks), v_hi = (int)((long)nvec * (kpart + 1) / ks);
#pragma unroll 2
    for (int v = v_lo + lane; v < v_hi; v += 32) {
      uint4 wv[SK_COLS];
#pragma unroll
      for (int c = 0; c < SK_COLS; ++c) wv[c] = ld_nc_v4(wr[c] + v);
#pragma unroll
      for (int m = 0; m < SK_MAXM; ++m) {
        if (m < p.M) {
          const uint4 xv = __ldg(reinterpret_cast<const uint4*>(p.x + (size_t)m * p.ldx) + v);
#pragma unroll
          for (int c = 0; c < SK_COLS; ++c) fma8(acc[m][c], xv, wv[c]);
        }
      }
    }
  }
  // lane (m * 4 + c) ends up with this warp's sum of (row m, column n0 + c)
  float mine = 0.f;
#pragma unroll
  for (int m = 0; m < SK_MAXM; ++m) {
#pragma unroll
    for (int c = 0; c < SK_COLS; ++c) {
      const float s = warp_sum(acc[m][c]);
      if (lane == m * SK_COLS + c) mine = s;
    }
  }
  if (ks > 1) {
    part[warp][lane] = mine;
    __syncthreads();
    if (kpart == 0)
      for (int j = 1; j < ks; ++j) mine += part[warp + j][lane];
  }
  const int m = lane >> 2, c = lane & 3, n = n0 + c;
  if (live && kpart == 0 && m < p.M && n < p.N) {
    float v = mine;
    if (p.bias) v += __bfloat162float(p.bias[n]);
    if (p.act == YMP_ACT_GELU_TANH) v = gelu_tanh(v);
    else if (p.act == YMP_ACT_GELU_ERF) v = gelu_erf(v);
    if (p.residual)
      v += p.res_f32 ? reinterpret_cast<const float*>(p.residual)[(size_t)m * p.ldr + n]
                     : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.residual)[(size_t)m * p.ldr + n]);
    if (p.out_f32) reinterpret_cast<float*>(p.y)[(size_t)m * p.ldy + n] = v;
    else reinterpret_cast<__nv_bfloat16*>(p.y)[(size_t)m * p.ldy + n] = __float2bfloat16(v);
  }
}

}  // namespace ymp

using namespace ymp;

extern "C" int ymp_gemm_skinny(const ymp_gemm_skinny_args* a, void* stream) {
  YMP_CHECK_ARG(a && a->x && a->w && a->y, "ymp_gemm_skinny: null pointer");
  YMP_CHECK_ARG(a->M >= 1 && a->M <= SK_MAXM && a->N > 0 && a->K > 0 && a->K % 8 == 0, "ymp_gemm_skinny: needs 1 <= M <= 8, K %% 8 == 0 (M=%d K=%d)", a->M, a->K);
  YMP_CHECK_ARG(a->ldx % 8 == 0 && a->ldw % 8 == 0 && a->ldx >= a->K && a->ldw >= a->K && aligned16(a->x) && aligned16(a->w), "ymp_gemm_skinny: x / w rows must be 16-byte aligned");
  YMP_CHECK_ARG(a->act >= 0 && a->act <= 2, "ymp_gemm_skinny: bad act");
  SkinnyParams p;
  p.x = (const __nv_bfloat16*)a->x; p.w = (const __nv_bfloat16*)a->w; p.bias = (const __nv_bfloat16*)a->bias;
  p.residual = a->residual; p.y = a->y;
  p.M = a->M; p.N = a->N; p.K = a->K; p.ldx = a->ldx; p.ldw = a->ldw; p.ldr = a->ldr; p.ldy = a->ldy;
  p.act = a->act; p.res_f32 = a->residual_dtype == YMP_DT_F32; p.out_f32 = a->out_dtype == YMP_DT_F32;
  // K slices per CTA: the fewest that still give >= 3 CTAs per SM (each warp keeps >= 32 16-byte chunks of its slice)
  int ks = 1;
  static const int force = [] { const char* e = getenv("YMP_SKINNY_KSPLIT"); return e ? atoi(e) : 0; }();
  while (ks < SK_WARPS && (a->N + (SK_WARPS / ks) * SK_COLS - 1) / ((SK_WARPS / ks) * SK_COLS) < 3 * num_sms() && a->K / (2 * ks) >= 256) ks *= 2;
  if (force == 1 || force == 2 || force == 4 || force == 8) ks = force;
  p.ksplit = ks;
  const int cols_per_cta = (SK_WARPS / ks) * SK_COLS;
  const int blocks = (a->N + cols_per_cta - 1) / cols_per_cta;
  gemm_skinny_kernel<<<blocks, SK_WARPS * 32, 0, (cudaStream_t)stream>>>(p);
  YMP_LAUNCH_CHECK();
  return YMP_OK;
}
