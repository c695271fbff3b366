// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA/TMEM).
// Everything here is architecture-specific on purpose: this library targets B200 only.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ymp {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
// 2D tiled load global -> shared, completion signalled on mbarrier (complete_tx::bytes)
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* desc, uint64_t* bar,
                                            int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)),
        "r"(c0), "r"(c1)
      : "memory");
}

// 5D tiled load (fused im2col of the patch embedding: coordinates {x, y, t, c, b} of a [B,C,T,H,W] video)
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const void* desc, uint64_t* bar, int32_t c0, int32_t c1,
                                            int32_t c2, int32_t c3, int32_t c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d_cta2(void* smem_dst, const void* desc, uint64_t* bar, int32_t c0, int32_t c1,
                                                 int32_t c2, int32_t c3, int32_t c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

// ---------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <- lane base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (sm_100 "version 1"), SWIZZLE_128B.
//   start address [0,14) >>4 | LBO [16,30) >>4 | SBO [32,46) >>4 | version=1 [46,48) | layout [61,64)=2
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor for kind::f16, bf16 inputs, fp32 accumulate.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major,
                                                       int b_mn_major) {
  return (1u << 4)                       // c_format = F32
         | (1u << 7)                     // a_format = BF16
         | (1u << 10)                    // b_format = BF16
         | ((uint32_t)a_mn_major << 15)  // a_major
         | ((uint32_t)b_mn_major << 16)  // b_major
         | ((uint32_t)(N >> 3) << 17)    // n_dim
         | ((uint32_t)(M >> 4) << 24);   // m_dim
}

// ---------------------------------------------------------------- 2-CTA (cta_group::2) variants
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load issued by either CTA of a pair; the bytes are accounted on the LEADER CTA's mbarrier
// (peer bit of the shared::cluster address cleared), where the single MMA issuer waits.
__device__ __forceinline__ void tma_load_2d_cta2(void* smem_dst, const void* desc, uint64_t* bar, int32_t c0,
                                                 int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar) & 0xFEFFFFFFu),
        "r"(c0), "r"(c1)
      : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_cta2(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_cta2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
// D[tmem of both CTAs, 256 x N] (+)= A (128 rows per CTA) * B (N/2 rows per CTA); issued by the leader
__device__ __forceinline__ void umma_bf16_cta2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of the pair's MMAs -> arrive on the mbarrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_cta2(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
// arrive on the mbarrier at this offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}

// ---------------------------------------------------------------- small math helpers
// programmatic dependent launch: wait until the preceding kernel of the stream has completed and its writes are visible
// (returns at once for an ordinary launch); allow the following kernel to start launching
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// streaming 128-bit load that does not allocate in L1 (weights read exactly once)
__device__ __forceinline__ uint4 ld_nc_v4(const uint4* p) {
  uint4 v;
  asm("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

// erf via Abramowitz-Stegun 7.1.26 (|abs err| < 1.5e-7: far below bf16 resolution)
__device__ __forceinline__ float fast_erf(float x) {
  float ax = fabsf(x);
  float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float poly = fmaf(fmaf(fmaf(fmaf(1.061405429f, t, -1.453152027f), t, 1.421413741f), t,
                         -0.284496736f),
                    t, 0.254829592f) *
               t;
  float r = 1.0f - poly * __expf(-ax * ax);
  return copysignf(r, x);
}
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// exact-erf GELU (nn.GELU default) and its derivative, MUFU-light: Phi(x) = 0.5 + x Q(x^2) with a
// degree-9 near-minimax Q on |x| <= 4.5 (clamped outside; |Phi error| < 1e-5, |gelu error| < 5e-5, far
// below the bf16 resolution of the stored activation), the normal pdf by one MUFU.EX2.  18 issue slots per
// element instead of 28 + a MUFU.RCP: the K=768 ViT MLP GEMM epilogue stops being the bottleneck.
__device__ __forceinline__ float norm_cdf_pdf(float x, float& pdf) {
  const float xc = fminf(fmaxf(x, -4.5f), 4.5f);
  const float u = xc * xc;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(u, -0.72134752044f, -1.32574806474f)));
  pdf = e;  // exp(-x^2/2) / sqrt(2 pi)
  float q = -1.6543631001e-12f;
  q = fmaf(q, u, 1.9532824653e-10f);
  q = fmaf(q, u, -1.0287317553e-08f);
  q = fmaf(q, u, 3.2170341066e-07f);
  q = fmaf(q, u, -6.7323919166e-06f);
  q = fmaf(q, u, 1.0108823657e-04f);
  q = fmaf(q, u, -1.1397043329e-03f);
  q = fmaf(q, u, 9.8841767687e-03f);
  q = fmaf(q, u, -6.6411978624e-02f);
  q = fmaf(q, u, 3.9892175804e-01f);
  return fmaf(xc, q, 0.5f);  // Phi(x)
}
__device__ __forceinline__ float gelu_erf(float x) {
  float e;
  return x * norm_cdf_pdf(x, e);
}
// ---- packed fp32 pairs (Blackwell FFMA2 / FMUL2 / FADD2: one issue slot for two IEEE operations, bit-identical to the
// scalar instructions).  The GEMM epilogues are issue-slot / dependency-latency bound on the activation polynomials.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  unsigned long long ra, rb, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
  float2 d;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
  return d;
}
__device__ __forceinline__ float2 splat2(float c) { return make_float2(c, c); }

// Phi and the normal pdf of EIGHT values in lockstep (four packed pairs): the same arithmetic as norm_cdf_pdf, element
// by element, written coefficient-major so that the four dependency chains interleave.
__device__ __forceinline__ void norm_cdf_pdf_x8(const float (&x)[8], float2 (&cdf)[4], float2 (&pdf)[4]) {
  float2 xc[4], u[4], q[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    xc[i] = make_float2(fminf(fmaxf(x[2 * i], -4.5f), 4.5f), fminf(fmaxf(x[2 * i + 1], -4.5f), 4.5f));
    u[i] = fmul2(xc[i], xc[i]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 a = ffma2(u[i], splat2(-0.72134752044f), splat2(-1.32574806474f));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pdf[i].x) : "f"(a.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pdf[i].y) : "f"(a.y));
    q[i] = ffma2(splat2(-1.6543631001e-12f), u[i], splat2(1.9532824653e-10f));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = ffma2(q[i], u[i], splat2(-1.0287317553e-08f));
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = ffma2(q[i], u[i], splat2(3.2170341066e-07f));
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = ffma2(q[i], u[i], splat2(-6.7323919166e-06f));
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = ffma2(q[i], u[i], splat2(1.0108823657e-04f));
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = ffma2(q[i], u[i], splat2(-1.1397043329e-03f));
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = ffma2(q[i], u[i], splat2(9.8841767687e-03f));
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = ffma2(q[i], u[i], splat2(-6.6411978624e-02f));
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = ffma2(q[i], u[i], splat2(3.9892175804e-01f));
#pragma unroll
  for (int i = 0; i < 4; ++i) cdf[i] = ffma2(xc[i], q[i], splat2(0.5f));
}
// v = gelu_erf(x), d = gelu_erf'(x) for eight values (bit-identical to gelu_erf_both element by element)
__device__ __forceinline__ void gelu_erf_both_x8(const float (&x)[8], float (&v)[8], float (&d)[8]) {
  float2 cdf[4], pdf[4];
  norm_cdf_pdf_x8(x, cdf, pdf);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 xv = make_float2(x[2 * i], x[2 * i + 1]);
    const float2 dd = ffma2(xv, pdf[i], cdf[i]), vv = fmul2(xv, cdf[i]);
    d[2 * i] = dd.x; d[2 * i + 1] = dd.y; v[2 * i] = vv.x; v[2 * i + 1] = vv.y;
  }
}
__device__ __forceinline__ void gelu_erf_x8(const float (&x)[8], float (&v)[8]) {
  float2 cdf[4], pdf[4];
  norm_cdf_pdf_x8(x, cdf, pdf);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 vv = fmul2(make_float2(x[2 * i], x[2 * i + 1]), cdf[i]);
    v[2 * i] = vv.x; v[2 * i + 1] = vv.y;
  }
}
// GELU and its derivative together (the derivative is what backward needs; it is stored in bf16 by the
// forward epilogue so that the backward epilogue is a plain multiply)
__device__ __forceinline__ float gelu_erf_both(float x, float& d) {
  float e;
  const float cdf = norm_cdf_pdf(x, e);
  d = fmaf(x, e, cdf);
  return x * cdf;
}
__device__ __forceinline__ float gelu_tanh_both(float x, float& d) {
  const float u = 0.79788456f * x * fmaf(0.044715f * x, x, 1.0f);
  const float t = tanh_fast(u);
  const float du = 0.79788456f * fmaf(0.134145f * x, x, 1.0f);
  const float hp = 0.5f * (1.0f + t);
  d = fmaf(0.5f * x * (1.0f - t * t), du, hp);
  return x * hp;
}
__device__ __forceinline__ float dgelu_erf(float x) {
  float e;
  const float cdf = norm_cdf_pdf(x, e);
  return fmaf(x, e, cdf);
}
// tanh-approximation GELU (Megatron bias_gelu) and its derivative
__device__ __forceinline__ float gelu_tanh(float x) {
  float u = 0.79788456f * x * fmaf(0.044715f * x, x, 1.0f);
  return 0.5f * x * (1.0f + tanh_fast(u));
}
__device__ __forceinline__ float dgelu_tanh(float x) {
  float u = 0.79788456f * x * fmaf(0.044715f * x, x, 1.0f);
  float t = tanh_fast(u);
  float du = 0.79788456f * fmaf(0.134145f * x, x, 1.0f);
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * du;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace ymp
