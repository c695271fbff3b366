// tcgen05 attention, head_dim 64 / 96 natively and 80 / 88 zero-padded to 96 inside shared memory (the 2.7B
// decoder and EVA-g): ViT spatial (197 keys), GPT causal (256, 384), the abstractor's 1570-key cross attention.
// Key ranges <= 256 take the single-pass kernel below; longer ranges run the same tile code in a loop over
// 256-key blocks with an online softmax (<LONG>: running row max / sum, O re-scaled in registers).
// Forward, one CTA (256 threads) = 128 query rows of one (sequence, head):
//   1. cp.async gathers Q [128 x HD], K, V [Nkv x HD] through the seqmap into shared memory laid out
//      exactly as the UMMA canonical SWIZZLE_128B (first 64 head-dim columns) / SWIZZLE_64B (columns
//      64..95 when HD = 96) tiles
//   2. one elected thread issues S[128 x Nkv] = Q K^T  (tcgen05.mma, K-major A and B, fp32 in TMEM)
//   3. two threads per query row (= TMEM lane), one per key-column half, do the exact softmax with two passes
//      of tcgen05.ld (row maxima / sums exchanged through shared memory) and write P as packed bf16 back into
//      TMEM columns they have already consumed (tcgen05.st)
//   4. O[128 x HD] = P V  with A = P read from TMEM and B = V as an MN-major smem operand (two issuer threads
//      for the 64- and 32-column parts)
//   5. tcgen05.ld O, scale by 1/l, 16-byte row stores + lse
// Short ranges: no online-softmax rescaling, no KV loop, no register-resident accumulators.  TMEM use is 256 columns
// (S, then P in [0,64) and [128,192), O in [64,128) and [192,224)), shared memory ~80-113 KB, so two CTAs
// share an SM and overlap each other's load / MMA / softmax phases.  The backward kernel is described below.
#include <math_constants.h>

#include "common.h"
#include "philox.cuh"
#include "ptx.cuh"

namespace ymp {

struct TcSeqMap {
  int seq_div, n_prefix, prefix_per_seq;
  long outer_stride, inner_stride, pos_stride, prefix_base, prefix_stride;
};
struct TcMat {
  const __nv_bfloat16* base;
  const __nv_bfloat16* prefix;
  long stride;
  int ld, n_prefix;
};
__device__ __forceinline__ TcMat tc_mat(const __nv_bfloat16* p, const TcSeqMap& m, int s, int ld, int col_off) {
  const int outer = s / m.seq_div, inner = s - outer * m.seq_div;
  TcMat r;
  r.base = p + ((long)outer * m.outer_stride + (long)inner * m.inner_stride) * ld + col_off;
  r.prefix = p + (m.prefix_base + (long)(m.prefix_per_seq ? s : outer) * m.prefix_stride) * ld + col_off;
  r.stride = m.pos_stride * ld;
  r.ld = ld;
  r.n_prefix = m.n_prefix;
  return r;
}
__device__ __forceinline__ const __nv_bfloat16* tc_row(const TcMat& m, int i) {
  return i < m.n_prefix ? m.prefix + (long)i * m.ld : m.base + (long)(i - m.n_prefix) * m.stride;
}

struct AttnTcParams {
  const __nv_bfloat16 *q, *k, *v;
  __nv_bfloat16* o;
  float* lse;
  int ldq, ldk, ldv, ldo, hsq, hsk, hsv, hso;
  TcSeqMap mq, mkv, mo;
  int n_seq, n_heads, s_q, s_kv, mask, mask_block;
  long total_rows;
  float scale_log2, scale;
  int kv_rows;  // rows of the K/V smem tiles = round32(min(s_kv, 256))
  int hd;       // real head_dim (<= HD): columns [hd, HD) are zero-filled in shared memory and never stored
  DropSpec drop;  // dropout of the probabilities (has_drop): row = (seq*heads + head)*s_q + query, column = key
  int has_drop;
};

__device__ __forceinline__ void cp16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
// byte offset of 16-byte chunk c16 of row r inside a SWIZZLE_128B block ([rows][64 bf16]) / SWIZZLE_64B
// block ([rows][32 bf16])
__device__ __forceinline__ uint32_t sw128_off(int r, int c16) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4));
}
__device__ __forceinline__ uint32_t sw64_off(int r, int c16) {
  return (uint32_t)((r >> 3) * 512 + (r & 7) * 64 + ((c16 ^ ((r >> 1) & 3)) << 4));
}
// smem descriptors: K-major / MN-major both use (SBO = one 8-row group, layout type)
__device__ __forceinline__ uint64_t desc_sw(uint32_t saddr, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;  // LBO (unused: single swizzle atom along the leading dimension)
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15, %16};"
      :
      : "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int TC_MASK_CAUSAL = 1, TC_MASK_BLOCK = 2;
constexpr int LAYOUT_SW128 = 2, LAYOUT_SW64 = 4;

// Load `rows` rows x HD columns into the swizzled blocks (rows beyond n_valid are zero-filled).
// thread -> (row = tid/4 within a 32-row pass, 16-byte chunks tid%4 + 4j): one row pointer per thread and
// pass, 4 consecutive lanes fetch 64 contiguous bytes.
#ifdef YMP_ATTN_DBG
__device__ unsigned long long ymp_attn_dbg_buf[256];
#define TDBG(k)                                                                                        \
  do {                                                                                                 \
    if (YMP_DBG_BLOCK && threadIdx.x == 0 && dbg_n < 256)                        \
      ymp_attn_dbg_buf[dbg_n++] = ((unsigned long long)(k) << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull); \
  } while (0)
#else
#define TDBG(k)
#endif

// 256-thread tile loader: thread -> (row = tid/4 of a 64-row pass, 16-byte chunks tid%4 + 4j).  Rows beyond
// n_valid and head-dim columns beyond `hd` (head_dim 80 / 88 padded to 96) are zero-filled.
template <int HD>
__device__ __forceinline__ void tc_load256(uint8_t* blk0, uint8_t* blk1, const TcMat& m, int r0, int rows, int n_valid, int hd) {
  constexpr int CPT = HD / 32;
  const int rl = threadIdx.x >> 2, c0 = threadIdx.x & 3;
  for (int rb = 0; rb < rows; rb += 64) {
    const int r = rb + rl;
    if (r >= rows) break;
    const bool rv = r0 + r < n_valid;
    const __nv_bfloat16* g = rv ? tc_row(m, r0 + r) : nullptr;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int c = c0 + 4 * j;
      uint8_t* dst = (c < 8) ? blk0 + sw128_off(r, c) : blk1 + sw64_off(r, c - 8);
      if (rv && c * 8 < hd) cp16(dst, g + c * 8);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
  }
}

template <int HD, bool LONG, bool DROP>
__global__ void __launch_bounds__(256) attn_tc_fwd_kernel(const AttnTcParams p) {
  static_assert(HD == 64 || HD == 96, "head_dim 64 or 96 (80 / 88 are padded to 96)");
  constexpr bool TWO = (HD == 96);
  // TMEM columns (256 allocated): S in [0, nkv); the two warpgroups own the key-column halves [0,128) / [128,256):
  // P (bf16 pairs) of half 0 lands in [0,64), of half 1 in [128,192) - always inside columns its own threads
  // have already consumed; O accumulates in the columns both halves have released: [64,128) (+ [192,224), HD = 96)
  // (short key ranges, <= 128 columns: the halves split at the middle chunk and O uses [128,224))
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const int kvr = p.kv_rows;
  uint8_t* q0s = smem;                                  // 128 x 128 B
  uint8_t* k0s = q0s + 128 * 128;                       // kvr x 128 B
  uint8_t* v0s = k0s + kvr * 128;                       // kvr x 128 B
  uint8_t* q1s = v0s + kvr * 128;                       // 128 x 64 B   (HD = 96 only)
  uint8_t* k1s = q1s + (TWO ? 128 * 64 : 0);
  uint8_t* v1s = k1s + (TWO ? kvr * 64 : 0);
  float* xch = reinterpret_cast<float*>(v1s + (TWO ? kvr * 64 : 0));  // [2][128]: row max, then row sum, per half
  uint64_t* bar = reinterpret_cast<uint64_t*>(xch + 256);              // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wq = warp & 3, half = warp >> 2;  // TMEM lane quarter (query rows), key-column half
  const int q0 = blockIdx.x * 128, h = blockIdx.y, s = blockIdx.z;
#ifdef YMP_ATTN_DBG
#define YMP_DBG_BLOCK (blockIdx.x == 0 && blockIdx.y == 3 && blockIdx.z == 100 && gridDim.z > 100 && blockDim.x == 256 && gridDim.x == 2)
  int dbg_n = 0;
#endif
  TDBG(0);
  int sq = p.s_q, skv = p.s_kv;
  if (p.total_rows > 0) {
    const long left = p.total_rows - (long)s * p.s_q;
    if (left < sq) sq = (int)left;
    if (left < skv) skv = (int)left;
  }
  if (q0 >= sq) return;
  // key range this tile needs
  int kv_end = skv;
  if (p.mask == TC_MASK_CAUSAL) kv_end = min(skv, q0 + 128 + (p.s_kv - p.s_q));
  const int nblk = LONG ? (kv_end + 255) >> 8 : 1;   // 256-key blocks (short ranges: exactly one)

  if (warp == 0) tmem_alloc<256>(tmem_ptr);
  if (threadIdx.x == 32) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], TWO ? 2 : 1);
    fence_mbar_init();
  }
  const TcMat Mk = tc_mat(p.k, p.mkv, s, p.ldk, h * p.hsk);
  const TcMat Mv = tc_mat(p.v, p.mkv, s, p.ldv, h * p.hsv);
  {
    const TcMat Mq = tc_mat(p.q, p.mq, s, p.ldq, h * p.hsq);
    const int kend0 = min(kv_end, 256);
    tc_load256<HD>(q0s, q1s, Mq, q0, 128, sq, p.hd);
    tc_load256<HD>(k0s, k1s, Mk, 0, (kend0 + 31) & ~31, kv_end, p.hd);
    tc_load256<HD>(v0s, v1s, Mv, 0, (kend0 + 31) & ~31, kv_end, p.hd);
  }
  TDBG(1);
  asm volatile("cp.async.wait_all;" ::: "memory");
  fence_proxy_async();  // smem written through the generic proxy is read by the tensor core's async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  TDBG(2);

  const int rl = wq * 32 + lane;  // row inside the tile
  const int row = q0 + rl;
  const uint32_t tl = tmem + ((uint32_t)(wq * 32) << 16);
  // valid key columns of this row form one interval [lo, hi): bounds + causal + block-diagonal masks
  int lo = 0, hi = kv_end;
  if (p.mask == TC_MASK_CAUSAL) hi = min(hi, row + 1 + (p.s_kv - p.s_q));
  if (p.mask == TC_MASK_BLOCK) { lo = (row / p.mask_block) * p.mask_block; hi = min(hi, lo + p.mask_block); }

  DropState ds = {};
  if (DROP) ds = drop_state(p.drop);
  const uint32_t drow = ((uint32_t)s * p.n_heads + h) * p.s_q + row;   // logical row of the probability matrix
  float m_run = -CUDART_INF_F, l_run = 0.f;   // running row max (raw scores) and row sum
  float oacc[LONG ? 2 : 1][LONG ? 32 : 1];    // LONG: this thread's O chunks (chunk c of HD/32 belongs to warpgroup c & 1)
  if (LONG) {
#pragma unroll
    for (int i = 0; i < 32; ++i) { oacc[0][i] = 0.f; oacc[LONG ? 1 : 0][i] = 0.f; }
  }
  uint32_t O0_COL = 64, O1_COL = 192;

#pragma unroll 1
  for (int kb = 0; kb < nblk; ++kb) {
    const int k0 = kb << 8;
    const int kend_b = min(kv_end - k0, 256);     // valid keys of this block
    const int nkv = (kend_b + 31) & ~31;          // MMA N / K extent (multiple of 32, <= 256)
    const uint32_t par = kb & 1;
    if (LONG && kb > 0) {
      // every thread has passed bar[1] of the previous block (its S and P V MMAs are complete, so K / V may be
      // overwritten) and has finished reading its O chunks out of TMEM
      tc_load256<HD>(k0s, k1s, Mk, k0, nkv, kv_end, p.hd);
      tc_load256<HD>(v0s, v1s, Mv, k0, nkv, kv_end, p.hd);
      asm volatile("cp.async.wait_all;" ::: "memory");
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
    }
    // ---- S = Q K^T
    if (threadIdx.x == 0) {
      const uint32_t idesc = make_idesc_bf16(128, nkv, 0, 0);
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        uint64_t ad, bd;
        if (ks < 4) {
          ad = desc_sw(smem_u32(q0s) + ks * 32, 1024, LAYOUT_SW128);
          bd = desc_sw(smem_u32(k0s) + ks * 32, 1024, LAYOUT_SW128);
        } else {
          ad = desc_sw(smem_u32(q1s) + (ks - 4) * 32, 512, LAYOUT_SW64);
          bd = desc_sw(smem_u32(k1s) + (ks - 4) * 32, 512, LAYOUT_SW64);
        }
        umma_bf16(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
      }
      umma_commit(&bar[0]);
    }
    TDBG(3);
    mbar_wait(&bar[0], par);
    tc_fence_after();
    TDBG(4);

    // ---- softmax of this block: two threads per query row (= TMEM lane), one per key-column half
    const int lo_b = lo - k0, hi_b = min(hi - k0, kend_b);   // valid interval in block-local columns
    const int nch = nkv / 32;
    const int c1 = nch > 4 ? 4 : (nch + 1) / 2;                  // first chunk of the second half
    const uint32_t P1_COL = 32 * c1;                              // where the second half's packed P starts
    O0_COL = nch > 4 ? 64 : 128;
    const int c_lo = half ? c1 : 0, c_hi = half ? nch : c1;       // this thread's 32-column chunks
    float mx = -CUDART_INF_F;
    for (int c = c_lo; c < c_hi; ++c) {
      // tcgen05.ld is .sync.aligned: the skip decision must be warp-uniform
      if (__all_sync(0xffffffffu, c * 32 >= hi_b || c * 32 + 32 <= lo_b)) continue;
      uint32_t r[32];
      tmem_ld32(tl + c * 32, r);
      tmem_ld_wait();
      if (c * 32 >= lo_b && c * 32 + 32 <= hi_b) {           // fully valid: no per-element masking
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int col = c * 32 + i;
          if (col >= lo_b && col < hi_b) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
    }
    xch[half * 128 + rl] = mx;
    __syncthreads();
    mx = fmaxf(mx, xch[(half ^ 1) * 128 + rl]);
    __syncthreads();  // both halves have read the maxima: the buffer is reused for the sums
    TDBG(5);
    const float m_new = fmaxf(m_run, mx);
    const float ms = (m_new == -CUDART_INF_F) ? 0.f : m_new * p.scale_log2;
    // re-scaling factor of what has been accumulated so far (online softmax; 1 block: unused)
    const float alpha = (m_run == -CUDART_INF_F) ? 0.f : exp2f(fmaf(m_run, p.scale_log2, -ms));
    float lsum = 0.f;
    for (int c = c_lo; c < c_hi; ++c) {
      uint32_t pk[16];
      if (__all_sync(0xffffffffu, c * 32 >= hi_b || c * 32 + 32 <= lo_b)) {   // warp-uniform
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
      } else {
        uint32_t r[32];
        tmem_ld32(tl + c * 32, r);
        tmem_ld_wait();
        float pv[32];
        if (c * 32 >= lo_b && c * 32 + 32 <= hi_b) {
#pragma unroll
          for (int i = 0; i < 32; ++i) pv[i] = exp2f(fmaf(__uint_as_float(r[i]), p.scale_log2, -ms));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int col = c * 32 + i;
            pv[i] = (col >= lo_b && col < hi_b) ? exp2f(fmaf(__uint_as_float(r[i]), p.scale_log2, -ms)) : 0.f;
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) lsum += pv[i];
        if (DROP) {  // O = dropout(P) V; the normaliser stays that of the undropped P (:772-780: softmax, then dropout)
#pragma unroll
          for (int j = 0; j < 8; ++j) drop4(ds, drow, (uint32_t)(k0 + c * 32 + 4 * j), pv[4 * j], pv[4 * j + 1], pv[4 * j + 2], pv[4 * j + 3]);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = pack_bf16(pv[2 * i], pv[2 * i + 1]);
      }
      // P (bf16 pairs) overwrites S columns this thread has already consumed
      tmem_st16(tl + half * P1_COL + (c - c_lo) * 16, pk);
    }
    xch[half * 128 + rl] = lsum;
    TDBG(6);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    lsum += xch[(half ^ 1) * 128 + rl];
    l_run = l_run * alpha + lsum;
    m_run = m_new;
    TDBG(7);

    // ---- O = P V   (A = P from TMEM, B = V MN-major): thread 0 issues the 64-column part, thread 32 the rest
    if (threadIdx.x == 0 || (TWO && threadIdx.x == 32)) {
      const bool second = threadIdx.x == 32;
      const uint32_t idesc = second ? make_idesc_bf16(128, 32, 0, 1) : make_idesc_bf16(128, 64, 0, 1);
      const uint32_t ocol = tmem + (second ? O1_COL : O0_COL);
      const uint32_t vb = smem_u32(second ? v1s : v0s);
      const int nks = nkv / 16;
      for (int ks = 0; ks < nks; ++ks) {
        const uint32_t acol = tmem + (ks < 2 * c1 ? ks * 8 : P1_COL + (ks - 2 * c1) * 8);  // 16 keys = 8 packed columns
        umma_ts(ocol, acol, second ? desc_sw(vb + ks * 1024, 512, LAYOUT_SW64) : desc_sw(vb + ks * 2048, 1024, LAYOUT_SW128), idesc,
                ks > 0 ? 1u : 0u);
      }
      umma_commit(&bar[1]);
    }
    TDBG(8);
    mbar_wait(&bar[1], par);
    tc_fence_after();
    TDBG(9);
    if (LONG) {
      // O_acc = O_acc * alpha + (P V of this block), kept in registers across the key blocks
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        if ((c & 1) != half) continue;
        uint32_t r[32];
        tmem_ld32(tl + (c < 2 ? O0_COL + c * 32 : O1_COL), r);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) oacc[LONG ? (c >> 1) : 0][LONG ? i : 0] = fmaf(oacc[LONG ? (c >> 1) : 0][LONG ? i : 0], alpha, __uint_as_float(r[i]));
      }
    }
  }

  // ---- epilogue: 32-column chunk c of O is read by warpgroup c & 1
  const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
  const bool valid = row < sq;
  __nv_bfloat16* orow = nullptr;
  if (valid) {
    const TcMat Mo = tc_mat(p.o, p.mo, s, p.ldo, h * p.hso);
    orow = const_cast<__nv_bfloat16*>(tc_row(Mo, row));
  }
#pragma unroll
  for (int c = 0; c < HD / 32; ++c) {
    if ((c & 1) != half) continue;
    float f[32];
    if (LONG) {
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = oacc[LONG ? (c >> 1) : 0][LONG ? i : 0];
    } else {
      uint32_t r[32];
      tmem_ld32(tl + (c < 2 ? O0_COL + c * 32 : O1_COL), r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(r[i]);
    }
    if (valid) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (c * 32 + j * 8 >= p.hd) break;   // padded head_dim columns are not stored
        uint4 o;
        o.x = pack_bf16(f[8 * j + 0] * inv, f[8 * j + 1] * inv);
        o.y = pack_bf16(f[8 * j + 2] * inv, f[8 * j + 3] * inv);
        o.z = pack_bf16(f[8 * j + 4] * inv, f[8 * j + 5] * inv);
        o.w = pack_bf16(f[8 * j + 6] * inv, f[8 * j + 7] * inv);
        *reinterpret_cast<uint4*>(orow + c * 32 + j * 8) = o;
      }
    }
  }
  if (valid && half == 0 && p.lse) p.lse[((size_t)s * p.n_heads + h) * p.s_q + row] = m_run * p.scale + logf(l_run);
  TDBG(10);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<256>(tmem);
  }
}

// ======================================================================================================
// Pair-tile forward for s_q, s_kv <= 256 (ViT spatial 197, GPT causal 256): one PERSISTENT CTA per SM walks over
// (sequence, head) items.  Per item K / V are fetched ONCE for both 128-row query tiles (the single-tile kernel
// above fetches them per tile: +33 % HBM traffic at 197 rows) into one of two shared-memory stages - the next
// item's K / V / Q arrive behind the current item's MMAs and softmax (cp.async issued a whole item ahead), so no
// load latency is exposed.  Warpgroup w owns query tile w: its own Q tile, its own 256 TMEM columns, its own
// mbarriers, one thread per query row over all key columns (S in [0,nkv), packed P over the consumed front half
// [0,nkv/2), O in [128,128+HD)); the two warpgroups drift apart, so one tile's softmax overlaps the other's MMAs.
// Output rows are staged in the item's (by then idle) K / V stage and written with coalesced 16-byte accesses.
// ======================================================================================================
template <int HD>
__device__ __forceinline__ void tc_load128(uint8_t* blk0, uint8_t* blk1, const TcMat& m, int r0, int rows, int n_valid, int hd, int t) {
  constexpr int CPT = HD / 32;
  const int rl = t >> 2, c0 = t & 3;
  for (int rb = 0; rb < rows; rb += 32) {
    const int r = rb + rl;
    if (r >= rows) break;
    const bool rv = r0 + r < n_valid;
    const __nv_bfloat16* g = rv ? tc_row(m, r0 + r) : nullptr;
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
      const int c = c0 + 4 * j;
      uint8_t* dst = (c < 8) ? blk0 + sw128_off(r, c) : blk1 + sw64_off(r, c - 8);
      if (rv && c * 8 < hd) cp16(dst, g + c * 8);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
  }
}
__device__ __forceinline__ void wg_sync(int w) { asm volatile("bar.sync %0, 128;" ::"r"(1 + w) : "memory"); }

template <int HD, bool DROP>
__global__ void __launch_bounds__(256, 1) attn_tc_fwd_pair_kernel(const AttnTcParams p) {
  static_assert(HD == 64 || HD == 96, "head_dim 64 or 96 (80 / 88 are padded to 96)");
  constexpr bool TWO = (HD == 96);
  constexpr int PITCH = HD * 2 + 16;                    // staging row pitch (the pad spreads the banks)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  const int kvr = p.kv_rows;
  const int QT0 = 128 * 128, QT1 = TWO ? 128 * 64 : 0;                 // bytes of one Q tile (128 B + 64 B column blocks)
  const int KB0 = kvr * 128, KB1 = TWO ? kvr * 64 : 0;                  // bytes of one K (or V) block pair
  const int STG = 2 * (KB0 + KB1);                                      // one stage: K0 K1 V0 V1
  uint8_t* qs = smem;                                                   // [2 tiles][QT0 + QT1]
  uint8_t* kvs = qs + 2 * (QT0 + QT1);                                  // [2 stages][STG]
  uint64_t* bar = reinterpret_cast<uint64_t*>(kvs + 2 * STG);          // bar_s[2], bar_o[2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int w = warp >> 2, wq = warp & 3, t = tid & 127;   // warpgroup = query tile, TMEM lane quarter, thread in group
  const int n_items = p.n_seq * p.n_heads;

  if (warp == 0) tmem_alloc<512>(tmem_ptr);
  if (tid == 32) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], (i >= 2 && TWO) ? 2 : 1);
    fence_mbar_init();
  }
  DropState ds = {};
  if (DROP) ds = drop_state(p.drop);

  // loads of one item: K, V by all 256 threads, the Q tile of warpgroup w by its 128 threads
  auto item_len = [&](int it, int& sq, int& skv) {
    sq = p.s_q; skv = p.s_kv;
    if (p.total_rows > 0) {
      const long left = p.total_rows - (long)(it / p.n_heads) * p.s_q;
      if (left < sq) sq = (int)max(left, 0l);
      if (left < skv) skv = (int)max(left, 0l);
    }
  };
  auto load_kv = [&](int it, int stage) {
    int sq, skv;
    item_len(it, sq, skv);
    const int s = it / p.n_heads, h = it - s * p.n_heads;
    const TcMat Mk = tc_mat(p.k, p.mkv, s, p.ldk, h * p.hsk), Mv = tc_mat(p.v, p.mkv, s, p.ldv, h * p.hsv);
    uint8_t* st = kvs + stage * STG;
    const int nk = (skv + 31) & ~31;
    tc_load256<HD>(st, st + KB0, Mk, 0, nk, skv, p.hd);
    tc_load256<HD>(st + KB0 + KB1, st + 2 * KB0 + KB1, Mv, 0, nk, skv, p.hd);
  };
  auto load_q = [&](int it) {
    int sq, skv;
    item_len(it, sq, skv);
    if (w * 128 >= sq) return;
    const int s = it / p.n_heads, h = it - s * p.n_heads;
    const TcMat Mq = tc_mat(p.q, p.mq, s, p.ldq, h * p.hsq);
    uint8_t* q0 = qs + w * (QT0 + QT1);
    tc_load128<HD>(q0, q0 + QT0, Mq, w * 128, 128, sq, p.hd, t);
  };

  int item = blockIdx.x;
  if (item < n_items) { load_kv(item, 0); load_q(item); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr + (uint32_t)(w * 256);                // this warpgroup's 256 columns
  const uint32_t tl = tmem + ((uint32_t)(wq * 32) << 16);
  const bool stage_ok = kvr * (HD * 2) >= 128 * PITCH;                  // the K (V) stage can hold a staged output tile

  uint32_t n0 = 0, n1 = 0;   // how many items so far had a tile 0 / a tile 1: the phases of bar_s / bar_o of each tile
  for (int k = 0; item < n_items; item += gridDim.x, ++k) {
    const int stage = k & 1;
    int sq, skv;
    item_len(item, sq, skv);
    const uint32_t par0 = n0 & 1, par1 = n1 & 1, par = w ? par1 : par0;
    n0 += sq > 0 ? 1 : 0;
    n1 += sq > 128 ? 1 : 0;
    const int s = item / p.n_heads, h = item - s * p.n_heads;
    uint8_t* st = kvs + stage * STG;
    uint8_t *k0s = st, *k1s = st + KB0, *v0s = st + KB0 + KB1, *v1s = st + 2 * KB0 + KB1;
    uint8_t *q0s = qs + w * (QT0 + QT1), *q1s = q0s + QT0;
    const bool have = w * 128 < sq;                                     // this warpgroup's tile exists
    const int q0 = w * 128;
    int kv_end = skv;
    if (p.mask == TC_MASK_CAUSAL) kv_end = min(skv, q0 + 128 + (p.s_kv - p.s_q));
    const int nkv = (kv_end + 31) & ~31;

    // ---- (A) this item's K / V / Q have landed; everybody has left the previous item
    asm volatile("cp.async.wait_all;" ::: "memory");
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // ---- (B) S = Q K^T of this tile, issued by the warpgroup's first thread
    if (have && t == 0) {
      const uint32_t idesc = make_idesc_bf16(128, nkv, 0, 0);
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        uint64_t ad, bd;
        if (ks < 4) {
          ad = desc_sw(smem_u32(q0s) + ks * 32, 1024, LAYOUT_SW128);
          bd = desc_sw(smem_u32(k0s) + ks * 32, 1024, LAYOUT_SW128);
        } else {
          ad = desc_sw(smem_u32(q1s) + (ks - 4) * 32, 512, LAYOUT_SW64);
          bd = desc_sw(smem_u32(k1s) + (ks - 4) * 32, 512, LAYOUT_SW64);
        }
        umma_bf16(tmem, ad, bd, idesc, ks > 0 ? 1u : 0u);
      }
      umma_commit(&bar[w]);
    }
    // the next item's K / V stream into the other stage behind everything that follows
    const int nxt = item + gridDim.x;
    if (nxt < n_items) load_kv(nxt, stage ^ 1);
    if (!have) {           // (warpgroup-uniform; the CTA-wide barrier of the next iteration still sees these threads)
      if (nxt < n_items) load_q(nxt);
      continue;
    }

    mbar_wait(&bar[w], par);
    tc_fence_after();
    if (nxt < n_items) load_q(nxt);                                     // Q tile consumed: fetch the next item's

    // ---- (C) softmax, one thread per query row
    const int rl = wq * 32 + lane, row = q0 + rl;
    int lo = 0, hi = kv_end;
    if (p.mask == TC_MASK_CAUSAL) hi = min(hi, row + 1 + (p.s_kv - p.s_q));
    const int nch = nkv / 32;
    const uint32_t drow = ((uint32_t)s * p.n_heads + h) * p.s_q + row;
    float mx = -CUDART_INF_F;
    for (int c = 0; c < nch; ++c) {
      if (__all_sync(0xffffffffu, c * 32 >= hi || c * 32 + 32 <= lo)) continue;
      uint32_t r[32];
      tmem_ld32(tl + c * 32, r);
      tmem_ld_wait();
      if (c * 32 >= lo && c * 32 + 32 <= hi) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int col = c * 32 + i;
          if (col >= lo && col < hi) mx = fmaxf(mx, __uint_as_float(r[i]));
        }
      }
    }
    const float ms = (mx == -CUDART_INF_F) ? 0.f : mx * p.scale_log2;
    float lsum = 0.f;
    for (int c = 0; c < nch; ++c) {
      uint32_t pk[16];
      if (__all_sync(0xffffffffu, c * 32 >= hi || c * 32 + 32 <= lo)) {
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = 0u;
      } else {
        uint32_t r[32];
        tmem_ld32(tl + c * 32, r);
        tmem_ld_wait();
        float pv[32];
        if (c * 32 >= lo && c * 32 + 32 <= hi) {
#pragma unroll
          for (int i = 0; i < 32; ++i) pv[i] = exp2f(fmaf(__uint_as_float(r[i]), p.scale_log2, -ms));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int col = c * 32 + i;
            pv[i] = (col >= lo && col < hi) ? exp2f(fmaf(__uint_as_float(r[i]), p.scale_log2, -ms)) : 0.f;
          }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) lsum += pv[i];
        if (DROP) {
#pragma unroll
          for (int j = 0; j < 8; ++j) drop4(ds, drow, (uint32_t)(c * 32 + 4 * j), pv[4 * j], pv[4 * j + 1], pv[4 * j + 2], pv[4 * j + 3]);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = pack_bf16(pv[2 * i], pv[2 * i + 1]);
      }
      tmem_st16(tl + c * 16, pk);   // packed P over the front half of S: chunk c lands inside chunk c/2, already consumed
    }
    tmem_st_wait();
    tc_fence_before();
    wg_sync(w);
    tc_fence_after();

    // ---- O = P V (A = P from TMEM, B = V MN-major): 64-column part by thread 0, the 32-column rest by thread 32
    if (t == 0 || (TWO && t == 32)) {
      const bool second = t == 32;
      const uint32_t idesc = second ? make_idesc_bf16(128, 32, 0, 1) : make_idesc_bf16(128, 64, 0, 1);
      const uint32_t ocol = tmem + (second ? 192 : 128);
      const uint32_t vb = smem_u32(second ? v1s : v0s);
      for (int ks = 0; ks < nkv / 16; ++ks)
        umma_ts(ocol, tmem + ks * 8, second ? desc_sw(vb + ks * 1024, 512, LAYOUT_SW64) : desc_sw(vb + ks * 2048, 1024, LAYOUT_SW128), idesc,
                ks > 0 ? 1u : 0u);
      umma_commit(&bar[2 + w]);
    }
    mbar_wait(&bar[2 + w], par);
    tc_fence_after();

    // ---- epilogue: rows staged in this item's idle K (tile 0) / V (tile 1) stage, then coalesced stores
    const float inv = lsum > 0.f ? 1.f / lsum : 0.f;
    const TcMat Mo = tc_mat(p.o, p.mo, s, p.ldo, h * p.hso);
    if (row < sq && p.lse) p.lse[((size_t)s * p.n_heads + h) * p.s_q + row] = mx * p.scale + logf(lsum);
    uint8_t* stg = w ? v0s : k0s;
    if (stage_ok && w == 0 && sq > 128) mbar_wait(&bar[1], par1);      // tile 1's S MMA reads K until its barrier flips
    if (stage_ok && w == 1) mbar_wait(&bar[2], par0);                   // tile 0's P V MMA reads V until its barrier flips
#pragma unroll
    for (int c = 0; c < HD / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tl + (c < 2 ? 128 + c * 32 : 192), r);
      tmem_ld_wait();
      if (stage_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<uint4*>(stg + rl * PITCH + c * 64 + g * 16) =
              make_uint4(pack_bf16(__uint_as_float(r[8 * g]) * inv, __uint_as_float(r[8 * g + 1]) * inv),
                         pack_bf16(__uint_as_float(r[8 * g + 2]) * inv, __uint_as_float(r[8 * g + 3]) * inv),
                         pack_bf16(__uint_as_float(r[8 * g + 4]) * inv, __uint_as_float(r[8 * g + 5]) * inv),
                         pack_bf16(__uint_as_float(r[8 * g + 6]) * inv, __uint_as_float(r[8 * g + 7]) * inv));
      } else if (row < sq) {
        __nv_bfloat16* orow = const_cast<__nv_bfloat16*>(tc_row(Mo, row));
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (c * 32 + g * 8 >= p.hd) break;
          *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) =
              make_uint4(pack_bf16(__uint_as_float(r[8 * g]) * inv, __uint_as_float(r[8 * g + 1]) * inv),
                         pack_bf16(__uint_as_float(r[8 * g + 2]) * inv, __uint_as_float(r[8 * g + 3]) * inv),
                         pack_bf16(__uint_as_float(r[8 * g + 4]) * inv, __uint_as_float(r[8 * g + 5]) * inv),
                         pack_bf16(__uint_as_float(r[8 * g + 6]) * inv, __uint_as_float(r[8 * g + 7]) * inv));
        }
      }
    }
    tc_fence_before();   // the next item's S MMA overwrites these TMEM columns: ordered by the CTA barrier at (A)
    if (stage_ok) {
      wg_sync(w);
      constexpr int CPT = HD / 32;
      const int srl = t >> 2, c0 = t & 3;
#pragma unroll
      for (int rb = 0; rb < 128; rb += 32) {
        const int r = rb + srl;
        if (q0 + r < sq) {
          __nv_bfloat16* g = const_cast<__nv_bfloat16*>(tc_row(Mo, q0 + r));
#pragma unroll
          for (int j = 0; j < CPT; ++j) {
            const int c = c0 + 4 * j;
            if (c * 8 < p.hd) *reinterpret_cast<uint4*>(g + c * 8) = *reinterpret_cast<const uint4*>(stg + r * PITCH + c * 16);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(*tmem_ptr);
  }
}

// ======================================================================================================
// tcgen05 attention backward, any s_kv, s_q <= 2048 (per-query statistics live in shared memory).
// One CTA (256 threads) = one (sequence, head).
// Everything is computed in the TRANSPOSED frame (TMEM lane = key row, TMEM column = query row) so that
// P^T and dS^T feed the dV / dK MMAs straight from TMEM:
//   for key block j (128 keys):  load K_j, V_j
//     for query block i (<= 128 queries; causal: i >= j):  load Q_i, dO_i
//       S^T  = K_j Q_i^T,  dP^T = V_j dO_i^T               (K-major smem operands, fp32 in TMEM)
//       P^T  = exp2(S^T scale - lse),  dS^T = P^T (dP^T - delta)   (8 warps, tcgen05.ld; bf16 pairs are
//              written back over consumed S^T / dP^T columns, dS^T additionally into smem)
//       dV_j += P^T dO_i,  dK_j += dS^T Q_i                 (A from TMEM, B = the same dO_i / Q_i tiles MN-major)
//       dQ_i(j) = dS K_j                                    (A = dS^T tile read MN-major, B = K_j MN-major)
//       dQ_i: the running partial over the key blocks seen so far is parked (bf16, unscaled) in the dq rows
//             themselves; every later key block reads it back (coalesced), adds its own partial and parks
//             it again - the last one applies the softmax scale
//     Q_i / dO_i tiles are double-buffered: the next step's tiles are fetched behind the current MMAs
//     store dK_j (x scale), dV_j
// delta = rowsum(dO o O) is computed in the prologue; no global workspace, no atomics, deterministic.
// TMEM (512 columns): S^T [0,128)  dP^T [128,256)  dV [256,256+HD)  dK [256+HD,256+2HD)  dQ [448,512)+[32,64)
// ======================================================================================================
struct AttnTcBwdParams {
  const __nv_bfloat16 *q, *k, *v, *o, *dout;
  __nv_bfloat16 *dq, *dk, *dv;
  const float* lse;
  int ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  int hsq, hsk, hsv, hso, hsdo, hsdq, hsdk, hsdv;
  TcSeqMap mq, mkv, mo, mdo, mdq, mdkv;
  int n_seq, n_heads, s_q, s_kv, mask;
  long total_rows;
  float scale_log2, scale;
  int hd;  // real head_dim (<= HD), see AttnTcParams
  DropSpec drop;
  int has_drop;
};

__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// MN-major / K-major descriptor with an explicit leading-dimension byte offset
__device__ __forceinline__ uint64_t desc_sw_lbo(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}

// Row-staging buffers ([128 rows][HD*2 + 16 B], the pad spreads the banks) make every global access of
// the dQ / dK / dV outputs coalesced: 4 consecutive lanes move 64 contiguous bytes of one row.
template <int HD>
__device__ __forceinline__ void tc_rows_store256(const uint8_t* stage, const TcMat& m, int r0, int n_valid, int hd) {
  constexpr int PITCH = HD * 2 + 16, CPT = HD / 32;
  const int rl = threadIdx.x >> 2, c0 = threadIdx.x & 3;
#pragma unroll
  for (int rb = 0; rb < 128; rb += 64) {
    const int r = rb + rl;
    if (r0 + r < n_valid) {
      __nv_bfloat16* g = const_cast<__nv_bfloat16*>(tc_row(m, r0 + r));
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int c = c0 + 4 * j;
        if (c * 8 < hd) *reinterpret_cast<uint4*>(g + c * 8) = *reinterpret_cast<const uint4*>(stage + r * PITCH + c * 16);
      }
    }
  }
}
template <int HD>
__device__ __forceinline__ void tc_rows_load256(uint8_t* stage, const TcMat& m, int r0, int n_valid, int hd) {
  constexpr int PITCH = HD * 2 + 16, CPT = HD / 32;
  const int rl = threadIdx.x >> 2, c0 = threadIdx.x & 3;
#pragma unroll
  for (int rb = 0; rb < 128; rb += 64) {
    const int r = rb + rl;
    if (r0 + r < n_valid) {
      const __nv_bfloat16* g = tc_row(m, r0 + r);
#pragma unroll
      for (int j = 0; j < CPT; ++j) {
        const int c = c0 + 4 * j;
        if (c * 8 < hd) cp16(stage + r * PITCH + c * 16, g + c * 8);
      }
    }
  }
}


template <int HD, bool DROP>
__global__ void __launch_bounds__(256, 1) attn_tc_bwd_kernel(const AttnTcBwdParams p) {
  static_assert(HD == 64 || HD == 96, "head_dim 64 or 96");
  constexpr bool TWO = (HD == 96);
  constexpr uint32_t C_ST = 0, C_DPT = 128, C_DV = 256, C_DK = 256 + HD, C_DQ0 = 448, C_DQ1 = 32;
  constexpr int T0 = 16384, T1 = TWO ? 8192 : 0;  // bytes of a 128-row SWIZZLE_128B / SWIZZLE_64B tile
  constexpr int PITCH = HD * 2 + 16;              // row pitch of the staging buffers
  static_assert(128 * PITCH <= 2 * T0, "dS area doubles as a staging buffer");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw & 1023u)) & 1023u);
  uint8_t* k0s = smem;                    // [128][128 B] SWIZZLE_128B, head-dim columns 0..63
  uint8_t* v0s = k0s + T0;
  uint8_t* q0s = v0s + T0;                // Q, two stages
  uint8_t* d0s = q0s + 2 * T0;            // dO, two stages
  uint8_t* dss = d0s + 2 * T0;            // dS^T: 2 blocks (queries 0..63 / 64..127) of [128 keys][128 B]
  uint8_t* k1s = dss + 2 * T0;            // [128][64 B] SWIZZLE_64B, head-dim columns 64..95 (HD = 96)
  uint8_t* v1s = k1s + T1;
  uint8_t* q1s = v1s + T1;
  uint8_t* d1s = q1s + 2 * T1;
  uint8_t* pst = d1s + 2 * T1;            // [128][PITCH] row staging: parked dQ read-back, dK store
  float2* stats = reinterpret_cast<float2*>(pst + 128 * PITCH);  // [round128(s_q)] (lse * log2e, delta)
  uint64_t* bar = reinterpret_cast<uint64_t*>(stats + ((p.s_q + 127) & ~127));  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wq = warp & 3, half = warp >> 2;  // TMEM lane quarter, column half
#ifdef YMP_ATTN_DBG
  int dbg_n = 0;
#endif
  TDBG(0);
  const int h = blockIdx.x, s = blockIdx.y;
  int sq = p.s_q, skv = p.s_kv;
  if (p.total_rows > 0) {
    const long left = p.total_rows - (long)s * p.s_q;
    if (left <= 0) return;
    if (left < sq) sq = (int)left;
    if (left < skv) skv = (int)left;
  }
  const bool causal = p.mask == TC_MASK_CAUSAL;
  const int nkb = (skv + 127) >> 7, nqb = (sq + 127) >> 7;

  if (warp == 0) tmem_alloc<512>(tmem_ptr);
  if (threadIdx.x == 32) {
    mbar_init(&bar[0], 2);  // S^T and dP^T issuers
    mbar_init(&bar[1], 3);  // dV, dK and dQ issuers
    fence_mbar_init();
  }
  const TcMat Mq = tc_mat(p.q, p.mq, s, p.ldq, h * p.hsq);
  const TcMat Mk = tc_mat(p.k, p.mkv, s, p.ldk, h * p.hsk);
  const TcMat Mv = tc_mat(p.v, p.mkv, s, p.ldv, h * p.hsv);
  const TcMat Mdo = tc_mat(p.dout, p.mdo, s, p.lddo, h * p.hsdo);
  const TcMat Mdq = tc_mat(p.dq, p.mdq, s, p.lddq, h * p.hsdq);
  // the first tiles are in flight while the per-query statistics are computed
  tc_load256<HD>(k0s, k1s, Mk, 0, 128, skv, p.hd);
  tc_load256<HD>(v0s, v1s, Mv, 0, 128, skv, p.hd);
  {
    const int nq0 = min(128, (sq + 31) & ~31);
    tc_load256<HD>(q0s, q1s, Mq, 0, nq0, sq, p.hd);
    tc_load256<HD>(d0s, d1s, Mdo, 0, nq0, sq, p.hd);
  }
  for (int r = threadIdx.x; r < nqb * 128; r += 256) {
    float2 st = make_float2(1e30f, 0.f);  // rows that do not exist: P = exp2(x - 1e30) = 0
    if (r < sq) {
      const TcMat Mo = tc_mat(p.o, p.mo, s, p.ldo, h * p.hso);
      const uint4* po = reinterpret_cast<const uint4*>(tc_row(Mo, r));
      const uint4* pd = reinterpret_cast<const uint4*>(tc_row(Mdo, r));
      // all 2 * HD / 8 row chunks are requested before the first one is used (one round trip instead of HD / 8)
      uint4 oa[HD / 8], ob[HD / 8];
#pragma unroll
      for (int c = 0; c < HD / 8; ++c) {
        const bool in = c * 8 < p.hd;
        oa[c] = in ? __ldg(po + c) : make_uint4(0, 0, 0, 0);
        ob[c] = in ? __ldg(pd + c) : make_uint4(0, 0, 0, 0);
      }
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < HD / 8; ++c) {
        const uint4 a = oa[c], b = ob[c];
        acc += bf16_lo(a.x) * bf16_lo(b.x) + bf16_hi(a.x) * bf16_hi(b.x) + bf16_lo(a.y) * bf16_lo(b.y) + bf16_hi(a.y) * bf16_hi(b.y) +
               bf16_lo(a.z) * bf16_lo(b.z) + bf16_hi(a.z) * bf16_hi(b.z) + bf16_lo(a.w) * bf16_lo(b.w) + bf16_hi(a.w) * bf16_hi(b.w);
      }
      st.x = p.lse[((size_t)s * p.n_heads + h) * p.s_q + r] * 1.4426950408889634f;
      st.y = acc;
    }
    stats[r] = st;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr;
  const uint32_t tl = tmem + ((uint32_t)(wq * 32) << 16);
  const int kr = wq * 32 + lane;  // this thread's TMEM lane: key row (softmax, dK/dV) or query row (dQ)
  DropState ds = {};
  if (DROP) ds = drop_state(p.drop);
  const uint32_t drow0 = ((uint32_t)s * p.n_heads + h) * p.s_q;  // logical row of query 0 in the probability matrix
  TDBG(1);

  int it = 0;
  for (int j = 0; j < nkb; ++j) {
    const int kj0 = j * 128;
    const int nk = min(128, (skv - kj0 + 31) & ~31);  // key extent of this block used as an MMA K dimension
    if (j > 0) {  // every MMA that read the previous key block has completed (bar[1] of its last step)
      tc_load256<HD>(k0s, k1s, Mk, kj0, 128, skv, p.hd);
      tc_load256<HD>(v0s, v1s, Mv, kj0, 128, skv, p.hd);
    }
    const int i0 = causal ? j : 0;
    for (int i = i0; i < nqb; ++i, ++it) {
      const int qi0 = i * 128, sb = it & 1;
      const int nq = min(128, (sq - qi0 + 31) & ~31);  // query extent (MMA N of S^T / K of dV, dK)
      uint8_t *qa = q0s + sb * T0, *qb = q1s + sb * T1, *da = d0s + sb * T0, *db = d1s + sb * T1;
      TDBG(10);
      asm volatile("cp.async.wait_all;" ::: "memory");
      TDBG(11);
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      TDBG(12);
      // ---- S^T = K_j Q_i^T (thread 0), dP^T = V_j dO_i^T (thread 32): independent accumulators, so two
      // threads of different warps issue them concurrently (one thread sustains ~1 MMA / 60 clk)
      if (threadIdx.x == 0 || threadIdx.x == 32) {
        const bool second = threadIdx.x == 32;
        const uint32_t idesc = make_idesc_bf16(128, nq, 0, 0);
        const uint32_t a0 = smem_u32(second ? v0s : k0s), a1 = smem_u32(second ? v1s : k1s);
        const uint32_t b0 = smem_u32(second ? da : qa), b1 = smem_u32(second ? db : qb);
        const uint32_t dcol = tmem + (second ? C_DPT : C_ST);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          uint64_t ad, bd;
          if (ks < 4) {
            ad = desc_sw(a0 + ks * 32, 1024, LAYOUT_SW128);
            bd = desc_sw(b0 + ks * 32, 1024, LAYOUT_SW128);
          } else {
            ad = desc_sw(a1 + (ks - 4) * 32, 512, LAYOUT_SW64);
            bd = desc_sw(b1 + (ks - 4) * 32, 512, LAYOUT_SW64);
          }
          umma_bf16(dcol, ad, bd, idesc, ks > 0 ? 1u : 0u);
        }
        umma_commit(&bar[0]);
      }
      TDBG(13);
      // ---- prefetch the next step's Q / dO into the other stage (its last readers finished a step ago)
      {
        int ni = i + 1, nj = j;
        if (ni >= nqb) { nj = j + 1; ni = causal ? nj : 0; }
        if (nj < nkb && ni < nqb) {
          const int nqn = min(128, (sq - ni * 128 + 31) & ~31);
          tc_load256<HD>(q0s + (sb ^ 1) * T0, q1s + (sb ^ 1) * T1, Mq, ni * 128, nqn, sq, p.hd);
          tc_load256<HD>(d0s + (sb ^ 1) * T0, d1s + (sb ^ 1) * T1, Mdo, ni * 128, nqn, sq, p.hd);
        }
      }
      // ---- dQ bookkeeping: the partial over key blocks 0..j-1 is parked (bf16, unscaled) in the dq output rows
      // themselves; every later key block reads it back (coalesced) into pst and adds its own
      const int jl = causal ? min(i, nkb - 1) : nkb - 1;
      const bool last_kb = (j == jl);
      if (j > 0) tc_rows_load256<HD>(pst, Mdq, qi0, sq, p.hd);
      TDBG(14);
      mbar_wait(&bar[0], it & 1);
      tc_fence_after();
      TDBG(15);

      // ---- P^T and dS^T: warp (wq, half) owns key rows wq*32.. and query columns half*64..half*64+63
      const bool diag = causal && (i == j);
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int c = half * 2 + cc;     // 32-column chunk of the 128 query columns
        if (c * 32 >= nq) break;         // warp-uniform
        uint32_t sr[32], dr[32];
        tmem_ld32(tl + C_ST + c * 32, sr);
        tmem_ld32(tl + C_DPT + c * 32, dr);
        tmem_ld_wait();
        uint32_t ppk[16], dpk[16];
        const float4* st4 = reinterpret_cast<const float4*>(stats + qi0 + c * 32);
        // dropout bits of (32 queries of this chunk) x (this lane's key): a Philox call yields the words of 4
        // consecutive KEYS of one query, so the 4 lanes of a key group split the 32 queries (8 calls each) and
        // exchange their packed keep-bits - 8 calls + 4 shuffles per chunk instead of 32 calls
        uint32_t kb[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
        if (DROP) {
          uint32_t mine = 0;
          const uint32_t kg = (uint32_t)(kj0 + kr) >> 2;
#pragma unroll
          for (int jq = 0; jq < 8; ++jq) {
            const uint4 w = drop_words(ds, drow0 + (uint32_t)(qi0 + c * 32 + (lane & 3) * 8 + jq), kg);
            const uint32_t bits = (w.x >= ds.thresh ? 1u : 0u) | (w.y >= ds.thresh ? 2u : 0u) | (w.z >= ds.thresh ? 4u : 0u) |
                                  (w.w >= ds.thresh ? 8u : 0u);
            mine |= bits << (4 * jq);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) kb[i] = __shfl_sync(0xffffffffu, mine, (lane & ~3) + i) >> (lane & 3);
        }
        const float dsc = DROP ? ds.scale : 1.f;
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float4 st = st4[e];  // (lse2, delta) of two consecutive queries, warp-broadcast
          float p0 = ex2_fast(fmaf(__uint_as_float(sr[2 * e]), p.scale_log2, -st.x));
          float p1 = ex2_fast(fmaf(__uint_as_float(sr[2 * e + 1]), p.scale_log2, -st.z));
          if (diag) {  // key kj0 + kr attends query qi0 + col only if key <= query
            if (c * 32 + 2 * e < kr) p0 = 0.f;
            if (c * 32 + 2 * e + 1 < kr) p1 = 0.f;
          }
          // query 2e (2e+1) of the chunk: word kb[(2e) >> 3], bit 4 * ((2e) & 7) (already shifted by this lane's key)
          const bool k0 = (kb[e >> 2] >> (4 * ((2 * e) & 7))) & 1u, k1 = (kb[e >> 2] >> (4 * ((2 * e + 1) & 7))) & 1u;
          const float dp0 = k0 ? __uint_as_float(dr[2 * e]) * dsc : 0.f, dp1 = k1 ? __uint_as_float(dr[2 * e + 1]) * dsc : 0.f;
          const float d0 = p0 * (dp0 - st.y);
          const float d1 = p1 * (dp1 - st.w);
          ppk[e] = pack_bf16(k0 ? p0 * dsc : 0.f, k1 ? p1 * dsc : 0.f);   // dV += dropout(P)^T dO
          dpk[e] = pack_bf16(d0, d1);
        }
        const uint32_t pc = half * 64 + cc * 16;  // packed pairs land inside this warp's consumed columns
        tmem_st16(tl + C_ST + pc, ppk);
        tmem_st16(tl + C_DPT + pc, dpk);
        uint8_t* dsb = dss + half * T0;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<uint4*>(dsb + sw128_off(kr, cc * 4 + g)) = make_uint4(dpk[4 * g], dpk[4 * g + 1], dpk[4 * g + 2], dpk[4 * g + 3]);
      }
      TDBG(16);
      tmem_st_wait();
      fence_proxy_async();
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      TDBG(17);

      // ---- dV_j += P^T dO_i (thread 0), dK_j += dS^T Q_i (thread 32), dQ_i(j) = dS K_j (thread 64)
      if (threadIdx.x == 0 || threadIdx.x == 32) {
        const bool second = threadIdx.x == 32;
        const uint32_t id64 = make_idesc_bf16(128, 64, 0, 1), id32 = make_idesc_bf16(128, 32, 0, 1);
        const uint32_t b0 = smem_u32(second ? qa : da), b1 = smem_u32(second ? qb : db);
        const uint32_t dcol = tmem + (second ? C_DK : C_DV), acolb = tmem + (second ? C_DPT : C_ST);
        const int nks = nq / 16;
        for (int ks = 0; ks < nks; ++ks) {
          const uint32_t acc = (i > i0 || ks > 0) ? 1u : 0u;
          const uint32_t acol = (ks < 4) ? ks * 8 : 64 + (ks - 4) * 8;  // packed pairs: queries 0..63 at +0, 64..127 at +64
          umma_ts(dcol, acolb + acol, desc_sw(b0 + ks * 2048, 1024, LAYOUT_SW128), id64, acc);
          if (TWO) umma_ts(dcol + 64, acolb + acol, desc_sw(b1 + ks * 1024, 512, LAYOUT_SW64), id32, acc);
        }
        umma_commit(&bar[1]);
      } else if (threadIdx.x == 64) {
        const uint32_t iq64 = make_idesc_bf16(128, 64, 1, 1), iq32 = make_idesc_bf16(128, 32, 1, 1);
        const int nkk = nk / 16;
        for (int ks = 0; ks < nkk; ++ks) {
          const uint64_t ad = desc_sw_lbo(smem_u32(dss) + ks * 2048, T0, 1024, LAYOUT_SW128);
          umma_bf16(tmem + C_DQ0, ad, desc_sw(smem_u32(k0s) + ks * 2048, 1024, LAYOUT_SW128), iq64, ks > 0 ? 1u : 0u);
          if (TWO) umma_bf16(tmem + C_DQ1, ad, desc_sw(smem_u32(k1s) + ks * 1024, 512, LAYOUT_SW64), iq32, ks > 0 ? 1u : 0u);
        }
        umma_commit(&bar[1]);
      }
      TDBG(18);
      mbar_wait(&bar[1], it & 1);
      tc_fence_after();
      TDBG(19);

      // ---- dQ_i: TMEM lane = query row; chunk c of HD/32 handled by warpgroup c & 1.  The rows are staged in
      // the (now idle) dS area and stored with coalesced 16-byte accesses.
      if (j > 0) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();  // parked rows were fetched by other threads
      }
      {
        const float sc = last_kb ? p.scale : 1.f;  // parked partials stay unscaled
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) {
          if ((c & 1) != half) continue;
          uint32_t r[32];
          tmem_ld32(tl + (c < 2 ? C_DQ0 + c * 32 : C_DQ1), r);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(r[8 * g + e]);
            if (j > 0) {
              const uint4 q4 = *reinterpret_cast<const uint4*>(pst + kr * PITCH + c * 64 + g * 16);
              f[0] += bf16_lo(q4.x); f[1] += bf16_hi(q4.x); f[2] += bf16_lo(q4.y); f[3] += bf16_hi(q4.y);
              f[4] += bf16_lo(q4.z); f[5] += bf16_hi(q4.z); f[6] += bf16_lo(q4.w); f[7] += bf16_hi(q4.w);
            }
            *reinterpret_cast<uint4*>(dss + kr * PITCH + c * 64 + g * 16) =
                make_uint4(pack_bf16(f[0] * sc, f[1] * sc), pack_bf16(f[2] * sc, f[3] * sc), pack_bf16(f[4] * sc, f[5] * sc), pack_bf16(f[6] * sc, f[7] * sc));
          }
        }
      }
      tc_fence_before();
      __syncthreads();
      tc_rows_store256<HD>(dss, Mdq, qi0, sq, p.hd);
    }
    TDBG(20);
    // ---- dV_j (warpgroup 0, staged in the dS area) and dK_j x scale (warpgroup 1, staged in pst)
    __syncthreads();  // the last dQ store pass has finished reading the dS area
    {
      uint8_t* stg = half ? pst : dss;
      const float sc = half ? p.scale : 1.f;
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t r[32];
        tmem_ld32(tl + (half ? C_DK : C_DV) + c * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g)
          *reinterpret_cast<uint4*>(stg + kr * PITCH + c * 64 + g * 16) =
              make_uint4(pack_bf16(__uint_as_float(r[8 * g]) * sc, __uint_as_float(r[8 * g + 1]) * sc), pack_bf16(__uint_as_float(r[8 * g + 2]) * sc, __uint_as_float(r[8 * g + 3]) * sc),
                         pack_bf16(__uint_as_float(r[8 * g + 4]) * sc, __uint_as_float(r[8 * g + 5]) * sc), pack_bf16(__uint_as_float(r[8 * g + 6]) * sc, __uint_as_float(r[8 * g + 7]) * sc));
      }
      tc_fence_before();
      __syncthreads();
      const TcMat Mdv = tc_mat(p.dv, p.mdkv, s, p.lddv, h * p.hsdv);
      const TcMat Mdk = tc_mat(p.dk, p.mdkv, s, p.lddk, h * p.hsdk);
      tc_rows_store256<HD>(dss, Mdv, kj0, skv, p.hd);
      tc_rows_store256<HD>(pst, Mdk, kj0, skv, p.hd);
    }
  }
  TDBG(21);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

static TcSeqMap tc_map(const ymp_seqmap& m) {
  TcSeqMap r;
  r.seq_div = m.seq_div > 0 ? m.seq_div : 1;
  r.n_prefix = m.n_prefix; r.prefix_per_seq = m.prefix_per_seq;
  r.outer_stride = m.outer_stride; r.inner_stride = m.inner_stride; r.pos_stride = m.pos_stride;
  r.prefix_base = m.prefix_base; r.prefix_stride = m.prefix_stride;
  return r;
}

template <int HD, bool LONG>
static int launch_tc(const AttnTcParams& p, cudaStream_t st) {
  const int kvr = p.kv_rows;
  const int smem = 128 * 128 + 2 * kvr * 128 + (HD == 96 ? 128 * 64 + 2 * kvr * 64 : 0) + 1024 + 64 + 1024;
  static DeviceMax opted;
  if (opted.raise(smem)) {
    YMP_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel<HD, LONG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    YMP_CUDA(cudaFuncSetAttribute(attn_tc_fwd_kernel<HD, LONG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  dim3 grid((p.s_q + 127) / 128, p.n_heads, p.n_seq);
  if (p.has_drop) attn_tc_fwd_kernel<HD, LONG, true><<<grid, 256, smem, st>>>(p);
  else attn_tc_fwd_kernel<HD, LONG, false><<<grid, 256, smem, st>>>(p);
  YMP_LAUNCH_CHECK();
  return YMP_OK;
}

template <int HD>
static int launch_tc_pair(const AttnTcParams& p, int smem, cudaStream_t st) {
  static DeviceMax opted;
  if (opted.raise(smem)) {
    YMP_CUDA(cudaFuncSetAttribute(attn_tc_fwd_pair_kernel<HD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    YMP_CUDA(cudaFuncSetAttribute(attn_tc_fwd_pair_kernel<HD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  const int items = p.n_seq * p.n_heads;
  if (p.has_drop) attn_tc_fwd_pair_kernel<HD, true><<<min(items, num_sms()), 256, smem, st>>>(p);
  else attn_tc_fwd_pair_kernel<HD, false><<<min(items, num_sms()), 256, smem, st>>>(p);
  YMP_LAUNCH_CHECK();
  return YMP_OK;
}

static bool tc_head_dim(int hd) { return hd == 64 || hd == 80 || hd == 88 || hd == 96; }

// Returns YMP_ENOSUP (without setting an error) when the configuration is outside this kernel's domain.
int attn_tc_fwd_try(const ymp_attn_args* a, cudaStream_t st) {
  if (!tc_head_dim(a->head_dim)) return YMP_ENOSUP;
  // a handful of query rows against a long cache (single-token decoding) is an HBM-bound GEMV: the 16-row
  // mma.sync tiles waste far less than a 128-row tcgen05 tile would
  if (a->s_q < 16 && a->s_kv > 256) return YMP_ENOSUP;
  // packed block-diagonal (temporal) sequences use 1/8 of each score tile: the mma.sync kernel,
  // which skips the masked chunks per warp, measured faster there (0.113 vs 0.128 ms)
  if (a->mask == YMP_MASK_BLOCK) return YMP_ENOSUP;
  if (a->q_head_stride % 8 || a->k_head_stride % 8 || a->v_head_stride % 8 || a->ldo % 8 || a->o_head_stride % 8) return YMP_ENOSUP;
  AttnTcParams p = {};
  p.q = (const __nv_bfloat16*)a->q; p.k = (const __nv_bfloat16*)a->k; p.v = (const __nv_bfloat16*)a->v;
  p.o = (__nv_bfloat16*)a->o; p.lse = a->lse;
  p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv; p.ldo = a->ldo;
  p.hsq = a->q_head_stride; p.hsk = a->k_head_stride; p.hsv = a->v_head_stride; p.hso = a->o_head_stride;
  p.mq = tc_map(a->map_q); p.mkv = tc_map(a->map_kv); p.mo = tc_map(a->map_o);
  p.n_seq = a->n_seq; p.n_heads = a->n_heads; p.s_q = a->s_q; p.s_kv = a->s_kv;
  p.mask = a->mask; p.mask_block = a->mask_block > 0 ? a->mask_block : 1; p.total_rows = a->total_rows;
  p.scale = a->scale; p.scale_log2 = a->scale * 1.4426950408889634f;
  p.hd = a->head_dim;
  p.has_drop = (a->drop.rng && a->drop.p > 0.f) ? 1 : 0;
  p.drop.rng = a->drop.rng; p.drop.site = a->drop.site; p.drop.p = a->drop.p;
  const bool lng = a->s_kv > 256;
  p.kv_rows = lng ? 256 : ((a->s_kv + 31) & ~31);
  // pair-tile persistent kernel: both query tiles of a (sequence, head) share one K / V fetch, two stages in flight
  static const bool no_pair = [] { const char* e = getenv("YMP_ATTN_NO_PAIR"); return e && e[0] == '1'; }();
  // measured (profiles/r02e_attn_probe_*.log): +3 % on the ViT spatial shape (197 rows: the second tile is mostly padding,
  // sharing K / V pays), -3 % on the causal GPT shape (tile 0 needs half the keys: the single-tile kernel loads less)
  static const bool pair_causal = [] { const char* e = getenv("YMP_ATTN_PAIR_CAUSAL"); return e && e[0] == '1'; }();
  if (!lng && !no_pair && a->s_q > 128 && a->s_q <= 256 && (a->mask == YMP_MASK_NONE || (pair_causal && a->mask == YMP_MASK_CAUSAL))) {
    const int HDp = a->head_dim == 64 ? 64 : 96;
    const int smem = 2 * 128 * HDp * 2 + 2 * 2 * p.kv_rows * HDp * 2 + 64 + 1024;
    if (smem <= 227 * 1024) return HDp == 64 ? launch_tc_pair<64>(p, smem, st) : launch_tc_pair<96>(p, smem, st);
  }
  if (a->head_dim == 64) return lng ? launch_tc<64, true>(p, st) : launch_tc<64, false>(p, st);
  return lng ? launch_tc<96, true>(p, st) : launch_tc<96, false>(p, st);
}

template <int HD>
static int launch_tc_bwd(const AttnTcBwdParams& p, cudaStream_t st) {
  const int smem = 8 * 16384 + (HD == 96 ? 6 * 8192 : 0) + 128 * (HD * 2 + 16) + ((p.s_q + 127) & ~127) * 8 + 64 + 1024;
  static DeviceMax opted;
  if (opted.raise(smem)) {
    YMP_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel<HD, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    YMP_CUDA(cudaFuncSetAttribute(attn_tc_bwd_kernel<HD, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  dim3 grid(p.n_heads, p.n_seq);
  if (p.has_drop) attn_tc_bwd_kernel<HD, true><<<grid, 256, smem, st>>>(p);
  else attn_tc_bwd_kernel<HD, false><<<grid, 256, smem, st>>>(p);
  YMP_LAUNCH_CHECK();
  return YMP_OK;
}

// Backward counterpart of attn_tc_fwd_try: YMP_ENOSUP when outside the kernel's domain.
int attn_tc_bwd_try(const ymp_attn_bwd_args* b, cudaStream_t st) {
  const ymp_attn_args* a = &b->fwd;
  if (!tc_head_dim(a->head_dim) || a->s_q > 2048) return YMP_ENOSUP;
  if (a->mask == YMP_MASK_BLOCK) return YMP_ENOSUP;
  if (a->mask == YMP_MASK_CAUSAL && a->s_q != a->s_kv) return YMP_ENOSUP;
  if (a->q_head_stride % 8 || a->k_head_stride % 8 || a->v_head_stride % 8 || a->ldo % 8 || a->o_head_stride % 8 || b->do_head_stride % 8 || b->dq_head_stride % 8 ||
      b->dk_head_stride % 8 || b->dv_head_stride % 8)
    return YMP_ENOSUP;
  AttnTcBwdParams p = {};
  p.q = (const __nv_bfloat16*)a->q; p.k = (const __nv_bfloat16*)a->k; p.v = (const __nv_bfloat16*)a->v;
  p.o = (const __nv_bfloat16*)a->o; p.dout = (const __nv_bfloat16*)b->dout; p.lse = a->lse;
  p.dq = (__nv_bfloat16*)b->dq; p.dk = (__nv_bfloat16*)b->dk; p.dv = (__nv_bfloat16*)b->dv;
  p.ldq = a->ldq; p.ldk = a->ldk; p.ldv = a->ldv; p.ldo = a->ldo;
  p.lddo = b->lddo; p.lddq = b->lddq; p.lddk = b->lddk; p.lddv = b->lddv;
  p.hsq = a->q_head_stride; p.hsk = a->k_head_stride; p.hsv = a->v_head_stride; p.hso = a->o_head_stride;
  p.hsdo = b->do_head_stride; p.hsdq = b->dq_head_stride; p.hsdk = b->dk_head_stride; p.hsdv = b->dv_head_stride;
  p.mq = tc_map(a->map_q); p.mkv = tc_map(a->map_kv); p.mo = tc_map(a->map_o);
  p.mdo = tc_map(b->map_do); p.mdq = tc_map(b->map_dq); p.mdkv = tc_map(b->map_dkv);
  p.n_seq = a->n_seq; p.n_heads = a->n_heads; p.s_q = a->s_q; p.s_kv = a->s_kv;
  p.mask = a->mask; p.total_rows = a->total_rows;
  p.scale = a->scale; p.scale_log2 = a->scale * 1.4426950408889634f;
  p.hd = a->head_dim;
  p.has_drop = (a->drop.rng && a->drop.p > 0.f) ? 1 : 0;
  p.drop.rng = a->drop.rng; p.drop.site = a->drop.site; p.drop.p = a->drop.p;
  return a->head_dim == 64 ? launch_tc_bwd<64>(p, st) : launch_tc_bwd<96>(p, st);
}

}  // namespace ymp

#ifdef YMP_ATTN_DBG
extern "C" int ymp_attn_dbg_read(unsigned long long* out) {
  return (int)cudaMemcpyFromSymbol(out, ymp::ymp_attn_dbg_buf, sizeof(unsigned long long) * 256);
}
#endif
