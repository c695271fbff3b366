/*
 * ymp.h - C ABI of libymp_b200.so: the sm_100a kernels behind the mPLUG-Video (Youku-mPLUG)
 * pre-training hot path.
 *
 * The reference (X-PLUG/Youku-mPLUG) has no native boundary of its own: every device op on the
 * path is a PyTorch library call made from models/{vision_transformer,modeling_distributed_gpt3,
 * distributed_gpt3}.py.  Each entry point below therefore cites the reference *call site* whose
 * device computation it replaces (file:line, relative to the reference repo root).
 *
 * Contract (all entry points):
 *   - plain pointers + sizes only; device pointers unless stated otherwise; no torch types
 *   - never allocates, never frees, never synchronises; work is enqueued on `stream`
 *     (a cudaStream_t passed as void*), so every call is CUDA-graph capturable
 *   - returns 0 on success, a negative YMP_E* code on failure; ymp_last_error() returns a
 *     thread-local human-readable message for the last failure on this thread
 *   - bf16 tensors are raw uint16 storage (__nv_bfloat16), fp32 tensors are float
 */
#ifndef YMP_H_
#define YMP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YMP_OK 0
#define YMP_EINVAL (-1)  /* bad argument (shape, alignment, null pointer) */
#define YMP_ECUDA (-2)   /* CUDA runtime / driver error (message has the CUDA string) */
#define YMP_ENOSUP (-3)  /* configuration not supported by the compiled kernels */

const char* ymp_last_error(void);
/* ABI version of this header; bumped on any struct change. */
int ymp_abi_version(void);
/* Number of kernel launches this process has enqueued through the library (for gpu_launches). */
uint64_t ymp_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * GEMM  D[M,N] = epilogue( alpha * op(A)[M,K] . op(B)[N,K]^T )          tcgen05 + TMA + TMEM
 *
 * Replaces: F.linear in ViT Attention qkv/proj (models/vision_transformer.py:171-176,205),
 * temporal_fc (:250), Mlp fc1/fc2 (:103-110), AttentionPool in/out proj + MLP (:368-374),
 * visual_fc (models/distributed_gpt3.py:136), GPT-3 query_key_value / dense / dense_h_to_4h /
 * dense_4h_to_h (models/modeling_distributed_gpt3.py:843-857,562-578), the tied LM head
 * (:1348-1350), and every dgrad / wgrad GEMM autograd derives from them.
 *
 * Operand storage (bf16):
 *   a_mn_major = 0: A is [M,K] row-major, row stride lda  ("K-major")
 *   a_mn_major = 1: A is stored transposed, [K,M] row-major, row stride lda ("MN-major")
 *   b_mn_major = 0: B is [N,K] row-major, row stride ldb  (a torch nn.Linear weight)
 *   b_mn_major = 1: B is [K,N] row-major, row stride ldb
 *   All leading dimensions must be multiples of 8 elements and base pointers 16-byte aligned.
 * Epilogue, applied in this order to v = alpha * acc:
 *   v += bias[n]                         (bias: bf16 [N], optional)
 *   if aux_out: aux_out[m,n] = bf16(v)   (pre-activation saved for backward, ld = ldd)
 *   v = act(v)                           (YMP_ACT_*)
 *   if aux_in:  v *= act'(aux_in[m,n])   (backward of an activation; act selects which GELU;
 *                                          with aux_in set, act() itself is NOT applied)
 *   v += residual[m,n]                   (bf16 [M,N], row stride ldr, optional)
 *   D[m,n] = v  (bf16 or fp32) ; or atomically D[m,n] += v (fp32, accumulate=1, used by split-K)
 * ------------------------------------------------------------------------------------------ */
#define YMP_ACT_NONE 0
#define YMP_ACT_GELU_ERF 1  /* nn.GELU() exact: ViT / abstractor MLP (vision_transformer.py:94) */
#define YMP_ACT_GELU_TANH 2 /* Megatron bias_gelu_impl: GPT-3 MLP (modeling_distributed_gpt3.py:586) */

#define YMP_DT_BF16 0
#define YMP_DT_F32 1

typedef struct ymp_gemm_args {
  const void* A;
  const void* B;
  void* D;
  int32_t M, N, K;
  int32_t lda, ldb, ldd;
  int32_t a_mn_major, b_mn_major;
  const void* bias;     /* bf16 [N] or NULL */
  const void* residual; /* bf16 [M,N] or NULL */
  int32_t ldr;
  int32_t act;          /* YMP_ACT_* */
  void* aux_out;        /* bf16 [M,N] (ld = ldd) or NULL */
  const void* aux_in;   /* bf16 [M,N] (ld = ldd) or NULL */
  int32_t out_dtype;    /* YMP_DT_BF16 | YMP_DT_F32 */
  int32_t accumulate;   /* 1: D (fp32) += result using atomics */
  int32_t split_k;      /* >=1; >1 requires accumulate=1 and out_dtype=F32; 0 = library picks */
  float alpha;
  int32_t tile_n;       /* 0 = auto, else 128 or 256 */
} ymp_gemm_args;

int ymp_gemm(const ymp_gemm_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* YMP_H_ */
