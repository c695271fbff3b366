"""Tensor-level wrappers over the C ABI (one function per entry point of include/ymp.h).

Inputs are torch CUDA tensors used purely as device buffers; every function enqueues exactly the
kernels of one ABI call on the current stream and returns the output tensor(s).
"""
import torch

from . import lib as L
from .lib import ACT_NONE, ACT_GELU_ERF, ACT_GELU_TANH, DT_BF16, DT_F32  # noqa: F401

bf16 = torch.bfloat16


def _chk2d(t, name):
    assert t.is_cuda and t.dim() == 2 and t.stride(1) == 1, f"{name}: need 2-D row-major CUDA tensor"


def gemm(a, b, *, a_t=False, b_t=False, bias=None, residual=None, act=ACT_NONE, aux_out=None,
         aux_in=None, out=None, out_dtype=bf16, accumulate=False, split_k=0, alpha=1.0, tile_n=0):
    """D[M,N] = epilogue(alpha * op(A) @ op(B)^T).

    a: [M,K] (or [K,M] when a_t)      b: [N,K] like nn.Linear.weight (or [K,N] when b_t)
    """
    _chk2d(a, "a"); _chk2d(b, "b")
    assert a.dtype == bf16 and b.dtype == bf16
    M, K = (a.shape[1], a.shape[0]) if a_t else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_t else (b.shape[0], b.shape[1])
    assert K == Kb, f"gemm: K mismatch {K} vs {Kb}"
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
        assert not accumulate, "accumulate needs an explicit (zeroed or running) output"
    _chk2d(out, "out")
    assert out.shape == (M, N)
    g = L.GemmArgs()
    g.A, g.B, g.D = a.data_ptr(), b.data_ptr(), out.data_ptr()
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb, g.ldd = a.stride(0), b.stride(0), out.stride(0)
    g.a_mn_major, g.b_mn_major = int(a_t), int(b_t)
    g.bias = L.ptr(bias)
    g.residual = L.ptr(residual)
    g.ldr = residual.stride(0) if residual is not None else 0
    g.act = act
    if aux_out is not None:
        assert aux_out.dtype == bf16 and aux_out.shape == (M, N) and aux_out.stride(0) == out.stride(0)
    if aux_in is not None:
        assert aux_in.dtype == bf16 and aux_in.shape == (M, N) and aux_in.stride(0) == out.stride(0)
    g.aux_out, g.aux_in = L.ptr(aux_out), L.ptr(aux_in)
    g.out_dtype = DT_F32 if out.dtype == torch.float32 else DT_BF16
    g.accumulate = int(accumulate)
    g.split_k = split_k
    g.alpha = alpha
    g.tile_n = tile_n
    L.call(L._gemm, g, "ymp_gemm")
    return out
