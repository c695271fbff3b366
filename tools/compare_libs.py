"""GPU probe: this repo's kernels next to the library kernels the reference's GPU path would call on the same
box (torch 2.11: cuBLAS matmul, SDPA / flash-attention, F.layer_norm) at the step's shapes.  Comparison
points only - nothing here is on the product path."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youku-mplug_b200"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from ymp import ops  # noqa: E402
from ymp.ops import TView  # noqa: E402

dev, bf16 = torch.device("cuda"), torch.bfloat16


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


def out(kind, **kw):
    print("CMP " + json.dumps(dict(kind=kind, **{k: (round(v, 1) if isinstance(v, float) else v) for k, v in kw.items()})))


def gemms():
    for (M, N, K) in [(8192, 8192, 2048), (8192, 2048, 8192), (8192, 6144, 2048), (8192, 2048, 2048), (50208, 3072, 768),
                      (50208, 768, 3072), (50208, 2304, 768), (50208, 768, 768), (8192, 51200, 2048)]:
        a = torch.randn(M, K, device=dev).to(bf16)
        b = torch.randn(N, K, device=dev).to(bf16)
        o = torch.empty(M, N, device=dev, dtype=bf16)
        t_y = timeit(lambda: ops.gemm(a, b, out=o))
        t_t = timeit(lambda: torch.matmul(a, b.t(), out=o))
        fl = 2.0 * M * N * K
        out("gemm", M=M, N=N, K=K, ymp_us=t_y, cublas_us=t_t, ymp_tflops=fl / t_y / 1e6, cublas_tflops=fl / t_t / 1e6)
    # wgrad-shaped (A^T B over a long K) - fp32 accumulate in ours, bf16 out in torch
    for (Mo, No, K) in [(768, 768, 50208), (3072, 768, 50208)]:
        a = torch.randn(K, Mo, device=dev).to(bf16)
        b = torch.randn(K, No, device=dev).to(bf16)
        o32 = torch.zeros(Mo, No, device=dev)
        t_y = timeit(lambda: ops.gemm(a, b, a_t=True, b_t=True, out=o32, accumulate=True))
        t_t = timeit(lambda: torch.matmul(a.t(), b))
        fl = 2.0 * Mo * No * K
        out("wgrad", M=Mo, N=No, K=K, ymp_us=t_y, cublas_us=t_t, ymp_tflops=fl / t_y / 1e6, cublas_tflops=fl / t_t / 1e6)


def attention():
    cases = [("vit_spatial", 256, 8, 197, 96, False), ("gpt_causal", 32, 32, 256, 64, True)]
    for name, n, heads, S, hd, causal in cases:
        C = heads * hd
        qkv = (torch.randn(n * S, 3 * C, device=dev) * 0.5).to(bf16)
        o = torch.empty(n * S, C, device=dev, dtype=bf16)
        m = ops.dense_map(S)
        q, k, v = (TView(qkv, i * C, hd, m) for i in range(3))
        kw = dict(n_seq=n, n_heads=heads, head_dim=hd, s_q=S, s_kv=S, causal=causal, scale=hd ** -0.5)
        lse = ops.attn_fwd(q, k, v, TView(o, 0, hd, m), **kw)
        t_f = timeit(lambda: ops.attn_fwd(q, k, v, TView(o, 0, hd, m), lse=lse, **kw))
        do = torch.randn_like(o)
        dqkv = torch.empty_like(qkv)
        dq, dk, dv = (TView(dqkv, i * C, hd, m) for i in range(3))
        t_b = timeit(lambda: ops.attn_bwd(q, k, v, TView(o, 0, hd, m), lse, TView(do, 0, hd, m), dq, dk, dv, **kw))
        # torch SDPA on [n, heads, S, hd] (flash backend), contiguous inputs = best case for the library
        qq, kk, vv = (torch.randn(n, heads, S, hd, device=dev).to(bf16).requires_grad_() for _ in range(3))
        t_sf = timeit(lambda: F.scaled_dot_product_attention(qq, kk, vv, is_causal=causal))
        oo = F.scaled_dot_product_attention(qq, kk, vv, is_causal=causal)
        g = torch.randn_like(oo)
        t_sb = timeit(lambda: torch.autograd.grad(oo, (qq, kk, vv), g, retain_graph=True))
        res = dict(ymp_fwd_us=t_f, ymp_bwd_us=t_b, sdpa_fwd_us=t_sf, sdpa_bwd_us=t_sb)
        try:
            from flash_attn import flash_attn_func
            q2, k2, v2 = (torch.randn(n, S, heads, hd, device=dev).to(bf16).requires_grad_() for _ in range(3))
            res["flash_attn_fwd_us"] = timeit(lambda: flash_attn_func(q2, k2, v2, causal=causal))
            o2 = flash_attn_func(q2, k2, v2, causal=causal)
            g2 = torch.randn_like(o2)
            res["flash_attn_bwd_us"] = timeit(lambda: torch.autograd.grad(o2, (q2, k2, v2), g2, retain_graph=True))
        except Exception as e:  # noqa: BLE001
            res["flash_attn"] = "unavailable: " + type(e).__name__
        out("attention", shape=name, **res)


def layernorm():
    for rows, D in [(50208, 768), (8192, 2048)]:
        x = torch.randn(rows, D, device=dev)
        g = torch.randn(D, device=dev).to(bf16)
        b = torch.randn(D, device=dev).to(bf16)
        y = torch.empty(rows, D, device=dev, dtype=bf16)
        t_y = timeit(lambda: ops.layernorm_fwd(x, g, b, 1e-5, out=y))
        xb = x.to(bf16)
        t_t = timeit(lambda: F.layer_norm(xb, (D,), g, b, 1e-5))
        t_t32 = timeit(lambda: F.layer_norm(x, (D,), g.float(), b.float(), 1e-5).to(bf16))
        out("layernorm_fwd", rows=rows, D=D, ymp_fp32_in_us=t_y, torch_bf16_in_us=t_t, torch_fp32_in_cast_us=t_t32)


if __name__ == "__main__":
    gemms()
    attention()
    layernorm()
