"""GPU probe: the path's attention shapes (ViT spatial / temporal, GPT causal, abstractor cross) forward
and backward, timed with CUDA events; run it under `ncu --set full -k regex:attn_` for stall data."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youku-mplug_b200"))
import torch  # noqa: E402
from ymp import ops  # noqa: E402
from ymp.ops import TView  # noqa: E402

dev = torch.device("cuda")
bf16 = torch.bfloat16


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def spatial(B=32, N=196, T=8, heads=8, hd=96, reps=5):
    D = heads * hd
    R, RB = B * N * T, B * N * T + B
    qkv = (torch.randn(RB, 3 * D, device=dev) * 0.5).to(bf16)
    att = torch.empty(RB + B * T, D, device=dev, dtype=bf16)
    m_in = ops.seqmap(seq_div=T, outer_stride=N * T, inner_stride=1, pos_stride=T, n_prefix=1, prefix_base=R, prefix_stride=1)
    m_out = ops.seqmap(seq_div=T, outer_stride=N * T, inner_stride=1, pos_stride=T, n_prefix=1, prefix_base=RB, prefix_stride=1, prefix_per_seq=1)
    q, k, v = (TView(qkv, i * D, hd, m_in) for i in range(3))
    kw = dict(n_seq=B * T, n_heads=heads, head_dim=hd, s_q=N + 1, s_kv=N + 1, causal=False, scale=hd ** -0.5)
    lse = ops.attn_fwd(q, k, v, TView(att, 0, hd, m_out), **kw)
    t_f = timeit(lambda: ops.attn_fwd(q, k, v, TView(att, 0, hd, m_out), lse=lse, **kw), reps)
    datt = torch.randn_like(att)
    dqkv = torch.empty(RB + B * T, 3 * D, device=dev, dtype=bf16)
    dq, dk, dv = (TView(dqkv, i * D, hd, m_out) for i in range(3))
    t_b = timeit(lambda: ops.attn_bwd(q, k, v, TView(att, 0, hd, m_out), lse, TView(datt, 0, hd, m_out), dq, dk, dv, **kw), reps)
    fl = 4.0 * (N + 1) ** 2 * hd * heads * B * T
    print("ATTN " + json.dumps(dict(shape="vit_spatial", fwd_ms=t_f, bwd_ms=t_b, fwd_tflops=fl / t_f / 1e9, bwd_tflops=2.5 * fl / t_b / 1e9)))


def temporal(B=32, N=196, T=8, heads=8, hd=96, reps=5):
    D = heads * hd
    R = B * N * T
    qkv = (torch.randn(R, 3 * D, device=dev) * 0.5).to(bf16)
    out = torch.empty(R, D, device=dev, dtype=bf16)
    lse = ops.attn_temporal_fwd(qkv, out, R=R, n_heads=heads, T=T, D=hd, scale=hd ** -0.5)
    t_f = timeit(lambda: ops.attn_temporal_fwd(qkv, out, R=R, n_heads=heads, T=T, D=hd, scale=hd ** -0.5), reps)
    dout = torch.randn_like(out)
    dqkv = torch.empty_like(qkv)
    t_b = timeit(lambda: ops.attn_temporal_bwd(qkv, out, lse, dout, dqkv, R=R, n_heads=heads, T=T, D=hd, scale=hd ** -0.5), reps)
    gb = (qkv.numel() + out.numel()) * 2 / 1e6
    print("ATTN " + json.dumps(dict(shape="vit_temporal", fwd_ms=t_f, bwd_ms=t_b, fwd_gbs=gb / t_f, bwd_gbs=(2 * qkv.numel() + 2 * out.numel()) * 2 / 1e6 / t_b)))


def gpt(B=32, S=256, heads=32, hd=64, reps=5):
    H = heads * hd
    qkv = (torch.randn(B * S, 3 * H, device=dev) * 0.5).to(bf16)
    att = torch.empty(B * S, H, device=dev, dtype=bf16)
    m = ops.dense_map(S)
    q, k, v = (TView(qkv, i * hd, 3 * hd, m) for i in range(3))
    kw = dict(n_seq=B, n_heads=heads, head_dim=hd, s_q=S, s_kv=S, causal=True, scale=hd ** -0.5)
    lse = ops.attn_fwd(q, k, v, TView(att, 0, hd, m), **kw)
    t_f = timeit(lambda: ops.attn_fwd(q, k, v, TView(att, 0, hd, m), lse=lse, **kw), reps)
    datt = torch.randn_like(att)
    dqkv = torch.empty_like(qkv)
    dq, dk, dv = (TView(dqkv, i * hd, 3 * hd, m) for i in range(3))
    t_b = timeit(lambda: ops.attn_bwd(q, k, v, TView(att, 0, hd, m), lse, TView(datt, 0, hd, m), dq, dk, dv, **kw), reps)
    fl = 4.0 * S * S * hd * heads * B  # full SxS as the reference computes it (algorithmic)
    print("ATTN " + json.dumps(dict(shape="gpt_causal", fwd_ms=t_f, bwd_ms=t_b, fwd_tflops=fl / t_f / 1e9, bwd_tflops=2.5 * fl / t_b / 1e9)))


if __name__ == "__main__":
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    spatial(reps=reps)
    temporal(reps=reps)
    gpt(reps=reps)
