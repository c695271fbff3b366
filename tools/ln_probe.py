"""GPU probe: LayerNorm forward / backward at the path's two shapes (ViT rows x 768 with affine grads,
GPT rows x 2048 frozen), CUDA-event timed; run under `ncu --set full -k regex:ln_` for pipe data."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youku-mplug_b200"))
import torch  # noqa: E402
from ymp import ops  # noqa: E402

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
    return e0.elapsed_time(e1) / n


def case(name, rows, D, wgrad, reps):
    x = torch.randn(rows, D, device=dev)
    gamma = torch.randn(D, device=dev).to(bf16)
    beta = torch.randn(D, device=dev).to(bf16)
    y, mean, rstd = ops.layernorm_fwd(x, gamma, beta, 1e-5)
    t_f = timeit(lambda: ops.layernorm_fwd(x, gamma, beta, 1e-5, out=y), reps)
    dy = torch.randn(rows, D, device=dev).to(bf16)
    add = torch.randn(rows, D, device=dev).to(bf16)
    dx = torch.empty(rows, D, device=dev, dtype=bf16)
    dg = torch.zeros(D, device=dev) if wgrad else None
    db = torch.zeros(D, device=dev) if wgrad else None
    t_b = timeit(lambda: ops.layernorm_bwd(dy, x, gamma, mean, rstd, add=add, dgamma=dg, dbeta=db, dx=dx), reps)
    fb = rows * D * (4 + 2) / 1e6
    bb = rows * D * (2 + 4 + 2 + 2) / 1e6
    print("LN " + json.dumps(dict(case=name, rows=rows, D=D, fwd_us=round(t_f * 1e3, 1), fwd_gbs=round(fb / t_f), bwd_us=round(t_b * 1e3, 1), bwd_gbs=round(bb / t_b))))


if __name__ == "__main__":
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    case("vit", 50208, 768, True, reps)
    case("gpt", 8192, 2048, False, reps)
