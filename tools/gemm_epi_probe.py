"""GPU probe: the epilogue-heavy ViT GEMM shapes of the step, timed with CUDA events."""
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


def report(name, M, N, K, ms):
    print("EPI " + json.dumps(dict(case=name, M=M, N=N, K=K, ms=round(ms, 4), tflops=round(2 * M * N * K / ms / 1e9, 1))))


M = 50208
x = (torch.randn(M, 768, device=dev) * 0.5).to(bf16)
w1 = (torch.randn(3072, 768, device=dev) * 0.03).to(bf16)
b1 = torch.randn(3072, device=dev).to(bf16)
h = torch.empty(M, 3072, device=dev, dtype=bf16)
hp = torch.empty_like(h)
report("fc1 fwd erf+act'", M, 3072, 768, timeit(lambda: ops.gemm(x, w1, bias=b1, act=1, out=h, aux_out=hp)))
report("fc1 fwd erf (no aux)", M, 3072, 768, timeit(lambda: ops.gemm(x, w1, bias=b1, act=1, out=h)))
report("fc1 fwd plain", M, 3072, 768, timeit(lambda: ops.gemm(x, w1, bias=b1, out=h)))
# check against torch
ref = torch.nn.functional.gelu(x[:4096].float() @ w1.float().t() + b1.float())
ops.gemm(x, w1, bias=b1, act=1, out=h, aux_out=hp)
err = (h[:4096].float() - ref).abs().max().item()
print("EPI " + json.dumps(dict(case="gelu max abs err vs torch fp32", err=err, ref_max=ref.abs().max().item())))
w2 = (torch.randn(768, 768, device=dev) * 0.03).to(bf16)
b2 = torch.randn(768, device=dev).to(bf16)
res = torch.randn(M, 768, device=dev)
out32 = torch.empty(M, 768, device=dev)
report("proj +bias +fp32 residual -> fp32", M, 768, 768, timeit(lambda: ops.gemm(x, w2, bias=b2, residual=res, out=out32)))
outb = torch.empty(M, 768, device=dev, dtype=bf16)
report("proj +bias -> bf16", M, 768, 768, timeit(lambda: ops.gemm(x, w2, bias=b2, out=outb)))
dy = (torch.randn(M, 768, device=dev) * 0.5).to(bf16)
dw = torch.zeros(768, 768, device=dev)
report("wgrad 768x768", 768, 768, M, timeit(lambda: ops.gemm(dy, x, a_t=True, b_t=True, out=dw, accumulate=True)))
dw3 = torch.zeros(3072, 768, device=dev)
report("wgrad 3072x768", 3072, 768, M, timeit(lambda: ops.gemm(h, x, a_t=True, b_t=True, out=dw3, accumulate=True)))
dh = torch.empty(M, 3072, device=dev, dtype=bf16)
w2t = (torch.randn(768, 3072, device=dev) * 0.03).to(bf16)
report("fc2 dgrad * act'", M, 3072, 768, timeit(lambda: ops.gemm(dy, w2t, b_t=True, out=dh, aux_in=hp, act=1)))
