"""One launch of each epilogue-heavy ViT GEMM (for `ncu --set full -k regex:gemm`)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youku-mplug_b200"))
import torch  # noqa: E402
from ymp import ops  # noqa: E402

dev, bf16 = torch.device("cuda"), torch.bfloat16
M = 50208
x = (torch.randn(M, 768, device=dev) * 0.5).to(bf16)
w1 = (torch.randn(3072, 768, device=dev) * 0.03).to(bf16)
b1 = torch.randn(3072, device=dev).to(bf16)
h = torch.empty(M, 3072, device=dev, dtype=bf16)
hp = torch.empty_like(h)
w2 = (torch.randn(768, 768, device=dev) * 0.03).to(bf16)
b2 = torch.randn(768, device=dev).to(bf16)
res = torch.randn(M, 768, device=dev)
out32 = torch.empty(M, 768, device=dev)
dy = (torch.randn(M, 768, device=dev) * 0.5).to(bf16)
dw = torch.zeros(768, 768, device=dev)
for _ in range(2):
    ops.gemm(x, w1, bias=b1, out=h)                                  # 0: plain K=768
    ops.gemm(x, w1, bias=b1, act=1, out=h, aux_out=hp)               # 1: erf + act'
    ops.gemm(x, w2, bias=b2, residual=res, out=out32)                # 2: fp32 residual stream
    ops.gemm(dy, x, a_t=True, b_t=True, out=dw, accumulate=True)     # 3: wgrad 768x768 split-K
torch.cuda.synchronize()
