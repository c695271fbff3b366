"""Where the 2-CTA GEMM's MMA warp waits (debug build only):
   make -C youku-mplug_b200/csrc EXTRA=-DYMP_GEMM_DBG (after touching gemm_tcgen05.cu); python tools/gemm_dbg.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youku-mplug_b200"))
import torch  # noqa: E402
from ymp import lib as L, ops  # noqa: E402

dev, bf16 = torch.device("cuda"), torch.bfloat16


def run(name, fn, n=5):
    buf = (ctypes.c_ulonglong * 16)()
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    L.lib.ymp_gemm_dbg_read(buf, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    L.lib.ymp_gemm_dbg_read(buf, 1)
    v = [int(x) for x in buf]
    tiles = max(v[3], 1)
    # v[0] is the MMA warp's span of the LAST launch; waits are summed over n launches
    print(f"{name:38s} {ms*1e3:7.1f} us  tiles/launch {tiles/n:5.1f}  span/tile {v[0]/(tiles/n):7.0f} clk | "
          f"wait acc {v[1]/tiles:6.0f}  wait smem {v[2]/tiles:6.0f} | producer wait-empty {v[4]/tiles:6.0f} | "
          f"epi wait-tfull {v[5]/tiles:6.0f} epi busy {v[6]/tiles:6.0f}")


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
outb = torch.empty(M, 768, device=dev, dtype=bf16)
dy = (torch.randn(M, 768, device=dev) * 0.5).to(bf16)
dw = torch.zeros(768, 768, device=dev)
w3 = (torch.randn(768, 3072, device=dev) * 0.03).to(bf16)
run("fc1 plain 50208x3072x768", lambda: ops.gemm(x, w1, bias=b1, out=h))
run("fc1 erf+act'", lambda: ops.gemm(x, w1, bias=b1, act=1, out=h, aux_out=hp))
run("fc2 50208x768x3072 +fp32 res", lambda: ops.gemm(h, w3, bias=b2, residual=res, out=out32))
run("proj 50208x768x768 +fp32 res", lambda: ops.gemm(x, w2, bias=b2, residual=res, out=out32))
run("proj 50208x768x768 bf16", lambda: ops.gemm(x, w2, bias=b2, out=outb))
run("wgrad 768x768x50208", lambda: ops.gemm(dy, x, a_t=True, b_t=True, out=dw, accumulate=True))
g = (torch.randn(8192, 2048, device=dev) * 0.5).to(bf16)
wg = (torch.randn(8192, 2048, device=dev) * 0.03).to(bf16)
og = torch.empty(8192, 8192, device=dev, dtype=bf16)
run("gpt fc1 8192x8192x2048 plain", lambda: ops.gemm(g, wg, out=og))
