import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "youku-mplug_b200"))
import torch
from ymp import ops
dev, bf16 = torch.device("cuda"), torch.bfloat16
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
K = 50208
for (Mo, No) in [(768, 768), (2304, 768), (3072, 768), (768, 3072)]:
    a = (torch.randn(K, Mo, device=dev) * 0.5).to(bf16)
    b = (torch.randn(K, No, device=dev) * 0.5).to(bf16)
    out = torch.zeros(Mo, No, device=dev)
    for sk in [0, 2, 4, 8, 16, 32]:
        for tn in ([512] if True else [512]):
            ms = timeit(lambda: ops.gemm(a, b, a_t=True, b_t=True, out=out, accumulate=True, split_k=sk, tile_n=tn))
            print(f"SPLITK M={Mo} N={No} split={sk} tile={tn} {ms*1e3:.1f} us {2*Mo*No*K/ms/1e9:.0f} TF/s")
