"""Phase timeline of one CTA of the tcgen05 attention backward (debug build only):
   make -C youku-mplug_b200/csrc clean all EXTRA=-DYMP_ATTN_DBG ; python tools/attn_dbg.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "youku-mplug_b200"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402
from ymp import lib as L  # noqa: E402
import attn_probe  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "spatial"
if which == "fwd":  # forward only (the backward would overwrite the trace buffer)
    from ymp import ops
    from ymp.ops import TView
    B, N, T, heads, hd = 32, 196, 8, 8, 96
    D, R, RB = heads * hd, B * N * T, B * N * T + B
    qkv = (torch.randn(RB, 3 * D, device="cuda") * 0.5).to(torch.bfloat16)
    att = torch.empty(RB + B * T, D, device="cuda", dtype=torch.bfloat16)
    m_in = ops.seqmap(seq_div=T, outer_stride=N * T, inner_stride=1, pos_stride=T, n_prefix=1, prefix_base=R, prefix_stride=1)
    m_out = ops.seqmap(seq_div=T, outer_stride=N * T, inner_stride=1, pos_stride=T, n_prefix=1, prefix_base=RB, prefix_stride=1, prefix_per_seq=1)
    q, k, v = (TView(qkv, i * D, hd, m_in) for i in range(3))
    for _ in range(3):
        ops.attn_fwd(q, k, v, TView(att, 0, hd, m_out), n_seq=B * T, n_heads=heads, head_dim=hd, s_q=N + 1, s_kv=N + 1, causal=False, scale=hd ** -0.5)
else:
    (attn_probe.spatial if which == "spatial" else attn_probe.gpt)(reps=1)
torch.cuda.synchronize()
buf = (ctypes.c_ulonglong * 256)()
rc = L.lib.ymp_attn_dbg_read(buf)
assert rc == 0, rc
ev = [(v >> 48, v & 0xFFFFFFFFFFFF) for v in buf if v]
t0 = ev[0][1]
prev = t0
for k, t in ev:
    print(f"{k:3d} t={t - t0:7d} (+{t - prev})")
    prev = t
