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
