"""GPU probe: KV-cache decoding latency at the 1.3B configuration (random weights): prefill of the
[128-query prefix | prompt] block and single-token beam steps (beam 5), CUDA-event timed."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "youku-mplug_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
os.environ["YMP_ALLOW_RANDOM_INIT"] = "1"
import torch  # noqa: E402
from helpers import make_model_dir  # noqa: E402
from bench import GCFG, VCFG_CLIP_B16  # noqa: E402  (model-shape constants; no oracle code on a probe's path)
import models.modeling_distributed_gpt3 as M  # noqa: E402

dev = torch.device("cuda")
td = make_model_dir(VCFG_CLIP_B16, dict(GCFG["1.3B"], tokens_to_generate=32))
with torch.device(dev):
    dec = M.DistributedGPT3(td, 0, megatron_cfg={}).to(torch.bfloat16).eval()
beam, Q, P, H = 5, 128, 16, 2048
qf = (torch.randn(beam, Q, H, device=dev) * 0.1).to(torch.bfloat16)
prompt = torch.randint(0, 51200, (beam, P), device=dev)


def run(n_steps):
    dec.inference_params = M.InferenceParams(beam, Q + P + n_steps + 1)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    with torch.no_grad():
        e[0].record()
        out = dec(tokens=prompt, query_embeds=qf)
        e[1].record()
        tok = out.logits[:, -1].argmax(-1, keepdim=True)
        for _ in range(n_steps):
            out = dec(tokens=tok)
            tok = out.logits[:, -1].argmax(-1, keepdim=True)
        e[2].record()
    torch.cuda.synchronize()
    return e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2]) / n_steps


wbytes = sum(p.numel() for p in dec.parameters()) * 2
MODES = os.environ.get("DECODE_PROBE_MODES", "0,1").split(",")
STEPS = int(os.environ.get("DECODE_PROBE_STEPS", "32"))
for mode in MODES:
    os.environ["YMP_DECODE_GRAPH"] = mode
    dec.__dict__.pop("_decode_pool", None)
    run(4)
    pre, step = run(STEPS)
    replay_ms = None
    ts = dec.inference_params.cache.token
    if mode == "1" and ts is not None and ts.graph is not None:
        # the captured step alone (device time, no host work between replays; the cache keeps growing - stop before max_len)
        n_rep = 8
        dec.inference_params = M.InferenceParams(beam, Q + P + STEPS + 1)
        with torch.no_grad():
            out = dec(tokens=prompt, query_embeds=qf)
            dec(tokens=out.logits[:, -1].argmax(-1, keepdim=True))
            dec(tokens=out.logits[:, -1].argmax(-1, keepdim=True))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_rep):
            ts.graph.replay()
        e1.record()
        torch.cuda.synchronize()
        replay_ms = round(e0.elapsed_time(e1) / n_rep, 3)
    print("DECODE " + json.dumps(dict(graph=int(mode), graph_replay_only_ms=replay_ms, beam=beam, prefill_rows=beam * (Q + P), prefill_ms=round(pre, 3),
                                      step_ms=round(step, 3), weight_gb=round(wbytes / 1e9, 2),
                                      hbm_floor_ms=round(wbytes / 6.4e12 * 1e3, 3))))

if os.environ.get("DECODE_PROBE_SKINNY", "1") == "0":
    sys.exit(0)
# the skinny GEMM alone at the decode shapes, cycling through enough weight copies to defeat the 126 MB L2
from ymp import ops  # noqa: E402
for name, N, K, kw in (("qkv", 6144, 2048, {}), ("dense", 2048, 2048, dict(res=True)), ("fc1", 8192, 2048, dict(act=2)),
                       ("fc2", 2048, 8192, dict(res=True)), ("lm_head", 51200, 2048, dict(f32=True))):
    copies = max(2, int(400e6 // (N * K * 2)))
    ws = [(torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16) for _ in range(copies)]
    x = torch.randn(beam, K, device=dev).to(torch.bfloat16)
    b = torch.randn(N, device=dev).to(torch.bfloat16)
    r = torch.randn(beam, N, device=dev) if kw.get("res") else None
    od = torch.float32 if (kw.get("res") or kw.get("f32")) else torch.bfloat16
    for skinny in (True, False):
        def call(w):
            if skinny:
                return ops.gemm_skinny(x, w, bias=b, residual=r, act=kw.get("act", 0), out_dtype=od)
            return ops.gemm(x, w, bias=b, residual=r, act=kw.get("act", 0), out_dtype=od)
        for w in ws:
            call(w)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record()
        for _ in range(reps):
            for w in ws:
                call(w)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / (reps * copies) * 1e3
        print("SKINNY " + json.dumps(dict(shape=name, N=N, K=K, rows=beam, kernel="skinny" if skinny else "tcgen05",
                                          us=round(us, 2), gbs=round(N * K * 2 / us / 1e3, 1))))
    del ws
