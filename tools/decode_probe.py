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
from oracle import port  # noqa: E402
import models.modeling_distributed_gpt3 as M  # noqa: E402

dev = torch.device("cuda")
td = make_model_dir(port.VCFG_CLIP_B16, dict(port.GCFG_1_3B, tokens_to_generate=32))
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


run(4)
pre, step = run(32)
wbytes = sum(p.numel() for p in dec.parameters()) * 2
print("DECODE " + json.dumps(dict(beam=beam, prefill_rows=beam * (Q + P), prefill_ms=round(pre, 3), step_ms=round(step, 3),
                                  weight_gb=round(wbytes / 1e9, 2), hbm_floor_ms=round(wbytes / 6.4e12 * 1e3, 3))))
