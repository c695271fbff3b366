"""CPU, world_size 2 over gloo: the host-side logic of the N>1 path (differentiable all-gather of the
contrastive features) and the optimizer parameter grouping."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "youku-mplug_b200"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from models.distributed_utils import all_gather, all_gather_cat, concat_all_gather
    torch.manual_seed(100 + rank)
    x = torch.randn(3, 4, requires_grad=True)
    full = all_gather_cat(x)                      # [world*3, 4]
    assert full.shape == (world * 3, 4)
    # every rank weights the gathered rows differently; backward must reduce-scatter (sum over ranks)
    w = torch.arange(world * 3 * 4, dtype=torch.float32).view(world * 3, 4) * (rank + 1)
    (full * w).sum().backward()
    expect = sum(torch.arange(world * 3 * 4, dtype=torch.float32).view(world * 3, 4)[rank * 3:(rank + 1) * 3] * (r + 1)
                 for r in range(world))
    ok_grad = torch.allclose(x.grad, expect)
    # forward content: rank r's slice equals rank r's tensor
    torch.manual_seed(100)
    r0 = torch.randn(3, 4)
    ok_fwd = torch.allclose(full[:3].detach(), r0)
    tup = all_gather(x.detach())
    ok_tuple = len(tup) == world and torch.equal(tup[rank], x.detach())
    idx = torch.tensor([rank * 10 + 1, rank * 10 + 2], dtype=torch.int64)
    ok_idx = torch.equal(concat_all_gather(idx), torch.tensor([1, 2, 11, 12]))
    q.put((rank, ok_grad, ok_fwd, ok_tuple, ok_idx))
    dist.destroy_process_group()


def test_all_gather_forward_backward_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert all(r[1:]), r


def test_param_groups_follow_reference_rule():
    from oracle import port
    from helpers import build_pretrain
    from ymp.train import default_param_groups
    m = build_pretrain(port.VCFG_TINY, port.GCFG_TINY, 8)
    groups = default_param_groups(m, 0.05, m.no_weight_decay())
    by_name = {n: g for g in groups for n in g["names"]}
    assert by_name["visual_fc.weight"]["weight_decay"] == 0.05
    assert by_name["visual_fc.bias"]["weight_decay"] == 0.0
    assert by_name["visual_encoder.blocks.0.norm1.weight"]["weight_decay"] == 0.0       # 1-D
    assert by_name["visual_encoder.pos_embed"]["weight_decay"] == 0.0                    # skip list
    assert by_name["learnable_queries"]["weight_decay"] == 0.05
    assert not any(n.startswith("text_decoder.") for n in by_name)                       # frozen decoder
    assert all(g["lr_scale"] == 1.0 for g in groups)
    scaled = default_param_groups(m, 0.05, m.no_weight_decay(), visual_backbone_scale=True)
    by_name = {n: g for g in scaled for n in g["names"]}
    assert by_name["visual_encoder.blocks.0.attn.qkv.weight"]["lr_scale"] == 0.1
    assert by_name["visual_encoder.blocks.0.temporal_fc.weight"]["lr_scale"] == 1.0
