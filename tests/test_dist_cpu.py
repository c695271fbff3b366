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


def _bucket_worker(rank, world, port, q):
    """The engine's reduction protocol on a CPU flat buffer: async all-reduce of each block's buckets as the (simulated)
    backward reports it final, in reverse block order, then the gaps - against one plain all-reduce of the buffer."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "youku-mplug_b200"))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ymp.train import gap_ranges, plan_buckets
    spans, pos = [], 0
    for blk in range(3):
        for nm, n in (("attn.qkv.weight", 5000), ("attn.proj.weight", 1700), ("norm1.weight", 8), ("mlp.fc1.weight", 4096)):
            spans.append((f"visual_encoder.blocks.{blk}.{nm}", pos, pos + n)); pos += n
        if blk == 1:   # a foreign parameter between two blocks splits the run
            spans.append(("visual_fc.weight", pos, pos + 3000)); pos += 3000
    spans.append(("learnable_queries", pos, pos + 100)); pos += 100
    buckets = plan_buckets(spans, min_elems=4000)
    torch.manual_seed(7 + rank)
    flat = torch.randn(pos)
    ref = flat.clone()
    dist.all_reduce(ref)
    pending, reduced = [], []
    for blk in (2, 1, 0):                         # the backward finishes the last block first
        for a, b in buckets.get(f"visual_encoder.blocks.{blk}.", ()):
            pending.append(dist.all_reduce(flat[a:b], async_op=True))
            reduced.append((a, b))
    for w in pending:
        w.wait()
    gaps = gap_ranges(reduced, flat.numel())
    for a, b in gaps:
        dist.all_reduce(flat[a:b])
    covered = sorted(reduced + gaps)
    tiles = covered[0][0] == 0 and covered[-1][1] == flat.numel() and all(x[1] == y[0] for x, y in zip(covered, covered[1:]))
    q.put((rank, torch.equal(flat, ref), tiles, {k: v for k, v in buckets.items()}))
    dist.destroy_process_group()


def test_bucketed_allreduce_protocol_equals_plain_allreduce_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, same, tiles, buckets in res:
        assert same and tiles, (rank, same, tiles)
        # one bucket per block: the whole block is one adjacent run of 10804 elements
        assert buckets == {"visual_encoder.blocks.0.": [(0, 10804)], "visual_encoder.blocks.1.": [(10804, 21608)],
                           "visual_encoder.blocks.2.": [(24608, 35412)]}, buckets


def test_plan_buckets_on_the_model_layout():
    """On the real parameter layout every bucket lies inside one block's parameters, buckets are disjoint, and a block
    whose ranges are all below the threshold simply has no bucket (its gradients go with the gaps)."""
    from oracle import port
    from helpers import build_pretrain
    from ymp.train import default_param_groups, gap_ranges, plan_buckets
    m = build_pretrain(port.VCFG_TINY, port.GCFG_TINY, 8)
    spans, pos = [], 0
    for g in default_param_groups(m, 0.05, m.no_weight_decay()):
        for n, p in zip(g["names"], g["params"]):
            spans.append((n, pos, pos + p.numel())); pos += p.numel()
    buckets = plan_buckets(spans, min_elems=1)
    where = {}
    for n, a, b in spans:
        for i in range(a, b, max(1, (b - a) // 3)):
            where[i] = n
    flat = sorted(r for v in buckets.values() for r in v)
    assert all(x[1] <= y[0] for x, y in zip(flat, flat[1:]))
    for prefix, ranges in buckets.items():
        for a, b in ranges:
            assert all(n.startswith(prefix) for n, sa, sb in spans if sa < b and sb > a), prefix
    assert set(buckets) == {f"visual_encoder.blocks.{i}." for i in range(port.VCFG_TINY["depth"])}
    gaps = gap_ranges(flat, pos)
    assert sum(b - a for a, b in flat) + sum(b - a for a, b in gaps) == pos
    assert plan_buckets(spans, min_elems=1 << 40) == {}
