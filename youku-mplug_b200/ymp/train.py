"""Data-parallel training engine for the drop-in models: the part of `deepspeed.initialize(...)`
the reference scripts rely on (run_pretrain_distributed_gpt3.py:257-267,134-138; utils.py:483-562),
rebuilt for one-process-per-GPU B200 training.

  engine = TrainEngine(model, optimizer_params, lr=..., betas=..., eps=..., clip_grad=...)
  loss, _ = engine(video, text); engine.backward(loss); engine.step()

* trainable parameters live in ONE flat bf16 buffer (module parameters are views of it) with fp32
  master weights and Adam moments beside it (ZeRO-free: 130 M trainable params = 1.6 GB of state);
* weight gradients are accumulated by the kernels directly into one flat fp32 buffer (the GEMM
  split-K epilogue adds into it), so there is no per-parameter .grad tensor and no copy;
* step(): one NCCL all-reduce of the flat gradient over NVLink/NVSwitch, device-side global-norm
  clipping, fused AdamW - no host synchronisation anywhere in backward/step.
"""
import contextlib

import torch
import torch.distributed as dist

from . import functional as YF
from . import ops


def default_param_groups(model, weight_decay, skip_list=(), visual_backbone_scale=False):
    """Same grouping rule as the reference's optim/optim_factory.py:219-265 (no decay for 1-D / bias /
    skip_list names; optional 0.1 lr scale for non-temporal visual_encoder weights)."""
    groups = {}
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        no_decay = p.dim() == 1 or name.endswith(".bias") or name in skip_list
        scaled = visual_backbone_scale and "visual_encoder." in name and "temporal" not in name
        key = (no_decay, scaled)
        g = groups.setdefault(key, dict(params=[], names=[], weight_decay=0.0 if no_decay else weight_decay,
                                        lr_scale=0.1 if scaled else 1.0))
        g["params"].append(p)
        g["names"].append(name)
    return list(groups.values())


class _Optimizer:
    """What the reference loop touches on `model.optimizer` (run_pretrain...py:46-53,88-96)."""

    def __init__(self, groups):
        self.param_groups = groups
        self.cur_scale = 1.0
        self._global_grad_norm = None


class TrainEngine:
    def __init__(self, model, optimizer_params=None, lr=1e-4, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.05,
                 clip_grad=3.0, process_group=None):
        self.module = model
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        if optimizer_params is None:
            optimizer_params = default_param_groups(model, weight_decay, getattr(model, "no_weight_decay", lambda: ())())
        self.clip_grad = clip_grad
        params = [p for g in optimizer_params for p in g["params"]]
        assert params, "no trainable parameters"
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("TrainEngine needs the model on a CUDA device (no CPU fallback)")
        total = sum((p.numel() + 7) // 8 * 8 for p in params)
        self.flat_param = torch.zeros(total, device=dev, dtype=torch.bfloat16)
        self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self._sink = {}
        off = 0
        groups = []
        for g in optimizer_params:
            start = off
            for p in g["params"]:
                n = p.numel()
                view = self.flat_param[off:off + n].view(p.shape)
                view.copy_(p.data)
                p.data = view
                self._sink[id(p)] = self.flat_grad[off:off + n]
                off += (n + 7) // 8 * 8
            groups.append(dict(params=g["params"], weight_decay=g.get("weight_decay", weight_decay),
                               lr_scale=g.get("lr_scale", 1.0), lr=lr * g.get("lr_scale", 1.0), betas=list(betas),
                               eps=eps, _range=(start, off)))
        self.master = self.flat_param.float()
        self.exp_avg = torch.zeros_like(self.master)
        self.exp_avg_sq = torch.zeros_like(self.master)
        self._sumsq = torch.zeros(1, device=dev, dtype=torch.float32)
        self.optimizer = _Optimizer(groups)
        self.micro_steps = 0
        self.global_steps = 0
        self._graphs = {}

    # ---- nn.Module-like surface -----------------------------------------------------------
    def __call__(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def train(self, mode=True):
        self.module.train(mode)
        return self

    def eval(self):
        self.module.eval()
        return self

    def parameters(self):
        return self.module.parameters()

    def named_parameters(self):
        return self.module.named_parameters()

    def state_dict(self):
        return self.module.state_dict()

    # ---- training step ----------------------------------------------------------------------
    def backward(self, loss):
        with YF.grad_sink(self._sink):
            loss.backward()
        self.micro_steps += 1

    def zero_grad(self):
        self.flat_grad.zero_()

    def step(self):
        if self.world > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        self.global_steps += 1
        self._sumsq.zero_()
        if self.clip_grad and self.clip_grad > 0:
            ops.sumsq(self.flat_grad, self._sumsq)
        scale = 1.0 / self.world
        for g in self.optimizer.param_groups:
            a, b = g["_range"]
            if b == a:
                continue
            ops.adamw(self.master[a:b], self.flat_param[a:b], self.flat_grad[a:b], self.exp_avg[a:b],
                      self.exp_avg_sq[a:b], step=self.global_steps, lr=g["lr"], beta1=g["betas"][0], beta2=g["betas"][1],
                      eps=g["eps"], weight_decay=g["weight_decay"], grad_scale=scale,
                      max_grad_norm=self.clip_grad or 0.0, sumsq_t=self._sumsq if self.clip_grad else None)
        self.optimizer._global_grad_norm = _LazyNorm(self._sumsq.clone(), scale)
        self.flat_grad.zero_()

    def train_step(self, video, text, use_graph=True, graph_warmup=2):
        """One full iteration (forward + backward + step) and the loss tensor.

        Shapes are static in pre-training (`padding='max_length'`, run_pretrain_distributed_gpt3.py:100),
        so after `graph_warmup` eager iterations the ~900 kernel launches of forward+backward are
        captured ONCE into a CUDA graph per input signature and replayed; the all-reduce and the
        optimizer stay outside the graph.  Inputs are copied into the graph's static buffers (the copy
        also casts fp32 frames to bf16), so callers may pass fresh tensors every step."""
        key = (tuple(video.shape), tuple(text.input_ids.shape))
        st = self._graphs.setdefault(key, dict(calls=0))
        if not use_graph or st["calls"] < graph_warmup:
            loss, _ = self.module(video, text)
            self.backward(loss)
            self.step()
            st["calls"] += 1
            return loss.detach()
        if "graph" not in st:
            from models.modeling_distributed_gpt3 import BatchEncoding
            st["video"] = torch.empty(video.shape, device=video.device, dtype=torch.bfloat16)
            st["ids"] = torch.empty_like(text.input_ids)
            st["att"] = torch.empty_like(text.attention_mask)
            st["video"].copy_(video)
            st["ids"].copy_(text.input_ids)
            st["att"].copy_(text.attention_mask)
            static_text = BatchEncoding(dict(input_ids=st["ids"], attention_mask=st["att"]))
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss, _ = self.module(st["video"], static_text)
                with YF.grad_sink(self._sink):
                    loss.backward()
            st["graph"], st["loss"] = g, loss.detach()
        else:
            st["video"].copy_(video, non_blocking=True)
            st["ids"].copy_(text.input_ids, non_blocking=True)
            st["att"].copy_(text.attention_mask, non_blocking=True)
        st["graph"].replay()
        self.micro_steps += 1
        self.step()
        return st["loss"]

    # ---- checkpointing (utils.py:476-480,441-455) ----------------------------------------------
    def save_checkpoint(self, save_dir, tag=None, client_state=None):
        import os
        tag = tag or f"global_step{self.global_steps}"
        os.makedirs(os.path.join(save_dir, str(tag)), exist_ok=True)
        if not dist.is_initialized() or dist.get_rank() == 0:
            torch.save(dict(module=self.module.state_dict(), master=self.master, exp_avg=self.exp_avg,
                            exp_avg_sq=self.exp_avg_sq, global_steps=self.global_steps,
                            client_state=client_state or {}),
                       os.path.join(save_dir, str(tag), "mp_rank_00_model_states.pt"))
            with open(os.path.join(save_dir, "latest"), "w") as f:
                f.write(str(tag))

    def load_checkpoint(self, load_dir, tag=None):
        import os
        if tag is None:
            with open(os.path.join(load_dir, "latest")) as f:
                tag = f.read().strip()
        ck = torch.load(os.path.join(load_dir, str(tag), "mp_rank_00_model_states.pt"), map_location="cpu", weights_only=False)
        self.module.load_state_dict(ck["module"])
        self.master.copy_(ck["master"])
        self.exp_avg.copy_(ck["exp_avg"])
        self.exp_avg_sq.copy_(ck["exp_avg_sq"])
        self.global_steps = ck["global_steps"]
        return load_dir, ck.get("client_state", {})


class _LazyNorm:
    """Global grad norm that only synchronises when somebody actually reads it."""

    def __init__(self, sumsq, scale):
        self._s, self._scale = sumsq, scale

    def __float__(self):
        return float(self._s.sqrt().item() * self._scale)

    def item(self):
        return float(self)

    def __repr__(self):
        return f"{float(self):.4f}"
