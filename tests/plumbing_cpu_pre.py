"""--ymp-pre hook for the CPU plumbing test ONLY: the reference script hard-codes CUDA / NCCL in its start-up
(utils.py:285-286 `torch.cuda.set_device`, backend 'nccl').  On the GPU-less dev container those two calls are
redirected so that the script reaches the first device computation, which must then fail loudly
(no CPU fallback).  Nothing here touches the product."""
import torch
import torch.distributed as dist

torch.cuda.set_device = lambda *a, **k: None
_init = dist.init_process_group


def _init_gloo(backend=None, **kw):
    return _init(backend="gloo", **kw)


dist.init_process_group = _init_gloo
torch.distributed.init_process_group = _init_gloo
