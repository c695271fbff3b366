"""Parameter containers: nested nn.Modules whose state_dict keys equal the reference's."""
import math

import torch
import torch.nn as nn


class Holder(nn.Module):
    """A bare module node (only holds parameters / children)."""


class EmbedHolder(Holder):
    """word_embeddings node: callable like nn.Embedding (the task models call it directly,
    reference models/distributed_gpt3.py:155)."""

    def forward(self, input_ids):
        return torch.nn.functional.embedding(input_ids, self.weight)


def add_param(root, dotted, tensor, requires_grad=True):
    parts = dotted.split(".")
    node = root
    for p in parts[:-1]:
        if p not in node._modules:
            node.add_module(p, Holder())
        node = node._modules[p]
    node.register_parameter(parts[-1], nn.Parameter(tensor, requires_grad=requires_grad))


def named_param_list(module, prefix=""):
    """(keys, params) of every parameter under `module`, keys prefixed with `prefix`."""
    keys, params = [], []
    for k, p in module.named_parameters():
        keys.append(prefix + k)
        params.append(p)
    return keys, params


def trunc_normal(shape, std):
    return nn.init.trunc_normal_(torch.empty(*shape), std=std)


def xavier_uniform(shape):
    return nn.init.xavier_uniform_(torch.empty(*shape))


def linear_default(out_f, in_f):
    """nn.Linear's default init (kaiming_uniform(a=sqrt(5)) weight, uniform bias)."""
    w = torch.empty(out_f, in_f)
    nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    bound = 1.0 / math.sqrt(in_f)
    return w, torch.empty(out_f).uniform_(-bound, bound)
