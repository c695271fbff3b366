"""`import ruamel.yaml as yaml` for the scripts (`yaml.load(f, Loader=yaml.Loader)`, `yaml.dump(config, f)`,
run_pretrain_distributed_gpt3.py:400,425) on top of PyYAML.  PyYAML implements YAML 1.1, where `1e-4`
(no dot) is a *string*; ruamel implements 1.2 and reads it as a float - the configs rely on that
(`lr: 1e-4`, `opt_eps: 1e-6`, `min_lr: 1e-6`), so the loaders below carry the 1.2 float resolver."""
import re

import yaml as _yaml
from yaml import *  # noqa: F401,F403

_FLOAT_12 = re.compile(r"""^(?:[-+]?(?:[0-9][0-9_]*)\.[0-9_]*(?:[eE][-+]?[0-9]+)?
                          |[-+]?(?:[0-9][0-9_]*)(?:[eE][-+]?[0-9]+)
                          |\.[0-9_]+(?:[eE][-+]?[0-9]+)?
                          |[-+]?\.(?:inf|Inf|INF)
                          |\.(?:nan|NaN|NAN))$""", re.X)


def _with_float12(base):
    cls = type(base.__name__, (base,), {})
    cls.yaml_implicit_resolvers = {k: list(v) for k, v in base.yaml_implicit_resolvers.items()}
    cls.add_implicit_resolver("tag:yaml.org,2002:float", _FLOAT_12, list("-+0123456789."))
    return cls


Loader = _with_float12(_yaml.Loader)
SafeLoader = _with_float12(_yaml.SafeLoader)
FullLoader = _with_float12(_yaml.FullLoader)
RoundTripLoader = Loader


def load(stream, Loader=Loader):
    return _yaml.load(stream, Loader=Loader)


def safe_load(stream):
    return _yaml.load(stream, Loader=SafeLoader)
