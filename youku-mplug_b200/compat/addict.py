class Dict(dict):
    """addict.Dict as the reference uses it (models/modeling_distributed_gpt3.py:23): attribute access, None for
    missing keys."""
    __getattr__ = dict.get
    __setattr__ = dict.__setitem__
