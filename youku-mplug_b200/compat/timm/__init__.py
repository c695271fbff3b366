"""Stand-in for timm==0.6.7: the reference imports only init / regularisation helpers
(models/vision_transformer.py:19-21, utils.py:13, optim/optim_factory.py)."""
__version__ = "0.6.7+ymp_stub"


def create_model(*args, **kwargs):
    raise NotImplementedError("timm.create_model: hub checkpoints need network access; use pretrained_ckpt: clip/...")
