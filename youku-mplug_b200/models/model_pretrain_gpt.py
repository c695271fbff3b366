"""Signature shim for the reference's dead `models/model_pretrain_gpt.py` (MPLUG_COCA).

The reference class cannot be imported (it needs VisionTransformerForMaskedImageModeling, which
exists nowhere in the repo - SURVEY.md D1) and no script uses it.  BASELINE.json's north_star
names it, so the forward signature (models/model_pretrain_gpt.py:96) is kept here and routed to
the live pre-training model.
"""
from .distributed_gpt3 import DistributedGPT3_Pretrain


class MPLUG_COCA(DistributedGPT3_Pretrain):
    def forward(self, image, text, bool_masked_pos=None, image_target=None):
        if bool_masked_pos is not None or image_target is not None:
            raise NotImplementedError("masked-visual-modelling inputs belong to the reference's dead code path")
        return super().forward(image, text)
