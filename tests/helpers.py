import json
import os
import tempfile

import torch


def make_model_dir(vcfg, gcfg, dropout=(0.0, 0.0)):
    """dropout = (hidden_dropout, attention_dropout) of the decoder; 0 = parity mode (the reference default is 0.1)."""
    td = tempfile.mkdtemp(prefix="ymp_model_")
    with open(os.path.join(td, "config.json"), "w") as f:
        json.dump(dict(gcfg, hidden_dropout=dropout[0], attention_dropout=dropout[1]), f)
    with open(os.path.join(td, "vis.json"), "w") as f:
        json.dump(dict(vcfg, pretrained_ckpt=None, grad_ckpt=False), f)
    return td


def pretrain_config(td, Q, **extra):
    cfg = dict(visual_cfg=os.path.join(td, "vis.json"), text_cfg=os.path.join(td, "config.json"), text_decoder=td,
               megatron_cfg={"world_size": 1, "model_parallel_size": 1, "tensor_model_parallel_size": 1},
               num_learnable_token=Q, use_contrastive=False, freeze_text_decoder=True)
    cfg.update(extra)
    return cfg


def build_pretrain(vcfg, gcfg, Q, sd=None, device="cpu", dtype=None, cls_name="DistributedGPT3_Pretrain", dropout=(0.0, 0.0),
                   **extra):
    os.environ["YMP_ALLOW_RANDOM_INIT"] = "1"
    import models.distributed_gpt3 as D
    td = make_model_dir(vcfg, gcfg, dropout)
    model = getattr(D, cls_name)(config=pretrain_config(td, Q, **extra), tokenizer=None)
    if sd is not None:
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected, unexpected
        assert not [m for m in missing if not m.startswith(("vision_proj", "text_proj", "temp", "cls_head"))], missing
    model = model.to(device)
    if dtype is not None:
        model = model.to(dtype)
    return model


def build_pretrain_image(ecfg, gcfg, Q, sd=None, device="cpu", dtype=None):
    """DistributedGPT3_Pretrain_Image with an EVA encoder of the given (tiny) dims: create_eva_vit_g hard-codes the
    EVA-g sizes (reference models/eva_vit.py:413-424), so for tests the factory the constructor calls is pointed at
    a small instance of the same class - exactly what oracle/make_golden.py run_image does to the reference."""
    os.environ["YMP_ALLOW_RANDOM_INIT"] = "1"
    import models.distributed_gpt3 as D
    import models.eva_vit as E
    td = make_model_dir(dict(img_size=ecfg["img_size"], embed_dim=ecfg["embed_dim"], num_heads=ecfg["num_heads"],
                             mlp_ratio=ecfg["mlp_ratio"], drop_path=0), gcfg)
    orig = E.create_eva_vit_g
    E.create_eva_vit_g = lambda img_size, norm_layer, drop_path_rate, use_checkpoint: E.VisionTransformer(
        img_size=img_size, patch_size=ecfg["patch_size"], use_mean_pooling=False, embed_dim=ecfg["embed_dim"], depth=ecfg["depth"],
        num_heads=ecfg["num_heads"], mlp_ratio=ecfg["mlp_ratio"], qkv_bias=True, drop_path_rate=0., norm_layer=norm_layer)
    try:
        model = D.DistributedGPT3_Pretrain_Image(config=pretrain_config(td, Q, use_eva_g=True), tokenizer=None)
    finally:
        E.create_eva_vit_g = orig
    if sd is not None:
        missing, unexpected = model.load_state_dict(sd, strict=False)
        assert not unexpected and not missing, (missing, unexpected)
    model = model.to(device)
    if dtype is not None:
        model = model.to(dtype)
    return model
