"""A training script written against the SAME third-party surface as the reference's
run_pretrain_distributed_gpt3.py (which cannot travel to the GPU box): `ruamel.yaml`, `megatron_util.mpu`
aliases (:36-40), `deepspeed.add_config_arguments` / `deepspeed.initialize(args=, model=, model_parameters=,
dist_init_required=, mpu=)` (:385-396, 263-267), the ds_config.json of utils.py:483-562, the per-step body of
train_one_epoch (:84-152) and utils.save_model / auto_load_model's engine calls (:451-480).  Run through
youku-mplug_b200/launch.py --ymp-standalone by tests/test_launcher_gpu.py; prints one JSON line."""
import argparse
import json
import os

import ruamel.yaml as yaml
import torch
import torch.distributed as dist

from models.distributed_gpt3 import DistributedGPT3_Pretrain
from models.modeling_distributed_gpt3 import DistributedGPT3Tokenizer

from megatron_util import mpu
mpu.get_model_parallel_group = mpu.get_tensor_model_parallel_group
mpu.get_model_parallel_world_size = mpu.get_tensor_model_parallel_world_size
mpu.get_model_parallel_rank = mpu.get_tensor_model_parallel_rank
mpu.get_model_parallel_src_rank = mpu.get_tensor_model_parallel_src_rank


def get_loss_scale_for_deepspeed(model):
    optimizer = model.optimizer
    loss_scale = None
    if hasattr(optimizer, 'loss_scale'):
        loss_scale = optimizer.loss_scale
    elif hasattr(optimizer, 'cur_scale'):
        loss_scale = optimizer.cur_scale
    return loss_scale, optimizer._global_grad_norm


def create_ds_config(args, opt, world):
    args.deepspeed_config = os.path.join(args.output_dir, "ds_config.json")
    ds_config = {
        "train_batch_size": args.batch_size * args.update_freq * world,
        "train_micro_batch_size_per_gpu": args.batch_size,
        "steps_per_print": 1000,
        "optimizer": {"type": "Adam", "adam_w_mode": True,
                      "params": {"lr": opt["lr"], "betas": list(opt["opt_betas"]), "eps": opt["opt_eps"],
                                 "weight_decay": opt["weight_decay"], "bias_correction": True}},
        "fp16": {"enabled": not args.bf16, "loss_scale": 0, "initial_scale_power": 16, "loss_scale_window": 500,
                 "hysteresis": 2, "min_loss_scale": 1},
        "bf16": {"enabled": args.bf16},
        "gradient_clipping": opt["clip_grad"],
        "zero_optimization": {"stage": 1, "reduce_bucket_size": 5e8},
    }
    with open(args.deepspeed_config, "w") as f:
        f.write(json.dumps(ds_config, indent=2))


if __name__ == '__main__':
    parser = argparse.ArgumentParser()
    parser.add_argument('--config')
    parser.add_argument('--output_dir')
    parser.add_argument('--update_freq', default=1, type=int)
    parser.add_argument('--bf16', action='store_true')
    parser.add_argument('--enable_deepspeed', action='store_true', default=False)
    parser.add_argument('--iters', default=3, type=int)
    import deepspeed
    parser = deepspeed.add_config_arguments(parser)
    ds_init = deepspeed.initialize
    args = parser.parse_args()
    config = yaml.load(open(args.config, 'r'), Loader=yaml.Loader)
    os.makedirs(args.output_dir, exist_ok=True)
    args.batch_size, args.max_length = config["batch_size"], config["max_length"]
    yaml.dump(config, open(os.path.join(args.output_dir, 'config.yaml'), 'w'))

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group(backend="nccl", init_method="env://", world_size=world, rank=rank)
    create_ds_config(args, config["optimizer"], world)
    device = torch.device("cuda")

    tokenizer = DistributedGPT3Tokenizer(model_dir=config['text_decoder'])
    model = DistributedGPT3_Pretrain(config=config, tokenizer=tokenizer)
    from ymp.train import default_param_groups          # optim/optim_factory.get_parameter_groups' rule
    vcfg = json.load(open(config["visual_cfg"]))
    optimizer_params = default_param_groups(model, config["optimizer"]["weight_decay"], model.no_weight_decay(),
                                            visual_backbone_scale=vcfg.get("clip_model", False))
    model, optimizer, _, _ = ds_init(args=args, model=model, model_parameters=optimizer_params, dist_init_required=False, mpu=mpu)
    p0 = {k: v.detach().float().clone() for k, v in model.module.named_parameters() if v.requires_grad}
    master0 = model.master.clone()

    T, R = vcfg["num_frames"], vcfg["img_size"]
    g = torch.Generator().manual_seed(1 + rank)
    texts = ["hello world video", "a cat runs on the grass", "the dog", "video a b c d e f g hello world"]
    model.train()
    model.zero_grad()
    model.micro_steps = 0
    log = []
    for it in range(args.iters):
        for i, param_group in enumerate(optimizer.param_groups):
            param_group["lr"] = 1e-3 * (it + 1) / args.iters * param_group["lr_scale"]
            if param_group["weight_decay"] > 0:
                param_group["weight_decay"] = 0.05
        video = torch.randn(args.batch_size, 3, T, R, R, generator=g).to(device, non_blocking=True)
        text = [texts[(it + j) % len(texts)] for j in range(args.batch_size)]
        text_input = tokenizer(text, padding='max_length', truncation=True, max_length=args.max_length, return_tensors="pt",
                               add_special_tokens=True).to(device)
        video = video.bfloat16()
        loss_caption, loss_ita = model(video, text_input)
        loss = loss_caption + loss_ita
        loss_value = loss.item()
        loss_list = [torch.zeros_like(loss) for _ in range(dist.get_world_size())]
        dist.all_gather(loss_list, loss)
        loss /= args.update_freq
        model.backward(loss)
        model.step()
        loss_scale_value, grad_norm = get_loss_scale_for_deepspeed(model)
        torch.cuda.synchronize()
        log.append(dict(loss=loss_value, loss_caption=loss_caption.item(), loss_ita=loss_ita.item(),
                        grad_norm=float(grad_norm), loss_scale=loss_scale_value,
                        lr=max(gp["lr"] for gp in optimizer.param_groups)))
    model.save_checkpoint(save_dir=args.output_dir, tag="checkpoint-0", client_state={'epoch': 0})
    _, client_states = model.load_checkpoint(args.output_dir, tag='checkpoint-0')
    changed = sum(int(not torch.equal(v.detach().float(), p0[k])) for k, v in model.module.named_parameters() if v.requires_grad)
    master_changed = float((model.master != master0).float().mean())
    if rank == 0:
        print("MINI " + json.dumps(dict(log=log, changed=changed, trainable=len(p0), client=client_states, master_changed=master_changed,
                                        engine=type(model).__module__ + "." + type(model).__name__,
                                        model_file=os.path.abspath(__import__("models").__file__))), flush=True)
    dist.destroy_process_group()
