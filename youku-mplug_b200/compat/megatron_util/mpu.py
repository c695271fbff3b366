import torch.distributed as dist


def _world_group():
    return dist.group.WORLD if dist.is_initialized() else None


def get_tensor_model_parallel_group():
    return None


def get_tensor_model_parallel_world_size():
    return 1


def get_tensor_model_parallel_rank():
    return 0


def get_tensor_model_parallel_src_rank():
    return dist.get_rank() if dist.is_initialized() else 0


def get_data_parallel_group():
    return _world_group()


def get_data_parallel_world_size():
    return dist.get_world_size() if dist.is_initialized() else 1


def get_data_parallel_rank():
    return dist.get_rank() if dist.is_initialized() else 0


def get_pipeline_model_parallel_world_size():
    return 1


def get_pipeline_model_parallel_rank():
    return 0
