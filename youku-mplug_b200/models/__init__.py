"""Drop-in mirror of the reference's `models` package for the mPLUG-Video (GPT-3) hot path.

Put `youku-mplug_b200/` on PYTHONPATH ahead of the reference checkout and the reference's
run_pretrain_distributed_gpt3.py / downstream/run_*_gpt3.py import these classes instead
(same module paths, class names, constructor arguments, forward signatures and state_dict keys).
"""
