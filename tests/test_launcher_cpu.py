"""The drop-in boundary at script level (BASELINE config-1 "plumbing only"):
youku-mplug_b200/launch.py runs the reference's UNMODIFIED run_pretrain_distributed_gpt3.py with the B200
`models` package and the compat/ stand-ins.  Without a GPU the script must get through argument parsing, yaml,
ds_config, distributed init, dataset / loader / tokenizer / model construction and parameter grouping, and stop
exactly at deepspeed.initialize with the engine's "no CPU fallback" error.  (/root/reference only exists in the
dev container: the test is skipped elsewhere; the GPU half is tests/test_launcher_gpu.py.)"""
import os
import socket
import subprocess
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("YMP_REFERENCE", "/root/reference")
SCRIPT = os.path.join(REF, "run_pretrain_distributed_gpt3.py")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_compat_yaml_reads_yaml12_floats():
    sys.path.append(os.path.join(ROOT, "youku-mplug_b200", "compat"))
    try:
        import importlib
        y = importlib.import_module("ruamel.yaml")
        d = y.load("a: 1e-4\nb: 1e-6\nc: [0.9, 0.999]\nd: 3\ne: {x: 2e-5,}\nf: text\n", Loader=y.Loader)
        assert d == {"a": 1e-4, "b": 1e-6, "c": [0.9, 0.999], "d": 3, "e": {"x": 2e-5}, "f": "text"}
        assert isinstance(d["a"], float) and isinstance(d["d"], int)
    finally:
        sys.path.remove(os.path.join(ROOT, "youku-mplug_b200", "compat"))
        for k in [k for k in sys.modules if k == "ruamel" or k.startswith("ruamel.")]:
            del sys.modules[k]


@pytest.mark.skipif(not os.path.isfile(SCRIPT), reason="reference checkout not present")
def test_unmodified_pretrain_script_reaches_the_device_boundary():
    from plumbing import make_workspace
    td = tempfile.mkdtemp(prefix="ymp_plumb_")
    ws = make_workspace(td)
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()),
               YMP_ALLOW_RANDOM_INIT="1", PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests"), PYTHONWARNINGS="ignore")
    cmd = [sys.executable, os.path.join(ROOT, "youku-mplug_b200", "launch.py"), "--ymp-pre",
           os.path.join(ROOT, "tests", "plumbing_cpu_pre.py"), SCRIPT, "--config", ws["config"], "--output_dir", ws["output_dir"],
           "--enable_deepspeed", "--bf16", "--device", "cpu", "--no_auto_resume"]
    r = subprocess.run(cmd, cwd=REF, env=env, capture_output=True, text=True, timeout=600)
    out = r.stdout + r.stderr
    assert r.returncode != 0
    for marker in ("Creating dataset", "Creating model", "number of params (B):", "Param groups ="):
        assert marker in out, (marker, out[-3000:])
    assert "needs a CUDA device - the B200 path has no CPU fallback" in out, out[-3000:]
    # the model the script built is the B200 one, and the frozen decoder never reached the optimizer groups
    assert "text_decoder." not in out.split("Param groups =")[1].split("}")[0]
    assert os.path.isfile(os.path.join(ws["output_dir"], "ds_config.json"))       # utils.create_ds_config ran
    assert os.path.isfile(os.path.join(ws["output_dir"], "config.yaml"))          # yaml.dump through the stand-in


DOWNSTREAM = {"cls": "run_cls_distributed_gpt3.py", "caption": "run_caption_distributed_gpt3.py",
              "retrieval": "run_retrieval_distributed_gpt3.py", "retrieval_itm": "run_retrieval_distributed_gpt3_itm.py"}


@pytest.mark.skipif(not os.path.isfile(SCRIPT), reason="reference checkout not present")
@pytest.mark.parametrize("task", sorted(DOWNSTREAM))
def test_unmodified_downstream_scripts_reach_the_device_boundary(task):
    """BASELINE configs[0] ("plumbing only": downstream/run_cls_distributed_gpt3.py, 2 frames, batch 1-2, world size 1 on
    CPU) and the other three downstream scripts, UNMODIFIED, through launch.py: argument parsing, yaml, dataset / sampler /
    loader, tokenizer, the B200 model class of the task, parameter groups - then deepspeed.initialize must refuse to
    build the engine without a GPU."""
    from plumbing import make_downstream_workspace
    td = tempfile.mkdtemp(prefix=f"ymp_plumb_{task}_")
    # (the cls script gives its val / test loaders int(0.1 * batch_size) samples: batch_size must be >= 10 there)
    ws = make_downstream_workspace(td, task, num_videos=12, batch_size=10 if task == "cls" else 2)
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()),
               YMP_ALLOW_RANDOM_INIT="1", PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests"), PYTHONWARNINGS="ignore")
    cmd = [sys.executable, os.path.join(ROOT, "youku-mplug_b200", "launch.py"), "--ymp-pre",
           os.path.join(ROOT, "tests", "plumbing_cpu_pre.py"), os.path.join(REF, "downstream", DOWNSTREAM[task]),
           "--config", ws["config"], "--output_dir", ws["output_dir"], "--enable_deepspeed", "--bf16", "--device", "cpu",
           "--no_auto_resume"]
    r = subprocess.run(cmd, cwd=REF, env=env, capture_output=True, text=True, timeout=600)
    out = r.stdout + r.stderr
    assert r.returncode != 0
    for marker in ("Creating dataset", "Creating model", "number of params (B):"):
        assert marker in out, (marker, out[-3000:])
    assert "needs a CUDA device - the B200 path has no CPU fallback" in out, out[-3000:]
    cls_name = {"cls": "DistributedGPT3_Cls", "caption": "DistributedGPT3_Caption", "retrieval": "DistributedGPT3_Retrieval",
                "retrieval_itm": "DistributedGPT3_Retrieval_Cls"}[task]
    assert f"Model = {cls_name}(" in out, out[-3000:]
