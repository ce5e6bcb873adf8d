"""Multi-GPU tests (-m multigpu): the fused collectives across REAL GPU boundaries (reference: tests/python/gpu/
test_kvstore_gpu.py, test_device.py:51-69, test_nccl.py:38-49).  Each test launches tests/multigpu_worker.py under torchrun
with one rank per GPU on 127.0.0.1 and checks the JSON it prints."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.multigpu]
HERE = os.path.dirname(os.path.abspath(__file__))
_PORT = [29610]


def _run(scenario, world, env=None, timeout=600):
    _PORT[0] += 1
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(_PORT[0]), os.path.join(HERE, "multigpu_worker.py"), scenario]
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(cmd, env=e, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT ")]
    assert r.returncode == 0 and lines, r.stdout[-4000:]
    return json.loads(lines[-1][7:])


def _world():
    n = torch.cuda.device_count()
    return 8 if n >= 8 else (4 if n >= 4 else 2)


@pytest.mark.parametrize("nvls", ["1", "0"])
def test_fused_optimizers_across_gpus_match_fp32_oracle(nvls):
    out = _run("optimizers", _world(), {"MXB200_NVLS": nvls})
    bad = {k: v for k, v in out.items() if isinstance(v, float) and v > 1.0}
    assert not bad, "per-element tolerance violated (ratio > 1): %s" % bad


def test_trainer_save_resume_is_exact_and_state_is_sharded():
    out = _run("checkpoint", _world())
    assert out["fused"]
    assert out["resume_max_abs_diff"] == 0.0, out
    if out["sharded"]:
        assert out["state_elems"] < out["full_state_elems"], out   # ZeRO-1: each rank holds only its slice


def test_cuda_graph_train_step_across_gpus():
    out = _run("graph", 2)
    assert out["graph"]["replays"] >= 5 and out["eager"]["weights_identical"] and out["graph"]["weights_identical"], out
    assert out["max_loss_diff"] < 0.25, out
