"""The example scripts run end to end in tiny configurations (the reference's nightly / tutorial tests do the same for its
example/ tree).  CPU only."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(rel, *args, timeout=600):
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.join(ROOT, rel)] + list(args), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=timeout, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:]
    return r.stdout


def test_model_parallel_matrix_factorization_learns():
    out = _run("examples/model-parallel/matrix_factorization.py", "--epochs", "6", "--samples", "20000")
    line = [l for l in out.splitlines() if l.startswith("validation rmse")][-1]
    rmse, std = float(line.split()[2]), float(line.split("label std")[1].split(")")[0])
    assert rmse < 0.8 * std, line
    assert "cpu(0)" in line and "cpu(1)" in line            # the two halves were placed on different contexts


def test_sparse_linear_classification_row_sparse_pull():
    out = _run("examples/sparse/linear_classification.py", "--epochs", "2", "--samples", "8000")
    last = [l for l in out.splitlines() if "accuracy" in l][-1]
    assert float(last.split("accuracy")[1]) > 0.9, out[-500:]


def test_profiler_examples_write_traces(tmp_path):
    f1, f2 = str(tmp_path / "a.json"), str(tmp_path / "b.json")
    out = _run("examples/profiler/profiler_matmul.py", "--iters", "4", "--size", "64", "--file", f1)
    assert "dot_64x64" in out and "engine" in out and os.path.getsize(f1) > 100
    out = _run("examples/profiler/profiler_executor.py", "--iters", "2", "--file", f2)
    assert "forward" in out and "backward" in out and os.path.getsize(f2) > 100


def test_gluon_mnist_and_dcgan():
    out = _run("examples/gluon/mnist.py", "--epochs", "2", "--samples", "1500")
    assert float(out.strip().splitlines()[-1].split()[-1]) > 0.9, out[-500:]
    out = _run("examples/gluon/dcgan.py", "--iters", "6", "--batch-size", "8")
    assert "finite True" in out, out[-500:]


def test_distributed_training_example_single_process():
    out = _run("examples/distributed_training/cifar10_dist.py", "--kv-store", "local", "--gpus", "0", "--epochs", "1", "--samples", "64",
               "--batch-size", "16")
    assert "worker 0/1 epoch 0" in out, out[-500:]
