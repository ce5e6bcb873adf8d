"""Tools that ship with the framework (reference: tools/bandwidth/test_measure.py, tools/launch.py)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bandwidth_measure_cpu_two_devices():
    """Same check as the reference's tools/bandwidth/test_measure.py: results are produced and exact."""
    sys.path.insert(0, os.path.join(ROOT, "tools", "bandwidth"))
    try:
        import measure
    finally:
        sys.path.pop(0)
    args = measure.parse_args(["--cpu", "--gpus", "0,1", "--network", "resnet", "--num-layers", "18", "--num-batches", "2"])
    res = measure.run(args)
    assert len(res) == 2
    assert all(r.error == 0.0 and r.bandwidth > 0 for r in res)


def test_launch_help_runs():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch.py"), "--help"], stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True, timeout=120)
    assert out.returncode == 0 and "--elastic-training-enabled" in out.stdout


def _run(cmd, cwd, timeout=300):
    return subprocess.run([sys.executable] + cmd, cwd=cwd, timeout=timeout, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)


def test_example_dynamic_training_symbolic_cpu(tmp_path):
    """The fork's headline example (train_resnet.py) end to end on CPU with synthetic data: Module.fit, elastic data
    iterator hook, Speedometer, checkpoint."""
    r = _run([os.path.join(ROOT, "examples", "dynamic-training", "train_resnet.py"), "--model", "resnet18_v1", "--benchmark", "1",
              "--batch-size", "8", "--num-epochs", "1", "--gpus", "", "--image-size", "32", "--num-classes", "10",
              "--num-training-samples", "256", "--mode", "symbolic", "--kvstore", "local", "--log-interval", "2", "--lr", "0.01"],
             str(tmp_path))
    assert r.returncode == 0, r.stdout[-3000:]
    assert "Train-accuracy" in r.stdout and os.path.exists(os.path.join(str(tmp_path), "params", "imagenet-resnet18_v1-0001.params"))


def test_example_train_imagenet_and_benchmark_score_cpu(tmp_path):
    r = _run([os.path.join(ROOT, "examples", "image-classification", "train_imagenet.py"), "--network", "resnet", "--num-layers", "18",
              "--benchmark", "1", "--batch-size", "4", "--image-shape", "3,32,32", "--num-classes", "10", "--num-examples", "32",
              "--disp-batches", "2", "--kv-store", "local"], str(tmp_path))
    assert r.returncode == 0 and "Train-accuracy" in r.stdout, r.stdout[-3000:]
    r = _run([os.path.join(ROOT, "examples", "image-classification", "benchmark_score.py"), "--networks", "resnet18_v1",
              "--batch-sizes", "2", "--num-batches", "2"], str(tmp_path))
    assert r.returncode == 0 and "images/sec" in r.stdout, r.stdout[-3000:]


def test_im2rec_and_rec2idx_roundtrip(tmp_path):
    import numpy as np
    cv2 = __import__("pytest").importorskip("cv2")
    root = os.path.join(str(tmp_path), "img")
    for c in ("a", "b"):
        os.makedirs(os.path.join(root, c))
        for i in range(3):
            cv2.imwrite(os.path.join(root, c, "%d.jpg" % i), (np.random.rand(20, 24, 3) * 255).astype("uint8"))
    prefix = os.path.join(str(tmp_path), "p")
    assert _run([os.path.join(ROOT, "tools", "im2rec.py"), "--list", prefix, root], str(tmp_path)).returncode == 0
    r = _run([os.path.join(ROOT, "tools", "im2rec.py"), prefix, root, "--resize", "16"], str(tmp_path))
    assert r.returncode == 0 and "wrote 6 records" in r.stdout, r.stdout
    r = _run([os.path.join(ROOT, "tools", "rec2idx.py"), prefix + ".rec", prefix + "2.idx"], str(tmp_path))
    assert r.returncode == 0 and "indexed 6 records" in r.stdout, r.stdout
    sys.path.insert(0, ROOT)
    import mxb200 as mx
    it = mx.image.ImageIter(2, (3, 16, 16), path_imgrec=prefix + ".rec", path_imgidx=prefix + ".idx")
    b = next(iter(it))
    assert b.data[0].shape == (2, 3, 16, 16)


def test_reference_unittest_runner(tmp_path):
    """tools/run_reference_tests.py executes the reference's own unit tests against this package (API-parity probe)."""
    import os
    import subprocess
    import sys
    ref = os.environ.get("MXB200_REFERENCE_TESTS", "/root/reference/tests/python/unittest")
    if not os.path.exists(os.path.join(ref, "test_metric.py")):
        import pytest
        pytest.skip("reference test tree not mounted")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = str(tmp_path / "t.md")
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "run_reference_tests.py"), "--out", out, "test_metric", "test_attr"],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout
    assert "test_metric" in r.stdout and "failed   0" in r.stdout
    assert "| test_attr | 4 | 0 |" in open(out).read()
