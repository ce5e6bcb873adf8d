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


def test_launch_host_file_manager_tags_and_terminates(tmp_path):
    """The EC2 side effects of the elastic launcher (reference tools/launch.py:171-199) against a fake EC2 client: ADD_QUEUED ->
    ADDED on the scheduler's log line, REMOVAL_QUEUED -> REMOVED + terminate, base hosts immutable, atomic host file."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("mxb_launch", os.path.join(os.path.dirname(__file__), "..", "tools", "launch.py"))
    launch = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(launch)

    class FakeEC2(object):
        def __init__(self):
            self.inst = {}          # ip -> (id, tags)
            self.terminated = []

        def describe_instances(self, Filters):
            return {"Reservations": [{"Instances": [{"PrivateIpAddress": ip, "InstanceId": iid,
                                                      "Tags": [{"Key": k, "Value": v} for k, v in tags.items()]}
                                                     for ip, (iid, tags) in self.inst.items() if iid not in self.terminated]}]}

        def create_tags(self, Resources, Tags):
            for ip, (iid, tags) in self.inst.items():
                if iid in Resources:
                    for t in Tags:
                        tags[t["Key"]] = t["Value"]

        def terminate_instances(self, InstanceIds):
            self.terminated.extend(InstanceIds)

    hf = str(tmp_path / "host_worker")
    open(hf, "w").write("10.0.0.1\n10.0.0.2\n")
    ec2 = FakeEC2()
    mgr = launch.HostFileManager(hf, ["10.0.0.1", "10.0.0.2"], ec2=ec2, elastic_tag="elastic", mount=str(tmp_path))
    ec2.inst["10.0.0.9"] = ("i-9", {"elastic": "1", "NodeType": "Worker"})
    assert mgr.poll() == ["10.0.0.1", "10.0.0.2"]                       # data not staged yet: not admitted
    open(str(tmp_path / "prepare_data_success_10.0.0.9"), "w").write("ok")
    assert mgr.poll() == ["10.0.0.1", "10.0.0.2", "10.0.0.9"]
    assert ec2.inst["10.0.0.9"][1]["elastic_worker_status"] == "ADD_QUEUED"
    open(hf + "_log", "w").write("0 ADDED 10.0.0.9 123\n")
    mgr.poll()
    assert ec2.inst["10.0.0.9"][1]["elastic_worker_status"] == "ADDED"
    del ec2.inst["10.0.0.9"][1]["elastic"]                              # operator withdraws the instance
    assert mgr.poll() == ["10.0.0.1", "10.0.0.2"]
    assert ec2.inst["10.0.0.9"][1]["elastic_worker_status"] == "REMOVAL_QUEUED"
    open(hf + "_log", "a").write("1 REMOVED 10.0.0.9 456\n")
    mgr.poll()
    assert ec2.inst["10.0.0.9"][1]["elastic_worker_status"] == "REMOVED" and ec2.terminated == ["i-9"]
    assert open(hf).read().split() == ["10.0.0.1", "10.0.0.2"]
