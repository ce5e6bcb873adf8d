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
