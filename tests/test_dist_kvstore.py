"""Distributed kvstore + elastic membership on one host, the reference's "distributed without a cluster" strategy
(tests/nightly/dist_sync_kvstore.py launched with tools/launch.py --launcher local): scheduler, servers and workers
are real processes talking over loopback TCP.  The elastic scenarios have NO counterpart in the reference test
suite."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAUNCH = os.path.join(ROOT, "tools", "launch.py")


def _launch(args, cwd, timeout=240, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, LAUNCH] + args, cwd=cwd, timeout=timeout, capture_output=True, text=True, env=e)


def test_dist_sync_closed_form_three_workers_two_servers(tmp_path):
    r = _launch(["-n", "3", "-s", "2", "--launcher", "local", sys.executable, os.path.join(ROOT, "tests", "dist_worker.py")], str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("OK") == 3


def test_dist_sync_two_bit_compression(tmp_path):
    r = _launch(["-n", "2", "-s", "1", "--launcher", "local", sys.executable, os.path.join(ROOT, "tests", "dist_worker.py"),
                 "compressed"], str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr


def test_dist_sync_survives_dropped_messages_with_resend(tmp_path):
    """PS_DROP_MSG fault injection + PS_RESEND reliable delivery (van.cc:430-432,463-469; never driven by a reference test)."""
    r = _launch(["-n", "2", "-s", "1", "--launcher", "local", sys.executable, os.path.join(ROOT, "tests", "dist_worker.py")],
                str(tmp_path), env={"PS_DROP_MSG": "10", "PS_RESEND": "1", "PS_RESEND_TIMEOUT": "100"})
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("OK") == 2


def test_elastic_add_then_remove_worker(tmp_path):
    """2 base workers; 'local2' is appended to the host file during epoch 0 and removed during epoch 2.
    Invariants: changes only at epoch boundaries, removal/addition take effect at the next barrier,
    ranks stay dense, the newcomer starts at EPOCH_BEGIN with state pulled from the servers, aux keys are averaged
    over the LIVE worker count, the audit log has strictly increasing sequence numbers."""
    out = str(tmp_path)
    r = _launch(["-n", "2", "-s", "1", "--launcher", "local", "--elastic-training-enabled", "True", sys.executable,
                 os.path.join(ROOT, "tests", "elastic_worker.py"), out], out, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    logs = {t: json.load(open(os.path.join(out, "out_%s.json" % t))) for t in ("local0", "local1", "local2")}
    base = logs["local0"]["epochs"]
    assert [e["n"] for e in base] == [2, 3, 3, 2, 2]
    new = logs["local2"]
    assert new["new"] and new["begin"] == 1 and new["pulled_on_join"][0] == 2.0
    assert [e["epoch"] for e in new["epochs"]] == [1, 2]                 # joined at epoch 1, gone after epoch 2
    for t in ("local0", "local1"):
        assert [e["w"] for e in logs[t]["epochs"]] == [2.0, 5.0, 8.0, 10.0, 12.0]
        assert [e["aux"] for e in logs[t]["epochs"]] == [1.5, 2.0, 2.0, 1.5, 1.5]
    for ep in range(5):
        ranks = sorted(l["epochs"][i]["rank"] for l in logs.values() for i in range(len(l["epochs"])) if l["epochs"][i]["epoch"] == ep)
        assert ranks == list(range(len(ranks)))                            # dense ranks after every change
    lines = open(os.path.join(out, "host_worker_log")).read().split("\n")
    lines = [l.split() for l in lines if l]
    assert [(l[0], l[1], l[2]) for l in lines] == [("0", "ADDED", "local2"), ("1", "REMOVED", "local2")]
    assert int(lines[1][3]) > int(lines[0][3])


def test_scheduler_allgather_three_workers(tmp_path):
    """Small-blob all-gather through the scheduler (exchange of CUDA IPC handles when the worker set changes)."""
    r = _launch(["-n", "3", "-s", "1", "--launcher", "local", sys.executable, os.path.join(ROOT, "tests", "allgather_worker.py")],
                str(tmp_path))
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("OK") == 3
