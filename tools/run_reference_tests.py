#!/usr/bin/env python
"""Run the reference's python unit tests against mxb200 (API-parity probe).

The test files are executed where they are (``/root/reference/tests/python/unittest``, read-only; override with
MXB200_REFERENCE_TESTS); ``import mxnet`` is aliased to this package by tools/reftests/mxshim.py and ``nose`` by a small
stub.  Prints one line per file: passed / failed / skipped / errors, and optionally writes a markdown table.

    python tools/run_reference_tests.py [--out profiles/reference_unittests.md] [--timeout 120] [test_metric test_symbol ...]
"""
import argparse
import os
import re
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT = ["test_attr", "test_autograd", "test_base", "test_contrib_autograd", "test_contrib_krprod", "test_contrib_optimizer",
           "test_contrib_svrg_module", "test_contrib_svrg_optimizer", "test_contrib_text", "test_engine", "test_exc_handling",
           "test_executor", "test_gluon", "test_gluon_contrib", "test_gluon_data", "test_gluon_data_vision", "test_gluon_model_zoo",
           "test_gluon_rnn", "test_gluon_trainer", "test_image", "test_infer_shape", "test_infer_type", "test_init", "test_io",
           "test_kvstore", "test_loss", "test_metric", "test_model_parallel", "test_module", "test_multi_device_exec",
           "test_ndarray", "test_operator", "test_optimizer", "test_profiler", "test_random", "test_recordio", "test_rnn", "test_sparse_ndarray",
           "test_sparse_operator", "test_symbol", "test_thread_local", "test_viz"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("tests", nargs="*")
    ap.add_argument("--out", default="")
    ap.add_argument("--timeout", type=int, default=120)
    a = ap.parse_args()
    ref = os.environ.get("MXB200_REFERENCE_TESTS", "/root/reference/tests/python/unittest")
    shim = os.path.join(HERE, "reftests")
    rows = []
    work = tempfile.mkdtemp(prefix="mxb200_reftests_")
    for t in (a.tests or DEFAULT):
        path = os.path.join(ref, t + ".py")
        if not os.path.exists(path):
            rows.append((t, "missing", 0, 0, 0, 0))
            continue
        env = dict(os.environ, PYTHONPATH=shim + os.pathsep + os.environ.get("PYTHONPATH", ""))
        cmd = [sys.executable, "-m", "pytest", "-p", "mxshim", path, "-q", "-p", "no:cacheprovider", "--rootdir", work, "-c", os.devnull,
               "--timeout", str(a.timeout), "--no-header"]
        try:
            out = subprocess.run(cmd, cwd=work, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                                 timeout=a.timeout * 40).stdout
        except subprocess.TimeoutExpired:
            rows.append((t, "timeout", 0, 0, 0, 0))
            continue
        tail = [l for l in out.splitlines() if re.search(r"\d+ (passed|failed|skipped|error)", l)]
        line = tail[-1] if tail else ""
        get = lambda k: int((re.search(r"(\d+) " + k, line) or [0, 0])[1])   # noqa: E731
        rows.append((t, "ok" if line else "collect-error", get("passed"), get("failed"), get("skipped"), get("error")))
        print("%-32s passed %3d  failed %3d  skipped %3d  errors %d" % (t, rows[-1][2], rows[-1][3], rows[-1][4], rows[-1][5]), flush=True)
    tp, tf = sum(r[2] for r in rows), sum(r[3] for r in rows)
    print("TOTAL passed %d failed %d (of %d test functions run)" % (tp, tf, tp + tf))
    if a.out:
        with open(a.out, "w") as f:
            f.write("# Reference python unit tests run against mxb200 (tools/run_reference_tests.py)\n\n"
                    "The files are the reference's own `tests/python/unittest/*.py`, executed in place with `mxnet` aliased to\n"
                    "`mxb200`; failures include tests that need network downloads or GPU-only features.\n\n"
                    "| file | passed | failed | skipped | status |\n|---|---|---|---|---|\n")
            for r in rows:
                f.write("| %s | %d | %d | %d | %s |\n" % (r[0], r[2], r[3], r[4], r[1] if r[5] == 0 else "%s (%d errors)" % (r[1], r[5])))
            f.write("| **total** | **%d** | **%d** | | |\n" % (tp, tf))


if __name__ == "__main__":
    main()
