"""Dependency engine: randomized read/write workloads must equal serial execution; waits, stop/start,
exception propagation (reference tests: tests/cpp/engine/threaded_engine_test.cc:46-353, test_exc_handling.py)."""
import random
import threading
import time

import pytest

import mxb200 as mx
from mxb200.engine import Engine


@pytest.mark.parametrize("kind", ["ThreadedEngine", "NaiveEngine"])
def test_random_workload_equals_serial(kind):
    eng = Engine(kind=kind, num_workers=4)
    nvar, nops = 8, 300
    rng = random.Random(7)
    vars_ = [eng.new_variable() for _ in range(nvar)]
    data = [0] * nvar
    serial = [0] * nvar
    lock = threading.Lock()
    for i in range(nops):
        reads = rng.sample(range(nvar), rng.randint(0, 3))
        writes = rng.sample([v for v in range(nvar) if v not in reads], rng.randint(1, 2))
        k = rng.randint(1, 5)
        # serial oracle
        s = sum(serial[r] for r in reads)
        for w in writes:
            serial[w] = serial[w] * 3 + s + k

        def op(reads=reads, writes=writes, k=k):
            s = sum(data[r] for r in reads)
            time.sleep(0.0002 * (k % 3))
            for w in writes:
                data[w] = data[w] * 3 + s + k
        eng.push(op, [vars_[r] for r in reads], [vars_[w] for w in writes])
    eng.wait_all()
    assert data == serial
    assert eng.num_pending() == 0


def test_wait_for_var_and_priority():
    eng = Engine(num_workers=2)
    v = eng.new_variable()
    out = []
    eng.push(lambda: (time.sleep(0.05), out.append(1)), [], [v])
    eng.push(lambda: out.append(2), [], [v])
    eng.wait_for_var(v)
    assert out == [1, 2]


def test_concurrent_readers_run_in_parallel():
    eng = Engine(num_workers=4)
    v = eng.new_variable()
    active, peak = [0], [0]
    lock = threading.Lock()

    def reader():
        with lock:
            active[0] += 1
            peak[0] = max(peak[0], active[0])
        time.sleep(0.05)
        with lock:
            active[0] -= 1
    for _ in range(4):
        eng.push(reader, [v], [])
    eng.wait_all()
    assert peak[0] >= 2


def test_exception_propagates_at_wait():
    eng = Engine(num_workers=2)
    v = eng.new_variable()

    def boom():
        raise ValueError("bad op")
    eng.push(boom, [], [v])
    with pytest.raises(ValueError):
        eng.wait_all()
    eng.push(lambda: None, [], [v])     # engine keeps working afterwards
    eng.wait_all()


def test_stop_start_drain():
    eng = Engine(num_workers=2)
    v = eng.new_variable()
    hits = []
    for i in range(10):
        eng.push(lambda i=i: hits.append(i), [], [v])
    eng.stop()                           # drains everything (elastic re-init)
    assert hits == list(range(10))
    eng.start()
    eng.push(lambda: hits.append(99), [], [v])
    eng.wait_all()
    assert hits[-1] == 99
    mx.nd.waitall()


def test_engine_info_logging(monkeypatch, caplog):
    """MXNET_ENGINE_INFO=1 reports pushes / dispatches / waits (threaded_engine.h:302,350 of the reference)."""
    import logging
    from mxb200 import engine as E
    monkeypatch.setenv("MXNET_ENGINE_INFO", "1")
    eng = E.Engine()
    v = eng.new_variable()
    out = []
    with caplog.at_level(logging.INFO, logger="mxb200.engine"):
        eng.push(lambda: out.append(1), mutable_vars=[v] if not isinstance(v, object().__class__) else [], name="probe")
        eng.wait_all()
    assert out == [1]
    text = caplog.text
    assert "probe" in text and "wait_all" in text


def test_profiler_records_every_engine_operation_with_queue_time(tmp_path):
    """Engine hook of the profiler (reference: threaded_engine.h:338-347 + tests/python/unittest/test_profiler.py)."""
    import json
    import time
    import mxb200 as mx
    from mxb200 import engine, profiler
    eng = engine.get()
    if eng._lib is None:
        pytest.skip("native engine unavailable")
    eng.wait_all()
    profiler.set_config(filename=str(tmp_path / "trace.json"))
    profiler.set_state("run")
    v = eng.new_variable()
    eng.push(lambda: time.sleep(0.05), mutable_vars=[v], name="slow_writer")
    eng.push(lambda: None, const_vars=[v], name="reader_after_writer", priority=3)
    eng.push(lambda: None, name="independent")
    eng.wait_all()
    profiler.set_state("stop")
    table = profiler.dumps()
    assert "slow_writer" in table and "reader_after_writer" in table
    profiler.dump()
    evs = [e for e in json.load(open(str(tmp_path / "trace.json")))["traceEvents"] if e.get("cat") == "engine"]
    by = {e["name"]: e for e in evs}
    assert {"slow_writer", "reader_after_writer", "independent"} <= set(by)
    assert by["slow_writer"]["dur"] >= 45e3
    # the reader was queued behind the writer: its queue time covers the writer's run, and it starts after the writer ends
    assert by["reader_after_writer"]["args"]["queue_us"] >= 40e3 and by["reader_after_writer"]["args"]["priority"] == 3
    assert by["reader_after_writer"]["ts"] >= by["slow_writer"]["ts"] + by["slow_writer"]["dur"] - 1e3
    # nothing is recorded while the profiler is stopped
    eng.push(lambda: None, name="not_profiled")
    eng.wait_all()
    assert eng.drain_profile() == []
