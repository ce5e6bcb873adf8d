"""Profiler: Chrome-trace JSON + NVTX ranges + aggregate table (reference: python/mxnet/profiler.py,
src/profiler/profiler.{h,cc}: modes, run/stop/pause, Domain/Task/Frame/Event/Counter/Marker objects,
``dump`` writing {"traceEvents": [...]}, ``dumps`` printing per-name min/max/avg).

Device-side ranges are recorded with CUDA events so GPU durations are real device time, and mirrored as
NVTX ranges (visible in Nsight).  ``MXNET_PROFILER_AUTOSTART=1`` starts profiling at import.
"""
import json
import os
import threading
import time
from collections import defaultdict

import torch

from .base import getenv

__all__ = ["set_config", "set_state", "dump", "dumps", "pause", "resume", "Domain", "Task", "Frame", "Event", "Counter",
           "Marker", "profiler_set_config", "profiler_set_state", "dump_profile", "scope", "set_kvstore_handle"]

_cfg = {"filename": "profile.json", "profile_all": False, "profile_symbolic": False, "profile_imperative": False,
        "profile_memory": False, "profile_api": False, "aggregate_stats": False, "continuous_dump": False,
        "dump_period": 1.0, "profile_process": "worker"}
_state = {"running": False, "paused": False}
_events = []
_gpu_pending = []
_lock = threading.Lock()
_t0 = time.time()
_kv = [None]


def set_kvstore_handle(kv):
    _kv[0] = kv


def set_config(**kwargs):
    """profile_process='server' forwards the config to the kvstore servers (kSetProfilerParams,
    kvstore_dist.h:102-110)."""
    if kwargs.get("profile_process", "worker") == "server" and _kv[0] is not None:
        _kv[0]._send_command_to_servers(5, json.dumps({"cmd": "set_config", "args": kwargs}))
        return
    for k, v in kwargs.items():
        if k not in _cfg:
            raise ValueError("unknown profiler option %s" % k)
        _cfg[k] = v


profiler_set_config = lambda mode="symbolic", filename="profile.json": set_config(  # noqa: E731
    profile_symbolic=True, profile_imperative=mode == "all", filename=filename)


def set_state(state="stop", profile_process="worker"):
    if profile_process == "server" and _kv[0] is not None:
        _kv[0]._send_command_to_servers(5, json.dumps({"cmd": "set_state", "args": state}))
        return
    assert state in ("run", "stop")
    _state["running"] = state == "run"
    _engine_hook(state == "run")
    if state == "stop" and _cfg.get("continuous_dump"):
        dump()


profiler_set_state = set_state


def _engine_hook(on):
    """Per-operation records of the dependency engine (IO, checkpoint writes, host reductions): the engine brackets every
    executed operation while profiling is on (reference: src/engine/threaded_engine.h:338-347)."""
    from . import engine as _eng
    try:
        _eng.get().set_profiling(on)
    except Exception:  # noqa: BLE001  (host library unavailable)
        pass


def _drain_engine():
    from . import engine as _eng
    if _eng._engine[0] is None:
        return
    t0us = _t0 * 1e6
    for r in _eng._engine[0].drain_profile():
        _add({"name": r["name"] or "engine_op", "cat": "engine", "ph": "X", "ts": r["start_us"] - t0us,
              "dur": max(0, r["end_us"] - r["start_us"]), "pid": os.getpid(), "tid": "engine-%d" % r["thread"],
              "args": {"queue_us": r["start_us"] - r["push_us"], "priority": r["priority"], "seq": r["seq"]}})


def pause(profile_process="worker"):
    _state["paused"] = True


def resume(profile_process="worker"):
    _state["paused"] = False


def _active():
    return _state["running"] and not _state["paused"]


def _now_us():
    return (time.time() - _t0) * 1e6


def _add(ev):
    with _lock:
        _events.append(ev)


def _flush_gpu():
    with _lock:
        pend, _gpu_pending[:] = list(_gpu_pending), []
    for name, cat, e0, e1, host_ts, args in pend:
        e1.synchronize()
        dur = e0.elapsed_time(e1) * 1e3
        _add({"name": name, "cat": cat, "ph": "X", "ts": host_ts, "dur": dur, "pid": os.getpid(), "tid": "gpu", "args": args})


class scope(object):
    """``with profiler.scope('kvstore:push', device=True):`` records one complete event (host wall time, plus the
    device time between two CUDA events when device=True) and an NVTX range."""

    def __init__(self, name, cat="operator", device=False, **args):
        self.name, self.cat, self.device, self.args = name, cat, device and torch.cuda.is_available(), args

    def __enter__(self):
        if not _active():
            return self
        self.t = _now_us()
        if torch.cuda.is_available():
            torch.cuda.nvtx.range_push(self.name)
        if self.device:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *a):
        if not _active() or not hasattr(self, "t"):
            return
        if self.device:
            self.e1.record()
            with _lock:
                _gpu_pending.append((self.name, self.cat, self.e0, self.e1, self.t, self.args))
        if torch.cuda.is_available():
            torch.cuda.nvtx.range_pop()
        _add({"name": self.name, "cat": self.cat, "ph": "X", "ts": self.t, "dur": _now_us() - self.t, "pid": os.getpid(),
              "tid": threading.get_ident() % 100000, "args": self.args})


def dump(finished=True, profile_process="worker"):
    _flush_gpu()
    _drain_engine()
    with _lock:
        evs = list(_events)
        if finished:
            _events[:] = []
    with open(_cfg["filename"], "w") as f:
        json.dump({"traceEvents": evs, "displayTimeUnit": "ms"}, f)


dump_profile = dump


def dumps(reset=False):
    """Aggregate per-name statistics table (src/profiler/aggregate_stats.cc)."""
    _flush_gpu()
    _drain_engine()
    agg = defaultdict(list)
    with _lock:
        for e in _events:
            if e.get("ph") == "X":
                agg[(e["cat"], e["name"], e["tid"] == "gpu")].append(e["dur"])
        if reset:
            _events[:] = []
    lines = ["%-12s %-40s %8s %12s %12s %12s %12s" % ("Domain", "Name", "Count", "Total(us)", "Min(us)", "Max(us)", "Avg(us)")]
    for (cat, name, gpu), v in sorted(agg.items()):
        lines.append("%-12s %-40s %8d %12.1f %12.1f %12.1f %12.1f" % (cat + ("/gpu" if gpu else ""), name[:40], len(v), sum(v), min(v), max(v), sum(v) / len(v)))
    return "\n".join(lines)


class Domain(object):
    def __init__(self, name):
        self.name = name

    def __str__(self):
        return self.name

    def new_task(self, name):
        return Task(self, name)

    def new_frame(self, name):
        return Frame(self, name)

    def new_counter(self, name, value=None):
        return Counter(self, name, value)

    def new_marker(self, name):
        return Marker(self, name)


class _Span(object):
    def __init__(self, domain, name):
        self.domain, self.name = domain, name
        self._scope = None

    def start(self):
        self._scope = scope(self.name, cat=str(self.domain))
        self._scope.__enter__()

    def stop(self):
        if self._scope is not None:
            self._scope.__exit__(None, None, None)
            self._scope = None

    def __str__(self):
        return self.name


class Task(_Span):
    pass


class Frame(_Span):
    pass


class Event(_Span):
    def __init__(self, name):
        super(Event, self).__init__("event", name)


class Counter(object):
    def __init__(self, domain, name, value=None):
        self.domain, self.name, self.value = domain, name, 0
        if value is not None:
            self.set_value(value)

    def set_value(self, value):
        self.value = value
        if _active():
            _add({"name": self.name, "cat": str(self.domain), "ph": "C", "ts": _now_us(), "pid": os.getpid(), "args": {self.name: value}})

    def increment(self, delta=1):
        self.set_value(self.value + delta)

    def decrement(self, delta=1):
        self.set_value(self.value - delta)

    def __iadd__(self, v):
        self.increment(v)
        return self

    def __isub__(self, v):
        self.decrement(v)
        return self


class Marker(object):
    def __init__(self, domain, name):
        self.domain, self.name = domain, name

    def mark(self, scope_="process"):
        if _active():
            _add({"name": self.name, "cat": str(self.domain), "ph": "i", "ts": _now_us(), "pid": os.getpid(),
                  "s": {"global": "g", "process": "p", "thread": "t"}.get(scope_, "p")})


if getenv("MXNET_PROFILER_AUTOSTART", False):
    set_config(profile_all=True, filename="profile.json")
    set_state("run")
