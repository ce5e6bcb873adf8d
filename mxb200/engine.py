"""Dependency engine front-end (reference: include/mxnet/engine.h:95-318, src/engine/*).

The native engine (csrc/host/engine.cc) tracks read/write dependencies on versioned variables and runs
host-side operations (IO, checkpoint writes, kvstore host reductions, membership RPCs) on worker threads
-- the role the reference's ThreadedEngine plays for *every* op.  Device kernels do not go through it:
CUDA streams already order them, so ``wait_to_read`` on device data is a stream wait.  The engine is what
``mx.nd.waitall()`` and the elastic re-init drain (Stop/Start in the reference,
threaded_engine_perdevice.cc:70-91) synchronise with.

Engine types (MXNET_ENGINE_TYPE): ``NaiveEngine`` runs ops inline on the caller thread (debugging races),
``ThreadedEngine`` (default) uses the native worker pool.
"""
import ctypes
import threading

from . import _native as N
from .base import getenv, MXNetError

__all__ = ["get", "Engine", "bulk", "set_bulk_size"]

_CB = ctypes.CFUNCTYPE(None, ctypes.c_void_p)


def _log_info(fmt, *args):
    import logging
    logging.getLogger("mxb200.engine").info(fmt, *args)


class Engine(object):
    def __init__(self, kind=None, num_workers=None):
        self.kind = kind or getenv("MXNET_ENGINE_TYPE", "ThreadedEngine")
        self._lib = None
        self._h = None
        self._keep = {}
        self._keep_lock = threading.Lock()
        self._next = [0]
        self._exc = []
        # MXNET_ENGINE_INFO=1: log every push / dispatch / wait (threaded_engine.h:302,350 of the reference)
        self._info = bool(getenv("MXNET_ENGINE_INFO", False))
        if self.kind != "NaiveEngine":
            lib = N.host(required=False)
            if lib is not None and hasattr(lib, "MXBEngineCreate"):
                self._lib = lib
                h = ctypes.c_void_p()
                nw = num_workers or getenv("MXNET_CPU_WORKER_NTHREADS", 4)
                N.check(lib, lib.MXBEngineCreate(ctypes.c_int(int(nw)), ctypes.byref(h)))
                self._h = h

    # variables ---------------------------------------------------------------------------------
    def new_variable(self):
        if self._lib is None:
            return object()
        v = ctypes.c_void_p()
        N.check(self._lib, self._lib.MXBEngineNewVar(self._h, ctypes.byref(v)))
        return v

    # push -----------------------------------------------------------------------------------------
    def push(self, fn, const_vars=(), mutable_vars=(), priority=0, name=None):
        """Schedule ``fn()`` after all earlier writers of const_vars and all earlier users of mutable_vars."""
        if self._lib is None:
            if self._info:
                _log_info("run (NaiveEngine) %s", name or getattr(fn, "__name__", "fn"))
            fn()
            return
        with self._keep_lock:
            self._next[0] += 1
            key = self._next[0]
        if self._info:
            _log_info("push #%d %s const=%d mutable=%d prio=%d", key, name or getattr(fn, "__name__", "fn"),
                      len(const_vars), len(mutable_vars), priority)

        def tramp(_arg, key=key, fn=fn):
            try:
                if self._info:
                    _log_info("dispatch #%d on %s", key, threading.current_thread().name)
                fn()
            except BaseException as e:  # noqa: BLE001 -- propagated at the next wait (test_exc_handling.py semantics)
                self._exc.append(e)
            finally:
                with self._keep_lock:
                    self._keep.pop(key, None)

        cb = _CB(tramp)
        with self._keep_lock:
            self._keep[key] = cb
        cv = (ctypes.c_void_p * max(1, len(const_vars)))(*[v.value for v in const_vars])
        mv = (ctypes.c_void_p * max(1, len(mutable_vars)))(*[v.value for v in mutable_vars])
        label = (name or getattr(fn, "__name__", "fn")).encode("utf-8", "replace")
        N.check(self._lib, self._lib.MXBEnginePushNamed(self._h, cb, None, cv, ctypes.c_int(len(const_vars)), mv,
                                                        ctypes.c_int(len(mutable_vars)), ctypes.c_int(priority), label))

    # profiler hook (src/engine/threaded_engine.h:338-347: every executed operation is bracketed when profiling) ----------
    def set_profiling(self, on):
        if self._lib is not None:
            N.check(self._lib, self._lib.MXBEngineSetProfiling(self._h, ctypes.c_int(1 if on else 0)))

    def drain_profile(self):
        """Executed operations since the last call: dicts with name, seq, push_us / start_us / end_us (unix epoch
        microseconds), thread, priority."""
        if self._lib is None:
            return []
        need = ctypes.c_int64(0)
        buf = ctypes.create_string_buffer(1 << 16)
        N.check(self._lib, self._lib.MXBEngineDrainProfile(self._h, buf, ctypes.c_int64(len(buf)), ctypes.byref(need)))
        if need.value > len(buf):
            buf = ctypes.create_string_buffer(need.value)
            N.check(self._lib, self._lib.MXBEngineDrainProfile(self._h, buf, ctypes.c_int64(len(buf)), ctypes.byref(need)))
        out = []
        for line in buf.raw[:need.value].decode("utf-8", "replace").splitlines():
            f = line.split("\t")
            if len(f) == 7:
                out.append({"name": f[0], "seq": int(f[1]), "push_us": int(f[2]), "start_us": int(f[3]), "end_us": int(f[4]),
                            "thread": int(f[5]), "priority": int(f[6])})
        return out

    def _raise_pending(self):
        if self._exc:
            e = self._exc.pop(0)
            self._exc.clear()
            raise e

    def wait_for_var(self, var):
        if self._info:
            _log_info("wait_for_var")
        if self._lib is not None:
            N.check(self._lib, self._lib.MXBEngineWaitForVar(self._h, var))
        self._raise_pending()

    def wait_all(self):
        if self._info:
            _log_info("wait_all")
        if self._lib is not None:
            N.check(self._lib, self._lib.MXBEngineWaitAll(self._h))
        self._raise_pending()

    def stop(self):
        """Drain everything and park the workers (elastic re-init / fork safety)."""
        self.wait_all()
        if self._lib is not None:
            N.check(self._lib, self._lib.MXBEngineStop(self._h))

    def start(self):
        if self._lib is not None:
            N.check(self._lib, self._lib.MXBEngineStart(self._h))

    def num_pending(self):
        if self._lib is None:
            return 0
        n = ctypes.c_int64()
        N.check(self._lib, self._lib.MXBEnginePending(self._h, ctypes.byref(n)))
        return n.value


_engine = [None]
_lock = threading.Lock()


def get():
    with _lock:
        if _engine[0] is None:
            _engine[0] = Engine()
            import atexit
            atexit.register(_drain_at_exit)
        return _engine[0]


def _drain_at_exit():
    e = _engine[0]
    if e is not None:
        try:
            e.wait_all()                       # asynchronous checkpoint writes must reach the disk
        except BaseException:  # noqa: BLE001
            pass


_bulk = threading.local()


def set_bulk_size(size):
    old = getattr(_bulk, "size", 0)
    _bulk.size = size
    return old


class bulk(object):
    """``with mx.engine.bulk(n):`` (python/mxnet/engine.py) -- kept for API parity; device work is already
    batched by stream order / CUDA graphs."""

    def __init__(self, size):
        self.size = size

    def __enter__(self):
        self.old = set_bulk_size(self.size)

    def __exit__(self, *a):
        set_bulk_size(self.old)
