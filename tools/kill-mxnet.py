#!/usr/bin/env python
"""Stop a job started with tools/launch.py on the hosts of a host file (reference: tools/kill-mxnet.py).

    python tools/kill-mxnet.py hostfile [user] [prog]

Processes are selected by the DMLC_ROLE environment variable of jobs started by this user and the given program name
(exact PIDs are collected first and then signalled; nothing is killed by command-line pattern)."""
import os
import signal
import subprocess
import sys


def local_pids(prog):
    pids = []
    for pid in os.listdir("/proc"):
        if not pid.isdigit() or int(pid) == os.getpid():
            continue
        try:
            if os.stat("/proc/" + pid).st_uid != os.getuid():
                continue
            env = open("/proc/%s/environ" % pid, "rb").read().split(b"\0")
            cmd = open("/proc/%s/cmdline" % pid, "rb").read().split(b"\0")
        except OSError:
            continue
        if any(e.startswith(b"DMLC_ROLE=") for e in env) and (not prog or any(prog.encode() in c for c in cmd)):
            pids.append(int(pid))
    return pids


if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit("usage: kill-mxnet.py hostfile [user] [prog]")
    prog = sys.argv[3] if len(sys.argv) > 3 else ""
    hosts = [h.strip() for h in open(sys.argv[1]) if h.strip()]
    me = os.path.abspath(__file__)
    for h in hosts:
        if h in ("localhost", "127.0.0.1") or h.startswith("local"):
            for pid in local_pids(prog):
                print("stopping pid", pid)
                os.kill(pid, signal.SIGTERM)
        else:
            user = (sys.argv[2] + "@") if len(sys.argv) > 2 and sys.argv[2] else ""
            subprocess.call(["ssh", "-o", "StrictHostKeyChecking=no", user + h, sys.executable, me, "/dev/stdin", "", prog],
                            stdin=subprocess.DEVNULL)
