#!/usr/bin/env python
"""Job launcher (reference: tools/launch.py:40-324 + the dmlc tracker it drives, which is MISSING from the fork's
snapshot -- SURVEY.md section 0; its contract is rebuilt here from the consumers of its environment variables).

    python tools/launch.py -n 4 [-s 1] [--launcher local|ssh] [-H hostfile] [--elastic-training-enabled True]
                           [--env K:V]* command ...

* starts one scheduler, ``-s`` servers and ``-n`` workers (local subprocesses, or over ssh round-robin on ``-H``),
  exporting DMLC_ROLE / DMLC_PS_ROOT_URI / DMLC_PS_ROOT_PORT / DMLC_NUM_WORKER / DMLC_NUM_SERVER;
* with ``--elastic-training-enabled True`` also exports ELASTIC_TRAINING_ENABLED, WORKER_HOST_FILE (``host_worker``
  beside the host file; the first ``-n`` lines are the immutable base workers), MXNET_LAUNCH_SCRIPT_PATH,
  TRAINING_CMD and INSTANCE_POOL=DEFAULT to the scheduler, which then adds / removes workers at epoch boundaries
  by re-reading that file (csrc/host/pslite.cc);
* ``--launch-worker True --host H`` starts ONE extra worker (this is what the scheduler invokes for newcomers);
* the optional EC2 manager thread (WORKER_LAUNCH_TEMPLATE_ID + ELASTIC_WORKER_TAG set) mirrors the reference's
  tag-driven host-file manager (launch.py:88-235); it needs boto3 and network access, both absent here.
"""
import argparse
import os
import signal
import socket
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def parse_env(items):
    env = {}
    for it in items or []:
        if ":" in it:
            k, v = it.split(":", 1)
            env[k] = v
        elif it in os.environ:
            env[it] = os.environ[it]
    return env


def spawn(cmd, env, host=None, launcher="local"):
    full = dict(os.environ)
    full.update(env)
    full["PYTHONPATH"] = ROOT + os.pathsep + full.get("PYTHONPATH", "")
    if launcher == "ssh" and host not in (None, "127.0.0.1", "localhost") and not host.startswith("local"):
        exports = " ".join("export %s=%s;" % (k, v) for k, v in env.items())
        return subprocess.Popen(["ssh", "-o", "StrictHostKeyChecking=no", host, exports + " cd %s; %s" % (os.getcwd(), " ".join(cmd))])
    return subprocess.Popen(cmd, env=full)


class HostFileManager(threading.Thread):
    """EC2 tag driven manager of ``host_worker`` (reference launch.py:88-235): base hosts are immutable, instances are
    admitted when their ``prepare_data_success_<ip>`` marker exists, the file is replaced atomically (.tmp + rename)
    and ``host_worker_log`` lines are mirrored back into instance tags.  Requires boto3."""

    def __init__(self, host_file, base_hosts, interval=5):
        super(HostFileManager, self).__init__(daemon=True)
        self.host_file, self.base_hosts, self.interval = host_file, list(base_hosts), interval

    def run(self):
        try:
            import boto3  # noqa: F401
        except ImportError:
            sys.stderr.write("launch.py: boto3 is not available; EC2 host-file manager disabled\n")
            return
        import boto3
        ec2 = boto3.client("ec2", region_name=os.environ.get("AWS_REGION", "us-east-1"))
        tag = os.environ["ELASTIC_WORKER_TAG"]
        seen_log = 0
        while True:
            resp = ec2.describe_instances(Filters=[{"Name": "tag-key", "Values": [tag]}, {"Name": "instance-state-name", "Values": ["running"]}])
            hosts = list(self.base_hosts)
            for r in resp.get("Reservations", []):
                for inst in r.get("Instances", []):
                    ip = inst.get("PrivateIpAddress")
                    status = {t["Key"]: t["Value"] for t in inst.get("Tags", [])}.get("elastic_worker_status", "")
                    ok = os.path.exists(os.path.join(os.path.dirname(self.host_file), "prepare_data_success_%s" % ip))
                    if ip and ip not in hosts and status in ("ADD_QUEUED", "ADDED") and ok:
                        hosts.append(ip)
            tmp = self.host_file + ".tmp"
            with open(tmp, "w") as f:
                f.write("\n".join(hosts) + "\n")
            os.rename(tmp, self.host_file)
            log = self.host_file + "_log"
            if os.path.exists(log):
                lines = open(log).read().splitlines()
                for line in lines[seen_log:]:
                    _seq, what, ip, _t = line.split()
                    # mirror ADDED / REMOVED into tags; terminate removed instances
                    sys.stderr.write("launch.py: %s %s\n" % (what, ip))
                seen_log = len(lines)
            time.sleep(self.interval)


def main():
    ap = argparse.ArgumentParser(description="Launch a distributed (optionally elastic) job")
    ap.add_argument("-n", "--num-workers", required=True, type=int)
    ap.add_argument("-s", "--num-servers", type=int)
    ap.add_argument("-H", "--hostfile", type=str)
    ap.add_argument("--launcher", type=str, default="local", choices=["local", "ssh", "mpi", "sge", "yarn"])
    ap.add_argument("--sync-dst-dir", type=str)
    ap.add_argument("--env-server", action="append", default=[])
    ap.add_argument("--env-worker", action="append", default=[])
    ap.add_argument("--env", action="append", default=[])
    ap.add_argument("--elastic-training-enabled", type=str, default="False")
    ap.add_argument("--launch-worker", type=str, default="False")
    ap.add_argument("--host", type=str)
    ap.add_argument("--port", type=str)
    ap.add_argument("command", nargs=argparse.REMAINDER)
    args = ap.parse_args()
    if args.launcher in ("mpi", "sge", "yarn"):
        sys.exit("launch.py: launcher '%s' needs a cluster scheduler that does not exist in this environment" % args.launcher)
    elastic = args.elastic_training_enabled in ("True", "true", "1")
    ns = args.num_servers if args.num_servers is not None else args.num_workers
    cmd = args.command
    if cmd and cmd[0] == "--":
        cmd = cmd[1:]
    common = parse_env(args.env)

    if args.launch_worker in ("True", "true", "1"):
        env = dict(common)
        env.update({"DMLC_ROLE": "worker", "MXB200_WORKER_TAG": args.host or "", "DMLC_NODE_HOST": "127.0.0.1"})
        p = spawn(cmd, env, args.host, args.launcher)
        sys.exit(p.wait())

    hosts = []
    if args.hostfile:
        hosts = [h.strip() for h in open(args.hostfile) if h.strip()]
    port = int(args.port or free_port())
    base = {"DMLC_PS_ROOT_URI": "127.0.0.1", "DMLC_PS_ROOT_PORT": str(port), "DMLC_NUM_WORKER": str(args.num_workers),
            "DMLC_NUM_SERVER": str(ns)}
    base.update(common)
    worker_tags = [(hosts[i % len(hosts)] if hosts else "local%d" % i) for i in range(args.num_workers)]
    host_worker = None
    if elastic:
        d = os.path.dirname(os.path.abspath(args.hostfile)) if args.hostfile else os.getcwd()
        host_worker = os.path.join(d, "host_worker")
        with open(host_worker, "w") as f:
            f.write("\n".join(worker_tags) + "\n")
        open(host_worker + "_log", "w").close()
        base["ELASTIC_TRAINING_ENABLED"] = "1"
    procs = []
    sched_env = dict(base, DMLC_ROLE="scheduler")
    if elastic:
        sched_env.update({"WORKER_HOST_FILE": host_worker, "INSTANCE_POOL": "DEFAULT",
                          "MXNET_LAUNCH_SCRIPT_PATH": os.path.abspath(__file__), "TRAINING_CMD": " ".join(cmd)})
        if os.environ.get("WORKER_LAUNCH_TEMPLATE_ID") and os.environ.get("ELASTIC_WORKER_TAG"):
            HostFileManager(host_worker, worker_tags).start()
    server_boot = [sys.executable, "-m", "mxb200.kvstore_server"]
    procs.append(spawn(server_boot, sched_env))
    for _ in range(ns):
        procs.append(spawn(server_boot, dict(base, DMLC_ROLE="server", **parse_env(args.env_server))))
    workers = []
    for i in range(args.num_workers):
        env = dict(base, DMLC_ROLE="worker", MXB200_WORKER_TAG=worker_tags[i], **parse_env(args.env_worker))
        workers.append(spawn(cmd, env, worker_tags[i], args.launcher))

    def kill_all(*_a):
        for p in procs + workers:
            if p.poll() is None:
                p.terminate()
        sys.exit(1)
    signal.signal(signal.SIGINT, kill_all)
    signal.signal(signal.SIGTERM, kill_all)
    rc = 0
    for w in workers:
        rc |= w.wait()
    deadline = time.time() + 60
    for p in procs:
        try:
            p.wait(timeout=max(1, deadline - time.time()))
        except subprocess.TimeoutExpired:
            p.terminate()
    sys.exit(rc)


if __name__ == "__main__":
    main()
