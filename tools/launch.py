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
    """EC2-tag driven manager of the worker host file: the elastic control loop of the reference's launcher
    (tools/launch.py:88-235 there), restructured as a small state machine around one ``poll()`` so that it can be driven by a
    fake EC2 client in tests (tests/test_tools.py).

    Protocol (tag ``elastic_worker_status`` on worker instances of one launch template):
      * the first ``fixed`` lines of the host file are the immutable base workers;
      * an instance carrying the elastic tag joins the candidate set once its ``prepare_data_success_<ip>`` marker exists
        on the shared mount; a new candidate is written to the host file and tagged ADD_QUEUED; when the scheduler's audit
        log (``<hostfile>_log``: SEQ ADDED|REMOVED IP TIME) reports ADDED the tag becomes ADDED;
      * an elastic host whose instance lost the elastic tag is dropped from the host file and tagged REMOVAL_QUEUED; when the
        log reports REMOVED the tag becomes REMOVED and the instance is TERMINATED;
      * the host file is replaced atomically (.tmp + rename): the scheduler diffs it at the next epoch boundary
        (csrc/host/pslite.cc, removal before addition)."""

    STATUS = "elastic_worker_status"
    ADD_QUEUED, ADDED, REMOVAL_QUEUED, REMOVED = "ADD_QUEUED", "ADDED", "REMOVAL_QUEUED", "REMOVED"

    def __init__(self, host_file, base_hosts, interval=5, ec2=None, elastic_tag=None, template_id=None, mount=None):
        super(HostFileManager, self).__init__(daemon=True)
        self.host_file, self.fixed, self.interval = host_file, len(list(base_hosts)), interval
        self.ec2 = ec2
        self.elastic_tag = elastic_tag or os.environ.get("ELASTIC_WORKER_TAG", "")
        self.template_id = template_id or os.environ.get("WORKER_LAUNCH_TEMPLATE_ID", "")
        self.mount = mount or os.environ.get("EFS_MOUNT") or os.path.dirname(os.path.abspath(host_file))
        self.add_queue, self.remove_queue = {}, {}      # ip -> instance id, waiting for the scheduler's ADDED / REMOVED
        self.seen_seq = -1
        self.actions = []                               # audit trail of (verb, ip, instance id): what the manager did

    def _tag(self, iid, value, ip):
        self.ec2.create_tags(Resources=[iid], Tags=[{"Key": self.STATUS, "Value": value}])
        self.actions.append((value, ip, iid))

    def _instances(self):
        flt = [{"Name": "tag:NodeType", "Values": ["Worker"]}, {"Name": "instance-state-name", "Values": ["running"]}]
        if self.template_id:
            flt.insert(0, {"Name": "tag:aws:ec2launchtemplate:id", "Values": [self.template_id]})
        resp = self.ec2.describe_instances(Filters=flt)
        for r in resp.get("Reservations", []):
            for inst in r.get("Instances", []):
                if inst.get("PrivateIpAddress"):
                    yield inst["PrivateIpAddress"], inst["InstanceId"], {t["Key"]: t["Value"] for t in inst.get("Tags", [])}

    def poll(self):
        """One pass of the control loop; returns the host list it wrote."""
        if not os.path.exists(self.host_file):
            return None
        with open(self.host_file) as f:
            lines = [ln.strip() for ln in f if ln.strip()]
        base, elastic_now = lines[:self.fixed], lines[self.fixed:]
        candidates, lost_tag = {}, {}
        for ip, iid, tags in self._instances():
            status = tags.get(self.STATUS)
            has_elastic = self.elastic_tag in tags
            if status == self.REMOVAL_QUEUED:
                self.remove_queue[ip] = iid
            elif status == self.ADD_QUEUED:
                self.add_queue[ip] = iid
            if has_elastic and status is None:
                if os.path.exists(os.path.join(self.mount, "prepare_data_success_%s" % ip)):
                    candidates[ip] = iid                 # data staged: may join at the next epoch boundary
            elif has_elastic:
                candidates[ip] = iid
            elif status != self.REMOVAL_QUEUED:
                lost_tag[ip] = iid                       # operator removed the elastic tag: leave at the next boundary
        # the scheduler's audit log acknowledges membership changes
        terminated = []
        log = self.host_file + "_log"
        if os.path.exists(log):
            with open(log) as f:
                for ln in f:
                    parts = ln.split()
                    if len(parts) < 3 or int(parts[0]) <= self.seen_seq:
                        continue
                    seq, what, ip = int(parts[0]), parts[1], parts[2]
                    if what == "ADDED" and ip in self.add_queue:
                        self._tag(self.add_queue.pop(ip), self.ADDED, ip)
                    elif what == "REMOVED" and ip in self.remove_queue:
                        iid = self.remove_queue.pop(ip)
                        self._tag(iid, self.REMOVED, ip)
                        terminated.append(iid)
                    self.seen_seq = seq
        if terminated:
            self.ec2.terminate_instances(InstanceIds=terminated)
            self.actions.extend(("TERMINATED", None, i) for i in terminated)
        keep = [h for h in elastic_now if h in candidates]
        leaving = [h for h in elastic_now if h not in candidates]
        joining = [h for h in candidates if h not in elastic_now]
        hosts = base + keep + joining
        tmp = self.host_file + ".tmp"
        with open(tmp, "w") as f:
            f.write("\n".join(hosts) + "\n")
        os.rename(tmp, self.host_file)
        for h in leaving:
            if h in lost_tag:
                self._tag(lost_tag[h], self.REMOVAL_QUEUED, h)
                self.remove_queue[h] = lost_tag[h]
        for h in joining:
            self._tag(candidates[h], self.ADD_QUEUED, h)
            self.add_queue[h] = candidates[h]
        return hosts

    def run(self):
        if self.ec2 is None:
            try:
                import boto3
            except ImportError:
                sys.stderr.write("launch.py: boto3 is not available; EC2 host-file manager disabled\n")
                return
            self.ec2 = boto3.client("ec2", region_name=os.environ.get("AWS_REGION", "us-east-1"))
        while True:
            try:
                self.poll()
            except Exception as e:  # noqa: BLE001 -- the manager must outlive transient API errors
                sys.stderr.write("launch.py: host-file manager: %s\n" % e)
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
