#!/usr/bin/env python
"""Environment report for bug reports (reference: tools/diagnose.py): platform, python, torch/CUDA, GPUs, native
libraries of this framework and the MXNET_* / DMLC_* environment."""
import os
import platform
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":
    print("----------Python Info----------")
    print("Version      :", platform.python_version(), "| Compiler:", platform.python_compiler())
    print("----------System Info----------")
    print("Platform     :", platform.platform(), "| machine:", platform.machine(), "| cpus:", os.cpu_count())
    print("----------Framework Info----------")
    import torch
    print("torch        :", torch.__version__, "| CUDA:", torch.version.cuda, "| available:", torch.cuda.is_available())
    import mxb200 as mx
    from mxb200 import build as B
    print("mxb200       :", getattr(mx, "__version__", "?"), "at", os.path.dirname(mx.__file__))
    for kind in ("host", "cuda"):
        p = B.lib_path(kind)
        print("  lib %-5s: %s  exists=%s fresh=%s" % (kind, p, os.path.exists(p), B.is_fresh(kind) if os.path.exists(p) else False))
    if torch.cuda.is_available():
        for i in range(torch.cuda.device_count()):
            pr = torch.cuda.get_device_properties(i)
            print("  GPU %d      : %s sm_%d%d %d SMs %.0f GB" % (i, pr.name, pr.major, pr.minor, pr.multi_processor_count, pr.total_memory / 2**30))
        try:
            print(subprocess.run(["nvidia-smi", "topo", "-m"], stdout=subprocess.PIPE, text=True, timeout=20).stdout)
        except Exception as e:  # noqa: BLE001
            print("nvidia-smi topo failed:", e)
    print("----------Environment----------")
    for k in sorted(os.environ):
        if k.startswith(("MXNET_", "MXB200_", "DMLC_", "PS_", "NCCL_", "CUDA_", "OMP_")) or k in ("WORKER_HOST_FILE", "ELASTIC_TRAINING_ENABLED"):
            print("%s=%s" % (k, os.environ[k]))
