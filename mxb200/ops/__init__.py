"""Kernel launchers (python side of csrc/cuda/*)."""
from . import functional, optim  # noqa: F401


REPLAYED = [0]     # kernels of ours executed through CUDA-graph replays (parallel/graph.py)


def launch_count():
    """Number of our own kernels launched so far in this process (bench.py's gpu_launches)."""
    from . import functional as f, optim as o
    n = f.LAUNCHES[0] + o.LAUNCHES[0] + REPLAYED[0]
    try:
        from ..parallel import fused as pf
        n += pf.LAUNCHES[0]
    except Exception:  # noqa: BLE001
        pass
    return n
