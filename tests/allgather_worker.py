"""Worker for test_scheduler_allgather: every worker contributes a blob, all get the list in rank order."""
import faulthandler
import os
import sys

faulthandler.dump_traceback_later(90, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mxb200.parallel.membership import ElasticGroup  # noqa: E402
from mxb200.parallel.peer import pslite_exchange  # noqa: E402

eg = ElasticGroup()
ex = pslite_exchange(eg.node)
for round_ in range(3):
    got = ex(bytes([eg.rank, round_, 0, 255, 10]))      # includes NUL / newline bytes: the transport hex-encodes
    assert len(got) == eg.world, got
    assert [b[0] for b in got] == list(range(eg.world)) and all(b[1:] == bytes([round_, 0, 255, 10]) for b in got), got
eg.close()
print("OK", flush=True)
