#!/bin/bash
# last GPU sanity pass of round 2 (after the python-side parity work): smoke, the GPU tests that touch the executor /
# operators changed since the full run, a short bench
mkdir -p gpurun_out
L=gpurun_out/r2_last.log
: > $L
( timeout 90 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" ) >> $L 2>&1
( timeout 150 python -m pytest tests -m gpu -q -x --timeout 60 -k "executor or module or symbol or batchnorm or batch_norm or leaky or pooling or upsampl or roi or ctc or custom or sparse" 2>&1 | tail -8 ) >> $L 2>&1
( timeout 120 python bench.py --steps 20 --warmup 5 2>&1 | tail -2 ) >> $L 2>&1
tail -30 $L
