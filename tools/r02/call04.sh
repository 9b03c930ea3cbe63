#!/bin/bash
# r02 GPU pass 4: diagnostic of the few-lane DOS moments at full size; sweep goldens; e2e timing split
mkdir -p gpurun_out
timeout 1200 python tools/r02/diag_r4.py 300 1000 > gpurun_out/r02_diag_r4.log 2>&1; echo "diag exit $?"
cat gpurun_out/r02_diag_r4.log | cut -c1-400
timeout 300 python -m pytest tests/test_parallel.py -m gpu -q 2>&1 | tail -5
