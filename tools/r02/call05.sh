#!/bin/bash
# r02 GPU pass 5: which knob removes the run-to-run differences of the few-lane staged kernel at 38 M sites
mkdir -p gpurun_out
timeout 1500 python tools/r02/diag_r4b.py > gpurun_out/r02_diag_r4b.log 2>&1; echo "diag exit $?"
cat gpurun_out/r02_diag_r4b.log
timeout 300 python -m pytest tests/test_parallel.py -m gpu -q 2>&1 | tail -15
