#!/bin/bash
# GPU pass 24: A/B of the ordering variants inside ONE call (same box, same thermal state): full-length moments phases.
mkdir -p gpurun_out
for cfg in "PBK_MACRO=0" "PBK_MACRO=256 PBK_COARSE=1" "PBK_MACRO=256 PBK_COARSE=16" "PBK_MACRO=0"; do
  env $cfg timeout 400 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e 2>> gpurun_out/ab_order.err | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$cfg', 'value %.4e' % d['value'], 'ms_per_step %.1f' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'], 'launch_ms %.3f' % d['roofline']['launch_ms'], 'sm_mhz', d['clocks']['sm_mhz'])" >> gpurun_out/ab_order.log
done
cat gpurun_out/ab_order.log; tail -n 3 gpurun_out/ab_order.err
