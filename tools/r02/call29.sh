#!/bin/bash
# r02 GPU pass 29: LDOS at configs[2] by the number of light-cone sub-systems advanced per launch (L2 residency vs launch count)
mkdir -p gpurun_out
: > gpurun_out/r02_ldos_group_sweep.log
for g in 0 1 2 4 8 16; do
  echo "# PBK_CONE_GROUP=$g" >> gpurun_out/r02_ldos_group_sweep.log
  PBK_CONE_GROUP=$g timeout 600 python bench.py --workload graphene_500nm_c128_ldos --steps 1 --warmup 1 > /tmp/l.json 2>/dev/null
  python -c "
import json;d=json.load(open('/tmp/l.json'));print(d['ms_per_step'], d['config']['moments_device_ms'], d['roofline']['achieved'], d['gpu_launches'], d['parity']['parity_max_rel'])" >> gpurun_out/r02_ldos_group_sweep.log
done
cat gpurun_out/r02_ldos_group_sweep.log
