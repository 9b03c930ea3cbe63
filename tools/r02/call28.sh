#!/bin/bash
# r02 GPU pass 28: new tests (GEMM tile shapes, descriptor window); configs[0] timing variance (persistent kernel vs graph replay, with and without the clock sampler)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_properties.py -m gpu -q -k "kubo_gemm_tile_shapes or descriptor_window" 2>&1 | tail -4
for i in 1 2 3; do
  timeout 300 python bench.py --workload graphene_40nm_f32_dos --steps 50 --warmup 5 --no-cpu --no-e2e > gpurun_out/r02_bench_40nm_v4_$i.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/r02_bench_40nm_v4_$i.json'));print('persist run $i', d['ms_per_step'], d['roofline']['persistent_launches'], d['clocks']['sm_mhz'], d['clocks']['samples'])"
done
PBK_PERSIST=0 timeout 300 python bench.py --workload graphene_40nm_f32_dos --steps 50 --warmup 5 --no-cpu --no-e2e > gpurun_out/r02_bench_40nm_v4_graph.json 2>/dev/null
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_40nm_v4_graph.json'));print('graph replay', d['ms_per_step'], d['roofline']['persistent_launches'], d['clocks']['sm_mhz'])"
python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0, '.')
import pybinding_b200 as pb
m = pb.graphene_rectangle(40.0, dtype=np.float32)
k = pb.kpm(m, energy_range=(-8.5, 8.5), silent=True)
for rep in range(3):
    ts = []
    for _ in range(30):
        k.impl.moments_dos(1026, 1); ts.append(k.stats.moments_device_ms)
    print("no sampler: device ms per phase: median %.3f min %.3f max %.3f" % (np.median(ts), min(ts), max(ts)))
PY
