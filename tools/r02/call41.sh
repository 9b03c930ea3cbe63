#!/bin/bash
# r02 GPU pass 41: resident-tile kernel with six prefetched halo rows per thread again (window base derived from the loop index): tests, sweep, configs[4] line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_properties.py tests/test_gpu_baseline_configs.py -m gpu -q -k "resident or cubic or config4 or window" 2>&1 | tail -2
timeout 600 python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 64 --reps 1 PBK_RES=1 PBK_RES=1,PBK_RES_STAGES=2 2>&1 | grep -v pbkpm | cut -c1-230
timeout 900 python bench.py --workload cubic_256_f32_dos --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_bench_cubic_v3.json 2> gpurun_out/r02_bench_cubic_v3.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_cubic_v3.json'));print('cubic', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks']['sm_mhz'], d['parity']['parity_max_rel'])"
