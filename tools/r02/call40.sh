#!/bin/bash
# r02 GPU pass 40: the reference arm as the driver runs it (CPU port on the box's host cores); configs[4] line at HEAD
mkdir -p gpurun_out
( time timeout 800 python bench.py --impl reference --gpus 1 --steps 2 --warmup 3 > gpurun_out/r02_bench_reference_v1.json 2> gpurun_out/r02_bench_reference_v1.err ) 2>&1 | grep real
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_reference_v1.json'));print(d['impl'], d['value'], d['ms_per_step'], d['cpu_baseline']['cores'], d['cpu_baseline']['asymptotic_value'], d['config']['samples_run'])"
timeout 900 python bench.py --workload cubic_256_f32_dos --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_bench_cubic_v2.json 2> gpurun_out/r02_bench_cubic_v2.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_cubic_v2.json'));print('cubic', d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'], d['e2e']['set_model_seconds_each'], d['clocks']['sm_mhz'], d['parity']['parity_max_rel'])"
