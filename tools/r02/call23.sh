#!/bin/bash
# r02 GPU pass 23: the headline bench line (configs[1], N = 1) with the sustained-copy probe; configs[0] line; launch list of the headline step
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench_full_v2.json 2> gpurun_out/r02_bench_full_v2.err; echo "bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_full_v2.json'))
print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline'].get('sustained_copy'),d['roofline'].get('frac_of_sustained_copy'))
print(d['e2e']); print(d['clocks']); print(d['parity']); print(d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"
tail -3 gpurun_out/r02_bench_full_v2.err
timeout 300 python bench.py --workload graphene_40nm_f32_dos --steps 20 --warmup 3 > gpurun_out/r02_bench_40nm_v3.json 2> gpurun_out/r02_bench_40nm_v3.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_40nm_v3.json'));print(d['value'],d['ms_per_step'],d['e2e']['seconds'],d['clocks'],d['parity']['parity_max_rel'])"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_full_v1.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > gpurun_out/r02_launches_full_v1.log 2>&1
tail -3 gpurun_out/r02_launches_full_v1.csv | cut -c1-200
