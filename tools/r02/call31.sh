#!/bin/bash
# r02 GPU pass 31 (final check of HEAD): full GPU test suite, smoke, the headline line, ncu --set full of the headline step, configs[0] / [3] lines
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_v5.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_v5.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_v2.log 2>&1; tail -1 gpurun_out/r02_smoke_v2.log | cut -c1-200
timeout 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench_full_v3.json 2> gpurun_out/r02_bench_full_v3.err; echo "bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_full_v3.json'))
print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline'].get('sustained_copy',{}).get('gbs'))
print(d['e2e']['value'], d['e2e']['seconds_each'], d['e2e']['set_model_seconds_each']); print(d['clocks']); print(d['parity']['parity_max_rel'], d['cpu_baseline']['value'])"
timeout 300 python bench.py --workload graphene_40nm_f32_dos --steps 20 --warmup 3 > gpurun_out/r02_bench_40nm_v5.json 2> gpurun_out/r02_bench_40nm_v5.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_40nm_v5.json'));print('40nm', d['value'],d['ms_per_step'],d['e2e']['seconds'],d['clocks'],d['parity']['parity_max_rel'])"
timeout 600 python bench.py --workload graphene_200nm_f64_conductivity --steps 3 --warmup 2 > gpurun_out/r02_bench_graphene_200nm_f64_conductivity_v6.json 2> gpurun_out/r02_bench_graphene_200nm_f64_conductivity_v6.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_graphene_200nm_f64_conductivity_v6.json'));print('sigma', d['value'],d['ms_per_step'],d['config']['step_seconds'],d['roofline']['frac'],d['clocks'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cheb_step_bulk -s 8 -c 1 -f -o /tmp/step_bulk_r02 \
    python tools/step_sweep.py --workload graphene_1000nm_c64_dos --moments 34 --reps 0 PBK_TILE=256 > gpurun_out/r02_ncu_step_bulk_full_r64.log 2>&1
ncu -i /tmp/step_bulk_r02.ncu-rep --page raw --csv > gpurun_out/r02_ncu_step_bulk_full_r64_raw.csv 2>> gpurun_out/r02_ncu_step_bulk_full_r64.log
tail -2 gpurun_out/r02_ncu_step_bulk_full_r64.log | cut -c1-200
