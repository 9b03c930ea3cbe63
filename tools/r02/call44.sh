#!/bin/bash
# r02 GPU pass 44 (final check of HEAD): full GPU suite, smoke, the headline line, configs[4] and configs[0] lines
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_v8.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_v8.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_v3.log 2>&1; tail -1 gpurun_out/r02_smoke_v3.log | cut -c1-160
timeout 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/r02_bench_full_v4.json 2> gpurun_out/r02_bench_full_v4.err; echo "bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_full_v4.json'))
print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['roofline']['traffic'],d['roofline'].get('sustained_copy',{}).get('gbs'))
print(d['e2e']['value'], d['e2e']['seconds_each'], d['e2e']['set_model_seconds_each']); print(d['clocks']); print(d['parity']['parity_max_rel'], d['cpu_baseline']['value'])"
timeout 900 python bench.py --workload cubic_256_f32_dos --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_bench_cubic_v4.json 2> gpurun_out/r02_bench_cubic_v4.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_cubic_v4.json'));print('cubic', d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks']['sm_mhz'], d['parity']['parity_max_rel'])"
timeout 300 python bench.py --workload graphene_40nm_f32_dos --steps 20 --warmup 3 > gpurun_out/r02_bench_40nm_v7.json 2>/dev/null
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_40nm_v7.json'));print('40nm', d['value'],d['ms_per_step'],d['e2e']['seconds'])"
