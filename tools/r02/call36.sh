#!/bin/bash
# r02 GPU pass 36: device block cache (set_hamiltonian without cudaFree / cudaMalloc in steady state): tests, e2e iterations, LDOS after the finish_sums change
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02_pytest_gpu_v7.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_v7.log
PBK_TIMING=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu --no-parity > gpurun_out/r02_bench_timing_v5.json 2> gpurun_out/r02_host_timing_v5.log; echo "bench exit $?"
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_timing_v5.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['e2e']['seconds_each'],d['e2e']['set_model_seconds_each'])"; grep "set_hamiltonian" gpurun_out/r02_host_timing_v5.log | tail -3
timeout 600 python bench.py --workload graphene_500nm_c128_ldos --steps 2 --warmup 1 > gpurun_out/r02_bench_graphene_500nm_c128_ldos_v3.json 2>/dev/null
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_graphene_500nm_c128_ldos_v3.json'));print('ldos', d['ms_per_step'], d['config']['step_seconds'], d['roofline']['frac'], d['parity']['parity_max_rel'])"
timeout 600 python bench.py --workload graphene_500nm_c128_greens --steps 3 --warmup 1 > gpurun_out/r02_bench_graphene_500nm_c128_greens_v3.json 2>/dev/null
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_graphene_500nm_c128_greens_v3.json'));print('greens', d['ms_per_step'], d['config']['step_seconds'], d['roofline']['frac'], d['parity']['parity_max_rel'])"
timeout 600 compute-sanitizer --tool memcheck python tools/sanitizer_cases.py > gpurun_out/r02_sanitizer_memcheck_v3.log 2>&1; tail -2 gpurun_out/r02_sanitizer_memcheck_v3.log
