#!/bin/bash
# r02 GPU pass 33: warp-shuffle fast path of finish_sums -- full GPU suite, mid-size step timing, configs[0] line with the bracketed clocks
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_v6.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_v6.log
timeout 600 python tools/step_sweep.py --workload graphene_200nm_c64_dos --moments 258 --reps 2 PBK_TILE=256 MB=8 MB=16 2>&1 | grep -v "pbkpm" | cut -c1-200
timeout 300 python bench.py --workload graphene_40nm_f32_dos --steps 20 --warmup 3 > gpurun_out/r02_bench_40nm_v6.json 2> gpurun_out/r02_bench_40nm_v6.err
python -c "
import json;d=json.load(open('gpurun_out/r02_bench_40nm_v6.json'));print('40nm', d['value'],d['ms_per_step'],d['config']['step_seconds'][:6],d['e2e']['seconds'],d['clocks'])"
PBK_PERSIST=0 timeout 300 python bench.py --workload graphene_40nm_f32_dos --steps 20 --warmup 3 --no-cpu --no-e2e > /tmp/g.json 2>/dev/null; python -c "
import json;d=json.load(open('/tmp/g.json'));print('40nm graph replay', d['ms_per_step'])"
timeout 600 python tools/sanitizer_cases.py > /dev/null 2>&1; timeout 900 compute-sanitizer --tool racecheck python tools/sanitizer_cases.py > gpurun_out/r02_sanitizer_racecheck_v3.log 2>&1; tail -3 gpurun_out/r02_sanitizer_racecheck_v3.log
