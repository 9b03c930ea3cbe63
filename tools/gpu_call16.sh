#!/bin/bash
# GPU pass 16: full GPU suite + smoke + full bench + launch list + one --set full capture of the step kernel at HEAD (two-level ordering on).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_full.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cheb_step_bulk -s 8 -c 1 -f -o gpurun_out/step_bulk_v5 \
  python tools/step_sweep.py --workload graphene_1000nm_c64_dos --moments 34 --reps 0 PBK_MACRO=256 > gpurun_out/ncu_step.log 2>&1
ncu -i gpurun_out/step_bulk_v5.ncu-rep --page raw --csv > gpurun_out/step_bulk_v5_raw.csv 2>/dev/null
tail -n 5 gpurun_out/pytest_gpu.log; tail -n 2 gpurun_out/smoke.log; cat gpurun_out/bench_full.json; tail -n 3 gpurun_out/bench_full.err; tail -n 2 gpurun_out/launches_full.csv; tail -n 3 gpurun_out/ncu_step.log
