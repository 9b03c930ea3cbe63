#!/bin/bash
# GPU pass 2 (round 1, v2): FP64 MMA probe, parity tests, headline bench, ncu launch list + full capture of the staged kernel.
mkdir -p gpurun_out
timeout 120 tools/probe/build/dmma_probe > gpurun_out/dmma_probe.jsonl 2>&1
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 1200 python bench.py --steps 2 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_full.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:cheb_step_bulk -s 8 -c 1 -f -o gpurun_out/step_bulk_full_r64 \
  python tools/step_sweep.py --workload graphene_1000nm_c64_dos --moments 34 --reps 0 PBK_BULK=4 > gpurun_out/ncu_full.log 2>&1
timeout 600 python bench.py --workload graphene_40nm_f32_dos --steps 3 --warmup 3 > gpurun_out/bench_40nm.json 2> gpurun_out/bench_40nm.err
timeout 900 python bench.py --workload cubic_256_f32_dos --steps 1 --warmup 3 --no-e2e > gpurun_out/bench_cubic.json 2> gpurun_out/bench_cubic.err
cat gpurun_out/dmma_probe.jsonl; tail -12 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench_full.json; tail -n 3 gpurun_out/bench_full.err; cat gpurun_out/bench_40nm.json gpurun_out/bench_cubic.json; tail -n 2 gpurun_out/ncu_full.log; ls -la gpurun_out
