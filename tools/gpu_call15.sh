#!/bin/bash
# GPU pass 15: compute-sanitizer over the new kernels, two-level locality ordering sweep, configs[2] LDOS with the specialised cone kernel.
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitizer_cases.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitizer_cases.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 python tools/sanitizer_cases.py > gpurun_out/sanitizer_synccheck.log 2>&1; echo "synccheck exit $?" >> gpurun_out/sanitizer_synccheck.log
timeout 900 python tools/step_sweep.py --workload graphene_1000nm_c64_dos --moments 66 --reps 1 \
  PBK_MACRO=0 PBK_MACRO=16 PBK_MACRO=64 PBK_MACRO=256 PBK_MACRO=1024 PBK_MACRO=64,MB=32 > gpurun_out/sweep_macro_full.log 2>&1
timeout 600 python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 128 --reps 1 \
  MB=64 MB=64,PBK_MACRO=16 MB=64,PBK_MACRO=64 MB=64,PBK_MACRO=256 > gpurun_out/sweep_macro_cubic.log 2>&1
timeout 600 python tools/config_bench.py ldos > gpurun_out/cfg_ldos.json 2> gpurun_out/cfg_ldos.err
tail -n 5 gpurun_out/sanitizer_memcheck.log; tail -n 5 gpurun_out/sanitizer_racecheck.log; tail -n 5 gpurun_out/sanitizer_synccheck.log
cat gpurun_out/sweep_macro_full.log gpurun_out/sweep_macro_cubic.log gpurun_out/cfg_ldos.json; tail -n 3 gpurun_out/cfg_ldos.err
