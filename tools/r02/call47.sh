#!/bin/bash
# r02 GPU pass 47: is the ~16 GB/s per SM of the GEMM's operand copies a per-thread / per-warp or a per-CTA limit?  Copies issued by lane 0 of every active warp
mkdir -p gpurun_out
PBK_KUBO_ISSUERS=8 timeout 600 python -m pytest tests/test_gpu_properties.py -m gpu -q -k "kubo_gemm_tile_shapes" 2>&1 | tail -1
: > gpurun_out/r02_kubo_gemm_issuers.log
for cfg in "PBK_KUBO_ISSUERS=1 --moments 514" "PBK_KUBO_ISSUERS=8 --moments 514" "PBK_KUBO_ISSUERS=1 --moments 520" "PBK_KUBO_ISSUERS=8 --moments 520" "PBK_KUBO_ISSUERS=8 --moments 512" "PBK_KUBO_ISSUERS=8 --moments 600"; do
  echo "# $cfg" >> gpurun_out/r02_kubo_gemm_issuers.log
  env ${cfg%% *} timeout 300 python tools/kubo_bench.py --reps 1 --vectors 2 ${cfg#* } >> gpurun_out/r02_kubo_gemm_issuers.log 2>&1
done
cut -c1-110 gpurun_out/r02_kubo_gemm_issuers.log
