#!/bin/bash
# r02 GPU pass 43: prefetched halo rows per thread 5 / 4 / 3 with the two-window descriptor cache
for h in 5 4 3; do
  cd pybinding_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function --diag-suppress 177 -DPBK_RES_HPT=$h -c kernels_res.cu -o build/kernels_res.o && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libpbkpm.so build/*.o -ldl && cd ../..
  echo "# HPT=$h"; timeout 600 python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 64 --reps 2 PBK_RES=1 PBK_RES=1,PBK_RES_TILE=320 2>&1 | grep -v "pbkpm\|# model" | cut -c1-130
done
