#!/bin/bash
# r02 GPU pass 42: branch-free two-window descriptor cache; 6 vs 5 prefetched halo rows per thread (12 bytes of spills vs late index loads)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_properties.py -m gpu -q -k "resident or window" 2>&1 | tail -2
echo "# HPT=6"; timeout 600 python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 64 --reps 2 PBK_RES=1 2>&1 | grep -v pbkpm | cut -c1-200
cd pybinding_b200/csrc && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function --diag-suppress 177 -DPBK_RES_HPT=5 -c kernels_res.cu -o build/kernels_res.o && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libpbkpm.so build/*.o -ldl && cd ../..
echo "# HPT=5"; timeout 600 python tools/step_sweep.py --workload cubic_256_f32_dos --moments 34 --vectors 64 --reps 2 PBK_RES=1 2>&1 | grep -v pbkpm | cut -c1-200
