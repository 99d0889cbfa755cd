#!/bin/bash
# second ncu pass of round 2: the marching kernel with bulk-TMA operand rows, the multi-dot with 8 virtual blocks
NCU="ncu --set full --clock-control none --import-source on"
python benchmarks/pma2_jvp_micro.py 2048 && $NCU -k regex:mesh_march_kernel --launch-skip 8 -c 2 -f -o gpurun_out/prof_marchtma2048_r2 python benchmarks/pma2_jvp_micro.py 2048 > gpurun_out/ncu_marchtma.log 2>&1
python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/plain_b2.log 2>&1 && $NCU -k regex:mdot_kernel --launch-skip 40 -c 1 -f -o gpurun_out/prof_mdotdet16384_r2 python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/ncu_mdotdet.log 2>&1
ls -la gpurun_out/prof_marchtma2048_r2.ncu-rep gpurun_out/prof_mdotdet16384_r2.ncu-rep
