#!/bin/bash
# ncu captures of round 2 (inside a 1-GPU gpurun call); every target has exited 0 without ncu in the same call first
NCU="ncu --set full --clock-control none --import-source on"
python benchmarks/ncu_targets.py sh && $NCU -k regex:sh_cycle_kernel --launch-skip 6 -c 1 -f -o gpurun_out/prof_shcycle64_r2 python benchmarks/ncu_targets.py sh > gpurun_out/ncu_shcycle.log 2>&1
python benchmarks/ncu_targets.py droplet && $NCU -k regex:mesh_cycle_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_meshcycle91x61_r2 python benchmarks/ncu_targets.py droplet > gpurun_out/ncu_meshcycle.log 2>&1
$NCU -k regex:pma_relax_band_kernel --launch-skip 1 -c 1 -f -o gpurun_out/prof_relaxband91x61_r2 python benchmarks/ncu_targets.py droplet > gpurun_out/ncu_relaxband.log 2>&1
# launch list of one headline step (per-launch times are cold-cache and serialised: shares only)
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r2.csv | tail -8
