#!/bin/bash
# inside a 1-GPU gpurun call: the whole -m gpu suite, smoke, the headline bench with the fingerprints the multi-GPU parity
# object compares against, the single-GPU SpMV sweep and the small-grid configurations
TAG=$1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1; tail -4 gpurun_out/pytest_gpu_${TAG}.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 300 python bench.py --steps 3 --warmup 3 --write-fingerprint > gpurun_out/bench_${TAG}_k3w3.json 2> gpurun_out/bench_${TAG}.err; cut -c1-300 gpurun_out/bench_${TAG}_k3w3.json
timeout 400 python bench.py --steps 20 --warmup 5 --write-fingerprint --no-cpu-baseline > gpurun_out/bench_${TAG}_k20w5.json 2>> gpurun_out/bench_${TAG}.err; cut -c1-300 gpurun_out/bench_${TAG}_k20w5.json
cp profiles/sh16384_fingerprint_r2.json gpurun_out/sh16384_fingerprint_r2.json
timeout 200 python benchmarks/spmv_sweep.py 2>&1 | grep '^{' > gpurun_out/spmv_sweep_${TAG}_n1.jsonl; cut -c1-200 gpurun_out/spmv_sweep_${TAG}_n1.jsonl
timeout 120 python benchmarks/config1_small_grid.py 2>/dev/null | grep '^{' > gpurun_out/config1_${TAG}.jsonl; cut -c1-330 gpurun_out/config1_${TAG}.jsonl
timeout 120 python benchmarks/droplet_step_micro.py 30 2>/dev/null | grep '^{' > gpurun_out/droplet_micro_${TAG}.jsonl; cut -c1-400 gpurun_out/droplet_micro_${TAG}.jsonl
