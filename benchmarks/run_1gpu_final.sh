#!/bin/bash
# inside a 1-GPU gpurun call: the whole -m gpu suite, smoke, the headline bench and the other configurations' records
TAG=$1
timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu_${TAG}_full.log 2>&1; tail -3 gpurun_out/pytest_gpu_${TAG}_full.log; grep -E "passed|failed" gpurun_out/pytest_gpu_${TAG}_full.log > gpurun_out/pytest_gpu_${TAG}.log
grep DROPLET_SUMMARY gpurun_out/pytest_gpu_${TAG}_full.log | grep '"backend": "cuda"' | sed 's/^DROPLET_SUMMARY //' > gpurun_out/droplet_100steps_${TAG}.json
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 300 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_k3w3.json 2> gpurun_out/bench_${TAG}.err; cut -c1-200 gpurun_out/bench_${TAG}_k3w3.json
timeout 120 python benchmarks/config1_small_grid.py 2>/dev/null | grep '^{' > gpurun_out/config1_${TAG}.jsonl; JFNK_CYCLE_PROF=1 timeout 120 python benchmarks/config1_small_grid.py 2>&1 | grep "cycle prof" >> gpurun_out/config1_${TAG}.jsonl; cut -c1-200 gpurun_out/config1_${TAG}.jsonl
timeout 120 python benchmarks/droplet_step_micro.py 30 2>/dev/null | grep '^{' > gpurun_out/droplet_micro_${TAG}.jsonl; cut -c1-300 gpurun_out/droplet_micro_${TAG}.jsonl
timeout 300 python benchmarks/reference_sizes.py 20 2>/dev/null | grep '^{' > gpurun_out/reference_sizes_${TAG}.jsonl; cut -c1-160 gpurun_out/reference_sizes_${TAG}.jsonl
timeout 300 python benchmarks/pma2_synthetic.py 2>/dev/null | grep '^{' > gpurun_out/pma2_${TAG}.jsonl; tail -1 gpurun_out/pma2_${TAG}.jsonl | cut -c1-300
