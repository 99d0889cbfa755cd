#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "512-0-1" 2>&1 | tail -5
echo "== bench P2P=1"
JFNK_P2P=1 timeout 500 $TR --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 2>&1 | grep '^{' > gpurun_out/bench_8gpu_p2p1.json
echo "== bench P2P=0"
JFNK_P2P=0 timeout 400 $TR --master-port 29522 bench.py --gpus 8 --steps 3 --warmup 2 --no-e2e 2>&1 | grep '^{' > gpurun_out/bench_8gpu_p2p0.json
for p in 1 0; do python -c "
import json; d=json.load(open('gpurun_out/bench_8gpu_p2p$p.json')); print(d['value'], d['ms_per_step'], d['config'].get('collectives'), d['spmv_GBps'], d['spmv_frac_of_peak']); print({k:(v['launches'],v['ms']) for k,v in d['kernels'].items()})"; done
echo "== sweep P2P=1"
JFNK_P2P=1 timeout 300 $TR --master-port 29523 benchmarks/spmv_sweep.py --sizes 4096 8192 16384 32768 2>&1 | grep '^{' > gpurun_out/spmv_sweep_n8_p2p1.jsonl
cat gpurun_out/spmv_sweep_n8_p2p1.jsonl | cut -c1-400
