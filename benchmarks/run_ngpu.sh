#!/bin/bash
# usage: benchmarks/run_ngpu.sh N TAG   (inside a gpurun --gpus N call): multi-rank tests, the headline bench and the SpMV sweep
N=$1; TAG=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi_${TAG}.log 2>&1; tail -3 gpurun_out/pytest_multi_${TAG}.log; fi
timeout 400 $TR --master-port 29561 bench.py --gpus $N --steps 3 --warmup 3 2> gpurun_out/bench_${TAG}.err | grep '^{' > gpurun_out/bench_${TAG}.json
python -c "
import json; d=json.load(open('gpurun_out/bench_${TAG}.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'] if d.get('e2e') else None, d['config'].get('collectives'), d['spmv_frac_of_peak'], d.get('parity',{}).get('vs_n1')); print({k:(v['launches'],round(v['ms'],1)) for k,v in d['kernels'].items()})"
timeout 300 $TR --master-port 29562 benchmarks/spmv_sweep.py 2>&1 | grep '^{' > gpurun_out/spmv_sweep_${TAG}.jsonl
cut -c1-260 gpurun_out/spmv_sweep_${TAG}.jsonl
