#!/bin/bash
# inside a 2-GPU gpurun call: the rank-count-independent sums -- 1-GPU fingerprint, then the 2-rank run compared with it
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --steps 3 --warmup 3 --write-fingerprint --no-cpu-baseline > gpurun_out/bench_det_n1.json 2> gpurun_out/bench_det.err
python -c "
import json; d=json.load(open('gpurun_out/bench_det_n1.json')); print('N=1', d['value'], d['e2e']['value'], {k:(v['launches'],round(v['GBps'])) for k,v in d['kernels'].items()})"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --steps 3 --warmup 3 2>> gpurun_out/bench_det.err | grep '^{' > gpurun_out/bench_det_n2.json
python -c "
import json; d=json.load(open('gpurun_out/bench_det_n2.json')); print('N=2', d['value'], d['e2e']['value'], d['parity']['vs_n1'])"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -5
