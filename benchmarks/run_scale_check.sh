#!/bin/bash
# usage: run_scale_check.sh N : the driver's SCALE command at N GPUs (--steps 20 --warmup 5), parity object vs the committed N=1 run
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/bench_scale_n$N.err | grep '^{' > gpurun_out/bench_r2_16384_n${N}_k20w5.json
python -c "
import json; d=json.load(open('gpurun_out/bench_r2_16384_n${N}_k20w5.json')); print('N=$N', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['vs_n1'], d['spmv_frac_of_peak'], d['clocks'])"
