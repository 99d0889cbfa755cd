#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -15
for p in 1 0; do
  echo "== JFNK_P2P=$p"
  JFNK_P2P=$p timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --no-e2e 2>&1 | grep '^{' > gpurun_out/bench_2gpu_p2p$p.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_2gpu_p2p$p.json')); print(d['value'], d['ms_per_step'], d['config'].get('collectives'), d['spmv_GBps'], d['spmv_frac_of_peak']); print({k:(v['launches'],v['ms']) for k,v in d['kernels'].items()})"
done
