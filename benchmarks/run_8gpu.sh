TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29561 bench.py --gpus 8 --steps 3 --warmup 3 2>&1 | grep '^{' > gpurun_out/bench_8gpu_fused.json
python -c "
import json; d=json.load(open('gpurun_out/bench_8gpu_fused.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('collectives'), d['spmv_GBps'], d['spmv_frac_of_peak']); print({k:(v['launches'],v['ms']) for k,v in d['kernels'].items()})"
timeout 200 $TR --master-port 29562 benchmarks/spmv_sweep.py --sizes 4096,8192,16384,32768 2>&1 | grep '^{' > gpurun_out/spmv_sweep_n8_fused.jsonl
cut -c1-260 gpurun_out/spmv_sweep_n8_fused.jsonl
