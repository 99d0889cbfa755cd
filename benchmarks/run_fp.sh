#!/bin/bash
# 1-GPU: the single-GPU fingerprints (driver's argument sets) for the multi-GPU parity object
timeout 300 python bench.py --steps 3 --warmup 3 --write-fingerprint > gpurun_out/bench_fp_k3w3.json 2> gpurun_out/bench_fp.err; cut -c1-200 gpurun_out/bench_fp_k3w3.json
timeout 400 python bench.py --steps 20 --warmup 5 --write-fingerprint --no-cpu-baseline > gpurun_out/bench_fp_k20w5.json 2>> gpurun_out/bench_fp.err; cut -c1-200 gpurun_out/bench_fp_k20w5.json
cp profiles/sh16384_fingerprint_r2.json gpurun_out/sh16384_fingerprint_r2.json
python -c "
import json
for f in ('gpurun_out/bench_fp_k3w3.json','gpurun_out/bench_fp_k20w5.json'):
    d=json.load(open(f)); print(d['value'], d['e2e']['value'], d['clocks']['sm_mhz'], d['roofline']['kernel'], round(d['roofline']['frac'],3), d['spmv_frac_of_peak'])"
