#!/bin/bash
L=iterative-solvers-summer-2020_b200/csrc/libjfnk.so
cp $L /tmp/base.so
for v in base t128 t256_3 t128_6; do
  if [ $v = base ]; then cp /tmp/base.so $L; else cp benchmarks/_scratch/libs/lib_$v.so $L; fi
  echo "== $v"
  timeout 200 python benchmarks/pma2_synthetic.py --steps 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['steps_per_s'], d['kernels']['mesh'])"
done
cp /tmp/base.so $L
