#!/bin/bash
# run on the GPU box: bench every variant library (kernel-only numbers)
cd $GRAFT_REPO_ROOT
cp systemlevelcontrol.jl_b200/lib/libslsb200.so /tmp/orig.so
for v in scratch/variants/lib*.so; do
  cp $v systemlevelcontrol.jl_b200/lib/libslsb200.so
  echo "== $v"
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e $AB_ARGS 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['kernel_ms'])"
done
cp /tmp/orig.so systemlevelcontrol.jl_b200/lib/libslsb200.so
