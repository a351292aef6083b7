#!/bin/bash
# A/B timing of library variants on the same GPU box: [CONFIG=c3] tools/ab_bench.sh ab/libA.so ab/libB.so ...
# (each variant is copied over the in-tree library and bench.py is run; variants are alternated twice)
set -e
orig=gr-dvbt2ll_b200/libdvbt2ll_cuda.so
cp $orig /tmp/lib_orig.so
for round in 1 2; do
  for v in "$@"; do
    cp $v $orig
    python bench.py --config ${CONFIG:-c3} --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['roofline']['stage_ms']
print('$v', round(d['ms_per_step'],4), {k:round(x,4) for k,x in s.items()})"
  done
done
cp /tmp/lib_orig.so $orig
