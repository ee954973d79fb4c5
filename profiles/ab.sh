#!/bin/bash
# A/B of library builds under profiles/_build: usage profiles/ab.sh <variant> ... ("base" = the in-tree library); prints solves/s
for v in "$@"; do
  if [ "$v" = base ]; then unset CMPC_B200_LIB; else export CMPC_B200_LIB=$PWD/profiles/_build/libcmpc_$v.so; fi
  for args in "" "--batch 4144" "--workload ergocub_b65536 --batch 8288"; do
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline $args 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', '$args', round(d['value']), d['config']['converged'], round(d['config']['mean_iterations'],2), 'lat', round(d['latency']['p50_single_solve_ms'],2))"
  done
done
