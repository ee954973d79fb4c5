#!/bin/bash
# A/B of library builds on the bench workload: every .so under profiles/_build plus the in-tree library.
# usage (GPU box): bash profiles/ab.sh [extra bench.py arguments]
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
P=paper_romualdi_2022_icra_centroidal-mpc-walking_b200
for lib in $P/libcmpc_b200.so profiles/_build/*.so; do
  [ -f "$lib" ] || continue
  name=$(basename $lib .so)
  for b in 0 4144; do
    CMPC_B200_LIB=$PWD/$lib timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 8 --batch $b "$@" > gpurun_out/ab_${name}_$b.json 2> gpurun_out/ab_${name}_$b.err || tail -3 gpurun_out/ab_${name}_$b.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_${name}_$b.json").read().strip().split("\n")[-1])
    print("%-28s batch %5s: %7.0f solves/s  e2e %7.0f  kernel %.3f ms  iters %.2f  conv %s  lat %.3f ms" % ("$name", "$b" if "$b"!="0" else "1024", d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms_per_launch"], d["solve_stats"]["mean_iterations"], d["solve_stats"]["converged"], d["latency"]["p50_single_solve_ms"]))
except Exception as e: print("$name $b ERR", e)
PY
  done
done
