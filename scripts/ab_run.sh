#!/bin/bash
# Same-box A/B: runs bench.py --quick on each variant library given as arguments (names under glsl-pathtracer_b200/ab/), twice, alternating.
# usage (on the GPU box): scripts/ab_run.sh out.jsonl [--workload X] -- 000 111 ...
out=$1; shift
args=()
while [ "$1" != "--" ]; do args+=("$1"); shift; done; shift
for rep in 1 2; do for n in "$@"; do
  PTB200_LIB=$PWD/glsl-pathtracer_b200/ab/libptb200_$n.so python bench.py --quick --steps 6 "${args[@]}" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print(json.dumps({'variant':'$n','rep':$rep,'spp_per_s':round(d['spp_per_s'],1),'ms_per_step':round(d['ms_per_step'],3),**{a:round(b,3) for a,b in k.items()}}))" | tee -a $out
done; done
