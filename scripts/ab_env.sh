#!/bin/bash
# Same-box A/B over run-time switches: scripts/ab_env.sh out.jsonl [bench args] -- "NAME=VAL ..." "NAME=VAL ..." (each quoted group is one variant)
out=$1; shift
args=()
while [ "$1" != "--" ]; do args+=("$1"); shift; done; shift
for rep in 1 2; do for v in "$@"; do
  env $v python bench.py --quick --steps 6 "${args[@]}" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print(json.dumps({'variant':'$v','rep':$rep,'spp_per_s':round(d['spp_per_s'],1),'ms_per_step':round(d['ms_per_step'],3),**{a:round(b,3) for a,b in k.items()}}))" | tee -a $out
done; done
