# A/B runs of bench.py under different tuning knobs (one GPU call)
for cfg in "PTB_SORT=1" "PTB_SORT=2"; do
  echo "== $cfg"
  env $cfg python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','spp_per_s','ms_per_step','gpu_launches')}, 'trace_ms', d['roofline']['trace_ms_per_step'], 'e2e', d['e2e']['spp_per_s'])"
done
