run() { echo "== $1 | $2"; env $1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','spp_per_s','ms_per_step','gpu_launches')}, 'trace_ms', d['roofline']['trace_ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['spp_per_s'])"; }
run "PTB_REFILL=0" ""
run "PTB_REFILL=4" ""
run "PTB_REFILL=8" ""
run "PTB_REFILL=16" ""
