#!/bin/bash
# One box: full GPU suite, both bench arms, ncu launch list + full capture of one step, compute-sanitizer.  Outputs under gpurun_out/ (TAG = $1, default r02).
tag=${1:-r03}
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/${tag}_pytest_gpu.txt
python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
python bench.py > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 54 --launch-count 18 -o gpurun_out/${tag}_full -f python bench.py --quick --steps 1 --warmup 3 > gpurun_out/${tag}_ncu_full.log 2>&1
(compute-sanitizer --tool memcheck python scripts/sanitize_general.py 2>&1 | tail -14) > gpurun_out/${tag}_sanitizer_memcheck_general.txt
(compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4) > gpurun_out/${tag}_sanitizer_memcheck_smoke.txt
tail -3 gpurun_out/${tag}_pytest_gpu.txt; tail -c 600 gpurun_out/${tag}_bench_1gpu.json; tail -2 gpurun_out/${tag}_sanitizer_memcheck_general.txt; ls -la gpurun_out/${tag}_full.ncu-rep
