#!/bin/bash
# final_measure.sh without the compute-sanitizer passes (TAG = $1)
tag=${1:-r03}
python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/${tag}_pytest_gpu.txt
python bench.py --impl reference > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
python bench.py > gpurun_out/${tag}_bench_1gpu.json 2> gpurun_out/${tag}_bench_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --quick --steps 2 --warmup 3 > gpurun_out/${tag}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on --launch-skip 54 --launch-count 18 -o gpurun_out/${tag}_full -f python bench.py --quick --steps 1 --warmup 3 > gpurun_out/${tag}_ncu_full.log 2>&1
tail -3 gpurun_out/${tag}_pytest_gpu.txt; tail -c 300 gpurun_out/${tag}_bench_1gpu.json; ls -la gpurun_out/${tag}_full.ncu-rep
