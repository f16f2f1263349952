# N = 2 and 4 for the non-headline multi-GPU configs, on a 4-GPU box -> gpurun_out/r01b_<N>gpu_<name>.json
for w in hyperion_sphere_light volume_cube; do
  for n in 2 4; do
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 6 --warmup 3 --workload $w 2> gpurun_out/r01b_${n}gpu_$w.err | tail -1 > gpurun_out/r01b_${n}gpu_$w.json
    python -c "import json; d=json.load(open('gpurun_out/r01b_${n}gpu_$w.json')); print('$w', d['n_gpus'], 'GPUs', round(d['spp_per_s'],1), 'spp/s', round(d['value']), 'Mseg/s', round(d['ms_per_step'],2), 'ms/step', 'e2e', round(d['e2e']['spp_per_s'],1))" || tail -5 gpurun_out/r01b_${n}gpu_$w.err
  done
done
# the same two workloads on one GPU with the same code, for the efficiency column
for w in hyperion_sphere_light volume_cube; do
  python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r01b_1gpu_$w.json
  python -c "import json; d=json.load(open('gpurun_out/r01b_1gpu_$w.json')); print('$w', 1, 'GPU', round(d['spp_per_s'],1), 'spp/s')"
done
