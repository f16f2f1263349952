# non-headline BASELINE configs on all 8 GPUs of one box (launched as the driver launches bench.py) -> gpurun_out/r01b_8gpu_<name>.json
for w in hyperion_sphere_light volume_cube instancing; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 6 --warmup 3 --workload $w 2> gpurun_out/r01b_8gpu_$w.err | tail -1 > gpurun_out/r01b_8gpu_$w.json
  python -c "import json; d=json.load(open('gpurun_out/r01b_8gpu_$w.json')); print('$w', d['n_gpus'], 'GPUs', round(d['spp_per_s'],1), 'spp/s', round(d['value']), 'Mseg/s', round(d['ms_per_step'],2), 'ms/step', 'e2e', round(d['e2e']['spp_per_s'],1))" || tail -5 gpurun_out/r01b_8gpu_$w.err
done
