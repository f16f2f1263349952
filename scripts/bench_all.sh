# every BASELINE config on one GPU -> gpurun_out/r01_cfg_<name>.json
for w in hyperion_rect_lights hyperion_sphere_light cornell_box_orig ibl_spheres volume_cube instancing; do
  python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/r01_cfg_$w.json
  python -c "import json; d=json.load(open('gpurun_out/r01_cfg_$w.json')); print('$w', round(d['spp_per_s'],1), 'spp/s', round(d['value']), 'Mseg/s', round(d['mrays_per_s']), 'Mrays/s', round(d['ms_per_step'],2), 'ms/step', 'frac', d['roofline']['frac'])" || tail -3 gpurun_out/r01_cfg_$w.json
done
