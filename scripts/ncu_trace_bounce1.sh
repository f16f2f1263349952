# ncu metrics of the bounce-1 k_trace launch (2nd k_trace of a wave) of one 8-pass 1080p hyperion wave; $1 = output tag
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_issued.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum
ncu --clock-control none -k regex:k_trace --launch-skip 1 --launch-count 1 --metrics $M --csv --log-file gpurun_out/ncu_trace1_$1.csv \
  python -c "
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import glsl_pathtracer_b200
from glsl_pathtracer_b200 import capi
from conftest import scene_at
ctx = capi.Context(scene_at('hyperion_rect_lights', 1920, 1080)); ctx.render_samples(1, 8); ctx.synchronize(); ctx.close()" > /dev/null 2>&1
python - <<PY
import csv
rows = [r for r in csv.reader(open('gpurun_out/ncu_trace1_$1.csv')) if len(r) > 10]
h = rows[0]; i_n, i_v = h.index('Metric Name'), h.index('Metric Value')
print('$1', {r[i_n]: r[i_v] for r in rows[1:]})
PY
