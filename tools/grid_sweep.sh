#!/bin/bash
# developer sweep of the obstacle-grid cell size (PP_GRID_CELL_SCALE / PP_GRID_EXT_SCALE, read at obstacle upload):
# one short bench run per setting, the grid-dependent workloads side by side.  Run under gpurun.
for cfg in "1.0 0.75" "0.7 0.75" "0.5 0.75" "0.5 0.5" "0.35 0.5" "0.35 0.35" "0.25 0.25"; do
  set -- $cfg
  PP_GRID_CELL_SCALE=$1 PP_GRID_EXT_SCALE=$2 python bench.py --skip-cpu --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); w=d['workloads']
print('cell $1 ext $2 :', 'extend %.4f (collide %.4f)' % (w['extend']['ms_per_step'], w['extend']['collide_kernel_ms']), 'fused %.4f' % w['extend_fused']['ms_per_step'], 'extend_dubins %.4f' % w['extend_dubins']['ms_per_step'], 'c5 %.4f' % w['dubins_rrt']['ms_per_step'], 'nohit %.4f' % w['dubins_rrt_nohit']['ms_per_step'], 'grid_nohit %.4f' % w['collide_grid_nohit']['collide_kernel_ms'])"
done
