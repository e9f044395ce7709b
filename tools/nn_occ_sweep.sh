#!/bin/bash
# developer sweep of the node grid's nodes-per-cell target (PP_NN_GRID_OCC, read when the grid is built).  Run under gpurun.
for occ in 4 3 2 1.5 1 0.7 0.5; do
  PP_NN_GRID_OCC=$occ python bench.py --skip-cpu --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); w=d['workloads']
print('occ $occ :', 'extend %.4f (nn %.4f collide %.4f)' % (w['extend']['ms_per_step'], w['extend']['nn_kernel_ms'], w['extend']['collide_kernel_ms']), 'fused %.4f' % w['extend_fused']['ms_per_step'], 'extend_dubins %.4f' % w['extend_dubins']['ms_per_step'], 'c5 %.4f' % w['dubins_rrt']['ms_per_step'], 'build %.3f' % w['extend'].get('nn_grid_build_ms_after_upload', 0))"
done
