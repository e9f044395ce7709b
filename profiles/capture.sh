#!/bin/bash
# ncu evidence for one round (run under gpurun, 1 GPU).  Usage: bash profiles/capture.sh r01
# Every ncu run is preceded by the same command line without ncu (B200_PROFILING.md).
set -u
R=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --skip-cpu --profile"
$CMD > gpurun_out/${R}_plain.json 2> gpurun_out/${R}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches.csv $CMD \
    > gpurun_out/${R}_launches_run.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-secondary --profile"
$CMD2 > gpurun_out/${R}_plain2.json 2> gpurun_out/${R}_plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:pp_dubins_eval_kernel -s 3 -c 1 \
    -o gpurun_out/${R}_dubins_eval $CMD2 > gpurun_out/${R}_dubins_eval_run.log 2>&1
echo "dubins_eval capture rc=$?"
if [ "${2:-}" = "all" ]; then
  $CMD > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"pp_nn_bucketed_kernel|pp_nn_grid_kernel|pp_collide_segments_grid_kernel|pp_collide_segments_bucketed_kernel|pp_verify_polylines_kernel|pp_dubins_plan_kernel" \
      -c 24 -o gpurun_out/${R}_rrt $CMD > gpurun_out/${R}_rrt_run.log 2>&1
  echo "rrt capture rc=$?"
fi
