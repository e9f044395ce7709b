#!/bin/bash
# ncu evidence for one round (run under gpurun, 1 GPU).  Usage: bash profiles/capture.sh r02 [all]
# Every ncu run is preceded by the same command line without ncu (B200_PROFILING.md).
set -u
R=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --skip-cpu --profile"
$CMD > gpurun_out/${R}_plain.json 2> gpurun_out/${R}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${R}_launches.csv $CMD \
    > gpurun_out/${R}_launches_run.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-secondary --profile"
$CMD2 > gpurun_out/${R}_plain2.json 2> gpurun_out/${R}_plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:pp_dubins_eval_kernel -s 3 -c 1 \
    -o gpurun_out/${R}_dubins_eval $CMD2 > gpurun_out/${R}_dubins_eval_run.log 2>&1
echo "dubins_eval capture rc=$?"
if [ "${2:-}" = "all" ]; then
  $CMD > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"pp_rrt_extend_fused_kernel|pp_extend_bin_kernel|pp_extend_scatter_kernel|pp_nn_grid_kernel|pp_collide_segments_grid_kernel|pp_verify_polylines_kernel|pp_dubins_plan_kernel|pp_dubins_fill_kernel" \
      -c 28 -o gpurun_out/${R}_rrt $CMD > gpurun_out/${R}_rrt_run.log 2>&1
  echo "rrt capture rc=$?"
fi
for f in gpurun_out/${R}_dubins_eval.ncu-rep gpurun_out/${R}_rrt.ncu-rep; do
  [ -f "$f" ] && ncu -i "$f" --page raw --csv > "${f%.ncu-rep}_raw.csv" 2>/dev/null
done
ls -la gpurun_out | tail -20
