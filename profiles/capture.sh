#!/bin/bash
# ncu evidence for one round (run under gpurun, 1 GPU).  Usage: bash profiles/capture.sh r02 [all]
# Every ncu run is preceded by the same command line without ncu (B200_PROFILING.md).  The reports are exported to
# CSV on the box (raw page + CUDA/SASS source page) and large .ncu-rep files are deleted: gpurun copies back <= 64 MiB.
set -u
R=${1:-r02}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --skip-cpu --profile"
$CMD > gpurun_out/${R}_plain.json 2> gpurun_out/${R}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${R}_launches.csv $CMD \
    > gpurun_out/${R}_launches_run.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --steps 1 --warmup 3 --skip-cpu --skip-secondary --profile"
$CMD2 > gpurun_out/${R}_plain2.json 2> gpurun_out/${R}_plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:pp_dubins_eval_kernel -s 3 -c 1 \
    -o gpurun_out/${R}_dubins_eval $CMD2 > gpurun_out/${R}_dubins_eval_run.log 2>&1
echo "dubins_eval capture rc=$?"
cap() {  # name, kernel regex, launch-skip, launch-count
  ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o gpurun_out/${R}_$1 $CMD \
      > gpurun_out/${R}_$1_run.log 2>&1
  echo "$1 capture rc=$?"
  ncu -i gpurun_out/${R}_$1.ncu-rep --page raw --csv > gpurun_out/${R}_$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/${R}_$1.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${R}_$1_source.csv 2>/dev/null
}
if [ "${2:-}" = "fillplan" ]; then  # late round 2: only the two kernels that changed after the r02b captures
  $CMD > /dev/null 2>&1
  cap fill "pp_dubins_fill_kernel" 0 1
  cap plan "pp_dubins_plan_kernel" 5 1
fi
if [ "${2:-}" = "verifyplan" ]; then  # the kernels touched by the plan-side path test (final build of round 2)
  $CMD > /dev/null 2>&1
  cap verify_c5 "pp_verify_polylines_kernel" 5 1
  cap verify_nohit "pp_verify_polylines_kernel" 7 1
  cap verify_extend "pp_verify_polylines_kernel" 3 1
  cap plan "pp_dubins_plan_kernel" 5 1
  cap plan_extend "pp_dubins_plan_kernel" 3 1
fi
if [ "${2:-}" = "all" ]; then
  $CMD > /dev/null 2>&1
  # the extend step: default pair of launches (first instances), then the binned fused kernel and its binning
  cap extend "pp_nn_grid_kernel|pp_collide_segments_grid_kernel|pp_rrt_extend_fused_kernel|pp_extend_bin_kernel|pp_extend_scan_kernel|pp_extend_scatter_kernel" 0 8
  # verify kernel: instances in bench order are strong-C5 x3, extend_dubins x2, C5 slice x2, C5 no-hit x2
  cap verify_c5 "pp_verify_polylines_kernel" 5 1
  cap verify_nohit "pp_verify_polylines_kernel" 7 1
  cap verify_extend "pp_verify_polylines_kernel" 3 1
  cap fill "pp_dubins_fill_kernel" 0 1
  cap plan "pp_dubins_plan_kernel" 5 1
fi
ncu -i gpurun_out/${R}_dubins_eval.ncu-rep --page raw --csv > gpurun_out/${R}_dubins_eval_raw.csv 2>/dev/null
ncu -i gpurun_out/${R}_dubins_eval.ncu-rep --page source --csv --print-source cuda,sass \
    > gpurun_out/${R}_dubins_eval_source.csv 2>/dev/null
find gpurun_out -name "*.ncu-rep" -size +8M -delete
du -sh gpurun_out; ls -la gpurun_out | tail -30
