#!/bin/bash
# ncu evidence for one round (run under gpurun, 1 GPU).  Usage: bash profiles/capture.sh r02 [all]
# Every ncu run is preceded by the same command line without ncu (B200_PROFILING.md).  The reports are exported to
# CSV on the box (raw page + CUDA/SASS source page) and the .ncu-rep files are deleted when large: gpurun copies back
# at most 64 MiB.
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
if [ "${2:-}" = "all" ]; then
  $CMD > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on \
      -k regex:"pp_rrt_extend_fused_kernel|pp_extend_bin_kernel|pp_extend_scatter_kernel|pp_verify_polylines_kernel|pp_dubins_fill_kernel" \
      -c 16 -o gpurun_out/${R}_rrt $CMD > gpurun_out/${R}_rrt_run.log 2>&1
  echo "rrt capture rc=$?"
fi
for f in gpurun_out/${R}_dubins_eval.ncu-rep gpurun_out/${R}_rrt.ncu-rep; do
  [ -f "$f" ] || continue
  ncu -i "$f" --page raw --csv > "${f%.ncu-rep}_raw.csv" 2>/dev/null
done
[ -f gpurun_out/${R}_dubins_eval.ncu-rep ] && ncu -i gpurun_out/${R}_dubins_eval.ncu-rep --page source --csv \
    --print-source cuda,sass > gpurun_out/${R}_dubins_eval_source.csv 2>/dev/null
if [ -f gpurun_out/${R}_rrt.ncu-rep ]; then
  for k in pp_rrt_extend_fused_kernel pp_verify_polylines_kernel pp_dubins_fill_kernel; do
    ncu -i gpurun_out/${R}_rrt.ncu-rep --page source --csv --print-source cuda,sass -k regex:$k -c 1 \
        > gpurun_out/${R}_${k}_source.csv 2>/dev/null
  done
  # the no-hit C5 launch of the verify kernel is the last of its six instances
  ncu -i gpurun_out/${R}_rrt.ncu-rep --page source --csv --print-source cuda,sass -k regex:pp_verify_polylines_kernel -s 5 -c 1 \
      > gpurun_out/${R}_pp_verify_polylines_kernel_nohit_source.csv 2>/dev/null
fi
find gpurun_out -name "*.ncu-rep" -size +20M -delete
du -sh gpurun_out; ls -la gpurun_out | tail -20
