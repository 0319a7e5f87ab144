# Round-2 profiling session (run on the GPU box only after the plain commands exited 0):
#   1. launch lists (gpu__time_duration per launch) of the default bench with the eager train block, so that the kernels of a
#      training step are listed individually (a CUDA-graph replay shows as one graph launch);
#   2. one `--set full` capture each of the hot kernels of the render step and of the training step.
set -x
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-stress --train-steps 3 > gpurun_out/r02_plain_before_ncu.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-stress --train-steps 3 > gpurun_out/r02_ncu_launches.log 2>&1
# render kernels: skip the warm-up launches (3 warm-up steps x 3 class launches), capture one of each
# (-k matches the function name without its template arguments: the three class launches <8>, <4>, <2> of one view are captured together)
for spec in "field_tc_kernel:field:12:3" "query_kernel:query:4:1" "color_tc_kernel:color:4:1" "sample_select_kernel:select:4:1"; do
  k=${spec%%:*}; rest=${spec#*:}; name=${rest%%:*}; rest=${rest#*:}; skip=${rest%%:*}; cnt=${rest#*:}
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:"$k" -s $skip -c $cnt -o gpurun_out/r02_prof_$name -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train > gpurun_out/r02_ncu_full_$name.log 2>&1
done
# training kernels (eager launches)
for spec in "tile_gemm_kernel:tile_gemm:30" "wgrad_tc_kernel:wgrad:4" "scatter_kernel:scatter:4" "dp_adam_kernel:dp_adam:4"; do
  k=${spec%%:*}; rest=${spec#*:}; name=${rest%%:*}; skip=${rest#*:}
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:"$k" -s $skip -c 1 -o gpurun_out/r02_prof_$name -f \
      python bench.py --steps 2 --warmup 3 --workload train --no-graph > gpurun_out/r02_ncu_full_$name.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
