# one `--set full` capture each of the three hot kernels of the render bench (run only after bench.py exited 0 without ncu)
set -e
timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu-baseline > gpurun_out/plain_before_ncu.log 2>&1
for k in field_tc_kernel query_kernel color_tc_kernel; do
  timeout 500 ncu --set full --import-source on --clock-control none -k regex:$k -s 2 -c 1 -o gpurun_out/prof_r01_final_$k -f \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
