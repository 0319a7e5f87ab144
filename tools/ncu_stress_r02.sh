# --set full captures of the two K = 16 kernels of the configs[4] stress block (10 M points, 5x5x5 kernel); run on the GPU box
set -x
ARGS="--steps 1 --warmup 1 --no-train --no-cpu-baseline"
timeout 600 python bench.py $ARGS > gpurun_out/r02_stress_plain.log 2>&1 || exit 1
# (demangled names read `query_kernel<(int)16>`, `field_tc_kernel<(int)16, (bool)0>`)
for spec in 'query_kernel<.int.16>:query16' 'field_tc_kernel<.int.16,:field16'; do
  k=${spec%%:*}; name=${spec#*:}
  timeout 900 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"$k" -s 1 -c 1 \
      -o gpurun_out/r02_prof_$name -f python bench.py $ARGS > gpurun_out/r02_ncu_full_$name.log 2>&1
done
ls -la gpurun_out/r02_prof_query16* gpurun_out/r02_prof_field16*
