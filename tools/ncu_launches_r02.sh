# launch list of the default bench without the stress block (train block launched eagerly so that its kernels are listed one by one)
# + one --set full capture of the colour kernel; run on the GPU box after the plain command exited 0
set -x
ARGS="--steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-stress --train-steps 3"
timeout 600 python bench.py $ARGS > gpurun_out/r02_plain_before_ncu.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py $ARGS > gpurun_out/r02_ncu_launches.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:color_tc_kernel -s 4 -c 1 -o gpurun_out/r02_prof_color -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train --no-stress > gpurun_out/r02_ncu_full_color.log 2>&1
