# ncu launch list of a bench command: bash tools/ncu_launches.sh <out-name> <bench args...>   (after the plain run exited 0)
name=$1; shift
timeout 300 python bench.py "$@" --no-cpu-baseline > gpurun_out/plain_$name.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$name.csv python bench.py "$@" --no-cpu-baseline > gpurun_out/ncu_$name.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_$name.csv | head -34
