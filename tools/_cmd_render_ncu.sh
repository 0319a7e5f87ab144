timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_r01_v3.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01_bf16_v3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_r01_v3.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r01_bf16_v3.csv | head -32
