timeout 400 python -m pytest tests/test_gpu_tc.py -q -k "training" 2>&1 | tail -2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01_train_tc.csv python bench.py --workload train --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_train.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r01_train_tc.csv 2>&1 | head -40
