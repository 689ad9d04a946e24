set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
(time timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40) > gpurun_out/pytest.log 2>&1
timeout 300 python bench.py --workload train_step --steps 10 --warmup 3 > gpurun_out/train.json 2> gpurun_out/train.err
timeout 300 python bench.py --workload train_step --whole-canvas --steps 6 --warmup 3 > gpurun_out/train_whole.json 2> gpurun_out/train_whole.err
timeout 300 python bench.py --workload interp --steps 10 --warmup 3 > gpurun_out/interp.json 2> gpurun_out/interp.err
timeout 300 python bench.py > gpurun_out/gen.json 2> gpurun_out/gen.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train.csv python bench.py --workload train_step --device-only --warmup 3 > gpurun_out/ncu_train.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_interp.csv python bench.py --workload interp --device-only --warmup 3 > gpurun_out/ncu_interp.log 2>&1
tail -3 gpurun_out/pytest.log; cat gpurun_out/train.json gpurun_out/interp.json
