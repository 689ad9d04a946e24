set -x
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -30) > gpurun_out/pytest.log 2>&1
timeout 300 python bench.py --workload train_step --steps 20 --warmup 5 > gpurun_out/train.json 2> gpurun_out/train.err
TMX_NO_WGRAD_PACK=1 timeout 300 python bench.py --workload train_step --steps 20 --warmup 5 > gpurun_out/train_nopack.json 2> gpurun_out/train_nopack.err
TMX_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train.csv python bench.py --workload train_step --device-only --warmup 3 > gpurun_out/ncu_train.log 2>&1
tail -5 gpurun_out/pytest.log; cat gpurun_out/train.json; tail -5 gpurun_out/train.err
