set -x
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60) > gpurun_out/pytest.log 2>&1
timeout 300 python bench.py --workload train_step --steps 20 --warmup 5 > gpurun_out/train.json 2> gpurun_out/train.err
timeout 300 python bench.py > gpurun_out/gen.json 2> gpurun_out/gen.err
TMX_E2E_MB=16 timeout 300 python bench.py > gpurun_out/gen_mb16.json 2> gpurun_out/gen_mb16.err
TMX_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_train.csv python bench.py --workload train_step --device-only --warmup 3 > gpurun_out/ncu_train.log 2>&1
grep -E "passed|failed" gpurun_out/pytest.log
