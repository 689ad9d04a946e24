set -x
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40) > gpurun_out/pytest.log 2>&1
timeout 300 python bench.py --workload train_step --steps 20 --warmup 5 > gpurun_out/train.json 2> gpurun_out/train.err
timeout 300 python bench.py --workload train_step --steps 10 --warmup 3 > gpurun_out/train_b.json 2> gpurun_out/train_b.err
tail -5 gpurun_out/pytest.log; cat gpurun_out/train.json
