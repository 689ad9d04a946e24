set -x
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests/test_gpu_train_step.py tests/test_gpu_eg_loss.py -m gpu -q -p no:cacheprovider 2>&1 | tail -30) > gpurun_out/pytest.log 2>&1
timeout 300 python bench.py --workload train_step --steps 20 --warmup 5 > gpurun_out/train.json 2> gpurun_out/train.err
grep -E "passed|failed" gpurun_out/pytest.log
