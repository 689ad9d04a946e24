set -x
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -30) > gpurun_out/pytest.log 2>&1
timeout 300 python bench.py --workload train_step --steps 20 --warmup 5 > gpurun_out/train.json 2> gpurun_out/train.err
TMX_NO_GRAPH=1 timeout 300 python bench.py --workload train_step --steps 20 --warmup 5 > gpurun_out/train_nograph.json 2> gpurun_out/train_nograph.err
tail -5 gpurun_out/pytest.log; cat gpurun_out/train.json; tail -5 gpurun_out/train.err
