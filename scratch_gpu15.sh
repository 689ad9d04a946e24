set -x
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -60) > gpurun_out/pytest.log 2>&1
grep -E "passed|failed" gpurun_out/pytest.log
