set -x
mkdir -p gpurun_out
(time timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -30) > gpurun_out/pytest.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1
timeout 300 python bench.py > gpurun_out/gen.json 2> gpurun_out/gen.err
timeout 300 python bench.py --workload interp --steps 10 --warmup 3 > gpurun_out/interp.json 2> gpurun_out/interp.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_gen.csv python bench.py --device-only --steps 2 --warmup 3 > gpurun_out/ncu_gen.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_tc_kernel -c 2 -o gpurun_out/trunk python bench.py --device-only --steps 1 --warmup 3 > gpurun_out/ncu_trunk.log 2>&1
ncu -i gpurun_out/trunk.ncu-rep --page raw --csv > gpurun_out/trunk_raw.csv 2>/dev/null
tail -3 gpurun_out/pytest.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/gen.json
