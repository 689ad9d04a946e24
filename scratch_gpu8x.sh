set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --workload train_step --steps 15 --warmup 4 > gpurun_out/train_8gpu.json 2> gpurun_out/train_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --workload train_step --steps 15 --warmup 4 > gpurun_out/train_4gpu.json 2> gpurun_out/train_4gpu.err
tail -c 400 gpurun_out/train_8gpu.json; tail -3 gpurun_out/train_8gpu.err
