set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; tail -1 gpurun_out/bench_8gpu.json; tail -3 gpurun_out/bench_8gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err; tail -1 gpurun_out/bench_4gpu.json | cut -c1-200
