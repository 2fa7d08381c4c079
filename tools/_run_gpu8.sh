mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_r2.txt 2>&1
for N in 1 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N tools/pcie_scaling_probe.py 2>/dev/null | tail -1 >> gpurun_out/pcie_scaling_r2.jsonl
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 tools/pcie_scaling_probe.py bind 2>/dev/null | tail -1 >> gpurun_out/pcie_scaling_r2.jsonl
cut -c1-400 gpurun_out/pcie_scaling_r2.jsonl
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_r2_n8.json 2> gpurun_out/bench_r2_n8.err; tail -c 3000 gpurun_out/bench_r2_n8.json; tail -3 gpurun_out/bench_r2_n8.err
python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/pytest_gpu_multi_r2.log 2>&1; tail -3 gpurun_out/pytest_gpu_multi_r2.log
