nvidia-smi topo -m 2>/dev/null | head -14; lscpu | grep -E "NUMA|Socket|^CPU\(s\)" | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 20 --warmup 3 2> gpurun_out/bench_4gpu_numa.err | grep '^{' > gpurun_out/bench_4gpu_numa.json; python -c "
import json; d=json.load(open('gpurun_out/bench_4gpu_numa.json')); print('value %.4e e2e %.4e' % (d['value'], d['e2e']['value']), d['e2e'].get('host_affinity'))"
