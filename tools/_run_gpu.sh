set -x
(time python -m pytest tests/test_gpu_lunar.py -x -q) > gpurun_out/pytest_gpu_s2f.log 2>&1; tail -4 gpurun_out/pytest_gpu_s2f.log
MEASURE_ENVS=LunarLander-v2 python tools/measure_envs.py > gpurun_out/envs_s2f.jsonl 2>&1; cat gpurun_out/envs_s2f.jsonl | cut -c1-330
