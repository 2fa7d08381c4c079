set -x
python -m pytest tests/test_gpu_classic.py -x -q -k "Acrobot or sharding" > gpurun_out/pytest_gpu_s2j.log 2>&1; tail -3 gpurun_out/pytest_gpu_s2j.log
MEASURE_MODE=rollout MEASURE_ENVS=Acrobot-v1 python tools/measure_envs.py 2>&1 | cut -c1-260
