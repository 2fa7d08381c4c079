python -m pytest tests/test_gpu_classic.py -x -q -k "Acrobot or edge or staged" 2>&1 | tail -3
MEASURE_MODE=rollout MEASURE_ENVS=Acrobot-v1 python tools/measure_envs.py 2>&1 | python -c "import sys,json; [print(d['env'], '%.1f us frac %.3f %.3e' % (d['ms_per_launch']*1e3, d['frac_of_measured_hbm'], d['env_steps_per_s'])) for d in map(json.loads, sys.stdin)]"
