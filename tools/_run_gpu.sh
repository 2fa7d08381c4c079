python -m pytest tests/test_gpu_classic.py -x -q -k "Pendulum or MountainCar or pendulum or edge" 2>&1 | tail -3
MEASURE_MODE=rollout MEASURE_ENVS=Pendulum-v1,MountainCarContinuous-v0,MountainCar-v0 python tools/measure_envs.py 2>&1 | python -c "import sys,json; [print(d['env'], '%.1f us frac %.3f' % (d['ms_per_launch']*1e3, d['frac_of_measured_hbm'])) for d in map(json.loads, sys.stdin)]"
