export MEASURE_MODE=step MEASURE_ENVS=CartPole-v1,Pendulum-v1
for lib in libgymcuda exp_old exp_older; do for i in 1 2; do
GYMCUDA_LIB=$PWD/gym.net_b200/csrc/$lib.so python tools/measure_envs.py 2>&1 | python -c "import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if d['mode']=='step_device': print('$lib', d['env'], '%.2f us' % (d['ms_per_launch']*1e3))"
done; done
