set -x
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_s2c.log 2>&1; tail -5 gpurun_out/pytest_gpu_s2c.log
export MEASURE_MODE=rollout MEASURE_ENVS=CartPole-v1,Pendulum-v1,MountainCarContinuous-v0,MountainCar-v0,Acrobot-v1
python tools/measure_envs.py > gpurun_out/envs_s2c_new.jsonl 2>&1
for f in new; do echo == $f; python - <<PY
import json
for l in open("gpurun_out/envs_s2c_$f.jsonl"):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print("%-28s %8.1f us  %.3e steps/s  frac %.3f" % (d["env"], d["ms_per_launch"]*1e3, d["env_steps_per_s"], d["frac_of_measured_hbm"]))
PY
done
