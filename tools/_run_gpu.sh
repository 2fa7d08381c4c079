(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_final2.log 2>&1; tail -4 gpurun_out/pytest_gpu_final2.log | head -2
python tools/measure_envs.py > gpurun_out/envs_final2.jsonl 2>gpurun_out/envs_final2.err; python - <<PY
import json
for l in open("gpurun_out/envs_final2.jsonl"):
    d=json.loads(l); print("%-26s %-12s %9.1f us  %.3e steps/s  frac %.3f" % (d["env"], d["mode"], d["ms_per_launch"]*1e3, d["env_steps_per_s"], d["frac_of_measured_hbm"]))
PY
python bench.py > gpurun_out/bench_final2_n1.json 2> gpurun_out/bench_final2_n1.err; cut -c1-700 gpurun_out/bench_final2_n1.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -o gpurun_out/prof_pendulum_final -f python tools/rollout_probe.py Pendulum-v1:262144:128 > gpurun_out/prof_pendulum_final.log 2>&1; tail -1 gpurun_out/prof_pendulum_final.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -o gpurun_out/prof_acrobot_final -f python tools/rollout_probe.py Acrobot-v1:131072:128 > gpurun_out/prof_acrobot_final.log 2>&1; tail -1 gpurun_out/prof_acrobot_final.log
