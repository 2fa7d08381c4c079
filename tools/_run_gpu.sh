set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_s2i.log 2>&1; tail -3 gpurun_out/pytest_gpu_s2i.log
(timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_probe.py 2>&1 | tail -4) > gpurun_out/sanitizer_mem_s2i.log 2>&1; cat gpurun_out/sanitizer_mem_s2i.log
(timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_probe.py 2>&1 | tail -4) > gpurun_out/sanitizer_race_s2i.log 2>&1; cat gpurun_out/sanitizer_race_s2i.log
