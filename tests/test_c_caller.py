"""The boundary is a C ABI: include/gymcuda.h must compile as pedantic C99 and a plain C program must link
against libgymcuda.so and drive it (examples/c_driver.c).  Without a GPU the program has to fail loudly with
GYMCUDA_ECUDA (exit code 3) -- there is no CPU path to fall back to."""
import os
import subprocess

import pytest

from conftest import HAS_GPU

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_program_compiles_links_and_runs():
    ex = os.path.join(ROOT, "examples")
    subprocess.run(["make", "-C", ex, "-s", "clean"], check=True)
    subprocess.run(["make", "-C", ex, "-s"], check=True)
    r = subprocess.run([os.path.join(ex, "c_driver"), "2048", "20"], capture_output=True, text=True, timeout=300)
    if HAS_GPU:
        assert r.returncode == 0, r.stderr
        assert "env-steps/s" in r.stdout and "rollout_random" in r.stdout
    else:
        assert r.returncode == 3, (r.returncode, r.stderr)
        assert "no CPU path" in r.stderr


@pytest.mark.gpu
def test_c_program_steps_on_the_gpu():
    ex = os.path.join(ROOT, "examples")
    subprocess.run(["make", "-C", ex, "-s"], check=True)
    r = subprocess.run([os.path.join(ex, "c_driver"), "4096", "50"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "env-steps/s" in r.stdout and "rollout_random" in r.stdout
    episodes = int(r.stdout.split("steps,")[1].split("episodes")[0])
    assert episodes > 0      # alternating actions: an episode ends every ~40 steps
