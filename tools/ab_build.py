#!/usr/bin/env python
"""Build A/B variants of libgymcuda.so for one gpurun call, and print the command that measures them.

    python tools/ab_build.py name1="-DFOO=1" name2="-DFOO=2 -DBAR" [--lunar] [--envs CartPole-v1,Pendulum-v1] [--mode rollout]

Each variant is compiled (in parallel, without the LunarLander kernels unless --lunar: 25 s instead of 100 s) into
gym.net_b200/csrc/exp/libgymcuda_<name>.so; GYMCUDA_LIB=<path> makes the Python binding load it (gymnet_b200/_native.py).
The printed loop alternates the variants with the in-tree build so that box-to-box differences cancel:

    /usr/local/graft/bin/gpurun --timeout 900 -- 'bash gym.net_b200/csrc/exp/run_ab.sh'

Remove gym.net_b200/csrc/exp/ afterwards (it travels to the GPU box with every gpurun call).
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gym.net_b200", "csrc")
EXP = os.path.join(CSRC, "exp")
BASE = ["/usr/local/cuda/bin/nvcc", "-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-gencode", "arch=compute_100a,code=sm_100a",
        "--extended-lambda", "-Xcompiler", "-fPIC", "-shared"]


def main():
    args = sys.argv[1:]
    lunar = "--lunar" in args
    envs = "CartPole-v1"
    mode = "rollout"
    variants = []
    it = iter(args)
    for a in it:
        if a == "--lunar":
            continue
        if a == "--envs":
            envs = next(it); continue
        if a == "--mode":
            mode = next(it); continue
        name, _, flags = a.partition("=")
        variants.append((name, flags.split()))
    if not variants:
        raise SystemExit(__doc__)
    os.makedirs(EXP, exist_ok=True)
    procs = []
    for name, flags in variants:
        out = os.path.join(EXP, "libgymcuda_%s.so" % name)
        cmd = BASE + (["-DGYMCUDA_WITH_LUNAR"] if lunar else []) + flags + ["-o", out, "gymcuda.cu", "-ldl"]
        procs.append((name, subprocess.Popen(cmd, cwd=CSRC, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, p in procs:
        log = p.communicate()[0]
        if p.returncode != 0:
            raise SystemExit("variant %s failed to build:\n%s" % (name, log[-3000:]))
        print("built", name)
    libs = ["libgymcuda"] + ["exp/libgymcuda_%s" % n for n, _ in variants]
    script = os.path.join(EXP, "run_ab.sh")
    with open(script, "w") as f:
        f.write("export MEASURE_MODE=%s MEASURE_ENVS=%s\nfor rep in 1 2; do for lib in %s; do\n" % (mode, envs, " ".join(libs)))
        f.write("echo \"== $lib\"\nGYMCUDA_LIB=$PWD/gym.net_b200/csrc/$lib.so python tools/measure_envs.py 2>&1 | python -c \"import sys,json\n"
                "for l in sys.stdin:\n    d=json.loads(l); print('  ', d['env'], d['mode'], '%.1f us  frac %.3f' % (d['ms_per_launch']*1e3, d['frac_of_measured_hbm']))\"\n")
        f.write("done; done\n")
    print("/usr/local/graft/bin/gpurun --timeout 900 -- 'bash gym.net_b200/csrc/exp/run_ab.sh'")


if __name__ == "__main__":
    main()
