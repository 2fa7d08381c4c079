"""What each open semantic of Aether.Physics2D does to the reference's PID episode (tests/Gym.Tests/Envs/Aether/
LunarLanderEnvironment.cs:38-150; golden for NumSharp seed 1000: 1547 steps, return 184.01764), on the generic
oracle (oracle/world2d).  The NumSharp stream is unreproducible, so this reports DISTRIBUTIONS over many terrains /
initial kicks drawn from the engine's Philox stream, not the golden itself.  TEST / ANALYSIS TOOL ONLY.

    python tests/tools/lunar_pid_study.py [episodes]  ->  profiles/lunar_pid_study_r2.txt
"""
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import world2d_lib as W  # noqa: E402


def pid(s):
    angle_targ = min(max(s[0] * 0.5 + s[2] * 1.0, -0.4), 0.4)
    hover_targ = 0.55 * abs(s[0])
    angle_todo = (angle_targ - s[4]) * 0.5 - s[5] * 1.0
    hover_todo = (hover_targ - s[1]) * 0.5 - s[3] * 0.5
    if s[6] > 0 or s[7] > 0:
        angle_todo = 0.0
        hover_todo = -s[3] * 0.5
    a = 0
    if hover_todo > abs(angle_todo) and hover_todo > 0.05:
        a = 2
    elif angle_todo < -0.05:
        a = 3
    elif angle_todo > 0.05:
        a = 1
    return a


def episode(seed, gid, max_steps=5000, **opts):
    w = W.LunarWorld(**opts)
    s = w.reset(W.reset_draws(seed, gid, 0), W.step_draws(seed, gid, 0))
    total, steps = np.float32(0), 0
    while True:
        s, r, d = w.step(pid(s), W.step_draws(seed, gid, steps))
        total = np.float32(total + r)
        steps += 1
        if d or steps > max_steps:
            break
    toi = w.toi_events()
    w.close()
    return steps, float(total), float(r), toi


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    lines = ["PID episodes on oracle/world2d, %d terrains (Philox seed 1000, env ids 0..%d); reference golden (NumSharp seed 1000): 1547 steps, 184.01764" % (n, n - 1),
             "columns: begin_contact_false / TOI / force_at_origin | landed asleep (+100) | crashed (-100) | timeout | steps median [p10, p90] of landed | return mean of landed | episodes with 1200..1900 steps and return 150..220 | TOI events per episode"]
    for bcf, toi, fo in itertools.product((0, 1, 2), (0, 1), (0, 1)):
        res = [episode(1000, g, begin_contact_false=bcf, continuous_physics=toi, force_at_origin=fo) for g in range(n)]
        landed = [(s, t) for s, t, last, _ in res if last == 100.0]
        crashed = [(s, t) for s, t, last, _ in res if last == -100.0]
        timeout = [1 for s, t, last, _ in res if s > 5000]
        near = [1 for s, t, last, _ in res if 1200 <= s <= 1900 and 150 <= t <= 220]
        ls = np.array([s for s, _ in landed]) if landed else np.array([0])
        lt = np.array([t for _, t in landed]) if landed else np.array([0.0])
        lines.append("  %d / %d / %d | %3d | %3d | %3d | %5.0f [%5.0f, %5.0f] | %7.2f | %3d | %.2f" % (
            bcf, toi, fo, len(landed), len(crashed), len(timeout), np.median(ls), np.percentile(ls, 10), np.percentile(ls, 90), lt.mean(), len(near),
            np.mean([x[3] for x in res])))
        print(lines[-1], flush=True)
    out = os.path.join(ROOT, "profiles", "lunar_pid_study_r2.txt")
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
