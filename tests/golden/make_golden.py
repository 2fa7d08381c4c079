"""Generates tests/golden/*.npz.  The reference is C# and cannot run in this image (no dotnet/mono),
so these vectors are NOT reference outputs: they come from an independent pure-Python float64
restatement of the same source lines (CartPoleEnv.cs:24-36,137-186) and of the upstream gym 0.26
formulas, written separately from oracle/classic.hpp.  They pin the C oracle against transcription
slips and give the GPU tests a frozen set of teacher-forced transitions.

    python tests/golden/make_golden.py
"""
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
f32 = np.float32

# ---- CartPole: C# `const float` folded in float32 (CartPoleEnv.cs:24-36)
GRAVITY = float(f32(9.8)); MASSPOLE = float(f32(0.1)); TOTAL_MASS = float(f32(0.1) + f32(1.0))
LENGTH = 0.5; POLEMASS_LENGTH = float(f32(0.1) * f32(0.5)); FORCE_MAG = 10.0; TAU = float(f32(0.02))
THETA_THR = float(f32(12 * 2 * math.pi / 360)); X_THR = float(f32(2.4))


def cartpole_step(state, action, sbd):
    x, x_dot, theta, theta_dot = state
    force = FORCE_MAG if action == 1 else -FORCE_MAG
    costheta, sintheta = math.cos(theta), math.sin(theta)
    temp = (force + POLEMASS_LENGTH * theta_dot * theta_dot * sintheta) / TOTAL_MASS
    thetaacc = (GRAVITY * sintheta - costheta * temp) / (LENGTH * (4.0 / 3.0 - MASSPOLE * costheta * costheta / TOTAL_MASS))
    xacc = temp - POLEMASS_LENGTH * thetaacc * costheta / TOTAL_MASS
    x = x + TAU * x_dot
    x_dot = x_dot + TAU * xacc
    theta = theta + TAU * theta_dot
    theta_dot = theta_dot + TAU * thetaacc
    done = x < -X_THR or x > X_THR or theta < -THETA_THR or theta > THETA_THR
    if not done:
        reward = 1.0
    elif sbd == -1:
        sbd, reward = 0, 1.0
    else:
        sbd, reward = sbd + 1, 0.0
    return (x, x_dot, theta, theta_dot), reward, done, sbd


def pendulum_step(state, u):
    th, thdot = state
    g, m, l, dt = 10.0, 1.0, 1.0, 0.05
    u = min(max(u, -2.0), 2.0)
    an = ((th + math.pi) % (2 * math.pi)) - math.pi
    costs = an ** 2 + 0.1 * thdot ** 2 + 0.001 * (u ** 2)
    newthdot = thdot + (3 * g / (2 * l) * math.sin(th) + 3.0 / (m * l ** 2) * u) * dt
    newthdot = min(max(newthdot, -8.0), 8.0)
    newth = th + newthdot * dt
    return (newth, newthdot), -costs, False


def mountaincar_step(state, action):
    position, velocity = state
    velocity += (action - 1) * 0.001 + math.cos(3 * position) * (-0.0025)
    velocity = min(max(velocity, -0.07), 0.07)
    position += velocity
    position = min(max(position, -1.2), 0.6)
    if position == -1.2 and velocity < 0:
        velocity = 0
    done = bool(position >= 0.5 and velocity >= 0)
    return (position, velocity), -1.0, done


def mountaincar_cont_step(state, a):
    position, velocity = state
    force = min(max(a, -1.0), 1.0)
    velocity += force * 0.0015 - 0.0025 * math.cos(3 * position)
    velocity = min(velocity, 0.07); velocity = max(velocity, -0.07)
    position += velocity
    position = min(position, 0.6); position = max(position, -1.2)
    if position == -1.2 and velocity < 0:
        velocity = 0
    done = bool(position >= 0.45 and velocity >= 0)
    reward = 100.0 if done else 0.0
    reward -= math.pow(a, 2) * 0.1
    return (position, velocity), reward, done


def _acro_dsdt(s, a):
    m1 = m2 = l1 = 1.0; lc1 = lc2 = 0.5; I1 = I2 = 1.0; g = 9.8
    theta1, theta2, dtheta1, dtheta2 = s
    d1 = m1 * lc1 ** 2 + m2 * (l1 ** 2 + lc2 ** 2 + 2 * l1 * lc2 * math.cos(theta2)) + I1 + I2
    d2 = m2 * (lc2 ** 2 + l1 * lc2 * math.cos(theta2)) + I2
    phi2 = m2 * lc2 * g * math.cos(theta1 + theta2 - math.pi / 2.0)
    phi1 = (-m2 * l1 * lc2 * dtheta2 ** 2 * math.sin(theta2) - 2 * m2 * l1 * lc2 * dtheta2 * dtheta1 * math.sin(theta2)
            + (m1 * lc1 + m2 * l1) * g * math.cos(theta1 - math.pi / 2) + phi2)
    ddtheta2 = (a + d2 / d1 * phi1 - m2 * l1 * lc2 * dtheta1 ** 2 * math.sin(theta2) - phi2) / (m2 * lc2 ** 2 + I2 - d2 ** 2 / d1)
    ddtheta1 = -(d2 * ddtheta2 + phi1) / d1
    return np.array([dtheta1, dtheta2, ddtheta1, ddtheta2])


def _wrap(x, m, M):
    diff = M - m
    while x > M:
        x -= diff
    while x < m:
        x += diff
    return x


def acrobot_step(state, action):
    a = float(action - 1); dt = 0.2
    y0 = np.array(state, dtype=np.float64)
    k1 = _acro_dsdt(y0, a); k2 = _acro_dsdt(y0 + dt / 2 * k1, a); k3 = _acro_dsdt(y0 + dt / 2 * k2, a); k4 = _acro_dsdt(y0 + dt * k3, a)
    ns = y0 + dt / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
    ns[0] = _wrap(ns[0], -math.pi, math.pi); ns[1] = _wrap(ns[1], -math.pi, math.pi)
    ns[2] = min(max(ns[2], -4 * math.pi), 4 * math.pi); ns[3] = min(max(ns[3], -9 * math.pi), 9 * math.pi)
    done = bool(-math.cos(ns[0]) - math.cos(ns[1] + ns[0]) > 1.0)
    return tuple(ns), (0.0 if done else -1.0), done


def main():
    rng = np.random.default_rng(20260925)
    n = 2048
    out = {}
    # CartPole: float32-representable states (the engine's storage type), incl. post-done ones
    s = rng.uniform([-2.6, -3, -0.25, -3.5], [2.6, 3, 0.25, 3.5], size=(n, 4)).astype(f32)
    a = rng.integers(0, 2, n).astype(np.int32)
    sbd = rng.integers(-1, 3, n).astype(np.int32)
    ns = np.empty((n, 4)); r = np.empty(n, f32); d = np.empty(n, np.uint8); nsbd = np.empty(n, np.int32)
    for i in range(n):
        st, rw, dn, sb = cartpole_step([float(v) for v in s[i]], int(a[i]), int(sbd[i]))
        ns[i], r[i], d[i], nsbd[i] = st, rw, dn, sb
    np.savez_compressed(os.path.join(HERE, "cartpole.npz"), state=s, action=a, sbd=sbd, next_state=ns, reward=r, done=d, next_sbd=nsbd)

    def gen(name, fn, lo, hi, act):
        s = rng.uniform(lo, hi, size=(n, len(lo))).astype(f32)
        ns = np.empty((n, len(lo))); r = np.empty(n); d = np.empty(n, np.uint8)
        for i in range(n):
            st, rw, dn = fn([float(v) for v in s[i]], act[i].item())
            ns[i], r[i], d[i] = st, rw, dn
        np.savez_compressed(os.path.join(HERE, name + ".npz"), state=s, action=act, next_state=ns, reward=r, done=d)

    gen("pendulum", pendulum_step, [-10, -8], [10, 8], rng.uniform(-2.5, 2.5, n).astype(f32))
    gen("mountaincar", mountaincar_step, [-1.2, -0.07], [0.6, 0.07], rng.integers(0, 3, n).astype(np.int32))
    gen("mountaincar_cont", mountaincar_cont_step, [-1.2, -0.07], [0.6, 0.07], rng.uniform(-1.2, 1.2, n).astype(f32))
    gen("acrobot", acrobot_step, [-3.14, -3.14, -6, -12], [3.14, 3.14, 6, 12], rng.integers(0, 3, n).astype(np.int32))


if __name__ == "__main__":
    main()
