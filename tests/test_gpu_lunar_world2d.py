"""GPU parity of the LunarLander kernel against the structurally independent oracle (oracle/world2d), through the C ABI:
teacher-forced single steps -- set_state, one step, get_state -- from >= 10^4 states the generic oracle reaches under the
reference test's PID heuristic mixed with random actions (free flight, near the ground, one leg, two legs, belly, asleep).

  * det mode  (the engine's float32 sincos plugged into the generic engine): every word of the state, the observation, the
    reward and done must agree BIT FOR BIT -- any transcription error of the specialised kernel shows here;
  * ref mode  ((float)Math.Sin((double)a), the reference's rotation arithmetic): integer words (contact flags, touching
    masks, limit states, contact ids, broad-phase pair lists) and done identical, floats within world2d_lib.TOLERANCES.
The committed fixture tests/golden/lunar_world2d.npz (generator beside it) is checked the same way."""
import os

import numpy as np
import pytest

import gymnet_b200 as G
import world2d_lib as W

pytestmark = pytest.mark.gpu
SEED, T = 2024, 1000


def gpu_step(tr, continuous=False, chunk=None):
    n = len(tr["action"])
    env = G.LunarLanderVecEnv(n, continuous=continuous, seed=SEED)
    env.ResetBatch()
    aux = np.zeros((n, env.aux_dim), np.int32)
    aux[:, :W.AUX_DIM] = tr["aux0"]
    env.SetState(tr["state0"], aux, T)
    obs, rew, done = env.StepBatch(tr["action"])
    st, ax, t = env.GetState()
    assert t == T + 1
    env.Close()
    return st, ax, obs, rew, done


@pytest.mark.parametrize("continuous", [False, True])
def test_kernel_equals_generic_oracle_bit_for_bit_with_engine_sincos(continuous):
    tr = W.generate_transitions(6000 if not continuous else 2400, seed=SEED, T=T, landers=16 if not continuous else 8, continuous=continuous, det_sincos=1)
    cat = W.categories(tr["aux0"], tr["state0"])
    assert {"free", "near", "legs", "two_legs", "belly"} <= set(cat), sorted(set(cat))
    st, ax, obs, rew, done = gpu_step(tr, continuous)
    W.compare_transitions("det sincos", tr, st, ax, obs, rew, done, exact=True)


def test_kernel_within_tolerance_of_generic_oracle_on_reference_sincos_10k_states():
    tr = W.generate_transitions(12000, seed=SEED, T=T, landers=24, det_sincos=0)
    cat = W.categories(tr["aux0"], tr["state0"])
    counts = {k: int((cat == k).sum()) for k in sorted(set(cat))}
    assert counts.get("legs", 0) + counts.get("two_legs", 0) > 500 and counts.get("belly", 0) > 20 and counts.get("free", 0) > 2000, counts
    st, ax, obs, rew, done = gpu_step(tr)
    W.compare_transitions("ref sincos %s" % counts, tr, st, ax, obs, rew, done, exact=False)


def test_golden_fixture_from_the_generic_oracle():
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "lunar_world2d.npz"))
    for name, exact in (("det", True), ("ref", False)):
        tr = {k.split("/", 1)[1]: fx[k] for k in fx.files if k.startswith(name + "/")}
        st, ax, obs, rew, done = gpu_step(tr)
        W.compare_transitions("fixture " + name, tr, st, ax, obs, rew, done, exact=exact)


def test_reset_matches_the_generic_oracle_construction():
    """LunarLanderEnv.Reset built from CreateBody / CreateFixture / RevoluteJoint calls (hulls, mass data, joint frames from the
    vertices) against the kernel's hard-coded topology: terrain, initial kick (force at the body origin), proxy boxes, zero step."""
    n = 256
    env = G.LunarLanderVecEnv(n, seed=SEED)
    obs = env.ResetBatch()
    st, ax, t = env.GetState()
    for g in range(n):
        wi, ti = W.ctor_draws(SEED, g)
        w = W.LunarWorld(det_sincos=1, wind_idx=wi, torque_idx=ti)
        o = w.reset(W.reset_draws(SEED, g, 0), W.step_draws(SEED, g, 0))
        s, a = w.export_state()
        assert np.array_equal(o, obs[g]), g
        assert np.array_equal(s, st[g, :W.STATE_DIM].astype(np.float32)) and np.array_equal(a, ax[g, :W.AUX_DIM]), g
        w.close()
    env.Close()
