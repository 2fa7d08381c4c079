// Gym.Environments.Vector.CudaVecEnv -- the drop-in VecEnv over libgymcuda.
// Derives from the reference's VecEnv (src/Gym/Envs/VecEnv.cs:12-93) and keeps IVecEnv
// (src/Gym/Envs/IVecEnv.cs:8-19) source-compatible: Reset() -> NDArray[], Step(int) -> Step[],
// Seed(int), Seed(int[]), Close().  The batched overloads are what a learner should call.
// NOT compiled here (no .NET SDK in the build image); mirrored 1:1 by gym.net_b200/vector.py, which
// is what the parity tests drive.
using System;
using System.Collections.Generic;
using System.Runtime.InteropServices;
using Gym.Envs;
using Gym.Observations;
using Gym.Spaces;
using NumSharp;

namespace Gym.Environments.Vector {
    // IVecEnv is re-implemented explicitly: VecEnv.Seed(int) / Seed(int[]) are NOT virtual (VecEnv.cs:44-53) and loop over the
    // (here empty) Environments list, so `public new void Seed` alone would be reached only through a CudaVecEnv-typed
    // reference -- through an IVecEnv or VecEnv reference the base would run and silently seed nothing.  Listing IVecEnv
    // again makes the interface map to the members below; a VecEnv-typed (base class) reference still binds statically to
    // the base methods, which is why INTEGRATION.md proposes `virtual` on the two base methods as the one-line upstream patch.
    public class CudaVecEnv : VecEnv, IVecEnv, IDisposable {
        private readonly GymCudaHandle _h;
        private readonly GymCudaSpaceInfo _info;
        private readonly float[] _obs, _reward;
        private readonly byte[] _done;
        // The result arrays handed back by Step are pinned for the GC for the env's lifetime AND page-locked for
        // CUDA (gymcuda_host_register): the step kernel then writes them in place over PCIe (42 us per 65 536-env
        // CartPole step) instead of the library staging pageable memory (160 us).  A caller that steps with the
        // same action array every time can give it the same treatment with PinActions().
        // (The data of a managed array is 8-byte aligned; when _obs lands on 8 mod 16 the library stages that one
        // buffer instead of writing it in place -- same results, see INTEGRATION.md "Host buffers".)
        private readonly List<GCHandle> _pins = new List<GCHandle>();

        public int ObsDim => _info.ObsDim;
        public int ActDim => _info.ActDim;
        public bool AutoReset { get; }

        protected CudaVecEnv(GymCudaEnvKind kind, int numEnvs, Space observationSpace, Space actionSpace, ulong seed = 0,
                             int device = 0, uint envIdOffset = 0, bool autoReset = false, int timeLimit = 0,
                             Action<GymCudaConfigBox> configure = null)
            : base(numEnvs, observationSpace, actionSpace) {
            Native.Check(Native.gymcuda_config_default(out var cfg, (int) kind, numEnvs));
            cfg.Seed = seed; cfg.Device = device; cfg.EnvIdOffset = envIdOffset; cfg.TimeLimit = timeLimit;
            cfg.Flags = autoReset ? (uint) GymCudaFlags.AutoReset : 0u;
            if (configure != null) { var box = new GymCudaConfigBox { Value = cfg }; configure(box); cfg = box.Value; }
            Native.Check(Native.gymcuda_create(ref cfg, out _h));
            Native.Check(Native.gymcuda_space(_h, out _info));
            AutoReset = autoReset;
            _obs = new float[numEnvs * _info.ObsDim];
            _reward = new float[numEnvs];
            _done = new byte[numEnvs];
            Pin(_obs, sizeof(float) * _obs.Length);
            Pin(_reward, sizeof(float) * _reward.Length);
            Pin(_done, _done.Length);
        }

        private void Pin(Array a, int bytes) {
            var h = GCHandle.Alloc(a, GCHandleType.Pinned);
            _pins.Add(h);
            Native.Check(Native.gymcuda_host_register(h.AddrOfPinnedObject(), (UIntPtr) (ulong) bytes));
        }

        /// <summary>Pins and page-locks an action array the caller refills in place before every Step.</summary>
        public void PinActions(int[] actions) { Pin(actions, sizeof(int) * actions.Length); }
        public void PinActions(float[] actions) { Pin(actions, sizeof(float) * actions.Length); }

        // ---- IVecEnv -------------------------------------------------------------------------
        public override NDArray[] Reset() {
            Native.Check(Native.gymcuda_reset(_h, _obs));
            return SplitObservations();
        }

        /// <summary>IVecEnv.Step(int action): ONE action broadcast to all envs (VecEnvWrapper.cs:22-24).</summary>
        public override Step[] Step(int action) {
            Native.Check(Native.gymcuda_step_broadcast(_h, action, _obs, _reward, _done));
            var obs = SplitObservations();
            var steps = new Step[NumberOfEnvironments];
            for (int i = 0; i < steps.Length; i++) steps[i] = new Step(obs[i], _reward[i], _done[i] != 0, null);
            return steps;
        }

        public override void Close() {
            if (_terminalPin.IsAllocated) { Native.gymcuda_host_unregister(_terminalPin.AddrOfPinnedObject()); _terminalPin.Free(); }
            foreach (var h in _pins) { Native.gymcuda_host_unregister(h.AddrOfPinnedObject()); h.Free(); }
            _pins.Clear();
            _h.Dispose();
        }
        public void Dispose() { Close(); }

        // Seed widening (the same rule in gym.net_b200/vector.py): the 32-bit pattern of the int, zero-extended -- Seed(-1) is
        // 0x00000000FFFFFFFF in both hosts.
        public new void Seed(int seed) { Native.Check(Native.gymcuda_seed(_h, (ulong) (uint) seed)); }
        public new void Seed(int[] seed) { Native.Check(Native.gymcuda_seed_each(_h, seed, seed.Length)); }   // length check -> ArgumentException, as VecEnv.cs:49
        void IVecEnv.Seed(int seed) { Seed(seed); }
        void IVecEnv.Seed(int[] seed) { Seed(seed); }

        // ---- batched API ---------------------------------------------------------------------
        // `actions`: a pageable managed array is marshalled (pinned for the call) and staged by the library (one H2D copy); an
        // array registered once with PinActions() is read in place by the kernel -- no per-call pin, no copy.
        public (float[] obs, float[] reward, byte[] done) Step(int[] actions) {
            Native.Check(Native.gymcuda_step(_h, actions, _obs, _reward, _done));
            return (_obs, _reward, _done);
        }
        public (float[] obs, float[] reward, byte[] done) Step(float[] actions) {
            Native.Check(Native.gymcuda_step(_h, actions, _obs, _reward, _done));
            return (_obs, _reward, _done);
        }
        /// <summary>k steps with caller-supplied actions [k][numEnvs] in one launch (gymcuda_step_many); outputs [k][numEnvs]...</summary>
        public void StepMany(int kSteps, int[] actions, float[] obs, float[] reward, byte[] done) {
            Native.Check(Native.gymcuda_step_many(_h, kSteps, actions, obs, reward, done));
        }
        public void StepMany(int kSteps, float[] actions, float[] obs, float[] reward, byte[] done) {
            Native.Check(Native.gymcuda_step_many(_h, kSteps, actions, obs, reward, done));
        }
        /// <summary>Under auto-reset: the array (numEnvs * ObsDim, kept pinned until replaced) receives the observation of the
        /// TERMINAL state of every env whose step returns done; null turns the side buffer off.</summary>
        public void SetTerminalObservations(float[] buffer) {
            if (_terminalPin.IsAllocated) {
                Native.Check(Native.gymcuda_set_terminal_obs(_h, IntPtr.Zero));
                Native.gymcuda_host_unregister(_terminalPin.AddrOfPinnedObject()); _terminalPin.Free();
            }
            if (buffer == null) return;
            if (buffer.Length != NumberOfEnvironments * _info.ObsDim) throw new ArgumentException("expected numEnvs * ObsDim floats", nameof(buffer));
            _terminalPin = GCHandle.Alloc(buffer, GCHandleType.Pinned);
            Native.Check(Native.gymcuda_host_register(_terminalPin.AddrOfPinnedObject(), (UIntPtr) (ulong) (sizeof(float) * buffer.Length)));
            Native.Check(Native.gymcuda_set_terminal_obs(_h, _terminalPin.AddrOfPinnedObject()));
        }
        private GCHandle _terminalPin;

        // Device-resident callers (a learner holding CUDA pointers, e.g. through TorchSharp): one asynchronous step on the handle's
        // stream; dObs = NoObservationCopy skips the observation copy, the observations of the env kinds whose observation is their
        // state vector (CartPoleEnv.cs:183) are then read in place at ObservationViewDevice.
        public static readonly IntPtr NoObservationCopy = Native.NoObs;
        public void StepDevice(IntPtr dActions, IntPtr dObs, IntPtr dReward, IntPtr dDone) {
            Native.Check(Native.gymcuda_step_device(_h, dActions, dObs, dReward, dDone));
        }
        public IntPtr ObservationViewDevice { get { Native.Check(Native.gymcuda_obs_view_device(_h, out IntPtr p)); return p; } }

        public float[] Reset(byte[] mask) { Native.Check(Native.gymcuda_reset_masked(_h, mask, _obs)); return _obs; }
        public void RolloutRandom(int kSteps, float[] obs, float[] reward, byte[] done, int[] actions) {
            Native.Check(Native.gymcuda_rollout_random(_h, kSteps, obs, reward, done, actions));
        }
        public int[] DoneIndices() {
            var idx = new int[NumberOfEnvironments];
            Native.Check(Native.gymcuda_done_indices(_h, idx, out int count));
            Array.Resize(ref idx, count);
            return idx;
        }
        /// <summary>Env.Render for `envIds` (null: the first `count` envs), rasterised on the device: RGB8 [count][height][width][3].</summary>
        public byte[] Render(int[] envIds, int count, int width = 600, int height = 400) {
            var rgb = new byte[count * width * height * 3];
            Native.Check(Native.gymcuda_render(_h, envIds, count, width, height, rgb));
            return rgb;
        }
        public GymCudaStats Stats(bool reset = false) { Native.Check(Native.gymcuda_get_stats(_h, out var s, reset ? 1 : 0)); return s; }

        // ---- observation / reward normalisation on device (what callers otherwise keep by hand, BasePlaySession.cs:58-69)
        public void NormalizeConfig(float gamma = 0.99f, float epsilon = 1e-8f, float clipObs = 10f, float clipReward = 10f) {
            Native.Check(Native.gymcuda_normalize_config(_h, gamma, epsilon, clipObs, clipReward));
        }
        /// <summary>Normalises the arrays in place; update: add this batch to the running statistics first.</summary>
        public void Normalize(float[] obs, float[] reward, byte[] done, bool update = true) {
            Native.Check(Native.gymcuda_normalize(_h, obs, reward, done, update ? 1 : 0));
        }
        public (double[] obsMean, double[] obsVar, double returnVar, double count) NormalizeStats() {
            var mean = new double[_info.ObsDim]; var variance = new double[_info.ObsDim];
            Native.Check(Native.gymcuda_normalize_get(_h, mean, variance, out double rv, out double c));
            return (mean, variance, rv, c);
        }
        public void NormalizeReset() { Native.Check(Native.gymcuda_normalize_reset(_h)); }

        private NDArray[] SplitObservations() {
            var res = new NDArray[NumberOfEnvironments];
            int d = _info.ObsDim;
            for (int i = 0; i < res.Length; i++) {
                var row = new float[d];
                Array.Copy(_obs, i * d, row, 0, d);
                res[i] = np.array(row);
            }
            return res;
        }
    }

    public sealed class GymCudaConfigBox { public GymCudaConfig Value; }

    /// <summary>Batched CartPoleEnv: spaces as CartPoleEnv.cs:46-48.</summary>
    public sealed class CartPoleVecEnv : CudaVecEnv {
        public CartPoleVecEnv(int numEnvs, ulong seed = 0, int device = 0, uint envIdOffset = 0, bool autoReset = false, int timeLimit = 0)
            : base(GymCudaEnvKind.CartPole, numEnvs,
                   new Box(-1 * High(), High(), np.float32), new Discrete(2), seed, device, envIdOffset, autoReset, timeLimit) { }
        private static NDArray High() { return np.array(2.4f * 2, float.MaxValue, (float) (12 * 2 * Math.PI / 360) * 2, float.MaxValue); }
    }

    /// <summary>Batched LunarLanderEnv: spaces as LunarLanderEnv.cs:412-422.</summary>
    public sealed class LunarLanderVecEnv : CudaVecEnv {
        public LunarLanderVecEnv(int numEnvs, bool continuous = false, float gravity = -10f, bool enableWind = false,
                                 float windPower = 15f, float turbulencePower = 1.5f, ulong seed = 0, int device = 0,
                                 uint envIdOffset = 0, bool autoReset = false, int timeLimit = 0)
            : base(continuous ? GymCudaEnvKind.LunarLanderContinuous : GymCudaEnvKind.LunarLander, numEnvs,
                   new Box(np.array(new float[] {-1.5f, -1.5f, -5f, -5f, (float) -Math.PI, -5f, 0f, 0f}),
                           np.array(new float[] {1.5f, 1.5f, 5f, 5f, (float) Math.PI, 5f, 1f, 1f})),
                   continuous ? (Space) new Box(-1f, 1f, new Shape(2)) : new Discrete(4),
                   seed, device, envIdOffset, autoReset, timeLimit,
                   c => { c.Value.Gravity = gravity; c.Value.EnableWind = enableWind ? 1 : 0; c.Value.WindPower = windPower; c.Value.TurbulencePower = turbulencePower; }) { }
    }

    public sealed class PendulumVecEnv : CudaVecEnv {
        public PendulumVecEnv(int numEnvs, ulong seed = 0, int device = 0, uint envIdOffset = 0, bool autoReset = false, int timeLimit = 0)
            : base(GymCudaEnvKind.Pendulum, numEnvs, new Box(np.array(-1f, -1f, -8f), np.array(1f, 1f, 8f), np.float32),
                   new Box(-2f, 2f, new Shape(1)), seed, device, envIdOffset, autoReset, timeLimit) { }
    }

    public sealed class MountainCarVecEnv : CudaVecEnv {
        public MountainCarVecEnv(int numEnvs, bool continuous = false, ulong seed = 0, int device = 0, uint envIdOffset = 0, bool autoReset = false, int timeLimit = 0)
            : base(continuous ? GymCudaEnvKind.MountainCarContinuous : GymCudaEnvKind.MountainCar, numEnvs,
                   new Box(np.array(-1.2f, -0.07f), np.array(0.6f, 0.07f), np.float32),
                   continuous ? (Space) new Box(-1f, 1f, new Shape(1)) : new Discrete(3), seed, device, envIdOffset, autoReset, timeLimit) { }
    }

    public sealed class AcrobotVecEnv : CudaVecEnv {
        public AcrobotVecEnv(int numEnvs, ulong seed = 0, int device = 0, uint envIdOffset = 0, bool autoReset = false, int timeLimit = 0)
            : base(GymCudaEnvKind.Acrobot, numEnvs,
                   new Box(np.array(-1f, -1f, -1f, -1f, (float) (-4 * Math.PI), (float) (-9 * Math.PI)),
                           np.array(1f, 1f, 1f, 1f, (float) (4 * Math.PI), (float) (9 * Math.PI)), np.float32),
                   new Discrete(3), seed, device, envIdOffset, autoReset, timeLimit) { }
    }
}
