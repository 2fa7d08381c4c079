// P/Invoke surface of libgymcuda (include/gymcuda.h).  One [DllImport] per exported symbol; the
// declarations are mechanical so they can be diffed against the header.  NOT compiled in this
// repository's CI: the build image has no .NET SDK (see INTEGRATION.md).
using System;
using System.Runtime.InteropServices;

namespace Gym.Environments.Vector {
    public enum GymCudaStatus : int { Ok = 0, EInval = -1, EAction = -2, ECuda = -3, ENccl = -4, ENoMem = -5, EState = -6 }

    public enum GymCudaEnvKind : int {
        CartPole = 0, Pendulum = 1, MountainCar = 2, MountainCarContinuous = 3, Acrobot = 4,
        LunarLander = 5, LunarLanderContinuous = 6
    }

    [Flags] public enum GymCudaFlags : uint { None = 0, AutoReset = 1, EpisodeStats = 2, DoneBits = 4 }

    [StructLayout(LayoutKind.Sequential)]
    public struct GymCudaConfig {
        public uint StructSize;
        public int EnvKind;
        public int NumEnvs;
        public int Device;
        public ulong Seed;
        public uint EnvIdOffset;
        public uint Flags;
        public int TimeLimit;
        public float Gravity;
        public int EnableWind;
        public float WindPower;
        public float TurbulencePower;
    }

    [StructLayout(LayoutKind.Sequential)]
    public unsafe struct GymCudaSpaceInfo {
        public int ObsDim, ActDim, ActN, StateDim, AuxDim, TimeLimit;
        public fixed float ObsLow[8];
        public fixed float ObsHigh[8];
        public fixed float ActLow[2];
        public fixed float ActHigh[2];
    }

    [StructLayout(LayoutKind.Sequential)]
    public struct GymCudaStats { public ulong EnvSteps, Episodes, InvalidActions; public double ReturnSum; public ulong LengthSum; }

    /// <summary>Owns the native handle; released by gymcuda_destroy.</summary>
    public sealed class GymCudaHandle : SafeHandle {
        public GymCudaHandle() : base(IntPtr.Zero, true) { }
        public override bool IsInvalid => handle == IntPtr.Zero;
        protected override bool ReleaseHandle() { return Native.gymcuda_destroy(handle) == 0; }
    }

    internal static unsafe class Native {
        private const string Lib = "gymcuda";   // libgymcuda.so / gymcuda.dll on the loader path

        [DllImport(Lib)] internal static extern int gymcuda_version();
        [DllImport(Lib)] internal static extern IntPtr gymcuda_last_error();
        [DllImport(Lib)] internal static extern int gymcuda_device_count(out int count);
        [DllImport(Lib)] internal static extern int gymcuda_config_default(out GymCudaConfig cfg, int envKind, int numEnvs);
        [DllImport(Lib)] internal static extern int gymcuda_create(ref GymCudaConfig cfg, out GymCudaHandle env);
        [DllImport(Lib)] internal static extern int gymcuda_destroy(IntPtr env);
        [DllImport(Lib)] internal static extern int gymcuda_space(GymCudaHandle env, out GymCudaSpaceInfo info);
        [DllImport(Lib)] internal static extern int gymcuda_num_envs(GymCudaHandle env);
        [DllImport(Lib)] internal static extern int gymcuda_seed(GymCudaHandle env, ulong seed);
        [DllImport(Lib)] internal static extern int gymcuda_seed_each(GymCudaHandle env, int[] seeds, int n);
        [DllImport(Lib)] internal static extern int gymcuda_reset(GymCudaHandle env, float[] obsOut);
        [DllImport(Lib)] internal static extern int gymcuda_reset_masked(GymCudaHandle env, byte[] mask, float[] obsOut);
        [DllImport(Lib)] internal static extern int gymcuda_step(GymCudaHandle env, int[] actions, float[] obs, float[] reward, byte[] done);
        [DllImport(Lib)] internal static extern int gymcuda_step(GymCudaHandle env, float[] actions, float[] obs, float[] reward, byte[] done);
        [DllImport(Lib)] internal static extern int gymcuda_step_device(GymCudaHandle env, IntPtr dActions, IntPtr dObs, IntPtr dReward, IntPtr dDone);
        [DllImport(Lib)] internal static extern int gymcuda_step_broadcast(GymCudaHandle env, int action, float[] obs, float[] reward, byte[] done);
        [DllImport(Lib)] internal static extern int gymcuda_step_many_device(GymCudaHandle env, int kSteps, IntPtr dActions, IntPtr dObs, IntPtr dReward, IntPtr dDone);
        [DllImport(Lib)] internal static extern int gymcuda_step_many(GymCudaHandle env, int kSteps, int[] actions, float[] obs, float[] reward, byte[] done);
        [DllImport(Lib)] internal static extern int gymcuda_step_many(GymCudaHandle env, int kSteps, float[] actions, float[] obs, float[] reward, byte[] done);
        [DllImport(Lib)] internal static extern int gymcuda_set_terminal_obs(GymCudaHandle env, IntPtr buffer);
        internal static readonly IntPtr NoObs = new IntPtr(1);   // GYMCUDA_NO_OBS: gymcuda_step_device writes no observation copy
        [DllImport(Lib)] internal static extern int gymcuda_obs_view_device(GymCudaHandle env, out IntPtr dObs);
        [DllImport(Lib)] internal static extern int gymcuda_box_sample_device(int device, IntPtr cudaStream, ulong seed, ulong index, IntPtr dLow, IntPtr dHigh, int dim, int count, int asInt, IntPtr dOut);
        [DllImport(Lib)] internal static extern int gymcuda_box_sample(int device, ulong seed, ulong index, float[] low, float[] high, int dim, int count, int asInt, float[] output);
        [DllImport(Lib)] internal static extern int gymcuda_rollout_random_device(GymCudaHandle env, int kSteps, IntPtr dObs, IntPtr dReward, IntPtr dDone, IntPtr dActions);
        [DllImport(Lib)] internal static extern int gymcuda_rollout_random(GymCudaHandle env, int kSteps, float[] obs, float[] reward, byte[] done, int[] actions);
        [DllImport(Lib)] internal static extern int gymcuda_sample_actions(GymCudaHandle env, byte[] mask, int[] actionsOut);
        [DllImport(Lib)] internal static extern int gymcuda_sample_actions(GymCudaHandle env, byte[] mask, float[] actionsOut);
        [DllImport(Lib)] internal static extern int gymcuda_sample_actions_device(GymCudaHandle env, IntPtr dMask, IntPtr dActionsOut);
        [DllImport(Lib)] internal static extern int gymcuda_done_indices(GymCudaHandle env, int[] idx, out int count);
        [DllImport(Lib)] internal static extern int gymcuda_done_indices_device(GymCudaHandle env, out IntPtr dIdx, out IntPtr dCount);
        [DllImport(Lib)] internal static extern int gymcuda_get_state(GymCudaHandle env, float[] state, int[] aux, out ulong t);
        [DllImport(Lib)] internal static extern int gymcuda_set_state(GymCudaHandle env, float[] state, int[] aux, ulong t);
        [DllImport(Lib)] internal static extern int gymcuda_observe(GymCudaHandle env, float[] obs);
        [DllImport(Lib)] internal static extern int gymcuda_render_device(GymCudaHandle env, IntPtr dEnvIds, int count, int width, int height, IntPtr dRgb);
        [DllImport(Lib)] internal static extern int gymcuda_render(GymCudaHandle env, int[] envIds, int count, int width, int height, byte[] rgb);
        [DllImport(Lib)] internal static extern int gymcuda_get_stats(GymCudaHandle env, out GymCudaStats stats, int resetCounters);
        [DllImport(Lib)] internal static extern int gymcuda_normalize_config(GymCudaHandle env, float gamma, float epsilon, float clipObs, float clipReward);
        [DllImport(Lib)] internal static extern int gymcuda_normalize(GymCudaHandle env, float[] obs, float[] reward, byte[] done, int update);
        [DllImport(Lib)] internal static extern int gymcuda_normalize_device(GymCudaHandle env, IntPtr dObs, IntPtr dReward, IntPtr dDone, int update);
        [DllImport(Lib)] internal static extern int gymcuda_normalize_get(GymCudaHandle env, double[] obsMean, double[] obsVar, out double returnVar, out double count);
        [DllImport(Lib)] internal static extern int gymcuda_normalize_reset(GymCudaHandle env);
        [DllImport(Lib)] internal static extern int gymcuda_set_stream(GymCudaHandle env, IntPtr cudaStream);
        [DllImport(Lib)] internal static extern int gymcuda_set_device_clock(GymCudaHandle env, int on);
        [DllImport(Lib)] internal static extern int gymcuda_sync(GymCudaHandle env);
        [DllImport(Lib)] internal static extern int gymcuda_host_alloc(out IntPtr ptr, UIntPtr bytes);
        [DllImport(Lib)] internal static extern int gymcuda_host_free(IntPtr ptr);
        [DllImport(Lib)] internal static extern int gymcuda_host_register(IntPtr ptr, UIntPtr bytes);
        [DllImport(Lib)] internal static extern int gymcuda_host_unregister(IntPtr ptr);
        [DllImport(Lib)] internal static extern int gymcuda_nccl_load([MarshalAs(UnmanagedType.LPStr)] string path);
        [DllImport(Lib)] internal static extern int gymcuda_nccl_unique_id(byte[] id128);
        [DllImport(Lib)] internal static extern int gymcuda_comm_init(GymCudaHandle env, byte[] id128, int rank, int worldSize);
        [DllImport(Lib)] internal static extern int gymcuda_allgather_obs(GymCudaHandle env, IntPtr dObs, IntPtr dOut);
        [DllImport(Lib)] internal static extern int gymcuda_gather_create(GymCudaHandle env, int rank, int worldSize, byte[] handle64);
        [DllImport(Lib)] internal static extern int gymcuda_gather_open(GymCudaHandle env, byte[] handles);
        [DllImport(Lib)] internal static extern int gymcuda_step_gather_device(GymCudaHandle env, IntPtr dActions, IntPtr dReward, IntPtr dDone, out IntPtr dGathered);
        [DllImport(Lib)] internal static extern int gymcuda_gather_wait(GymCudaHandle env);

        /// <summary>Error convention of the boundary: status code -> the reference's exception types.</summary>
        internal static void Check(int status) {
            if (status == 0) return;
            string msg = Marshal.PtrToStringAnsi(gymcuda_last_error()) ?? "gymcuda error";
            switch ((GymCudaStatus) status) {
                case GymCudaStatus.EAction: throw new Gym.Exceptions.InvalidActionError(msg);   // src/Gym/Exceptions/InvalidActionError.cs
                case GymCudaStatus.EInval: throw new ArgumentException(msg);
                case GymCudaStatus.ENoMem: throw new OutOfMemoryException(msg);
                case GymCudaStatus.EState: throw new InvalidOperationException(msg);
                default: throw new ExternalException(msg, status);
            }
        }
    }
}
