// ReferenceReplay -- dumps what the UNMODIFIED reference computes, in the CSV layout tests/tools/replay_reference_dump.py reads.
//
//   cartpole_reference.csv      teacher-forced single steps of CartPoleEnv.Step (CartPoleEnv.cs:137-186): the private fields
//                               `state` / `steps_beyond_done` (:40-41) are set by reflection, so every row is one transition
//                               state, action, sbd -> next state (17 significant digits), reward, done, next sbd
//   cartpole_loop_seed<S>.csv   the reference test's loop (tests/Gym.Tests/Envs/Classic/CartpoleEnvironment.cs:19-26): i % 2
//                               actions, reset on done; with the Reset() states, so the replay is independent of NumSharp's stream
//   lunar_<mode>_seed<S>.csv    LunarLanderEnv episodes (random policy and the PID loop of
//                               tests/Gym.Tests/Envs/Aether/LunarLanderEnvironment.cs:102-150) TOGETHER WITH the random draws the
//                               env consumed: a twin NumPyRandom with the same seed is advanced in lockstep with the env's own
//                               (ctor: two randint, LunarLanderEnv.cs:409-410; Reset: fx, fy, 12 heights, :496,507; every Step: two
//                               uniforms, :611-612), so the replay feeds the very same numbers to the engines under test and the
//                               comparison pins Aether.Physics2D's World.Step alone, not NumSharp's generator.
//
// Nothing here is part of the product or of the reference; it only calls the reference's public API (+ two private fields).
using System;
using System.Globalization;
using System.IO;
using System.Reflection;
using System.Text;
using Gym.Environments.Envs.Aether;
using Gym.Environments.Envs.Classic;
using Gym.Rendering;   // NullEnvViewer
using NumSharp;

internal static class Program {
    private static readonly CultureInfo Inv = CultureInfo.InvariantCulture;
    private static string R(double v) { return v.ToString("R", Inv); }
    private static string R(float v) { return v.ToString("R", Inv); }

    private static void CartPoleTeacherForced(string dir, int rows, int seed) {
        var env = new CartPoleEnv(NullEnvViewer.Factory);
        var fState = typeof(CartPoleEnv).GetField("state", BindingFlags.NonPublic | BindingFlags.Instance);
        var fSbd = typeof(CartPoleEnv).GetField("steps_beyond_done", BindingFlags.NonPublic | BindingFlags.Instance);
        var rng = new Random(seed);
        var sb = new StringBuilder("x,x_dot,theta,theta_dot,action,sbd,nx,nx_dot,ntheta,ntheta_dot,reward,done,next_sbd\n");
        env.Reset();
        for (int i = 0; i < rows; i++) {
            // float32 states (what the engine stores), spread over and beyond the non-terminal box; a third sits next to a threshold
            double[] s = {
                (float) ((rng.NextDouble() * 2 - 1) * 2.6), (float) ((rng.NextDouble() * 2 - 1) * 3.0),
                (float) ((rng.NextDouble() * 2 - 1) * 0.25), (float) ((rng.NextDouble() * 2 - 1) * 3.5)};
            if (i % 3 == 1) s[0] = (float) ((rng.Next(2) * 2 - 1) * 2.4f - 0.02f * s[1] + (rng.NextDouble() - 0.5) * 1e-6);
            if (i % 3 == 2) s[2] = (float) ((rng.Next(2) * 2 - 1) * 0.20943952f - 0.02f * s[3] + (rng.NextDouble() - 0.5) * 1e-7);
            int action = rng.Next(2), sbd = rng.Next(-1, 3);
            fState.SetValue(env, np.array(s));
            fSbd.SetValue(env, sbd);
            var (obs, reward, done, _) = env.Step(action);
            sb.Append(R(s[0])).Append(',').Append(R(s[1])).Append(',').Append(R(s[2])).Append(',').Append(R(s[3])).Append(',')
              .Append(action).Append(',').Append(sbd).Append(',')
              .Append(R(obs.GetDouble(0))).Append(',').Append(R(obs.GetDouble(1))).Append(',').Append(R(obs.GetDouble(2))).Append(',').Append(R(obs.GetDouble(3))).Append(',')
              .Append(R(reward)).Append(',').Append(done ? 1 : 0).Append(',').Append((int) fSbd.GetValue(env)).Append('\n');
        }
        File.WriteAllText(Path.Combine(dir, "cartpole_reference.csv"), sb.ToString());
    }

    private static void CartPoleLoop(string dir, int seed, int steps) {
        var env = new CartPoleEnv(NullEnvViewer.Factory);
        env.Seed(seed);
        var sb = new StringBuilder("step,action,reset,x,x_dot,theta,theta_dot,reward,done\n");
        var o = env.Reset();
        sb.Append("-1,-1,1,").Append(R(o.GetDouble(0))).Append(',').Append(R(o.GetDouble(1))).Append(',').Append(R(o.GetDouble(2))).Append(',').Append(R(o.GetDouble(3))).Append(",0,0\n");
        for (int i = 0; i < steps; i++) {
            var (obs, reward, done, _) = env.Step(i % 2);
            sb.Append(i).Append(',').Append(i % 2).Append(",0,").Append(R(obs.GetDouble(0))).Append(',').Append(R(obs.GetDouble(1))).Append(',')
              .Append(R(obs.GetDouble(2))).Append(',').Append(R(obs.GetDouble(3))).Append(',').Append(R(reward)).Append(',').Append(done ? 1 : 0).Append('\n');
            if (done) {
                o = env.Reset();
                sb.Append(i).Append(",-1,1,").Append(R(o.GetDouble(0))).Append(',').Append(R(o.GetDouble(1))).Append(',').Append(R(o.GetDouble(2))).Append(',').Append(R(o.GetDouble(3))).Append(",0,0\n");
            }
        }
        File.WriteAllText(Path.Combine(dir, "cartpole_loop_seed" + seed + ".csv"), sb.ToString());
    }

    // the heuristic of the reference's own test (LunarLanderEnvironment.cs:102-150), discrete branch
    private static int Pid(float[] s) {
        float angleTarg = Math.Max(-0.4f, Math.Min(0.4f, s[0] * 0.5f + s[2] * 1f));
        float hoverTarg = 0.55f * Math.Abs(s[0]);
        float angleTodo = (angleTarg - s[4]) * 0.5f - s[5] * 1f;
        float hoverTodo = (hoverTarg - s[1]) * 0.5f - s[3] * 0.5f;
        if (s[6] > 0f || s[7] > 0f) { angleTodo = 0f; hoverTodo = -s[3] * 0.5f; }
        if (hoverTodo > Math.Abs(angleTodo) && hoverTodo > 0.05f) return 2;
        if (angleTodo < -0.05f) return 3;
        if (angleTodo > 0.05f) return 1;
        return 0;
    }

    private static void Lunar(string dir, int seed, bool pid, int maxSteps) {
        const float W = 600f / 30f, H = 400f / 30f;
        var envRng = np.random.RandomState(seed);
        var twin = np.random.RandomState(seed);   // advanced in lockstep: records what the env draws
        var env = new LunarLanderEnv(NullEnvViewer.Factory, random_state: envRng);
        int windIdx = twin.randint(-9999, 9999), torqueIdx = twin.randint(-9999, 9999);   // :409-410
        var sb = new StringBuilder();
        sb.Append("# seed=").Append(seed).Append(" policy=").Append(pid ? "pid" : "random").Append(" wind_idx=").Append(windIdx).Append(" torque_idx=").Append(torqueIdx).Append('\n');
        sb.Append("kind,action,d0,d1,o0,o1,o2,o3,o4,o5,o6,o7,reward,done,extra...\n");
        var actRng = new Random(seed + 1);
        var obs = (float[]) env.Reset();
        {   // Reset draws (:496, :504-508), then the zero step's two dispersion draws (:611-612)
            float fx = (float) twin.uniform(-1000f, 1000f), fy = (float) twin.uniform(-1000f, 1000f);
            sb.Append("reset,0,");
            var heights = new float[12];
            for (int i = 0; i < 12; i++) heights[i] = (float) twin.uniform(0, H / 2);
            float d0 = (float) twin.uniform(-1.0, 1.0), d1 = (float) twin.uniform(-1.0, 1.0);
            sb.Append(R(d0)).Append(',').Append(R(d1));
            foreach (var o in obs) sb.Append(',').Append(R(o));
            sb.Append(",0,0,").Append(R(fx)).Append(',').Append(R(fy));
            foreach (var h in heights) sb.Append(',').Append(R(h));
            sb.Append('\n');
        }
        float total = 0f;
        for (int step = 0; step < maxSteps; step++) {
            int a = pid ? Pid(obs) : actRng.Next(4);
            float d0 = (float) twin.uniform(-1.0, 1.0), d1 = (float) twin.uniform(-1.0, 1.0);
            var (o, reward, done, _) = env.Step(a);
            obs = (float[]) o;
            total += reward;
            sb.Append("step,").Append(a).Append(',').Append(R(d0)).Append(',').Append(R(d1));
            foreach (var v in obs) sb.Append(',').Append(R(v));
            sb.Append(',').Append(R(reward)).Append(',').Append(done ? 1 : 0).Append('\n');
            if (done) break;
        }
        sb.Append("# total_reward=").Append(R(total)).Append('\n');
        File.WriteAllText(Path.Combine(dir, "lunar_" + (pid ? "pid" : "random") + "_seed" + seed + ".csv"), sb.ToString());
        env.CloseEnvironment();
    }

    private static int Main(string[] args) {
        string dir = args.Length > 0 ? args[0] : "reference_dump";
        Directory.CreateDirectory(dir);
        CartPoleTeacherForced(dir, 200000, 7);
        CartPoleLoop(dir, 0, 1000);
        Lunar(dir, 1000, true, 5000);    // the reference's golden: 1547 steps, 184.01764 (LunarLanderEnvironment.cs:32-34)
        foreach (int seed in new[] {1, 2, 3, 4, 5, 6, 7, 8}) { Lunar(dir, seed, true, 5000); Lunar(dir, seed, false, 1000); }
        Console.WriteLine("wrote " + dir);
        return 0;
    }
}
