"""Device-time split of one training iteration (CUDA events): rollout (24 x PPO.act + env.step + store), returns, update."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    env, fg, runner, state, pool_host, pool_dev = bench.build_world(N, 0, "cuda:0")
    runner.learn(3)
    alg, wenv = runner.alg, runner.env
    obs_dict = wenv.get_observations()
    obs, priv, hist = obs_dict["obs"], obs_dict["privileged_obs"], obs_dict["obs_history"]
    ev = lambda: torch.cuda.Event(enable_timing=True)
    out = {"rollout": [], "act": [], "env": [], "returns": [], "update": []}
    for it in range(4):
        e = [ev() for _ in range(4)]
        ea, eb = [], []
        e[0].record()
        with torch.inference_mode():
            for i in range(24):
                a0, a1, a2 = ev(), ev(), ev()
                a0.record()
                actions = alg.act(obs, priv, hist, obs_dict["base_vel"], None)
                a1.record()
                obs_dict, rewards, dones, infos = wenv.step(actions)
                obs, priv, hist = obs_dict["obs"], obs_dict["privileged_obs"], obs_dict["obs_history"]
                alg.process_env_step(rewards, dones, next_obs=obs_dict["obs"], infos=infos)
                a2.record()
                ea.append((a0, a1)); eb.append((a1, a2))
            e[1].record()
            alg.compute_returns(obs, priv, obs_dict["base_vel"])
        e[2].record()
        alg.update()
        e[3].record()
        torch.cuda.synchronize()
        out["rollout"].append(e[0].elapsed_time(e[1])); out["returns"].append(e[1].elapsed_time(e[2])); out["update"].append(e[2].elapsed_time(e[3]))
        out["act"].append(sum(a.elapsed_time(b) for a, b in ea)); out["env"].append(sum(a.elapsed_time(b) for a, b in eb))
    print(json.dumps({k: round(sorted(v)[len(v) // 2], 3) for k, v in out.items()}))

if __name__ == "__main__":
    main()
