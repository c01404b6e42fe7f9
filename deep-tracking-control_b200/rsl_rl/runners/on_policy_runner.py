"""`OnPolicyRunner` (rsl_rl/rsl_rl/runners/on_policy_runner.py:47-273): wiring + the learn loop of SURVEY.md 3.1.

The loop keeps the reference's order of calls.  Differences are confined to host/device traffic: the per-step
`.cpu().numpy().tolist()` episode bookkeeping (:130-140) is done on the device and read once per iteration, and
`update()` reads its statistics once.  Logged quantities (Perf/total_fps etc.) are computed as in :165-247.
"""
import os
import statistics
import time
from collections import deque

import collections

import torch

from ..algorithms import PPO
from ..env.vec_env import check_env
from ..env.wrappers import HistoryWrapper
from ..modules import ActorCriticDecoder


def _nvtx_push(name):
    """Timeline ranges around the three phases of an iteration (the C ABI adds one per entry point below them)."""
    if torch.cuda.is_available():
        torch.cuda.nvtx.range_push(name)


def _nvtx_pop():
    if torch.cuda.is_available():
        torch.cuda.nvtx.range_pop()


class OnPolicyRunner:
    def __init__(self, env, train_cfg, log_dir=None, device="cuda:0"):
        self.cfg = train_cfg["runner"]
        self.alg_cfg = train_cfg["algorithm"]
        self.policy_cfg = train_cfg["policy"]
        self.device = device
        self.env = HistoryWrapper(check_env(env))
        num_critic_obs = self.env.num_privileged_obs if self.env.num_privileged_obs is not None else self.env.num_obs
        actor_critic = ActorCriticDecoder(self.env.num_obs, num_critic_obs, self.env.num_actions, **self.policy_cfg).to(self.device)
        actor_critic.seed = int(getattr(env, "seed", 0))  # per-rank noise stream under data parallelism
        self.alg = PPO(actor_critic, device=self.device, **self.alg_cfg)
        self.num_steps_per_env = self.cfg["num_steps_per_env"]
        self.save_interval = self.cfg["save_interval"]
        self.alg.init_storage(self.env.num_envs, self.num_steps_per_env, [self.env.num_obs], [self.env.num_privileged_obs],
                              [self.env.num_obs_history], [self.env.num_actions])
        self.log_dir = log_dir
        self.writer = None
        self.tot_timesteps = 0
        self.tot_time = 0
        self.current_learning_iteration = 0
        self.last_perf = {}
        # the rollout (T x {PPO.act, env.step, process_env_step}: ~28 launches per step) as ONE CUDA graph, captured after a first
        # eager iteration and replayed from then on; used when nothing in the loop needs the host (no logging, device-side draws)
        self.cuda_graph = bool(self.cfg.get("cuda_graph", True))
        self._graph, self._graph_obs, self._graph_launches, self._eager_iters = None, None, 0, 0
        self.env.reset()

    def learn(self, num_learning_iterations, init_at_random_ep_len=False):
        if self.log_dir is not None and self.writer is None:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self.writer = SummaryWriter(log_dir=self.log_dir, flush_secs=10)
            except Exception:
                self.writer = None
        if init_at_random_ep_len:
            # lands on the wrapper object, exactly as with gym.Wrapper in the reference (:91)
            self.env.episode_length_buf = torch.randint_like(self.env.env.episode_length_buf, high=int(self.env.max_episode_length))
        obs_dict = self.env.get_observations()
        obs, privileged_obs, obs_history = obs_dict["obs"], obs_dict["privileged_obs"], obs_dict["obs_history"]
        self.alg.actor_critic.train()
        ep_infos = []
        rewbuffer, lenbuffer = deque(maxlen=100), deque(maxlen=100)
        N, T, dev = self.env.num_envs, self.num_steps_per_env, self.device
        cur_reward_sum = torch.zeros(N, dtype=torch.float, device=dev)
        cur_episode_length = torch.zeros(N, dtype=torch.float, device=dev)
        rec_rew = torch.zeros(T, N, device=dev)
        rec_len = torch.zeros(T, N, device=dev)
        rec_done = torch.zeros(T, N, dtype=torch.bool, device=dev)
        tot_iter = self.current_learning_iteration + num_learning_iterations
        for it in range(self.current_learning_iteration, tot_iter):
            start = time.time()
            rew_buf = self.env.get_reward_buf()
            with torch.inference_mode():
                _nvtx_push("rollout")
                if self._use_graph():
                    obs_dict = self._graph_rollout(obs_dict, rew_buf, T)
                    obs, privileged_obs, obs_history = obs_dict["obs"], obs_dict["privileged_obs"], obs_dict["obs_history"]
                else:
                    self._eager_iters += 1
                    for i in range(T):
                        actions = self.alg.act(obs, privileged_obs, obs_history, obs_dict["base_vel"], rew_buf)
                        obs_dict, rewards, dones, infos = self.env.step(actions)
                        obs, privileged_obs, obs_history = obs_dict["obs"], obs_dict["privileged_obs"], obs_dict["obs_history"]
                        self.alg.process_env_step(rewards, dones, next_obs=obs_dict["obs"], infos=infos)
                        if self.log_dir is not None:
                            if "episode" in infos:
                                ep_infos.append(infos["episode"])
                            cur_reward_sum += rewards
                            cur_episode_length += 1
                            d = dones > 0
                            rec_rew[i], rec_len[i], rec_done[i] = cur_reward_sum, cur_episode_length, d
                            cur_reward_sum.masked_fill_(d, 0)
                            cur_episode_length.masked_fill_(d, 0)
                _nvtx_pop()
                stop = time.time()
                collection_time = stop - start
                start = stop
                _nvtx_push("compute_returns")
                self.alg.compute_returns(obs, privileged_obs, obs_dict["base_vel"])
                _nvtx_pop()
            _nvtx_push("update")
            (mean_value_loss, mean_surrogate_loss, mean_adaptation_module_loss, mean_decoder_loss, mean_recons_loss, mean_vel_loss,
             mean_kld_loss) = self.alg.update()
            _nvtx_pop()
            stop = time.time()
            learn_time = stop - start
            if self.log_dir is not None:
                m = rec_done.flatten()
                rewbuffer.extend(rec_rew.flatten()[m].tolist())
                lenbuffer.extend(rec_len.flatten()[m].tolist())
                self.log(locals())
                if it % self.save_interval == 0:
                    self.save(os.path.join(self.log_dir, "model_{}.pt".format(it)))
            ep_infos.clear()
        self.current_learning_iteration += num_learning_iterations
        if self.log_dir is not None:
            self.save(os.path.join(self.log_dir, "model_{}.pt".format(self.current_learning_iteration)))

    # ------------------------------------------------------------------ CUDA graph of the rollout
    def _use_graph(self):
        env, ac = self.env.env, self.alg.actor_critic
        ok = (self.cuda_graph and self.log_dir is None and self._eager_iters >= 1 and getattr(env, "graph_supported", lambda: False)()
              and ac._inject is None and self.alg._inject is None)
        if not ok and self._graph is not None:
            self._graph = None  # conditions changed (e.g. a test injected draws): fall back to eager launches for good
            self.cuda_graph = False
        return ok

    def reset_graph(self):
        """Drop the captured rollout (call after changing anything the capture baked in, e.g. where the simulator tensors come
        from): the next iteration runs eagerly, the one after it captures again."""
        self._graph, self._graph_obs, self._eager_iters = None, None, 0

    def _graph_rollout(self, obs_dict, rew_buf, T):
        from ... import _lib as B
        env, ac, st = self.env.env, self.alg.actor_critic, self.alg.storage
        gym = getattr(env, "gym", None)
        if self._graph is None:
            torch.cuda.synchronize(self.device)
            env.graph_capture_begin()
            ac.graph_capture_begin()
            l0 = B.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                od = obs_dict
                for i in range(T):
                    if gym is not None and hasattr(gym, "graph_step"):
                        gym.graph_step = (i, T)
                    actions = self.alg.act(od["obs"], od["privileged_obs"], od["obs_history"], od["base_vel"], rew_buf)
                    od, rewards, dones, infos = self.env.step(actions)
                    self.alg.process_env_step(rewards, dones, next_obs=od["obs"], infos=infos)
                env.graph_capture_end(T)
                ac.graph_capture_end(T)
            if gym is not None and hasattr(gym, "graph_step"):
                gym.graph_step = None
            self._graph, self._graph_obs, self._graph_launches = g, od, B.launch_count() - l0
            # the capture pass advanced the host-side counters (storage step, common_step_counter, act calls) exactly as a real
            # rollout does but launched nothing: the first replay does this iteration's work
            env.graph_before_replay()
            g.replay()
        else:
            env.graph_before_replay()
            self._graph.replay()
            B.lib().dtc_count_launches(self._graph_launches)
            env.graph_replayed(T)
            ac.graph_replayed(T)
            st.step = T
        env.graph_after_replay()
        return self._graph_obs

    def log(self, locs, width=80, pad=35):
        self.tot_timesteps += self.num_steps_per_env * self.env.num_envs
        self.tot_time += locs["collection_time"] + locs["learn_time"]
        iteration_time = locs["collection_time"] + locs["learn_time"]
        fps = int(self.num_steps_per_env * self.env.num_envs / (locs["collection_time"] + locs["learn_time"]))
        mean_std = self.alg.actor_critic.std.mean().item()
        scal = {"Loss/value_function": locs["mean_value_loss"], "Loss/surrogate": locs["mean_surrogate_loss"],
                "Loss/Reconstruction": locs["mean_recons_loss"], "Loss/Vel_estimation": locs["mean_vel_loss"],
                "Loss/KL_div": locs["mean_kld_loss"], "Loss/learning_rate": self.alg.learning_rate,
                "Policy/mean_noise_std": mean_std, "Perf/total_fps": fps, "Perf/collection time": locs["collection_time"],
                "Perf/learning_time": locs["learn_time"]}
        ep_string = ""
        if locs["ep_infos"]:
            for key in locs["ep_infos"][0]:
                vals = [torch.as_tensor(ep[key]).float().reshape(-1) for ep in locs["ep_infos"]]
                value = torch.cat(vals).mean().item()
                scal["Episode/" + key] = value
                ep_string += f"""{f'Mean episode {key}:':>{pad}} {value:.4f}\n"""
        if len(locs["rewbuffer"]) > 0:
            scal["Train/mean_reward"] = statistics.mean(locs["rewbuffer"])
            scal["Train/mean_episode_length"] = statistics.mean(locs["lenbuffer"])
        self.last_perf = scal
        if self.writer is not None:
            for k, v in scal.items():
                self.writer.add_scalar(k, v, locs["it"])
        head = f" \033[1m Learning iteration {locs['it']}/{locs['tot_iter']} \033[0m "
        log_string = (f"""{'#' * width}\n{head.center(width, ' ')}\n\n"""
                      f"""{'Computation:':>{pad}} {fps:.0f} steps/s (collection: {locs['collection_time']:.3f}s, learning {locs['learn_time']:.3f}s)\n"""
                      f"""{'Value function loss:':>{pad}} {locs['mean_value_loss']:.4f}\n"""
                      f"""{'Surrogate loss:':>{pad}} {locs['mean_surrogate_loss']:.4f}\n"""
                      f"""{'Mean action noise std:':>{pad}} {mean_std:.2f}\n""")
        if len(locs["rewbuffer"]) > 0:
            log_string += (f"""{'Mean reward:':>{pad}} {scal['Train/mean_reward']:.2f}\n"""
                           f"""{'Mean episode length:':>{pad}} {scal['Train/mean_episode_length']:.2f}\n""")
        log_string += ep_string
        log_string += (f"""{'-' * width}\n{'Total timesteps:':>{pad}} {self.tot_timesteps}\n"""
                       f"""{'Iteration time:':>{pad}} {iteration_time:.2f}s\n{'Total time:':>{pad}} {self.tot_time:.2f}s\n""")
        print(log_string)

    def save(self, path, infos=None):
        """Same checkpoint dict as the reference (:249-255): model + main optimizer only (the VAE optimizer state is
        not saved there either)."""
        osd = self.alg.optimizer.state_dict()  # torch.optim.Adam's own layout (per-parameter exp_avg / exp_avg_sq / step)
        osd["state"] = {i: {k: v.cpu() for k, v in st.items()} for i, st in osd["state"].items()}
        torch.save({"model_state_dict": collections.OrderedDict((k, v.cpu()) for k, v in self.alg.actor_critic.state_dict().items()),
                    "optimizer_state_dict": osd, "iter": self.current_learning_iteration, "infos": infos}, path)

    def load(self, path, load_optimizer=True):
        loaded_dict = torch.load(path, map_location="cpu", weights_only=True)
        self.alg.actor_critic.load_state_dict(loaded_dict["model_state_dict"])
        if load_optimizer:
            self.alg.optimizer.load_state_dict(loaded_dict["optimizer_state_dict"])
        self.current_learning_iteration = loaded_dict["iter"]
        self.alg.sync_replicas()
        return loaded_dict["infos"]

    def get_inference_policy(self, device=None):
        self.alg.actor_critic.eval()
        return self.alg.actor_critic.act_inference
