"""HistoryWrapper (rsl_rl/rsl_rl/env/wrappers/history_wrapper.py:6-53).

The 5-frame shift-concat of `step` (:23) is fused into the environment's observation kernel (dtc_env_observe), which
keeps the history in `env.obs_history`; this wrapper only hands out the dict.  `get_observations` shifts once more
and `reset` clears the history, exactly like the reference (both are called once per run, torch ops are fine).
As in the reference, attribute writes on the wrapper (e.g. the runner's `episode_length_buf = ...`,
on_policy_runner.py:91) land on the wrapper object, not on the environment."""
import torch


class HistoryWrapper:
    def __init__(self, env):
        self.env = env
        self.obs_history_length = self.env.cfg.env.num_observation_history
        self.num_obs_history = self.obs_history_length * self.env.num_obs

    def __getattr__(self, name):
        if name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def obs_history(self):
        return self.env.obs_history

    def _dict(self, obs, privileged_obs):
        return {"obs": obs, "privileged_obs": privileged_obs, "obs_history": self.env.obs_history,
                "base_vel": self.env.get_base_vel()}

    def step(self, action):
        obs, privileged_obs, rew, done, info = self.env.step(action)
        return self._dict(obs, privileged_obs), rew, done, info

    def get_observations(self):
        obs = self.env.get_observations()
        privileged_obs = self.env.get_privileged_observations()
        h = self.env.obs_history
        h.copy_(torch.cat((h[:, self.env.num_obs:], obs), dim=-1))
        return self._dict(obs, privileged_obs)

    def reset_idx(self, env_ids):
        ret = self.env.reset_idx(env_ids)
        self.env.obs_history[env_ids, :] = 0
        return ret

    def reset(self):
        ret = self.env.reset()
        privileged_obs = self.env.get_privileged_observations()
        self.env.obs_history[:, :] = 0
        return {"obs": ret, "privileged_obs": privileged_obs, "obs_history": self.env.obs_history,
                "base_vel": self.env.get_base_vel()}

    def get_reward_buf(self):
        return self.env.get_reward_buf()
