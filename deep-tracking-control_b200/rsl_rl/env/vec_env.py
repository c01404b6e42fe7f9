"""What `OnPolicyRunner`, `HistoryWrapper` and `PPO` require of an environment - the duck-typed surface of SURVEY.md 8b
(reference: rsl_rl/rsl_rl/env/vec_env.py:36-59 for the names; history_wrapper.py:6-53 and on_policy_runner.py:57-128 for how
they are used).  `VecEnv` keeps the reference's name so `isinstance` / subclass checks written against it still work;
`check_env` is what this package's runner calls to fail early, with the missing member named, instead of deep inside a
CUDA launch."""
from abc import ABC, abstractmethod

import torch

# attribute -> meaning, as consumed by the training loop
ENV_ATTRIBUTES = {
    "num_envs": "environments stepped in lock-step (rows of every buffer)",
    "num_obs": "width of obs_buf (53 for Lite3 DTC)",
    "num_privileged_obs": "width of privileged_obs_buf (1389), or None",
    "num_actions": "width of the action tensor (12)",
    "max_episode_length": "steps before a time-out (runner randomises episode_length_buf up to it)",
    "obs_buf": "float32 [num_envs, num_obs], overwritten by every step()",
    "privileged_obs_buf": "float32 [num_envs, num_privileged_obs] or None",
    "rew_buf": "float32 [num_envs]",
    "reset_buf": "done flags [num_envs]",
    "episode_length_buf": "int64 [num_envs], current episode duration (written by the runner)",
    "extras": "dict with 'time_outs' and optionally 'episode'",
    "device": "torch.device of all buffers",
}
ENV_METHODS = ("step", "reset", "get_observations", "get_privileged_observations")


class VecEnv(ABC):
    """Abstract base with the four calls of the contract; the attributes of ENV_ATTRIBUTES are plain instance members."""

    @abstractmethod
    def step(self, actions: torch.Tensor):
        """-> (obs, privileged_obs | None, rewards, dones, extras); the returned tensors are views that the next step overwrites."""

    @abstractmethod
    def reset(self, env_ids=None):
        """Resets the given (default: all) environments and returns the first observations."""

    @abstractmethod
    def get_observations(self) -> torch.Tensor:
        """Current obs_buf (a dict with obs / privileged_obs / obs_history / base_vel behind HistoryWrapper)."""

    @abstractmethod
    def get_privileged_observations(self):
        """Current privileged_obs_buf, or None when the task has none."""


def check_env(env):
    """Raises TypeError naming every member of the contract that `env` lacks."""
    missing = [a for a in ENV_ATTRIBUTES if not hasattr(env, a)] + [m + "()" for m in ENV_METHODS if not callable(getattr(env, m, None))]
    if missing:
        raise TypeError(f"{type(env).__name__} does not satisfy the VecEnv contract: missing {', '.join(missing)}")
    return env
