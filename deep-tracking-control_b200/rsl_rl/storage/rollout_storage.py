"""`RolloutStorage` with the reference's surface (rsl_rl/rsl_rl/storage/rollout_storage.py:36-214) over a packed,
GEMM-ready device layout (include/dtc_b200.h: dtc_storage):

    hist    [T,N,268]  observation_histories 265 + 3 zero
    priv_a  [T,N,696]  privileged_observations[:, :693] + 3 zero        (terrain-encoder input)
    xc      [T,N,752]  privileged_observations[:, 693:] | observations 53 | base_vel 3   (critic input)

so a transition is written once and read by the first-layer GEMMs without a concatenation copy.  The reference's
public tensor attributes are exposed as views (or, for `privileged_observations`, materialised on access).
"""
import ctypes as C

import torch

from ... import _lib as B


class RolloutStorage:
    class Transition:
        def __init__(self):
            self.observations = None
            self.next_observations = None
            self.privileged_observations = None
            self.observation_histories = None
            self.critic_observations = None
            self.actions = None
            self.rewards = None
            self.dones = None
            self.values = None
            self.actions_log_prob = None
            self.action_mean = None
            self.action_sigma = None
            self.hidden_states = None
            self.base_vel = None

        def clear(self):
            self.__init__()

    def __init__(self, num_envs, num_transitions_per_env, obs_shape, privileged_obs_shape, obs_history_shape, actions_shape,
                 device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise B.DtcError("RolloutStorage lives on a CUDA device only (no CPU fallback)")
        if [list(obs_shape), list(privileged_obs_shape), list(obs_history_shape), list(actions_shape)] != [[53], [1389], [265], [12]]:
            raise ValueError("the packed layout is specialised for the DTC shapes [53], [1389], [265], [12]")
        self.obs_shape, self.privileged_obs_shape = obs_shape, privileged_obs_shape
        self.obs_history_shape, self.actions_shape = obs_history_shape, actions_shape
        T, N = num_transitions_per_env, num_envs
        self.num_transitions_per_env, self.num_envs = T, N
        f = lambda *s: torch.zeros(*s, device=self.device, dtype=torch.float32)
        self.hist, self.priv_a, self.xc, self.next_obs_p = f(T, N, 268), f(T, N, 696), f(T, N, 752), f(T, N, 56)
        # 3xTF32 companions of the GEMM-input arrays (csrc/dtc_gemm_tc.cu), maintained by the kernels
        self._hist_lo, self._priv_a_lo, self._xc_lo = f(T, N, 268), f(T, N, 696), f(T, N, 752)
        self.actions, self.mu, self.sigma = f(T, N, 12), f(T, N, 12), f(T, N, 12)
        self.rewards, self.values, self.returns = f(T, N, 1), f(T, N, 1), f(T, N, 1)
        self.advantages, self.actions_log_prob = f(T, N, 1), f(T, N, 1)
        self.dones = torch.zeros(T, N, 1, device=self.device, dtype=torch.uint8)
        self.saved_hidden_states_a = self.saved_hidden_states_c = None
        self.step = 0
        self._scratch = torch.zeros(4, device=self.device, dtype=torch.float64)
        self._batch = None
        self._inject_perm = None  # tests: permutation used by the next minibatch pass
        self.generator = None
        self._c = self._make_struct()

    def _make_struct(self):
        s = B.Storage()
        for name, t in (("hist", self.hist), ("priv_a", self.priv_a), ("xc", self.xc), ("next_obs", self.next_obs_p),
                        ("actions", self.actions), ("mu", self.mu), ("sigma", self.sigma), ("rewards", self.rewards),
                        ("values", self.values), ("returns", self.returns), ("advantages", self.advantages),
                        ("logp", self.actions_log_prob), ("dones", self.dones), ("hist_lo", self._hist_lo),
                        ("priv_a_lo", self._priv_a_lo), ("xc_lo", self._xc_lo)):
            assert t.is_contiguous()
            setattr(s, name, t.data_ptr())
        s.T, s.N = self.num_transitions_per_env, self.num_envs
        return s

    # ------------------------------------------------------------------ reference-shaped views
    @property
    def observations(self):
        return self.xc[..., 696:749]

    @property
    def base_vel(self):
        return self.xc[..., 749:752]

    @property
    def observation_histories(self):
        return self.hist[..., :265]

    @property
    def next_observations(self):
        return self.next_obs_p[..., :53]

    @property
    def privileged_observations(self):
        return torch.cat((self.priv_a[..., :693], self.xc[..., :696]), dim=-1)

    # ------------------------------------------------------------------ reference methods
    def add_transitions(self, transition):
        """Generic path with torch copies (API parity).  The training loop writes transitions through the fused
        kernels instead: dtc_policy_act stores the act-time half, dtc_store_transition the env-time half."""
        if self.step >= self.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        s, t = self.step, transition
        self.hist[s, :, :265].copy_(t.observation_histories)
        self.priv_a[s, :, :693].copy_(t.privileged_observations[:, :693])
        self.xc[s, :, :696].copy_(t.privileged_observations[:, 693:])
        self.xc[s, :, 696:749].copy_(t.observations)
        self.xc[s, :, 749:752].copy_(t.base_vel)
        for x, lo in ((self.hist, self._hist_lo), (self.priv_a, self._priv_a_lo), (self.xc, self._xc_lo)):
            r = x[s] - (x[s].view(torch.int32) & -8192).view(torch.float32)            # x - trunc_tf32(x)
            lo[s] = ((r.view(torch.int32) + 0x1000) & -8192).view(torch.float32)     # rounded to TF32
        self.next_obs_p[s, :, :53].copy_(t.next_observations)
        self.actions[s].copy_(t.actions)
        self.rewards[s].copy_(t.rewards.view(-1, 1))
        self.dones[s].copy_(t.dones.view(-1, 1))
        self.values[s].copy_(t.values.view(-1, 1))
        self.actions_log_prob[s].copy_(t.actions_log_prob.view(-1, 1))
        self.mu[s].copy_(t.action_mean)
        self.sigma[s].copy_(t.action_sigma)
        self.step += 1

    def clear(self):
        self.step = 0

    def compute_returns(self, last_values, gamma, lam, group=None):
        """GAE scan + advantage normalisation in one or (data-parallel) two launches (rollout_storage.py:138-152)."""
        lv = last_values.reshape(-1).contiguous().float()
        B.require_cuda(lv, "last_values")
        st = B.stream_ptr(self.device)
        from ..utils import dp
        multi = dp.world_size(group) > 1
        B.check(B.lib().dtc_gae(C.byref(self._c), B.ptr(lv), gamma, lam, B.ptr(self._scratch), 1 if multi else 0, st), "dtc_gae")
        if multi:
            dp.combine_moments_(self._scratch[:3], group)
            B.check(B.lib().dtc_gae_normalize(C.byref(self._c), B.ptr(self._scratch), st), "dtc_gae_normalize")

    def get_statistics(self):
        done = self.dones.clone()
        done[-1] = 1
        flat_dones = done.permute(1, 0, 2).reshape(-1, 1)
        done_indices = torch.cat((flat_dones.new_tensor([-1], dtype=torch.int64), flat_dones.nonzero(as_tuple=False)[:, 0]))
        trajectory_lengths = done_indices[1:] - done_indices[:-1]
        return trajectory_lengths.float().mean(), self.rewards.mean()

    def gather(self, indices):
        """Packed copy of the rows `indices` (a permutation prefix) - the minibatch gather of :165-214 done once per
        update for all minibatches (the reference reuses one permutation for every epoch)."""
        rows = int(indices.numel())
        if self._batch is None or self._batch.num_envs != rows:
            self._batch = _Batch(rows, self.device)
        idx = indices.to(self.device, torch.int64).contiguous()
        B.check(B.lib().dtc_gather_minibatch(C.byref(self._c), C.byref(self._batch._c), B.ptr(idx), rows, B.stream_ptr(self.device)),
                "dtc_gather_minibatch")
        return self._batch

    def draw_permutation(self, n):
        if self._inject_perm is not None:
            p, self._inject_perm = self._inject_perm, None
            return p.to(self.device)
        return torch.randperm(n, requires_grad=False, device=self.device, generator=self.generator)

    def mini_batch_generator(self, num_mini_batches, num_epochs=8):
        batch_size = self.num_envs * self.num_transitions_per_env
        mini_batch_size = batch_size // num_mini_batches
        indices = self.draw_permutation(num_mini_batches * mini_batch_size)
        b = self.gather(indices)
        for _ in range(num_epochs):
            for i in range(num_mini_batches):
                sl = slice(i * mini_batch_size, (i + 1) * mini_batch_size)
                obs = b.xc[0, sl, 696:749]
                priv = torch.cat((b.priv_a[0, sl, :693], b.xc[0, sl, :696]), dim=-1)
                yield (obs, obs, priv, b.hist[0, sl, :265], b.actions[0, sl], b.values[0, sl], b.advantages[0, sl],
                       b.returns[0, sl], b.actions_log_prob[0, sl], b.mu[0, sl], b.sigma[0, sl], b.xc[0, sl, 749:752],
                       b.next_obs_p[0, sl, :53], (None, None), None, b.rewards[0, sl])


class _Batch(RolloutStorage):
    """Gather destination: a one-step storage with `rows` "environments"."""

    def __init__(self, rows, device):
        RolloutStorage.__init__(self, rows, 1, [53], [1389], [265], [12], device=device)
