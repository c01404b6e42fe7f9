"""`PPO` with the reference's surface (rsl_rl/rsl_rl/algorithms/ppo.py:42-357) over the fused learner kernels.

act()            -> dtc_policy_act: encoder + actor + critic + sampling + log-prob, written straight into the
                    rollout storage (the act-time half of add_transitions)
process_env_step -> dtc_store_transition (timeout bootstrap + env-time half of add_transitions)
compute_returns  -> dtc_policy_evaluate + dtc_gae
update()         -> one gather for all minibatches, then per minibatch dtc_vae_step + dtc_ppo_step (forward,
                    hand-written backward, clip_grad_norm_, Adam, adaptive-KL learning rate - all on the device; the
                    host never waits for a value inside the loop)

Data parallel: when torch.distributed is initialised with world_size > 1 each optimizer step stops
after backward, the flat gradient range (plus the KL sum riding behind it) is all-reduced over NCCL, and
dtc_optimizer_apply finishes with grad_scale = 1/world.  The advantage moments are all-reduced once per iteration.
"""
import ctypes as C
import os

import torch

from ... import _lib as B
from ..storage import RolloutStorage
from ..utils import dp


class _DeviceAdamHandle:
    """`alg.optimizer` / `alg.vae_optimizer` of the reference (torch.optim.Adam, ppo.py:78-79), reduced to what the runner's
    save/load touches.  `state_dict()` / `load_state_dict()` speak torch.optim.Adam's own format - parameter index i is the
    i-th entry of `actor_critic.parameters()` (main) / `actor_critic.vae.parameters()` (vae), i.e. of the state_dict key order
    - so `model_*.pt` files written by the reference's OnPolicyRunner.save (on_policy_runner.py:249-255) resume here and
    files written here resume there.  Only parameters that receive a gradient in that optimizer's step carry state, as in
    the reference (the others keep `.grad is None` and torch skips them)."""

    def __init__(self, alg, which):
        self._alg, self._which = alg, which

    def _keys(self):
        from ..modules.actor_critic_decoder import STATE_KEYS
        ac = self._alg.actor_critic
        keys = [k for k in STATE_KEYS if self._which == "main" or k.startswith("vae.")]
        b, e = ac._table.ranges["policy" if self._which == "main" else "vae"]
        live = []
        for i, k in enumerate(keys):
            idx = ac._table.index[k]
            lo, hi = int(idx.min()), int(idx.max())
            if lo >= b and hi < e:
                live.append((i, k))
        return keys, live

    def _mv(self):
        ac = self._alg.actor_critic
        return ac._adam[0 if self._which == "main" else 2], ac._adam[1 if self._which == "main" else 3]

    def _steps(self):
        ac = self._alg.actor_critic
        if self._which in self._alg._pending_steps:
            return int(self._alg._pending_steps[self._which])
        a, b = C.c_int64(), C.c_int64()
        if ac._h is not None:
            B.lib().dtc_learner_get_adam_steps(ac._h, C.byref(a), C.byref(b))
        return b.value if self._which == "main" else a.value

    def state_dict(self):
        ac = self._alg.actor_critic
        keys, live = self._keys()
        m, v = self._mv()
        steps = self._steps()
        state = {}
        if steps > 0:
            for i, k in live:
                idx = ac._idx[k]
                shape = ac._table.shape[k]
                state[i] = {"step": torch.tensor(float(steps)), "exp_avg": m[idx].view(shape).clone(),
                            "exp_avg_sq": v[idx].view(shape).clone()}
        lr = self._alg.learning_rate if self._which == "main" else 5.e-4
        group = {"lr": lr, "betas": (0.9, 0.999), "eps": 1e-08, "weight_decay": 0, "amsgrad": False, "maximize": False,
                 "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(keys)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        ac = self._alg.actor_critic
        m, v = self._mv()
        if sd.get("layout") == "dtc_b200.flat":  # files written by earlier builds of this package
            m.copy_(sd["exp_avg"])
            v.copy_(sd["exp_avg_sq"])
            self._alg._pending_steps[self._which] = int(sd["step"])
            if self._which == "main":
                self._alg.learning_rate = float(sd["lr"])
            return
        keys, live = self._keys()
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(keys):
            raise ValueError(f"optimizer state has {sum(len(g['params']) for g in groups)} parameters in {len(groups)} groups; "
                             f"expected one group of {len(keys)} (torch.optim.Adam over the reference's parameter list)")
        order = list(groups[0]["params"])
        state = sd["state"]
        m.zero_()
        v.zero_()
        steps = 0
        for i, k in live:
            st = state.get(order[i])
            if st is None:
                continue
            shape = ac._table.shape[k]
            if tuple(st["exp_avg"].shape) != shape:
                raise ValueError(f"optimizer state of {k}: shape {tuple(st['exp_avg'].shape)} != {shape}")
            idx = ac._idx[k]
            m[idx] = st["exp_avg"].to(m.device, torch.float32).reshape(-1)
            v[idx] = st["exp_avg_sq"].to(v.device, torch.float32).reshape(-1)
            steps = max(steps, int(float(st["step"])))
        self._alg._pending_steps[self._which] = steps
        if self._which == "main":
            self._alg.learning_rate = float(groups[0]["lr"])


class PPO:
    def __init__(self, actor_critic, num_learning_epochs=1, num_mini_batches=1, clip_param=0.2, gamma=0.998, lam=0.95,
                 value_loss_coef=1.0, entropy_coef=0.0, learning_rate=1e-3, max_grad_norm=1.0, use_clipped_value_loss=True,
                 schedule="fixed", desired_kl=0.01, device="cpu"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise B.DtcError("PPO runs on a CUDA device only (no CPU fallback)")
        self.desired_kl, self.schedule = desired_kl, schedule
        self.actor_critic = actor_critic
        if actor_critic.device is None:
            actor_critic.to(self.device)
        self.storage = None
        self.optimizer = _DeviceAdamHandle(self, "main")
        self.vae_optimizer = _DeviceAdamHandle(self, "vae")
        self._pending_steps = {}
        self.transition = RolloutStorage.Transition()
        self.clip_param = clip_param
        self.num_learning_epochs, self.num_mini_batches = num_learning_epochs, num_mini_batches
        self.value_loss_coef, self.entropy_coef = value_loss_coef, entropy_coef
        self.gamma, self.lam, self.max_grad_norm = gamma, lam, max_grad_norm
        self.use_clipped_value_loss = use_clipped_value_loss
        self._lr_host = float(learning_rate)
        self._lr_dirty = True
        self._inject = None  # tests: dict(perm=LongTensor, eps=[per optimizer step [M,16] tensors, vae/ppo alternating])
        self._update_calls = 0
        self._grad_tap = None  # tests: callable(which) invoked between backward and the optimizer step (0 = vae, 1 = policy)
        self.group = None  # torch.distributed process group for data parallel training (None = default group)
        self._peer = None  # rsl_rl.utils.dp.PeerAllReduce, created on first use (False: not available, NCCL instead)
        self._comm = None  # communication stream of the bucketed gradient all-reduce (created on first use)

    # ------------------------------------------------------------------ learning rate lives on the device
    @property
    def learning_rate(self):
        ac = self.actor_critic
        if not self._lr_dirty and ac._h is not None:
            self._lr_host = float(ac.stats()[8].item())
        return self._lr_host

    @learning_rate.setter
    def learning_rate(self, v):
        self._lr_host, self._lr_dirty = float(v), True

    def _push_lr(self, h):
        if self._lr_dirty:
            B.check(B.lib().dtc_learner_set_lr(h, self._lr_host, B.stream_ptr(self.device)), "dtc_learner_set_lr")
            self._lr_dirty = False

    def _hparams(self):
        hp = B.PPOHParams()
        hp.clip_param, hp.value_loss_coef, hp.entropy_coef = self.clip_param, self.value_loss_coef, self.entropy_coef
        hp.max_grad_norm = self.max_grad_norm
        hp.desired_kl = self.desired_kl if self.desired_kl is not None else 0.0
        hp.adaptive_lr = 1 if (self.desired_kl is not None and self.schedule == "adaptive") else 0
        hp.use_clipped_value_loss = 1 if self.use_clipped_value_loss else 0
        return hp

    # ------------------------------------------------------------------ reference methods
    def init_storage(self, num_envs, num_transitions_per_env, actor_obs_shape, privileged_obs_shape, obs_history_shape, action_shape):
        self.storage = RolloutStorage(num_envs, num_transitions_per_env, actor_obs_shape, privileged_obs_shape, obs_history_shape,
                                      action_shape, self.device)
        rows = (num_envs * num_transitions_per_env) // self.num_mini_batches
        self.actor_critic._learner(max(rows, num_envs))
        self.sync_replicas()

    def sync_replicas(self):
        """Data parallel: every rank must hold rank 0's parameters, Adam moments, step counters and learning rate; gradients are
        the only thing the update all-reduces.  Called at init_storage() and after load(); a no-op at world size 1."""
        if self._world() <= 1:
            return
        import torch.distributed as dist
        ac = self.actor_critic
        for t in (ac._flat, *ac._adam):
            dist.broadcast(t, 0, group=self.group)
        a, b = C.c_int64(), C.c_int64()
        if ac._h is not None:
            B.lib().dtc_learner_get_adam_steps(ac._h, C.byref(a), C.byref(b))
        meta = torch.tensor([self._pending_steps.get("vae", a.value), self._pending_steps.get("main", b.value), self.learning_rate],
                            dtype=torch.float64, device=self.device)
        dist.broadcast(meta, 0, group=self.group)
        self._pending_steps = {"vae": int(meta[0].item()), "main": int(meta[1].item())}
        self.learning_rate = float(meta[2].item())
        ac._params_written()

    def test_mode(self):
        self.actor_critic.eval()

    def train_mode(self):
        self.actor_critic.train()

    def act(self, obs, privileged_obs, obs_history, base_vel, rew_buf=None):
        st = self.storage
        if st.step >= st.num_transitions_per_env:
            raise AssertionError("Rollout buffer overflow")
        self.actor_critic._forward_act(obs, obs_history, privileged_obs, base_vel, storage=st._c, step=st.step, need_copies=False)
        self.transition.actions = st.actions[st.step]
        self.transition.values = st.values[st.step]
        return self.transition.actions

    def process_env_step(self, rewards, dones, next_obs, infos):
        st = self.storage
        to = infos["time_outs"] if "time_outs" in infos else None
        if to is not None and to.dtype != torch.uint8:
            to = to.view(torch.uint8) if to.dtype == torch.bool else to.to(torch.uint8)  # bool -> uint8 is a zero-copy view
        d = dones if dones.dtype == torch.uint8 else (dones.view(torch.uint8) if dones.dtype == torch.bool else dones.to(torch.uint8))
        B.check(B.lib().dtc_store_transition(C.byref(st._c), st.step, B.ptr(rewards), B.ptr(d), B.ptr(to), B.ptr(next_obs),
                                             next_obs.stride(0), self.gamma, B.stream_ptr(self.device)), "dtc_store_transition")
        st.step += 1
        self.transition.clear()
        self.actor_critic.reset(dones)

    def compute_returns(self, last_critic_obs, last_critic_privileged_obs, base_vel):
        last_values = self.actor_critic.evaluate(last_critic_obs, last_critic_privileged_obs, base_vel)
        self.storage.compute_returns(last_values, self.gamma, self.lam, group=self.group)

    def _world(self):
        return dp.world_size(self.group)

    def _allreduce_buckets(self, h, which):
        """SUM all-reduce of the step's flat gradient range in the two buckets the step publishes (include/dtc_b200.h:
        dtc_learner_grad_bucket), on a communication stream: bucket 0 starts as soon as its event fires, underneath the shared
        encoders' backward that is still running on the compute streams; the compute stream rejoins before the optimizer."""
        if self._world() <= 1:
            return
        lib, ac = B.lib(), self.actor_critic
        if self._comm is None:
            self._comm = torch.cuda.Stream(self.device)
            self._buckets = {}
            for w in (0, 1):
                for k in (0, 1):
                    b0, b1 = C.c_int64(), C.c_int64()
                    B.check(lib.dtc_learner_grad_bucket(w, k, C.byref(b0), C.byref(b1)), "dtc_learner_grad_bucket")
                    self._buckets[(w, k)] = (b0.value, b1.value)
        comm = self._comm
        # DTC_DP = p2p (default on one NVLink node: csrc/dtc_dp.cu, one peer-memory all-reduce of the whole range at step end) |
        #          nccl1 (one NCCL all-reduce at step end) | nccl2 (two NCCL buckets on a communication stream).
        # Measured, 2 x B200 (gpurun_out/r2r): nccl2 87.1 / 86.8 ms per iteration, nccl1 86.5 / 86.9 - the bucket that could overlap
        # does not (the persistent GEMM grids leave NCCL no SMs until they drain), so what matters is the latency of the one
        # exposed collective per optimizer step.
        mode = os.environ.get("DTC_DP", "p2p")
        lo = min(self._buckets[(which, 0)][0], self._buckets[(which, 1)][0])
        hi = max(self._buckets[(which, 0)][1], self._buckets[(which, 1)][1])
        if mode == "p2p":
            if self._peer is None:
                n = ac._grads.numel()
                ok = dp.PeerAllReduce.available(self.device, self.group) and lo % 4 == 0 and ac._grads.data_ptr() % 16 == 0
                self._peer = dp.PeerAllReduce((n + 3) // 4 * 4, self.device, self.group) if ok else False
                if self._peer and os.environ.get("DTC_DP_INPLACE", "1") != "0":
                    self._peer.register(ac._grads)  # all-reduce straight out of / into every rank's gradient buffer
            if self._peer:
                hi4 = min((hi + 3) // 4 * 4, ac._grads.numel() // 4 * 4)  # a few floats past the range may ride along: they are rewritten before use
                self._peer.allreduce_sum_(ac._grads[lo:hi4])
                return
            mode = "nccl1"
        if mode == "nccl1":
            dp.allreduce_sum_(ac._grads[lo:hi], self.group)
            return
        for k in (0, 1):
            b0, b1 = self._buckets[(which, k)]
            B.check(lib.dtc_learner_wait_bucket(h, which, k, C.c_void_p(comm.cuda_stream)), "dtc_learner_wait_bucket")
            with torch.cuda.stream(comm):
                dp.allreduce_sum_(ac._grads[b0:b1], self.group)
        torch.cuda.current_stream(self.device).wait_stream(comm)

    def update(self):
        ac, st, lib = self.actor_critic, self.storage, B.lib()
        T, N = st.num_transitions_per_env, st.num_envs
        mbs = (T * N) // self.num_mini_batches
        h = ac._learner(max(mbs, N))
        if self._pending_steps:
            a, b = C.c_int64(), C.c_int64()
            lib.dtc_learner_get_adam_steps(h, C.byref(a), C.byref(b))
            lib.dtc_learner_set_adam_steps(h, self._pending_steps.get("vae", a.value), self._pending_steps.get("main", b.value))
            self._pending_steps = {}
        self._push_lr(h)
        stream = B.stream_ptr(self.device)
        B.check(lib.dtc_learner_reset_stats(h, stream), "dtc_learner_reset_stats")
        inj = self._inject or {}
        self._inject = None
        if "perm" in inj:
            st._inject_perm = inj["perm"]
        perm = st.draw_permutation(self.num_mini_batches * mbs)
        batch = st.gather(perm)
        eps_list = list(inj.get("eps", []))
        hp = self._hparams()
        world = self._world()
        tap = self._grad_tap
        sync = 1 if (world > 1 or tap is not None or os.environ.get("DTC_FORCE_SYNC") == "1") else 0  # env: measurement switch
        tab = ac._table
        self._update_calls += 1
        k = 0
        for epoch in range(self.num_learning_epochs):
            for i in range(self.num_mini_batches):
                ctr = (self._update_calls << 16) + 2 * k
                e1 = eps_list[2 * k] if eps_list else None
                e2 = eps_list[2 * k + 1] if eps_list else None
                k += 1
                B.check(lib.dtc_vae_step(h, C.byref(batch._c), i * mbs, mbs, B.ptr(e1), ac.seed + 7919, ctr, C.byref(hp), sync, stream),
                        "dtc_vae_step")
                if sync:
                    if tap is not None:
                        tap(0)
                    self._allreduce_buckets(h, 0)
                    B.check(lib.dtc_optimizer_apply(h, 0, C.byref(hp), 1.0 / world, mbs * world, stream), "dtc_optimizer_apply")
                B.check(lib.dtc_ppo_step(h, C.byref(batch._c), i * mbs, mbs, B.ptr(e2), ac.seed + 7919, ctr + 1, C.byref(hp), sync, stream),
                        "dtc_ppo_step")
                if sync:
                    if tap is not None:
                        tap(1)
                    self._allreduce_buckets(h, 1)
                    B.check(lib.dtc_optimizer_apply(h, 1, C.byref(hp), 1.0 / world, mbs * world, stream), "dtc_optimizer_apply")
        s = ac.stats().tolist()  # the one device->host read of update()
        if self._peer:
            self._peer.check()  # no rank ever timed out waiting for a peer's flag
        n = self.num_learning_epochs * self.num_mini_batches
        self._lr_host = s[8]
        self.last_stats = dict(value=s[0] / n, surrogate=s[1] / n, recons=s[2] / n, vel=s[3] / n, kld=s[4] / n, height=s[5] / n,
                               entropy=s[6] / n, kl_mean=s[7], learning_rate=s[8], grad_norm_vae=s[9], grad_norm_policy=s[10])
        st.clear()
        return s[0] / n, s[1] / n, 0.0, 0, s[2] / n, s[3] / n, s[4] / n
