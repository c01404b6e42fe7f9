"""`ActorCriticDecoder` - the policy surface of the hot path, backed by the sm_100a learner kernels.

Keeps the reference's constructor and method names (rsl_rl/rsl_rl/modules/actor_critic_decoder.py:305-551) and its
state_dict key names (checkpoints are interchangeable).  All parameters live in ONE flat float32 CUDA
buffer laid out by the C library (include/dtc_b200.h: dtc_param_info); `state_dict()` / `load_state_dict()` translate
between that layout and the reference's per-layer tensors.  Parameter initialisation restates the reference's
construction order on the CPU so the same torch seed yields bit-identical parameters (:91-264, :305-376).

There is no CPU fallback: every forward runs through libdtc_b200.so on a CUDA device.
"""
import collections
import ctypes as C

import numpy as np
import torch
import torch.nn as nn

from ... import _lib as B


class AC_Args:
    """Hyper-dimensions of the live path (actor_critic_decoder.py:11-88); the `policy` config section is ignored by
    the reference (:312-314) and therefore here as well."""
    init_noise_std = 1.0
    terrain_latent = 512
    cenet_encoder = (265, [128], 64)
    cenet_decoder = (19 + 512, [64, 128], 53)
    terrain_encoder = (693, [512, 512], 512)
    terrain_decoder = (512, [512, 512], 693)
    memory_mlp = (265 + 512, [256, 128], 512)
    ga_decoder = (64, [128], 693)  # built and discarded by the reference (:212-233); it consumes RNG
    gb_encoder = (128, [128], 64)
    actor_hidden = [512, 256, 128]
    critic_hidden = [512, 256, 128]
    rnn_type, rnn_num_layers, rnn_hidden_size = "gru", 2, 50  # unused constants of the reference (:86-88)


STATE_KEYS = (
    ["std"]
    + [f"vae.cenet_encoder.{i}.{p}" for i in (0, 2) for p in ("weight", "bias")]
    + [f"vae.{n}.{p}" for n in ("latent_mu", "latent_var") for p in ("weight", "bias")]
    + [f"vae.{n}.{i}.{p}" for n in ("cenet_decoder", "terrain_encoder", "terrain_decoder", "memory_mlp") for i in (0, 2, 4)
       for p in ("weight", "bias")]
    + [f"vae.gb_encoder.{i}.{p}" for i in (0, 2) for p in ("weight", "bias")]
    + [f"{n}.{i}.{p}" for n in ("actor_body", "critic_body") for i in (0, 2, 4, 6) for p in ("weight", "bias")]
)


def _mlp_params(prefix, inp, hidden, out, sd):
    """nn.Linear stack in the reference's creation order: the first layer keeps torch's default init, all later ones
    are orthogonal(gain 0.01) with zero bias (layer_init, actor_critic_decoder.py:268-272)."""
    def put(i, layer):
        sd[f"{prefix}.{i}.weight"] = layer.weight.detach().clone()
        sd[f"{prefix}.{i}.bias"] = layer.bias.detach().clone()

    def ortho(layer):
        nn.init.orthogonal_(layer.weight, 0.01)
        nn.init.constant_(layer.bias, 0.0)
        return layer

    put(0, nn.Linear(inp, hidden[0]))
    for i in range(len(hidden)):
        nxt = out if i == len(hidden) - 1 else hidden[i + 1]
        put(2 * (i + 1), ortho(nn.Linear(hidden[i], nxt)))


def reference_init_state_dict(num_obs=53, num_critic_obs=1389, num_actions=12):
    """Fresh parameters drawn from torch's global CPU generator in the reference's order."""
    sd, A = {}, AC_Args
    _mlp_params("vae.cenet_encoder", *A.cenet_encoder, sd)
    for name, out in (("vae.latent_mu", 19), ("vae.latent_var", 16)):
        layer = nn.Linear(64, out)
        nn.init.orthogonal_(layer.weight, 0.01)
        nn.init.constant_(layer.bias, 0.0)
        sd[name + ".weight"], sd[name + ".bias"] = layer.weight.detach().clone(), layer.bias.detach().clone()
    _mlp_params("vae.cenet_decoder", *A.cenet_decoder, sd)
    _mlp_params("vae.terrain_encoder", *A.terrain_encoder, sd)
    _mlp_params("vae.terrain_decoder", *A.terrain_decoder, sd)
    _mlp_params("vae.memory_mlp", *A.memory_mlp, sd)
    _mlp_params("_discarded_ga_decoder", *A.ga_decoder, {})
    _mlp_params("vae.gb_encoder", *A.gb_encoder, sd)
    _mlp_params("actor_body", num_obs + 16 + 3 + A.terrain_latent, A.actor_hidden, num_actions, sd)
    _mlp_params("critic_body", 693 + num_obs + 3 + 15 + 12 - 24, A.critic_hidden, 1, sd)
    sd["std"] = A.init_noise_std * torch.ones(num_actions)
    return collections.OrderedDict((k, sd[k]) for k in STATE_KEYS)


class _ParamTable:
    """dtc_param_info rows -> flat-buffer index maps (one int64 index per reference element)."""
    _cache = None

    def __init__(self):
        lib = B.lib()
        self.total = int(lib.dtc_param_total_floats())
        self.index = {}
        self.shape = {}
        for i in range(lib.dtc_param_count()):
            p = B.ParamInfo()
            B.check(lib.dtc_param_get(i, C.byref(p)), "dtc_param_get")
            name = p.name.decode()
            col = np.zeros(p.cols, dtype=np.int64)
            for s in range(p.nseg):
                col[p.seg_src[s]:p.seg_src[s] + p.seg_len[s]] = np.arange(p.seg_dst[s], p.seg_dst[s] + p.seg_len[s])
            idx = p.offset + np.arange(p.rows, dtype=np.int64)[:, None] * p.ld + col[None, :]
            is_vec = name.endswith(".bias") or name == "std"
            self.shape[name] = (p.cols,) if is_vec else (p.rows, p.cols)
            self.index[name] = torch.from_numpy(idx.reshape(-1))
        self.ranges = {}
        for which, key in enumerate(("vae", "policy", "policy_sync", "all")):
            b, e = C.c_int64(), C.c_int64()
            B.check(lib.dtc_param_range(which, C.byref(b), C.byref(e)), "dtc_param_range")
            self.ranges[key] = (b.value, e.value)

    @classmethod
    def get(cls):
        if cls._cache is None:
            cls._cache = cls()
        return cls._cache


class _VaeView:
    """`actor_critic.vae` of the reference: only what callers of the hot path touch."""

    def __init__(self, ac):
        self._ac = ac

    def parameters(self):
        b, e = self._ac._table.ranges["vae"]
        return iter([self._ac._flat[b:e]])

    def state_dict(self):
        return collections.OrderedDict((k[4:], v) for k, v in self._ac.state_dict().items() if k.startswith("vae."))


class ActorCriticDecoder:
    is_recurrent = False

    def __init__(self, num_obs, num_critic_obs, num_actions, device=None, seed=0, **kwargs):
        if kwargs:
            print("ActorCriticDecoder.__init__ got unexpected arguments, which will be ignored: " + str([k for k in kwargs]))
        if (num_obs, num_critic_obs, num_actions) != (53, 1389, 12):
            raise ValueError("the fused kernels are specialised for the Lite3/X30 DTC shapes (53, 1389, 12)")
        self.num_obs, self.num_critic_obs, self.num_actions = num_obs, num_critic_obs, num_actions
        self._table = None
        self._init_sd = reference_init_state_dict(num_obs, num_critic_obs, num_actions)
        self._flat = None
        self._h = None
        self._max_rows = 0
        self.device = None
        self.seed = int(seed)
        self._calls = 0
        self._calls_offset = 0      # CUDA-graph support: dtc_policy_act gets _calls - _calls_offset, the device adds _calls_base
        self._calls_base = None
        self._inject = None  # tests: dict(eps_z=[M,16], eps_a=[M,12]) consumed by the next act()
        self.vae = _VaeView(self)
        self._out = None
        if device is not None:
            self.to(device)

    # ------------------------------------------------------------------ nn.Module-like plumbing
    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise B.DtcError("ActorCriticDecoder runs on a CUDA device only (no CPU fallback)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = device
        self._table = _ParamTable.get()
        n = self._table.total
        z = lambda: torch.zeros(n, device=device, dtype=torch.float32)
        self._flat, self._grads = z(), z()
        self._adam = [z(), z(), z(), z()]  # main m, main v, vae m, vae v
        self._idx = {k: v.to(device) for k, v in self._table.index.items()}
        self.load_state_dict(self._init_sd)
        self._init_sd = None
        return self

    def train(self, mode=True):
        return self

    def eval(self):
        return self

    def parameters(self):
        return iter([self._flat])

    def named_parameters(self):
        return iter(self.state_dict().items())

    def state_dict(self):
        self._need_device()
        out = collections.OrderedDict()
        for k in STATE_KEYS:
            out[k] = self._flat[self._idx[k]].view(self._table.shape[k]).clone()
        return out

    def load_state_dict(self, sd, strict=True):
        self._need_device()
        missing = [k for k in STATE_KEYS if k not in sd]
        if strict and missing:
            raise KeyError(f"missing keys in state_dict: {missing}")
        for k in STATE_KEYS:
            if k in sd:
                v = torch.as_tensor(sd[k]).to(self.device, torch.float32)
                if tuple(v.shape) != self._table.shape[k]:
                    raise ValueError(f"{k}: shape {tuple(v.shape)} != {self._table.shape[k]}")
                self._flat[self._idx[k]] = v.reshape(-1)
        self._params_written()
        return self

    def _params_written(self):
        """The flat buffer was written from outside the kernels: refresh the 3xTF32 companions of the parameters."""
        if self._h is not None:
            B.check(B.lib().dtc_learner_refresh_params(self._h, B.stream_ptr(self.device)), "dtc_learner_refresh_params")

    @property
    def std(self):
        """View into the flat parameter buffer (writes go through)."""
        o = int(self._table.index["std"][0])
        return self._flat[o:o + self.num_actions]

    def _need_device(self):
        if self._flat is None:
            raise B.DtcError("call .to('cuda') first: parameters live on the GPU only")

    def __del__(self):
        try:
            if self._h:
                B.lib().dtc_learner_destroy(self._h)
        except Exception:
            pass

    # ------------------------------------------------------------------ learner handle
    def _learner(self, rows):
        """(Re)creates the C-side learner with room for `rows` batch rows."""
        self._need_device()
        if self._h is not None and rows <= self._max_rows:
            return self._h
        lib = B.lib()
        steps = None
        if self._h is not None:
            a, b = C.c_int64(), C.c_int64()
            lib.dtc_learner_get_adam_steps(self._h, C.byref(a), C.byref(b))
            steps = (a.value, b.value)
            stats = self.stats()
            torch.cuda.synchronize(self.device)
            lib.dtc_learner_destroy(self._h)
            self._h = None
        with torch.cuda.device(self.device):
            nbytes = int(lib.dtc_learner_workspace_bytes(rows))
            self._ws = torch.empty(nbytes + 256, device=self.device, dtype=torch.uint8)
            off = (-self._ws.data_ptr()) % 256
            h = C.c_void_p()
            m1, v1, m2, v2 = self._adam
            B.check(lib.dtc_learner_create(rows, B.ptr(self._flat), B.ptr(self._grads), B.ptr(m1), B.ptr(v1), B.ptr(m2), B.ptr(v2),
                                           C.c_void_p(self._ws.data_ptr() + off), nbytes, B.stream_ptr(self.device), C.byref(h)),
                    "dtc_learner_create")
        self._h, self._max_rows = h, rows
        self._stats_ptr = lib.dtc_learner_stats(h)
        if self._calls_base is None:
            self._calls_base = torch.zeros(1, device=self.device, dtype=torch.int64)
        self._calls_base.fill_(self._calls_offset)
        B.check(lib.dtc_learner_set_act_counter_base(h, B.ptr(self._calls_base)), "dtc_learner_set_act_counter_base")
        if steps is not None:
            lib.dtc_learner_set_adam_steps(h, steps[0], steps[1])
            self._stats_tensor().copy_(stats)
        return h

    def _stats_tensor(self):
        """float64[16] view of the device-side learner statistics."""
        return _wrap_device_doubles(self._stats_ptr, 16, self.device)

    def stats(self):
        return self._stats_tensor().clone()

    # ------------------------------------------------------------------ forward surface
    def reset(self, dones=None):
        pass

    def forward(self):
        raise NotImplementedError

    def _buffers(self, M):
        if self._out is None or self._out["actions"].shape[0] != M:
            f = lambda *s: torch.empty(*s, device=self.device, dtype=torch.float32)
            self._out = dict(actions=f(M, 12), values=f(M), logp=f(M), mean=f(M, 12), sigma=f(M, 12))
        return self._out

    def _forward_act(self, obs, hist, priv, base_vel, storage=None, step=0, need_copies=True):
        """One fused pass: ActorCriticDecoder.act + evaluate + get_actions_log_prob (ppo.py:141-154)."""
        M = obs.shape[0]
        h = self._learner(M)
        for t, name in ((obs, "obs"), (hist, "obs_history"), (priv, "privileged_obs"), (base_vel, "base_vel")):
            B.require_cuda(t, name)
            if t.dtype != torch.float32 or t.stride(-1) != 1:
                raise ValueError(f"{name}: expected float32 rows with unit inner stride")
        inj = self._inject or {}
        self._inject = None
        ez, ea = inj.get("eps_z"), inj.get("eps_a")
        o = self._buffers(M) if need_copies or storage is None else None
        self._calls += 1
        args = [h, M, B.ptr(obs), obs.stride(0), B.ptr(hist), hist.stride(0), B.ptr(priv), priv.stride(0), B.ptr(base_vel),
                base_vel.stride(0), B.ptr(ez), B.ptr(ea), self.seed, self._calls - self._calls_offset,
                C.byref(storage) if storage is not None else None, step]
        if o is not None:
            args += [B.ptr(o["actions"]), B.ptr(o["values"]), B.ptr(o["logp"]), B.ptr(o["mean"]), B.ptr(o["sigma"])]
        else:
            args += [None] * 5
        B.check(B.lib().dtc_policy_act(*args, B.stream_ptr(self.device)), "dtc_policy_act")
        return o

    # ------------------------------------------------------------------ CUDA-graph capture of act() sequences (see LeggedRobotDTC)
    def graph_capture_begin(self):
        self._calls_offset = self._calls
        self._calls_base.fill_(self._calls_offset)

    def graph_capture_end(self, calls):
        B.check(B.lib().dtc_counter_add(B.ptr(self._calls_base), int(calls), B.stream_ptr(self.device)), "dtc_counter_add")
        self._calls_offset += int(calls)

    def graph_replayed(self, calls):
        self._calls += int(calls)
        self._calls_offset += int(calls)

    def act(self, observations, observation_history, privileged_observations, rew_buf=None, base_vel=None, **kwargs):
        if base_vel is None:
            base_vel = torch.zeros(observations.shape[0], 3, device=self.device)
        o = self._forward_act(observations, observation_history, privileged_observations, base_vel)
        return o["actions"]

    def update_distribution(self, observations, observation_history, privileged_observations):
        self.act(observations, observation_history, privileged_observations)

    @property
    def action_mean(self):
        return self._out["mean"]

    @property
    def action_std(self):
        return self._out["sigma"]

    @property
    def entropy(self):
        s = self._out["sigma"]
        return (0.5 + 0.5 * float(np.log(2 * np.pi)) + torch.log(s)).sum(dim=-1)

    def get_actions_log_prob(self, actions):
        if actions is self._out["actions"]:
            return self._out["logp"]
        mu, s = self._out["mean"], self._out["sigma"]
        return (-((actions - mu) ** 2) / (2 * s * s) - torch.log(s) - float(np.log(np.sqrt(2 * np.pi)))).sum(dim=-1)

    def evaluate(self, observations, privileged_observations, base_vel, **kwargs):
        M = observations.shape[0]
        h = self._learner(M)
        out = torch.empty(M, device=self.device, dtype=torch.float32)
        B.check(B.lib().dtc_policy_evaluate(h, M, B.ptr(observations), observations.stride(0), B.ptr(privileged_observations),
                                            privileged_observations.stride(0), B.ptr(base_vel), base_vel.stride(0), B.ptr(out),
                                            B.stream_ptr(self.device)), "dtc_policy_evaluate")
        return out.unsqueeze(1)

    def act_teacher(self, observations, observation_history, privileged_observations, **kwargs):
        M = observations.shape[0]
        h = self._learner(M)
        out = torch.empty(M, 12, device=self.device, dtype=torch.float32)
        B.check(B.lib().dtc_policy_act_teacher(h, M, B.ptr(observations), observations.stride(0), B.ptr(observation_history),
                                               observation_history.stride(0), B.ptr(privileged_observations),
                                               privileged_observations.stride(0), B.ptr(out), B.stream_ptr(self.device)),
                "dtc_policy_act_teacher")
        return out

    act_inference = act_teacher

    def debug_buffer(self, name):
        """Named activation / gradient buffer of the last step as a [rows, ld] tensor view (tests)."""
        p, r, c, ld = C.c_void_p(), C.c_int32(), C.c_int32(), C.c_int32()
        B.check(B.lib().dtc_learner_debug_buffer(self._h, name.encode(), C.byref(p), C.byref(r), C.byref(c), C.byref(ld)), name)
        return _wrap_device_floats(p.value, r.value * ld.value, self.device).view(r.value, ld.value)


class _CudaArray:
    def __init__(self, ptr, n, typestr, itemsize):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr), False), "version": 3,
                                         "strides": (itemsize,)}


def _wrap_device_doubles(ptr, n, device):
    with torch.cuda.device(device):
        return torch.as_tensor(_CudaArray(ptr, n, "<f8", 8), device=device)


def _wrap_device_floats(ptr, n, device):
    with torch.cuda.device(device):
        return torch.as_tensor(_CudaArray(ptr, n, "<f4", 4), device=device)
