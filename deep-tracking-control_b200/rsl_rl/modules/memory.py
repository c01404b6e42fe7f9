"""`Memory` of the reference (rsl_rl/rsl_rl/modules/actor_critic_decoder.py:584-614): an nn.GRU wrapper with per-env hidden
state, kept across `forward` calls in inference mode and zeroed for finished episodes by `reset(dones)`.

The reference declares it (AC_Args: rnn_type 'gru', 2 layers, hidden 50) but never instantiates it on the training path
(SURVEY.md 0.1); it is provided because the task names a GRU forward.  Forward only (inference / evaluation): the reference's
`is_recurrent` is False, so no optimizer ever reaches these weights.  Same constructor, same `forward(input, masks,
hidden_states)` / `reset(dones)` / `hidden_states` surface, same `state_dict` keys (`rnn.weight_ih_l0`, ...)."""
import collections
import ctypes as C
import math

import torch

from ... import _lib as B
from ..utils import unpad_trajectories


class Memory:
    def __init__(self, input_size, type="lstm", num_layers=1, hidden_size=256, device=None):
        if type.lower() != "gru":
            raise B.DtcError("Memory: only the GRU variant (AC_Args.rnn_type = 'gru') has a CUDA kernel; there is no CPU fallback")
        self.input_size, self.num_layers, self.hidden_size = int(input_size), int(num_layers), int(hidden_size)
        self.hidden_states = None
        self.device = None
        self._flat = None
        self._views = collections.OrderedDict()
        self._cpu_init = self._default_init()
        if device is not None:
            self.to(device)

    # ------------------------------------------------------------------ parameters (nn.GRU names, shapes and default init)
    def _shapes(self):
        H = self.hidden_size
        out = collections.OrderedDict()
        for l in range(self.num_layers):
            in_l = self.input_size if l == 0 else H
            out[f"rnn.weight_ih_l{l}"] = (3 * H, in_l)
            out[f"rnn.weight_hh_l{l}"] = (3 * H, H)
            out[f"rnn.bias_ih_l{l}"] = (3 * H,)
            out[f"rnn.bias_hh_l{l}"] = (3 * H,)
        return out

    def _default_init(self):
        k = 1.0 / math.sqrt(self.hidden_size)  # nn.RNNBase.reset_parameters: U(-1/sqrt(H), 1/sqrt(H)) for every tensor
        return collections.OrderedDict((name, (torch.rand(shape) * 2 - 1) * k) for name, shape in self._shapes().items())

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise B.DtcError("Memory runs on a CUDA device only (no CPU fallback)")
        src = self.state_dict() if self._flat is not None else self._cpu_init
        n = int(B.lib().dtc_gru_param_floats(self.input_size, self.hidden_size, self.num_layers))
        self._flat = torch.empty(n, device=device)
        self.device = device
        off = 0
        self._views = collections.OrderedDict()
        for name, shape in self._shapes().items():  # per layer: weight_ih | weight_hh | bias_ih | bias_hh (the C ABI's order)
            cnt = int(torch.tensor(shape).prod())
            self._views[name] = self._flat[off:off + cnt].view(shape)
            off += cnt
        assert off == n
        self.load_state_dict(src)
        if self.hidden_states is not None:
            self.hidden_states = self.hidden_states.to(device)
        return self

    def state_dict(self):
        return collections.OrderedDict((k, v.clone()) for k, v in self._views.items())

    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self._shapes() if k not in sd]
        if strict and missing:
            raise KeyError(f"missing keys in state_dict: {missing}")
        for k, v in self._views.items():
            if k in sd:
                t = torch.as_tensor(sd[k]).to(self.device, torch.float32)
                if tuple(t.shape) != tuple(v.shape):
                    raise ValueError(f"{k}: shape {tuple(t.shape)} != {tuple(v.shape)}")
                v.copy_(t)
        return self

    def parameters(self):
        return iter(self._views.values())

    # ------------------------------------------------------------------ forward / reset
    def _run(self, x, h):
        """x [T,N,in] -> out [T,N,H]; h [L,N,H] updated in place."""
        T, N, _ = x.shape
        out = torch.empty(T, N, self.hidden_size, device=self.device)
        B.check(B.lib().dtc_gru_forward(T, N, self.input_size, self.hidden_size, self.num_layers, B.ptr(self._flat), B.ptr(x),
                                        B.ptr(h), B.ptr(out), B.stream_ptr(self.device)), "dtc_gru_forward")
        return out

    def forward(self, input, masks=None, hidden_states=None):
        if self._flat is None:
            raise B.DtcError("Memory: call .to('cuda') first (no CPU fallback)")
        B.require_cuda(input, "input")
        if input.shape[-1] != self.input_size:
            raise ValueError(f"input feature size {input.shape[-1]} != {self.input_size}")
        batch_mode = masks is not None
        if batch_mode:
            # batch mode (policy update): saved hidden states, padded trajectories [T, n_traj, in]
            if hidden_states is None:
                raise ValueError("Hidden states not passed to memory module during policy update")
            h = hidden_states.to(self.device, torch.float32).contiguous().clone()
            out = self._run(input.contiguous().float(), h)
            return unpad_trajectories(out, masks)
        # inference mode (collection): hidden states of the last step
        x = input.contiguous().float().unsqueeze(0)
        N = x.shape[1]
        if self.hidden_states is None:
            self.hidden_states = torch.zeros(self.num_layers, N, self.hidden_size, device=self.device)
        elif self.hidden_states.shape[1] != N:
            raise RuntimeError(f"Expected hidden size ({self.num_layers}, {N}, {self.hidden_size}), got {list(self.hidden_states.shape)}")
        return self._run(x, self.hidden_states)

    __call__ = forward

    def reset(self, dones=None):
        if self.hidden_states is None:
            return
        if dones is None:
            raise TypeError("reset(dones): a boolean / uint8 [N] tensor is required, as for tensor indexing in the reference")
        d = dones.to(self.device).to(torch.uint8).contiguous()
        B.check(B.lib().dtc_gru_reset(self.hidden_states.shape[1], self.hidden_size, self.num_layers, B.ptr(self.hidden_states), B.ptr(d),
                                      B.stream_ptr(self.device)), "dtc_gru_reset")
