"""Trajectory helpers around the GRU `Memory` kernel: the same two operations the reference's recurrent path uses
(`rsl_rl/rsl_rl/utils/utils.py:33-71`, `split_and_pad_trajectories` / `unpad_trajectories`), written as index arithmetic
on the device instead of Python-side `torch.split` + `pad_sequence` (no host round trip for the trajectory lengths
except the one scalar that sizes the output)."""
import torch


def split_and_pad_trajectories(tensor, dones):
    """`tensor` [T, N, ...] holds N environments' rollouts, `dones` [T, N] marks the last step of an episode.
    Every (environment, episode-fragment) becomes one column of the result, ordered environment-major, fragments
    shorter than T are zero-padded at the end: returns (padded [T, n_traj, ...], masks bool [T, n_traj]) exactly as the
    reference does (a fragment always ends at the last rollout step).  The time axis is always T long; the reference's
    pad_sequence stops at the longest fragment, which only differs when every environment finished an episode inside the
    rollout - a case its own unpad_trajectories cannot handle."""
    T, N = dones.shape[0], dones.shape[1]
    ends = dones.reshape(T, N).to(torch.bool).clone()
    ends[-1] = True
    flat_end = ends.t().reshape(-1)                      # environment-major order of all T*N steps
    # fragment id of every step = number of fragment ends strictly before it; position inside the fragment by a running count
    frag = torch.cumsum(flat_end.to(torch.int64), 0) - flat_end.to(torch.int64)
    n_traj = int(frag[-1].item()) + 1
    start_flag = torch.ones_like(flat_end)
    start_flag[1:] = flat_end[:-1]
    idx = torch.arange(T * N, device=dones.device)
    start_idx = torch.where(start_flag, idx, torch.zeros_like(idx))
    pos = idx - torch.cummax(start_idx, 0).values
    src = tensor.transpose(0, 1).reshape(T * N, *tensor.shape[2:])
    padded = torch.zeros(T, n_traj, *tensor.shape[2:], dtype=tensor.dtype, device=tensor.device)
    padded[pos, frag] = src
    masks = torch.zeros(T, n_traj, dtype=torch.bool, device=tensor.device)
    masks[pos, frag] = True
    return padded, masks


def unpad_trajectories(trajectories, masks):
    """Inverse of `split_and_pad_trajectories`: [T, n_traj, F] + masks -> [T, N, F] (reference utils.py:67-71)."""
    T, F = trajectories.shape[0], trajectories.shape[-1]
    valid = trajectories.transpose(0, 1)[masks.transpose(0, 1)]      # environment-major list of all real steps
    return valid.reshape(-1, T, F).transpose(0, 1)
