"""Trajectory helpers of the reference's recurrent path (rsl_rl/rsl_rl/utils/utils.py:33-71); plain tensor plumbing around the
GRU `Memory` kernel (works on whatever device the tensors live on)."""
import torch


def split_and_pad_trajectories(tensor, dones):
    """Splits [time, envs, ...] trajectories at done indices, concatenates them and pads with zeros up to the longest one;
    returns (padded [time, n_traj, ...], masks [time, n_traj]) - utils.py:33-65."""
    dones = dones.clone()
    dones[-1] = 1
    flat_dones = dones.transpose(1, 0).reshape(-1, 1)
    done_indices = torch.cat((flat_dones.new_tensor([-1], dtype=torch.int64), flat_dones.nonzero()[:, 0]))
    trajectory_lengths = done_indices[1:] - done_indices[:-1]
    trajectories = torch.split(tensor.transpose(1, 0).flatten(0, 1), trajectory_lengths.tolist())
    padded = torch.nn.utils.rnn.pad_sequence(trajectories)
    masks = trajectory_lengths > torch.arange(0, tensor.shape[0], device=tensor.device).unsqueeze(1)
    return padded, masks


def unpad_trajectories(trajectories, masks):
    """Inverse of split_and_pad_trajectories (utils.py:67-71)."""
    return trajectories.transpose(1, 0)[masks.transpose(1, 0)].view(-1, trajectories.shape[0], trajectories.shape[-1]).transpose(1, 0)
