"""Mirror of utils/nn_distance.py (huber_loss :11-28, nn_distance :32-59): dense chamfer helper.
Broadcasting replaces the reference's two (B,N,M,C) .repeat() copies; values are identical."""
import torch


def huber_loss(error, delta=1.0):
    abs_error = torch.abs(error)
    quadratic = torch.clamp(abs_error, max=delta)
    linear = abs_error - quadratic
    return 0.5 * quadratic ** 2 + delta * linear


def nn_distance(pc1, pc2, l1smooth=False, delta=1.0, l1=False):
    """pc1 (B,N,C), pc2 (B,M,C) -> dist1 (B,N), idx1 (B,N), dist2 (B,M), idx2 (B,M)."""
    pc_diff = pc1.unsqueeze(2) - pc2.unsqueeze(1)
    if l1smooth:
        pc_dist = torch.sum(huber_loss(pc_diff, delta), dim=-1)
    elif l1:
        pc_dist = torch.sum(torch.abs(pc_diff), dim=-1)
    else:
        pc_dist = torch.sum(pc_diff ** 2, dim=-1)
    dist1, idx1 = torch.min(pc_dist, dim=2)
    dist2, idx2 = torch.min(pc_dist, dim=1)
    return dist1, idx1, dist2, idx2
