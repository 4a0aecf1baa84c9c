"""Multi-GPU plumbing for the tiled denoise loop (new design; the reference's inference path is single-GPU).
Within a scheduler step the tile forwards are independent (pipeline_FaithDiff_tlc.py:1330-1376); the only coupling is the
Gaussian blend (:1409), a linear weighted sum divided by a static count map.  So: shard tiles over ranks, every rank
accumulates its tiles into a full-size fp32 numerator canvas, ONE all-reduce(sum) per step (NCCL over NVLink for CUDA
tensors, gloo in the CPU tests), divide by the count map, re-tile locally."""
import torch


def shard_tiles(num_tiles, rank, world_size):
    """Round-robin tile ownership: rank r owns tiles r, r+W, r+2W, … (balanced to within one tile)."""
    return list(range(rank, num_tiles, world_size))


def plan_launch_groups(my_tiles, sizes):
    """Greedy split of a rank's tile list into launch groups using the available graph batch sizes (descending)."""
    groups, i = [], 0
    sizes = sorted(set(sizes), reverse=True)
    while i < len(my_tiles):
        left = len(my_tiles) - i
        nb = next((s for s in sizes if s <= left), sizes[-1])
        groups.append(my_tiles[i:i + nb])
        i += nb
    return groups


def exchange_canvas(numerator, group=None):
    """The per-step exchange of the overlap-region accumulation: in-place sum over ranks."""
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1:
        torch.distributed.all_reduce(numerator, op=torch.distributed.ReduceOp.SUM, group=group)
    return numerator
