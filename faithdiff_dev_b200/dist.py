"""Multi-GPU plumbing for the tiled denoise loop (new design; the reference's inference path is single-GPU).
Within a scheduler step the tile forwards are independent (pipeline_FaithDiff_tlc.py:1330-1376); the only coupling is the
Gaussian blend (:1409), a linear weighted sum divided by a static count map.  So: shard tiles over ranks, every rank
accumulates its tiles into a full-size fp32 numerator canvas, ONE all-reduce(sum) per step (NCCL over NVLink for CUDA
tensors, gloo in the CPU tests), divide by the count map, re-tile locally."""
import torch


def shard_tiles(num_tiles, rank, world_size):
    """Round-robin tile ownership: rank r owns tiles r, r+W, r+2W, … (balanced to within one tile)."""
    return list(range(rank, num_tiles, world_size))


# measured cost (ms) of one CUDA-graph replay of the UNet tile-step for n tiles per launch on B200 (profiles/): batching
# amortises prologues and improves wave quantisation (20.7 ms for a single tile, 15.8 ms per tile from 7 tiles up)
REPLAY_COST_MS = {1: 20.73, 2: 35.80, 3: 48.70, 4: 64.75, 5: 79.83, 6: 95.50, 7: 110.47, 8: 126.53}


def plan_launch_groups(my_tiles, sizes, costs=None):
    """Split a rank's tile list into launch groups (sizes drawn from `sizes`) minimising the summed replay cost
    (dynamic programme over the tile count); groups keep the tile order."""
    costs = costs or REPLAY_COST_MS
    sizes = sorted({s for s in sizes if s >= 1} | {1})
    n = len(my_tiles)
    best = [0.0] + [float("inf")] * n
    choice = [0] * (n + 1)
    for i in range(1, n + 1):
        for s in sizes:
            if s <= i:
                c = best[i - s] + costs.get(s, costs[max(costs)] * s / max(costs))
                if c < best[i] - 1e-9:
                    best[i], choice[i] = c, s
    groups, i = [], n
    while i > 0:
        groups.append(choice[i])
        i -= choice[i]
    out, pos = [], 0
    for g in sorted(groups, reverse=True):
        out.append(my_tiles[pos:pos + g])
        pos += g
    return out


def shard_units(num_tiles, rank, world_size):
    """(tile, CFG-branch) units u = 2*tile + branch, split into contiguous balanced ranges: rank r owns
    [r*U/W, (r+1)*U/W).  Contiguity keeps both branches of a tile on one rank wherever possible (shared stem)."""
    U = 2 * num_tiles
    lo, hi = rank * U // world_size, (rank + 1) * U // world_size
    return list(range(lo, hi))


def shard_images(num_images, rank, world_size):
    """BASELINE config 5 (a batch of independent images, e.g. 32 x 512 -> 2048): plain data parallelism, no exchange.
    The reference restores one image per call (`assert b == 1`, pipeline_FaithDiff_tlc.py:171,1220), so the batch is a list of
    calls; rank r takes the contiguous balanced range [r*N/W, (r+1)*N/W) and runs each through a single-GPU pipeline."""
    lo, hi = rank * num_images // world_size, (rank + 1) * num_images // world_size
    return list(range(lo, hi))


# measured replay cost (ms) by number of SAMPLES (= units) per launch (9, 11, 13, 15 interpolated)
UNIT_REPLAY_COST_MS = {1: 15.60, 2: 20.73, 3: 26.28, 4: 35.80, 5: 42.89, 6: 48.70, 7: 56.53, 8: 64.75, 9: 72.3, 10: 79.83,
                       11: 87.7, 12: 95.50, 13: 103.0, 14: 110.47, 15: 118.5, 16: 126.53}


def plan_unit_groups(units, max_samples=16):
    """Split a contiguous unit range into consecutive launch groups minimising the summed replay cost."""
    n = len(units)
    best = [0.0] + [float("inf")] * n
    choice = [0] * (n + 1)
    for i in range(1, n + 1):
        for sz in range(1, min(max_samples, i, max(UNIT_REPLAY_COST_MS)) + 1):
            c = best[i - sz] + UNIT_REPLAY_COST_MS[sz]
            if c < best[i] - 1e-9:
                best[i], choice[i] = c, sz
    sizes, i = [], n
    while i > 0:
        sizes.append(choice[i])
        i -= choice[i]
    sizes.reverse()
    out, pos = [], 0
    for sz in sizes:
        out.append(units[pos:pos + sz])
        pos += sz
    return out


def exchange_canvas(numerator, group=None):
    """The per-step exchange of the overlap-region accumulation: in-place sum over ranks."""
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(group) > 1:
        torch.distributed.all_reduce(numerator, op=torch.distributed.ReduceOp.SUM, group=group)
    return numerator
