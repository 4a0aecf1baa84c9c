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


class PeerCanvas:
    """Opt-in (FD_P2P_EXCHANGE=1; run on 2 and 8 B200s in round 2: correct, not faster than NCCL — DESIGN §6): the per-step
    exchange as ONE kernel over NVLink peer memory instead of an NCCL all-reduce + fd_canvas_finalize (csrc/peer_exchange.cu).  Every rank's fp32 numerator canvas lives in a
    cudaMalloc'd region that all ranks of the node map through CUDA IPC; `numerator(step)` is the slot this rank accumulates
    step `step` into, `exchange_finalize(step, count, dtype)` publishes it, waits for the peers, sums all canvases in rank
    order and divides by the count map.  torch.distributed only carries the 64-byte IPC handles, once."""

    def __init__(self, shape, rank, world_size, group=None, device=None):
        import ctypes as C
        from ._lib import check, lib
        self.shape, self.rank, self.world = tuple(shape), rank, world_size
        self.n = 1
        for d in self.shape:
            self.n *= d
        assert self.n % 4 == 0, "canvas size must be a multiple of 4 floats"
        self.n_pad = (self.n + 63) // 64 * 64
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self._lib, self._check, self._C, self._group = lib, check, C, group
        mine = C.c_void_p()
        check(lib.fd_peer_alloc(C.byref(mine), lib.fd_peer_region_bytes(self.n)), "fd_peer_alloc")
        self._mine = mine.value
        handle = C.create_string_buffer(64)
        check(lib.fd_peer_ipc_handle(C.c_void_p(self._mine), handle), "fd_peer_ipc_handle")
        handles = [None] * world_size
        torch.distributed.all_gather_object(handles, bytes(handle.raw), group=group)
        self._opened, ptrs = [], []
        for r in range(world_size):
            if r == rank:
                ptrs.append(self._mine)
            else:
                q = C.c_void_p()
                check(lib.fd_peer_ipc_open(C.create_string_buffer(handles[r], 64), C.byref(q)), "fd_peer_ipc_open")
                self._opened.append(q.value)
                ptrs.append(q.value)
        self._regions = (C.c_void_p * world_size)(*ptrs)
        torch.distributed.barrier(group=group)          # nobody signals before everyone has mapped everyone

    def numerator(self, step):
        """fp32 view [shape] of this rank's canvas slot for scheduler step `step` (1-based counter); zero it before accumulating."""
        iface = {"shape": self.shape, "typestr": "<f4", "version": 2,
                 "data": (self._mine + (step & 1) * self.n_pad * 4, False)}
        holder = type("_PeerSlot", (), {"__cuda_array_interface__": iface})()
        return torch.as_tensor(holder, device=self.device)

    def exchange_finalize(self, step, count, lowp_dtype):
        from .ops import _DT, _p, _stream
        out = torch.empty(self.shape, device=self.device, dtype=torch.float32)
        lowp = torch.empty(self.shape, device=self.device, dtype=lowp_dtype)
        self._check(self._lib.fd_canvas_exchange_finalize(self._regions, self.rank, self.world, step & 1, step, self.n, _p(count),
                                                          _p(out), _p(lowp), _DT[lowp_dtype], _stream()), "fd_canvas_exchange_finalize")
        return out, lowp

    def close(self):
        """Collective: every rank must call it.  A peer may still be inside its last exchange kernel reading THIS rank's region
        over P2P (completion of my own kernel only proves the peers have started theirs), so all ranks synchronise, meet at a
        barrier and synchronise again before any handle is closed or freed."""
        C = self._C
        torch.cuda.synchronize()
        if torch.distributed.is_initialized():
            torch.distributed.barrier(group=self._group)
            torch.cuda.synchronize()
        for q in self._opened:
            self._lib.fd_peer_ipc_close(C.c_void_p(q))
        self._opened = []
        if self._mine:
            self._lib.fd_peer_free(C.c_void_p(self._mine))
            self._mine = None
