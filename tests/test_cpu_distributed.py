"""-m "not gpu": the N>1 host logic on CPU with gloo, world_size 2: tile sharding + per-step canvas exchange reproduce
the single-process Gaussian blend (oracle arithmetic) up to fp32 summation order."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.distributed.init_process_group("gloo", rank=rank, world_size=world)
    from faithdiff_dev_b200.dist import exchange_canvas, shard_tiles
    from faithdiff_dev_b200.tiling import gaussian_weights, tile_origins
    h, w, k = 40, 56, 16
    idx, k1, k2 = tile_origins(h, w, k, k, 0.5)
    wts = torch.from_numpy(gaussian_weights(k2, k1))
    g = torch.Generator().manual_seed(3)
    tiles = torch.randn(len(idx), 4, k1, k2, generator=g).to(torch.bfloat16)      # same on every rank
    num = torch.zeros(1, 4, h, w)
    cnt = torch.zeros(1, 4, h, w)
    for t in shard_tiles(len(idx), rank, world):
        i, j = idx[t]
        num[0, :, i:i + k1, j:j + k2] += tiles[t] * wts
    for (i, j) in idx:
        cnt[0, :, i:i + k1, j:j + k2] += wts
    exchange_canvas(num)
    q.put((rank, (num / cnt).numpy()))          # numpy: plain pickling, no shared-memory handle that can outlive the worker
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_canvas_exchange_matches_single_process_blend():
    from oracle import faithdiff_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r: torch.from_numpy(a) for r, a in (q.get(timeout=240) for _ in range(2))}
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    lo = O.LocalAttention((16, 16), 0.5)
    lo.grids(torch.zeros(1, 4, 40, 56))
    g = torch.Generator().manual_seed(3)
    tiles = torch.randn(len(lo.idxes), 4, 16, 16, generator=g).to(torch.bfloat16)
    ref = lo.grids_inverse(tiles)
    assert torch.equal(res[0], res[1])                       # all ranks hold the same canvas after the exchange
    assert (res[0] - ref).abs().max() <= 2e-6 * ref.abs().max()


def test_shard_and_launch_plan():
    from faithdiff_dev_b200.dist import plan_launch_groups, shard_tiles
    for T, W in [(28, 1), (28, 2), (28, 4), (28, 8), (112, 8), (1, 8), (9, 4)]:
        shards = [shard_tiles(T, r, W) for r in range(W)]
        assert sorted(t for s in shards for t in s) == list(range(T))
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    assert plan_launch_groups([0, 4, 8, 12, 16, 20, 24], [2, 1]) == [[0, 4], [8, 12], [16, 20], [24]]
    assert plan_launch_groups([], [2, 1]) == []
    assert plan_launch_groups([5], [2]) == [[5]]
    for n in range(0, 40):
        g = plan_launch_groups(list(range(n)), [1, 2, 3, 4])
        assert [t for grp in g for t in grp] == list(range(n)) and all(1 <= len(x) <= 4 for x in g)
    assert sorted(len(x) for x in plan_launch_groups(list(range(28)), [1, 2, 3, 4])) == [4] * 7
    assert sorted(len(x) for x in plan_launch_groups(list(range(28)), range(1, 9))) == [7] * 4
    assert sorted(len(x) for x in plan_launch_groups(list(range(7)), [1, 2, 3, 4])) == [3, 4]


def test_shard_images_partitions_the_batch():
    from faithdiff_dev_b200.dist import shard_images
    for n, w in [(32, 8), (32, 1), (5, 8), (9, 4), (0, 2)]:
        parts = [shard_images(n, r, w) for r in range(w)]
        assert sorted(i for p in parts for i in p) == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_unit_shards_and_launch_groups_for_every_world_size():
    """(tile, CFG-branch) units of the 4K (28 tiles) and 8K (112 tiles) layouts over 1, 2, 4, 8 ranks: contiguous balanced ranges
    that cover every unit once, and launch groups that keep the order and the per-launch sample cap."""
    from faithdiff_dev_b200.dist import plan_unit_groups, shard_units
    for T in (28, 112, 9, 1):
        for W in (1, 2, 4, 8):
            shards = [shard_units(T, r, W) for r in range(W)]
            assert [u for s in shards for u in s] == list(range(2 * T))
            assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
            for s in shards:
                groups = plan_unit_groups(s, max_samples=16)
                assert [u for g in groups for u in g] == s and all(1 <= len(g) <= 16 for g in groups)
