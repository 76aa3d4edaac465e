"""Host-side multi-GPU logic on CPU: world_size-2 gloo group (no GPU needed)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deep_prob_feature_track_b200.sharding import gather_poses, max_over_ranks, shard_range


def test_shard_range_covers_everything_once():
    for n in (0, 1, 7, 64, 1024, 1025):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard_range(n_total, world, rank)
    rows = torch.arange(lo, hi, dtype=torch.float32).view(-1, 1).repeat(1, 12)   # stand-in for solved poses
    slow = max_over_ranks(10.0 + rank)
    parts = gather_poses(rows, n_total)
    q.put((rank, slow, torch.cat(parts)[:, 0].tolist()))
    dist.destroy_process_group()


def test_two_ranks_gloo():
    world, n_total = 2, 7
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, slow, order in results:
        assert slow == 11.0                      # the slower rank's time wins
        assert order == list(map(float, range(n_total)))   # every pair exactly once, in order
