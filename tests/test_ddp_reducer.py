"""FlatBucketReducer (deep_prob_feature_track_b200/ddp.py): gradients as views of one flat buffer, bucketed
all-reduce started from autograd hooks.  Two ranks on CPU with gloo (the N > 1 path of the training step; on the GPU
box the same code runs over NCCL)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from deep_prob_feature_track_b200.ddp import FlatBucketReducer, broadcast_parameters


def make_net(seed):
    torch.manual_seed(seed)
    return nn.Sequential(nn.Conv2d(3, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, 8, 3, padding=1), nn.ReLU(),
                         nn.Conv2d(8, 4, 1), nn.Flatten(), nn.Linear(4 * 8 * 8, 6))


def batch_of(rank):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn((5, 3, 8, 8), generator=g), torch.randn((5, 6), generator=g)


def worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = make_net(seed=rank)                 # different weights per rank until the broadcast
        broadcast_parameters(net)
        red = FlatBucketReducer(net.parameters(), n_buckets=3)
        assert 2 <= len(red.bounds) <= 3 and red.bounds[0][0] == 0 and red.bounds[-1][1] == red.flat.numel()
        assert all(a[1] == b[0] for a, b in zip(red.bounds, red.bounds[1:]))
        opt = torch.optim.SGD(net.parameters(), lr=0.1)
        x, y = batch_of(rank)
        for step in range(2):
            red.zero_grad()
            loss = ((net(x) - y) ** 2).mean()
            loss.backward()
            assert all(h is not None for h in red.handles)        # every bucket was launched from a hook
            red.finish()
            if step == 0:
                out.put((rank, red.flat.numpy().copy(), [p.detach().numpy().copy() for p in net.parameters()]))
            opt.step()
        out.put((rank + 10, None, [p.detach().numpy().copy() for p in net.parameters()]))
    finally:
        dist.destroy_process_group()


def test_two_ranks_average_their_gradients():
    world = 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2 * world):
        k, flat, params = out.get(timeout=60)      # numpy arrays: tensors would travel as file handles of a dead process
        got[k] = (None if flat is None else torch.from_numpy(flat), [torch.from_numpy(a) for a in params])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # the broadcast made the weights equal, the reducer made the gradients equal to the mean of the local ones
    assert all(torch.equal(a, b) for a, b in zip(got[0][1], got[1][1]))
    assert torch.equal(got[0][0], got[1][0])
    net = make_net(seed=0)
    grads = []
    for r in range(world):
        net.zero_grad()
        x, y = batch_of(r)
        ((net(x) - y) ** 2).mean().backward()
        grads.append(torch.cat([p.grad.flatten() for p in reversed(list(net.parameters()))]))
    assert torch.allclose(got[0][0], (grads[0] + grads[1]) / 2, rtol=1e-5, atol=1e-7)
    # and both ranks took the same optimizer steps
    assert all(torch.equal(a, b) for a, b in zip(got[10][1], got[11][1]))


def test_single_process_is_a_no_op():
    net = make_net(seed=3)
    red = FlatBucketReducer(net.parameters(), n_buckets=2)
    x, y = batch_of(0)
    red.zero_grad()
    ((net(x) - y) ** 2).mean().backward()
    red.finish()
    ref = make_net(seed=3)
    ((ref(x) - y) ** 2).mean().backward()
    for p, q in zip(net.parameters(), ref.parameters()):
        assert torch.equal(p.grad, q.grad)
    with pytest.raises(RuntimeError):
        net.zero_grad(set_to_none=True)
        red.zero_grad()
