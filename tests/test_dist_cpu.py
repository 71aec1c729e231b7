"""CPU, world_size 2, gloo: the N > 1 slice-partition + single all-reduce path
(host logic of optyx_b200.dist) with the kernels replaced by the numpy
descriptor emulator."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import emulate_plan, random_network
from oracle.contract import contract_network
from optyx_b200 import dist as b2dist
from optyx_b200.planner import plan_network


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, seed, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    net = random_network(np.random.default_rng(seed), n_leaves=9, n_inds=10, n_out=1)
    plan = plan_network(net, target_size=4)
    s0, s1 = b2dist.partition(plan.nslices, rank, world)
    part = emulate_plan(net, plan, slice_range=(s0, s1))[0]
    t = torch.from_numpy(np.ascontiguousarray(part))
    b2dist.allreduce_sum_(t)
    if rank == 0:
        q.put((t.numpy().copy(), plan.nslices))
    dist.destroy_process_group()


def test_partition_covers_all_slices():
    for n in (1, 5, 8, 33):
        for w in (1, 2, 3, 8):
            ranges = [b2dist.partition(n, r, w) for r in range(w)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            assert max(b - a for a, b in ranges) - min(b - a for a, b in ranges) <= 1


def test_two_ranks_gloo_sliced_sum_matches_oracle():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 3, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, nslices = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert nslices > 1
    net = random_network(np.random.default_rng(3), n_leaves=9, n_inds=10, n_out=1)
    assert np.allclose(got, contract_network(net), rtol=1e-9, atol=1e-9)
