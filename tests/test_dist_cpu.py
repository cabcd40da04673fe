"""CPU (gloo, world size 2): the multi-GPU host logic — pair / view sharding, the single gradient all-reduce,
variable-length gathers."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from starst3r_b200 import dist as sd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # pair sharding: disjoint, complete
        mine = sd.shard_pairs(8)
        everyone = [None] * world
        dist.all_gather_object(everyone, mine)
        flat = [p for part in everyone for p in part]
        assert sorted(flat) == sorted(sd.unordered_pairs(8)) and len(set(flat)) == 28
        assert abs(len(everyone[0]) - len(everyone[1])) <= 1
        # view sharding
        assert sd.shard_indices(8) == list(range(rank, 8, world))
        # gradient all-reduce: every rank ends with the sum
        g = torch.Generator().manual_seed(rank)
        N = 50
        grads = {"means": torch.randn(N, 3, generator=g), "quats": torch.randn(N, 4, generator=g),
                 "scales": torch.randn(N, 3, generator=g), "opacities": torch.randn(N, generator=g),
                 "sh": torch.randn(N, 4, 3, generator=g)}
        ref = {}
        for k in grads:
            parts = []
            for r in range(world):
                gr = torch.Generator().manual_seed(r)
                t = {"means": torch.randn(N, 3, generator=gr), "quats": torch.randn(N, 4, generator=gr),
                     "scales": torch.randn(N, 3, generator=gr), "opacities": torch.randn(N, generator=gr),
                     "sh": torch.randn(N, 4, 3, generator=gr)}
                parts.append(t[k])
            ref[k] = sum(parts)
        sd.allreduce_gradients(grads)
        for k in grads:
            assert torch.allclose(grads[k], ref[k], atol=1e-6), k
        # variable-length gather of correspondence lists
        t = torch.arange((rank + 1) * 3 * 2, dtype=torch.int64).reshape(-1, 2) + 100 * rank
        parts = sd.gather_varlen(t)
        assert [p.shape[0] for p in parts] == [3 * (r + 1) for r in range(world)]
        assert torch.equal(parts[rank], t)
        x = torch.full((4,), float(rank))
        sd.broadcast_tensors([x], src=0)
        assert x.eq(0).all()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_single_process_passthrough():
    assert sd.world() == (0, 1)
    assert sd.shard_indices(5) == [0, 1, 2, 3, 4]
    g = {k: torch.ones(2) for k in sd.GRAD_KEYS}
    assert sd.allreduce_gradients(g) is g
