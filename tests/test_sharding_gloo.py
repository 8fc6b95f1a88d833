"""Multi-GPU path on CPU: world_size 2, gloo.  Groups are sharded with no data-path collective;
the only communication is the host gather of consensus strings."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import OracleBackedContext, oracle_consensus_batch, pack_groups
from mandalorion_b200 import shard
from mandalorion_b200.synth import GroupConfig, make_groups

CFG = GroupConfig("shard", 23, 1, 9, 80, 400, "loguniform", 0.03, (0.3, 0.35, 0.35))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    gro, rbo, bases = pack_groups(make_groups(CFG))
    owner = shard.lpt_assign(shard.group_costs(gro, rbo), world)
    idx, g2, r2, b2 = shard.take_shard(gro, rbo, bases, owner, rank)
    out = OracleBackedContext().consensus_batch(packed=(g2, r2, b2))
    parts = [None] * world
    dist.all_gather_object(parts, (idx, out["cons"], out["status"]))
    if rank == 0:
        cons, status = shard.merge_shards(len(gro) - 1, parts)
        q.put((cons, status.tolist(), [len(p[0]) for p in parts]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_reproduces_the_single_rank_result():
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    cons, status, sizes = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    want = oracle_consensus_batch(make_groups(CFG))
    assert cons == want["cons"] and status == want["status"].tolist()
    assert sum(sizes) == 23 and min(sizes) >= 8          # balanced, nothing lost, nothing duplicated


def test_lpt_is_balanced_and_deterministic():
    rng = np.random.default_rng(0)
    costs = rng.lognormal(3, 1.5, 4000)
    o1, o2 = shard.lpt_assign(costs, 8), shard.lpt_assign(costs, 8)
    assert np.array_equal(o1, o2)
    load = np.bincount(o1, weights=costs, minlength=8)
    assert load.max() / load.mean() < 1.01


def test_take_and_merge_roundtrip():
    groups = [["ACGT", "ACG"], [], ["AAAAAAAA"], ["AC", "AC", "ACC"], ["G"]]
    gro, rbo, bases = pack_groups(groups)
    owner = np.array([0, 1, 1, 0, 1], dtype=np.int32)
    parts = []
    for r in range(2):
        idx, g2, r2, b2 = shard.take_shard(gro, rbo, bases, owner, r)
        sub = [[b2[r2[k]:r2[k + 1]].tobytes() for k in range(g2[i], g2[i + 1])] for i in range(len(idx))]
        assert sub == [[x.encode() for x in groups[g]] for g in idx]
        parts.append((idx, [b"|".join(s) for s in sub], np.zeros(len(idx), np.int32)))
    cons, status = shard.merge_shards(5, parts)
    assert cons == [b"|".join(x.encode() for x in g) for g in groups] and (status == 0).all()
