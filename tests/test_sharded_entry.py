"""The product's multi-GPU entry point on CPU: shard.consensus_batch_sharded() with one test double
per "device" (oracle-backed contexts, one host thread each).  Results must come back in input
order and equal the single-context run; nothing is exchanged between shards."""
import numpy as np

from helpers import OracleBackedContext, oracle_consensus_batch, pack_groups
from mandalorion_b200 import shard
from mandalorion_b200.synth import GroupConfig, make_groups

CFG = GroupConfig("shard2", 41, 0, 9, 60, 500, "loguniform", 0.03, (0.3, 0.35, 0.35))


def test_sharded_batch_equals_the_single_context_run():
    groups = make_groups(CFG)
    packed = pack_groups(groups)
    want = oracle_consensus_batch(packed=packed)
    for n_dev in (1, 2, 5):
        ctxs = {d: OracleBackedContext() for d in range(n_dev)}
        out = shard.consensus_batch_sharded(packed, devices=list(range(n_dev)), contexts=ctxs)
        assert out["cons"] == want["cons"] and out["status"].tolist() == want["status"].tolist()
        assert all(c.calls == 1 for c in ctxs.values())                      # one batch per device
        assert sum(s["band_cells"] for s in out["stats"]) == want["stats"]["band_cells"]
        if n_dev > 1:
            cost = shard.group_costs(packed[0], packed[1])
            load = np.bincount(out["owner"], weights=cost, minlength=n_dev)
            assert load.max() / load.mean() < 1.25                            # cost-balanced shards


def test_flags_travel_with_their_groups():
    groups = make_groups(CFG)[:12]
    packed = pack_groups(groups)
    flags = np.arange(12, dtype=np.uint8) % 2
    seen = {}

    class Spy(OracleBackedContext):
        def consensus_batch(self, groups=None, packed=None, trace=False, flags=None):
            seen[id(self)] = (len(packed[0]) - 1, None if flags is None else flags.tolist())
            return super().consensus_batch(packed=packed)

    ctxs = {0: Spy(), 1: Spy()}
    out = shard.consensus_batch_sharded(packed, devices=[0, 1], contexts=ctxs, flags=flags)
    for k, c in ctxs.items():
        n, f = seen[id(c)]
        idx = np.nonzero(out["owner"] == k)[0]
        assert n == len(idx) and f == flags[idx].tolist()
