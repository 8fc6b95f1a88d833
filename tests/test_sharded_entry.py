"""The product's multi-GPU entry point.  On CPU: the native shard plan (mpoa_shard_plan: host code,
no GPU).  On the GPU (-m gpu): mpoa_consensus_batch_multi and mpoa_batch_upload_subset against the
oracle -- several contexts on one device stand in for several GPUs, the code path is the same."""
import ctypes as C

import numpy as np
import pytest

from helpers import oracle_consensus_batch, pack_groups
from mandalorion_b200 import shard
from mandalorion_b200.poa import shard_plan
from mandalorion_b200.synth import GroupConfig, make_groups

CFG = GroupConfig("shard2", 41, 0, 9, 60, 500, "loguniform", 0.03, (0.3, 0.35, 0.35))


def test_native_plan_is_balanced_deterministic_and_complete(built):
    packed = pack_groups(make_groups(CFG))
    ng = len(packed[0]) - 1
    cost = shard.group_costs(packed[0], packed[1])
    for n in (1, 2, 5):
        owner = shard_plan(packed, n)
        assert len(owner) == ng and owner.min() >= 0 and owner.max() < n
        assert owner.tolist() == shard_plan(packed, n).tolist()
        if n > 1:
            load = np.bincount(owner, weights=cost, minlength=n)
            assert load.max() / load.mean() < 1.25
    assert shard_plan(pack_groups([]), 3).tolist() == []
    # a seeded group is cheaper than the same group unseeded (windowed band), never free
    long_groups = pack_groups([[b"ACGT" * 3000] * 4, [b"ACGT" * 3000] * 4, [b"ACGT" * 3000] * 4])
    assert shard_plan(long_groups, 2, flags=np.array([1, 1, 0], np.uint8)).tolist() == [1, 1, 0]
    assert shard_plan(long_groups, 2).tolist() == [0, 1, 0]


@pytest.mark.gpu
def test_multi_context_batch_equals_the_oracle(gpu_ctx):
    from mandalorion_b200 import PoaContext
    from mandalorion_b200.poa import consensus_batch_multi
    groups = make_groups(CFG) + [[], [b"ACGTTGCA" * 30]]
    packed = pack_groups(groups)
    want = oracle_consensus_batch(packed=packed)
    extra = [PoaContext(0) for _ in range(4)]
    try:
        for n in (1, 2, 5):
            out = consensus_batch_multi(([gpu_ctx] + extra)[:n], packed)
            assert out["cons"] == want["cons"] and out["status"].tolist() == want["status"].tolist()
            assert sum(s["band_cells"] for s in out["stats"]) == want["stats"]["band_cells"]
            assert sum(s["n_groups"] for s in out["stats"]) == len(groups)
            assert out["owner"].tolist() == shard_plan(packed, n).tolist()
        # the same context twice is refused
        from mandalorion_b200 import PoaError
        with pytest.raises(PoaError):
            consensus_batch_multi([gpu_ctx, gpu_ctx], packed)
        # more contexts than groups: empty shards are fine
        tiny = pack_groups(groups[:2])
        out = consensus_batch_multi([gpu_ctx] + extra, tiny)
        assert out["cons"] == oracle_consensus_batch(packed=tiny)["cons"]
    finally:
        for c in extra:
            c.close()


@pytest.mark.gpu
def test_flags_travel_with_their_groups(gpu_ctx):
    from mandalorion_b200 import PoaContext
    from mandalorion_b200.poa import consensus_batch_multi
    cfg = GroupConfig("shard_seed", 10, 3, 5, 1300, 1800, "uniform", 0.02, (0.3, 0.35, 0.35))
    packed = pack_groups(make_groups(cfg))
    flags = (np.arange(10) % 2).astype(np.uint8)
    want = oracle_consensus_batch(packed=packed, flags=flags)
    other = PoaContext(0)
    try:
        out = consensus_batch_multi([gpu_ctx, other], packed, flags=flags)
    finally:
        other.close()
    assert out["cons"] == want["cons"]
    assert sum(s["n_seed_groups"] for s in out["stats"]) == 5
    assert sum(s["band_cells"] for s in out["stats"]) == want["stats"]["band_cells"]


@pytest.mark.gpu
def test_subset_upload_runs_only_the_selected_groups(gpu_ctx):
    groups = make_groups(CFG)
    packed = pack_groups(groups)
    flags = np.zeros(len(groups), np.uint8)
    for sel in ([0], [3, 4, 5, 17, 40], list(range(len(groups))), []):
        gpu_ctx.upload(*packed, flags=flags, subset=sel)
        st = gpu_ctx.run()
        out = gpu_ctx.fetch()
        want = oracle_consensus_batch(packed=pack_groups([groups[g] for g in sel]))
        assert out["cons"] == want["cons"] and st["band_cells"] == want["stats"]["band_cells"]
    from mandalorion_b200 import PoaError
    for bad in ([5, 3], [1, 1], [len(groups)], [-1]):
        with pytest.raises(PoaError):
            gpu_ctx.upload(*packed, subset=bad)


@pytest.mark.gpu
def test_multi_reports_a_too_small_result_buffer(gpu_ctx):
    from mandalorion_b200.poa import _load
    lib = _load()
    gro, rbo, bases = pack_groups(make_groups(CFG)[:8])
    ng = len(gro) - 1
    off = np.zeros(ng + 1, np.int64)
    buf = np.zeros(4, np.uint8)
    status = np.zeros(ng, np.int32)
    handles = (C.c_void_p * 1)(gpu_ctx._h)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    rc = lib.mpoa_consensus_batch_multi(handles, 1, ng, p(gro), p(rbo), p(bases), None, p(off), p(buf), 4, p(status), None, None)
    assert rc == -4 and off[ng] == sum(len(c) for c in oracle_consensus_batch(packed=(gro, rbo, bases))["cons"])


@pytest.mark.gpu
def test_two_shards_on_one_gpu_through_the_sharded_entry(gpu_ctx):
    """a device named twice: two shards, two contexts, kernels taking turns -- same answers, input order"""
    from mandalorion_b200 import PoaContext
    from mandalorion_b200.shard import consensus_batch_sharded
    packed = pack_groups(make_groups(CFG))
    want = oracle_consensus_batch(packed=packed)
    other = PoaContext(0)
    try:
        out = consensus_batch_sharded(packed, devices=[0, 0], contexts={0: [gpu_ctx, other]})
        assert out["cons"] == want["cons"] and list(out["status"]) == list(want["status"])
        assert len(out["stats"]) == 2 and set(out["owner"].tolist()) == {0, 1}
        assert sum(s["band_cells"] for s in out["stats"]) == want["stats"]["band_cells"]
        again = consensus_batch_sharded(packed, devices=[0, 0])          # contexts created and closed by the call
        assert again["cons"] == want["cons"]
    finally:
        other.close()
