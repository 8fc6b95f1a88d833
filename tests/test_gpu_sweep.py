"""Large GPU-vs-oracle sweeps (-m gpu): thousands of groups of the named configs and a
hypothesis-driven fuzz over read length / depth / error rate / scoring parameters.  The oracle
(all host threads) is the checker; bit-exact consensus, scores, band-cell counts and
per-base node identities."""
import os

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from helpers import OracleParams, oracle_consensus_batch, pack_groups
from mandalorion_b200 import PoaContext, PoaParams
from mandalorion_b200.synth import GroupConfig, make_groups
from test_gpu_parity import assert_same

pytestmark = pytest.mark.gpu
THREADS = os.cpu_count() or 1


@pytest.mark.parametrize("cfg,n,first", [("cfg1", 1000, 0), ("cfg2", 2048, 20000), ("cfg4", 48, 100), ("cfg3", 48, 100)])
def test_parity_sweep_of_the_named_configs(gpu_ctx, cfg, n, first):
    groups = make_groups(cfg, n, first=first)
    packed = pack_groups(groups)
    want = oracle_consensus_batch(packed=packed, trace=True, n_threads=THREADS)
    got = gpu_ctx.consensus_batch(packed=packed, trace=True)
    assert_same(got, want, packed)
    assert (got["status"] == 0).all()


def test_junk_long_reads_leave_the_packed_path_and_still_match(gpu_ctx):
    """unrelated 7-9 kb reads: alignment scores fall far below the diagonal, the packed kernel hands the
    group to the int32 kernel (ST_RETRY_32) -- on the GPU, never on the CPU -- and the result is the oracle's"""
    rng = np.random.default_rng(4)
    groups = [["".join(rng.choice(list("ACGT"), int(rng.integers(7000, 9000)))) for _ in range(3)] for _ in range(3)]
    groups += make_groups("cfg3", 2)
    packed = pack_groups(groups)
    want = oracle_consensus_batch(packed=packed, trace=True, n_threads=THREADS)
    got = gpu_ctx.consensus_batch(packed=packed, trace=True)
    assert_same(got, want, packed)
    assert got["stats"]["n_retry_groups"] >= 3 and got["stats"]["n_kernel_launches"] >= 2


@settings(max_examples=25, deadline=None, suppress_health_check=list(HealthCheck), derandomize=True)
@given(seed=st.integers(0, 2 ** 31 - 1), len_lo=st.integers(20, 3000), spread=st.integers(0, 600),
       reads_hi=st.integers(3, 24), err=st.sampled_from([0.0, 0.002, 0.01, 0.03, 0.08, 0.2]),
       mix=st.sampled_from([(0.3, 0.35, 0.35), (1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0), (0.2, 0.4, 0.4)]),
       pk=st.sampled_from([{}, {}, {}, dict(simd_pn_i16=8, simd_pn_i32=4), dict(simd_pn_i16=32, simd_pn_i32=16),
                           dict(match=2), dict(wb=4, wf=0.0), dict(wb=60), dict(wb=130, simd_pn_i16=8),
                           dict(gap_open1=6, gap_ext1=3, gap_open2=30, gap_ext2=2, mismatch=6), dict(match=1, mismatch=9)]))
def test_fuzz_lengths_depths_errors_parameters(built, seed, len_lo, spread, reads_hi, err, mix, pk):
    cfg = GroupConfig("fz%d" % (seed % 1000), 6, 1, reads_hi, len_lo, len_lo + spread, "uniform", err, mix)
    rng = np.random.default_rng(seed)
    groups = make_groups(cfg, first=int(rng.integers(0, 10 ** 6)))
    packed = pack_groups(groups)
    want = oracle_consensus_batch(packed=packed, trace=True, params=OracleParams(**pk), n_threads=THREADS)
    with PoaContext(0, PoaParams(**pk)) as ctx:
        got = ctx.consensus_batch(packed=packed, trace=True)
    assert_same(got, want, packed)


@pytest.mark.parametrize("cfg,n", [("cfg3", 24), ("cfg1", 200), ("cfg2", 96), ("cfg4", 4)])
def test_seeded_groups_match_the_seeded_oracle(gpu_ctx, cfg, n):
    """`abpoa -S` (MPOA_FLAG_SEED): windowed alignment between minimizer anchors.  GPU == oracle bit for bit --
    consensus, per-read score (sum over windows + anchors), band cells, per-base nodes -- also in batches
    that mix seeded and unseeded groups."""
    groups = make_groups(cfg, n, first=50)
    packed = pack_groups(groups)
    for flags in (np.ones(n, np.uint8), (np.arange(n) % 3 == 0).astype(np.uint8)):
        want = oracle_consensus_batch(packed=packed, trace=True, n_threads=THREADS, flags=flags)
        got = gpu_ctx.consensus_batch(packed=packed, trace=True, flags=flags)
        assert_same(got, want, packed)
        assert got["stats"]["n_seed_groups"] == int(flags.sum()) == got["stats"]["n_seed_applied"]
        plain = gpu_ctx.consensus_batch(packed=packed, flags=flags)
        assert_same(plain, want, packed, check_trace=False)


def test_seeded_long_reads_with_structural_differences(gpu_ctx):
    """anchors next to long insertions / deletions, reads that share no anchor at all with their predecessor
    (one window = the whole graph), a junk read inside a seeded group"""
    rng = np.random.default_rng(21)
    groups = []
    for g in range(6):
        t = "".join(rng.choice(list("ACGT"), 9000))
        reads = []
        for r in range(5):
            s = list(t)
            for _ in range(int(rng.integers(0, 3))):
                pos = int(rng.integers(0, len(s) - 700))
                if rng.random() < 0.5:
                    del s[pos:pos + int(rng.integers(100, 600))]
                else:
                    s[pos:pos] = list(rng.choice(list("ACGT"), int(rng.integers(100, 600))))
            s = [("ACGT"[int(rng.integers(0, 4))] if rng.random() < 0.01 else c) for c in s]
            reads.append("".join(s))
        if g == 0:
            reads.insert(2, "".join(rng.choice(list("ACGT"), 8500)))       # unrelated read: no anchors, wide scores
        groups.append(reads)
    packed = pack_groups(groups)
    flags = np.ones(len(groups), np.uint8)
    want = oracle_consensus_batch(packed=packed, trace=True, n_threads=THREADS, flags=flags)
    got = gpu_ctx.consensus_batch(packed=packed, trace=True, flags=flags)
    assert_same(got, want, packed)


def test_seeded_three_stage_interface_and_many_chunks(gpu_ctx):
    """upload / run / fetch seed inside the upload (the caller's buffer is free when upload returns); the
    one-call entry seeds beside the kernels and publishes the anchors chunk by chunk (poa_capi.cu:
    publish_seeds) -- more groups than one chunk holds, so that launches really wait for later chunks.  Both
    must give the seeded oracle's answer."""
    n = 40
    groups = make_groups("cfg1", n, first=900)
    packed = pack_groups(groups)
    flags = (np.arange(n) % 5 != 0).astype(np.uint8)
    want = oracle_consensus_batch(packed=packed, n_threads=THREADS, flags=flags)
    one_call = gpu_ctx.consensus_batch(packed=packed, flags=flags)
    assert_same(one_call, want, packed, check_trace=False)
    gpu_ctx.upload(*packed, flags=flags)
    for _ in range(2):                                   # the anchors stay resident between runs
        st = gpu_ctx.run()
        got = gpu_ctx.fetch()
        got["stats"] = st
        assert_same(got, want, packed, check_trace=False)


def test_pipeline_of_contexts_gives_the_same_answers(built):
    """PoaPipeline: several batches in flight on one GPU (two contexts, each on its own stream and host
    thread).  Every batch must come back exactly as a lone PoaContext.consensus_batch() returns it."""
    from mandalorion_b200 import PoaPipeline
    batches = []
    for i, (cfg, n) in enumerate((("cfg1", 64), ("cfg2", 24), ("cfg3", 6), ("cfg1", 80), ("cfg4", 3), ("cfg2", 16))):
        packed = pack_groups(make_groups(cfg, n, first=100 * i))
        flags = (np.arange(n) % 2 == (i % 2)).astype(np.uint8) if cfg == "cfg3" else None
        batches.append((packed, flags))
    with PoaPipeline(device=0, depth=2) as pipe:
        futs = [pipe.submit(packed=p, flags=f) for p, f in batches]
        got = [f.result() for f in futs]
        again = list(pipe.map(batches[:2]))
    for (packed, flags), g in zip(batches + batches[:2], got + again):
        want = oracle_consensus_batch(packed=packed, n_threads=THREADS, flags=flags)
        assert_same(g, want, packed, check_trace=False)
