"""The orientation step (reference utils/SpliceDefineConsensus.py:895, :900-907) as the library
computes it: mpoa_orient_batch = minimizer sketch (k 15, w 10) + colinear chaining, the seed-chain
stage of minimap2's map-ont.  Host code: runs without a GPU.  When mappy is importable the two
are diffed; it is not installable in the build image, so that test skips here."""
import numpy as np
import pytest

from helpers import random_seq
from mandalorion_b200 import consensus as C
from mandalorion_b200.poa import orient_batch, pack_groups
from mandalorion_b200.synth import make_groups, revcomp


def flipped(groups, seed):
    rng = np.random.default_rng(seed)
    out, truth = [], []
    for g in groups:
        fl = [bool(rng.random() < 0.5) for _ in g]
        out.append([revcomp(r) if f else r for r, f in zip(g, fl)])
        truth.append([f != fl[0] for f in fl])
    return out, truth


@pytest.mark.parametrize("cfg,n", [("cfg1", 60), ("cfg2", 40), ("cfg4", 3), ("cfg3", 3)])
def test_strand_of_every_read_of_an_isoform(built, cfg, n):
    groups, truth = flipped(make_groups(cfg, n), 5)
    packed = pack_groups(groups)
    cnt, strand = orient_batch(packed)
    k = 0
    for tr in truth:
        for flip in tr:
            assert cnt[k] == 1 and strand[k, 0] == (-1 if flip else 1)
            k += 1
    # thread count does not change the answer
    cnt1, strand1 = orient_batch(packed, n_threads=1)
    assert np.array_equal(cnt, cnt1) and np.array_equal(strand, strand1)


def test_unmapped_reads_are_dropped_and_chimeras_written_twice(built):
    rng = np.random.default_rng(1)
    t = random_seq(rng, 1500)
    junk = random_seq(rng, 1200)
    chimera = t[:700] + revcomp(t[700:].encode()).decode()          # second half inverted: two primary hits
    with_n = t[:300] + "N" * 50 + t[350:]
    group = [t, junk, chimera, revcomp(t.encode()).decode(), t[:40], with_n, t.lower()]
    cnt, strand = orient_batch(pack_groups([group]))
    assert cnt.tolist() == [1, 0, 2, 1, 0, 1, 1]
    assert strand[0, 0] == 1 and strand[3, 0] == -1 and sorted(strand[2].tolist()) == [-1, 1]
    # through the host layer: the junk read and the 40-mer vanish, the chimera appears twice (:902-907)
    reads = [("r%d" % i, s) for i, s in enumerate(group)]
    pg = C.prepare_group(reads, rng=np.random.RandomState(0), orienter_factory=C.NativeOrienter)
    if pg.sequences[0] in (t, t.lower(), with_n, revcomp(t.encode()).decode()):
        assert len(pg.sequences) == 6 and len(pg.seq_lengths) == 7


def test_empty_and_tiny_inputs(built):
    cnt, strand = orient_batch(pack_groups([[], ["ACGT"], ["", "ACGT"]]))
    assert cnt.tolist() == [0, 0, 0]
    cnt, _ = orient_batch(pack_groups([]))
    assert len(cnt) == 0


def test_pending_orientation_is_resolved_in_one_batch(built, monkeypatch):
    calls = []
    real = C.native_hits
    monkeypatch.setattr(C, "native_hits", lambda groups, n_threads=None: (calls.append(len(groups)), real(groups, n_threads))[1])
    monkeypatch.setattr(C, "mappy_available", lambda: False)
    groups, _ = flipped(make_groups("cfg1", 12), 3)
    pgs = [C.prepare_group([("g%d_%d" % (g, i), r.decode()) for i, r in enumerate(reads)], rng=np.random.RandomState(g))
           for g, reads in enumerate(groups)]
    assert all(pg.sequences is None for pg in pgs)
    C.orient_pending(pgs)
    assert calls == [12]
    for pg in pgs:
        first = pg.sequences[0]
        assert all(C.NativeOrienter(first).hits(s) == [1] for s in pg.sequences[:3])


@pytest.mark.skipif(not C.mappy_available(), reason="mappy is not installable in the build image")
def test_agrees_with_mappy(built):
    rng = np.random.default_rng(9)
    groups, _ = flipped(make_groups("cfg1", 40) + make_groups("cfg2", 20), 7)
    t = random_seq(rng, 1400)
    groups.append([t.encode(), random_seq(rng, 900).encode(), (t[:600] + revcomp(t[600:].encode()).decode()).encode()])
    bad = 0
    for reads in groups:
        reads = [r.decode() for r in reads]
        m = C.MappyOrienter(reads[0])
        want = [sorted(m.hits(r)) for r in reads]
        got = [sorted(h) for h in C.native_hits([reads])[0]]
        bad += want != got
    assert bad == 0, f"{bad}/{len(groups)} groups: strands / primary-hit counts differ from mappy"
