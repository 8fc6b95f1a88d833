"""Host-side mirror of determine_consensus() and the module-D dispatch (no GPU: the CUDA context
is replaced by a test double backed by the oracle)."""
import os

import numpy as np
import pytest

from helpers import OracleBackedContext, oracle_consensus_batch, random_seq
from mandalorion_b200 import consensus as C
from mandalorion_b200.synth import GroupConfig, make_groups, revcomp


def reference_subsample(n):
    """The reference's RNG use, verbatim semantics (utils/SpliceDefineConsensus.py:884-885)."""
    return np.random.choice(np.arange(0, n), min(n, 100), replace=False)


@pytest.mark.parametrize("n", [1, 2, 3, 17, 100, 101, 250])
def test_subsample_consumes_the_global_rng_like_the_reference(n):
    reads = [(f"r{i}", "ACGT" * 10 + "A" * (i % 7)) for i in range(n)]
    np.random.seed(1234)
    want_idx = reference_subsample(n)
    want_next = np.random.random()
    np.random.seed(1234)
    pg = C.prepare_group(reads, orienter_factory=lambda first: type("O", (), {"hits": lambda self, s: [1]})())
    got_next = np.random.random()
    assert got_next == want_next                      # same stream position afterwards
    assert pg.names == [r[0] for r in reads]          # ALL names, not the subsample (:880-882, :931)
    assert pg.sequences == [reads[i][1] for i in want_idx]
    assert len(pg.sequences) == min(n, 100)           # the 100-read cap (:885)
    assert pg.bypass == (len(pg.sequences) <= 2)      # :911


def test_orientation_and_dropping():
    rng = np.random.default_rng(5)
    t = random_seq(rng, 600)
    other = random_seq(rng, 600)
    reads = [("a", t), ("b", revcomp(t.encode()).decode()), ("c", other), ("d", t[5:-7])]
    # a permutation whose first read belongs to the isoform (the reference aligns everything to read 0)
    seed = next(s for s in range(50) if np.random.RandomState(s).choice(np.arange(0, 4), 4, replace=False)[0] != 2)
    pg = C.prepare_group(reads, rng=np.random.RandomState(seed), orienter_factory=C.NativeOrienter)
    first = pg.sequences[0]
    # every kept read is in the orientation of the first one; the unrelated read is dropped,
    # but its length still counts for the -S decision (seq_lengths records every read, :901)
    assert len(pg.sequences) == 3 and len(pg.seq_lengths) == 4
    k = C.NativeOrienter(first)
    assert all(k.hits(s) == [1] for s in pg.sequences)
    # the batched path (orientation pending until the groups are flushed) gives the same group
    pg2 = C.prepare_group(reads, rng=np.random.RandomState(seed))
    if pg2.sequences is None:                          # mappy absent: deferred to one native call
        C.orient_pending([pg2])
    assert pg2.sequences == pg.sequences and pg2.bypass == pg.bypass and pg2.seq_lengths == pg.seq_lengths
    assert other not in pg.sequences and revcomp(other.encode()).decode() not in pg.sequences


def test_seed_flag_follows_the_median_length():
    long = [("r%d" % i, "A" * 8000) for i in range(3)]
    short = [("r%d" % i, "A" * 7999) for i in range(3)]
    one = lambda first: type("O", (), {"hits": lambda self, s: [1]})()   # noqa: E731
    assert C.prepare_group(long, orienter_factory=one).seed is True      # :916-919
    assert C.prepare_group(short, orienter_factory=one).seed is False


def test_determine_consensus_matches_the_reference_flow():
    groups = make_groups(GroupConfig("host", 5, 1, 9, 150, 300, "uniform", 0.03, (0.3, 0.35, 0.35)), with_names=True)
    ctx = OracleBackedContext()
    for reads in groups:
        np.random.seed(7)
        idx = reference_subsample(len(reads))
        np.random.seed(7)
        cons, names = C.determine_consensus(reads, "unused_root", "unused_abpoa", ctx=ctx, orienter_factory=C.NativeOrienter)
        assert names == [r[0] for r in reads]
        seqs = [reads[i][1] for i in idx]
        if len(seqs) <= 2:
            assert cons == seqs[0]
        else:
            want = oracle_consensus_batch([seqs])["cons"][0].decode()
            assert cons == (want or seqs[0])


def test_two_primary_hits_write_the_read_twice():
    # a read with two primary (supplementary) hits is written twice, re-reversed per '-' hit (:902-907)
    class TwoHits:
        def __init__(self, first):
            pass

        def hits(self, s):
            return [-1, -1] if s.startswith("TT") else [1]

    reads = [("a", "AACC"), ("b", "TTGG"), ("c", "AACC")]
    pg = C.prepare_group(reads, rng=np.random.RandomState(3), orienter_factory=TwoHits)
    assert pg.sequences.count("CCAA") == 1 and pg.sequences.count("TTGG") == 1
    assert len(pg.sequences) == 4


def test_empty_consensus_falls_back_to_first_read():
    class Failing(OracleBackedContext):
        def fetch(self, trace=False):
            out = super().fetch()
            out["status"] = np.ones_like(out["status"])
            out["cons"] = [b""] * len(out["cons"])
            return out

    reads = [("r%d" % i, "ACGTACGTAC") for i in range(4)]
    cons, _ = C.determine_consensus(reads, ctx=Failing(), rng=np.random.RandomState(1),
                                    orienter_factory=lambda f: type("O", (), {"hits": lambda self, s: [1]})())
    assert cons == "ACGTACGTAC"                          # :924-925


def test_locus_dispatch_and_writer(tmp_path):
    cfg = GroupConfig("host2", 7, 1, 6, 120, 200, "uniform", 0.02, (0.3, 0.35, 0.35))
    groups = make_groups(cfg, with_names=True)
    loci = [("chr1~100~900", {"1": groups[0], "2": groups[1], "3": groups[2]}),
            ("chr1~2000~2900", {"1": groups[3]}),
            ("chr2~50~700", {"1": groups[4], "2": groups[5], "3": groups[6]})]
    ctx = OracleBackedContext()
    np.random.seed(99)
    res = C.consensus_for_loci(loci, ctx=ctx, orienter_factory=C.NativeOrienter)
    assert ctx.calls <= 1                                # ONE batched call for all loci
    # the serial reference flow with the same seed gives the same IsoData
    np.random.seed(99)
    for root, seq_dict in loci:
        for isoform, reads in seq_dict.items():
            cons, names = C.determine_consensus(reads, ctx=OracleBackedContext(), orienter_factory=C.NativeOrienter)
            assert res[root][isoform] == [cons, names]
    n = C.write_isoform_files([r for r, _ in loci], res, str(tmp_path))
    assert n == 7
    fa = open(os.path.join(tmp_path, "Isoform_Consensi.fasta")).read().splitlines()
    assert fa[0] == ">Isoform1_%d" % len(groups[0]) and fa[1] == res["chr1~100~900"]["1"][0]
    assert fa[6] == ">Isoform4_%d" % len(groups[3])
    r2i = open(os.path.join(tmp_path, "reads2isoforms.txt")).read().splitlines()
    assert len(r2i) == sum(len(g) for g in groups)
    assert r2i[0] == "%s\tIsoform1_%d" % (groups[0][0][0], len(groups[0]))


def test_abpoa_cli_argument_protocol(tmp_path):
    from mandalorion_b200 import abpoa_cli
    p, seed, path = abpoa_cli.parse_args(["-M", "5", "-r", "0", "-S", "x.fasta"])
    assert (p.match, seed, path) == (5, True, "x.fasta")
    fa = tmp_path / "in.fasta"
    fa.write_text(">a\nACGT\nAC\n>b\nGGTT\n")
    assert abpoa_cli.read_fasta(str(fa)) == ["ACGTAC", "GGTT"]
    with pytest.raises(SystemExit):
        abpoa_cli.parse_args(["-r", "1", "x.fasta"])


def test_pipeline_host_logic_without_a_gpu():
    """PoaPipeline is host logic over objects with a consensus_batch() method: results come back per
    submitted batch, in submission order through map(), errors surface through the future, and contexts
    handed in by the caller are not closed."""
    from mandalorion_b200.poa import PoaPipeline
    from mandalorion_b200.synth import make_groups
    from helpers import pack_groups, oracle_consensus_batch

    class Ctx(OracleBackedContext):
        closed = False

        def close(self):
            self.closed = True

    batches = [(pack_groups(make_groups("cfg1", 6, first=10 * i)), None) for i in range(5)]
    ctxs = [Ctx(), Ctx()]
    with PoaPipeline(contexts=ctxs) as pipe:
        got = list(pipe.map(batches))
        bad = pipe.submit()                                               # neither groups nor packed
        with pytest.raises(Exception):
            bad.result(timeout=60)
        again = pipe.submit(packed=batches[0][0]).result(timeout=60)      # the pipeline survives a failed batch
    for (packed, _), g in zip(batches, got):
        assert g["cons"] == oracle_consensus_batch(packed=packed)["cons"]
        assert "wall" in g and g["wall"][1] >= g["wall"][0]
    assert again["cons"] == got[0]["cons"]
    assert sum(c.calls for c in ctxs) == 6 and not any(c.closed for c in ctxs)
