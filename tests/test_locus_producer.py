"""Group producer: mandalorion_b200.locus against the UNMODIFIED
reference functions on synthetic spliced loci.

With /root/reference present (build container) the reference's own process_locus() runs on the
same tmp_SS/<root>.psl with determine_consensus() replaced by a recorder, so what is compared is its
`seqDict` -- group numbering, members, order -- and the state of NumPy's global RNG afterwards (the
producer must consume it exactly like the reference: determine_consensus() subsamples from the same
stream next, utils/SpliceDefineConsensus.py:884).  The digests of those runs are frozen in
tests/golden/locus_producer.json, so the same cases are checked where the reference is absent
(regenerate with REGEN_LOCUS_GOLDEN=1)."""
import hashlib
import json
import os

import numpy as np
import pytest

from dstep_synth import write_locus, write_spliced_locus
from mandalorion_b200 import locus

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
GOLD = os.path.join(HERE, "golden", "locus_producer.json")
JUNCTIONS = "gtag,gcag,atac,ctac,ctgc,gtat".split(",")
HAVE_REF = os.path.isfile(os.path.join(REF, "defineIsoforms.py"))


def _mono_lines(chrom, start, n, rng, tag):
    out = []
    for r in range(n):
        a = start + int(rng.integers(0, 6))
        b = start + 400 - int(rng.integers(0, 6))
        cols = ["0"] * 24
        cols[8], cols[9], cols[10], cols[11], cols[12], cols[13] = "+", f"{tag}{r}", "400", "3", "397", chrom
        cols[15], cols[16], cols[17], cols[18], cols[19], cols[20] = str(a), str(b), "1", f"{b - a},", "0,", f"{a},"
        cols[21], cols[22], cols[23] = "0.95", "=" + "A" * (b - a), "ACGT" * 100
        out.append("\t".join(cols))
    return out


def build_case(name, path):
    """Returns (root, chrom, left_bounds, right_bounds, params) and writes the locus file."""
    seed = int(hashlib.sha1(name.encode()).hexdigest()[:6], 16)
    rng = np.random.Generator(np.random.PCG64(seed))
    none = {"5": [], "3": []}
    par = dict(splice_site_width=1, minimum_read_count=2, cutoff=0.1, upstream_buffer=10, downstream_buffer=50)
    if name == "plus_gene":
        root = write_spliced_locus(path, "chr1", 10000, rng)
    elif name == "minus_gene_deep":
        root = write_spliced_locus(path, "chr7", 250000, rng, n_reads=400, n_exons=7, strand="-")
    elif name == "noncanonical_intron":
        root = write_spliced_locus(path, "chr2", 5000, rng, n_reads=150, noncanonical=0)
    elif name == "noisy_wide_window":
        root = write_spliced_locus(path, "chr3", 90000, rng, n_reads=200, err=0.06)
        par.update(splice_site_width=3, minimum_read_count=3, upstream_buffer=5, downstream_buffer=20)
    elif name == "with_mono_and_foreign_reads":
        extra = _mono_lines("chr4", 700, 9, rng, "mono") + _mono_lines("chr4", 960, 4, rng, "tail") + \
            _mono_lines("chr9", 700, 3, rng, "foreign")
        root = write_spliced_locus(path, "chr4", 3000, rng, n_reads=90, extra_lines=extra)
    elif name == "underscore_chromosome":
        extra = _mono_lines("chrUn_KI270", 100, 6, rng, "m")
        root = write_spliced_locus(path, "chrUn_KI270", 2000, rng, n_reads=80, extra_lines=extra)
    elif name == "annotated_sites":
        root = write_spliced_locus(path, "chr5", 40000, rng, n_reads=160)
        # annotation near the first two junctions (taken from the reads themselves) plus a far-away cluster
        lf = locus.read_locus(os.path.join(path, root + ".psl"))
        ends = np.concatenate([(b + s)[:-1] for b, s in zip(lf.bstart, lf.bsize)])
        starts = np.concatenate([b[1:] for b in lf.bstart])
        top_l = int(np.bincount(ends - ends.min()).argmax() + ends.min())
        top_r = int(np.bincount(starts - starts.min()).argmax() + starts.min())
        return root, "chr5", {"5": [top_l, top_l + 1, top_l + 9], "3": [40010, 40011]}, {"5": [], "3": [top_r, top_r - 2]}, par
    elif name == "mono_exonic_only":
        root = write_locus(path, "chr6", 1000, [(0, 420, 6), (2000, 300, 3), (2100, 300, 5)], rng)
    elif name == "thin_evidence":
        root = write_spliced_locus(path, "chr8", 7000, rng, n_reads=9, n_exons=3)
    else:
        raise KeyError(name)
    return root, root.split("~")[0], none, none, par


CASES = ["plus_gene", "minus_gene_deep", "noncanonical_intron", "noisy_wide_window", "with_mono_and_foreign_reads",
         "underscore_chromosome", "annotated_sites", "mono_exonic_only", "thin_evidence"]
SEED = 5


def digest(groups):
    h = hashlib.sha1()
    for k, reads in groups.items():
        h.update(repr((k, [(n, s) for n, s in reads])).encode())
    return h.hexdigest()


def run_ours(path, root, chrom, lb, rb, par):
    np.random.seed(SEED)
    groups = locus.locus_groups(os.path.join(path, root + ".psl"), chrom, lb, rb, par["splice_site_width"],
                                par["minimum_read_count"], JUNCTIONS, par["cutoff"], par["upstream_buffer"],
                                par["downstream_buffer"])
    return groups, float(np.random.random())


def run_reference(path, root, chrom, lb, rb, par):
    from test_dstep_reference import load_reference_process_locus
    sdc, process_locus = load_reference_process_locus()
    process_locus.__globals__["upstream_buffer"] = par["upstream_buffer"]
    process_locus.__globals__["downstream_buffer"] = par["downstream_buffer"]
    orig = sdc.determine_consensus
    sdc.determine_consensus = lambda reads, root, abpoa: (reads, None)
    try:
        np.random.seed(SEED)
        _, start, end = root.split("~")
        iso = process_locus(path, root, chrom, lb, rb, int(start), int(end), par["splice_site_width"],
                            par["minimum_read_count"], JUNCTIONS, par["cutoff"], "unused")
    finally:
        sdc.determine_consensus = orig
    return {k: v[0] for k, v in iso.items()}, float(np.random.random())


@pytest.mark.parametrize("name", CASES)
def test_groups_match_frozen_reference_runs(name, tmp_path):
    root, chrom, lb, rb, par = build_case(name, str(tmp_path))
    groups, nxt = run_ours(str(tmp_path), root, chrom, lb, rb, par)
    gold = json.load(open(GOLD))[name]
    assert len(groups) == gold["n_groups"] and sum(map(len, groups.values())) == gold["n_reads"]
    assert digest(groups) == gold["digest"]
    assert nxt == gold["next_random"]          # the global RNG was consumed exactly like the reference does


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference is only present in the build container")
@pytest.mark.parametrize("name", CASES)
def test_groups_match_the_unmodified_reference(name, tmp_path, capsys):
    root, chrom, lb, rb, par = build_case(name, str(tmp_path))
    want, want_next = run_reference(str(tmp_path), root, chrom, lb, rb, par)
    got, got_next = run_ours(str(tmp_path), root, chrom, lb, rb, par)
    assert list(got) == list(want)
    for k in want:
        assert got[k] == want[k], k
    assert got_next == want_next
    if os.environ.get("REGEN_LOCUS_GOLDEN"):
        gold = json.load(open(GOLD)) if os.path.exists(GOLD) else {}
        gold[name] = dict(n_groups=len(want), n_reads=sum(map(len, want.values())), digest=digest(want), next_random=want_next)
        json.dump(gold, open(GOLD, "w"), indent=1, sort_keys=True)


def test_cases_are_not_trivial():
    gold = json.load(open(GOLD))
    assert gold["plus_gene"]["n_groups"] >= 3 and gold["minus_gene_deep"]["n_groups"] >= 3
    assert gold["mono_exonic_only"]["n_groups"] >= 2
    # a junction without an allowed motif is not called: its reads are dropped, fewer reads survive
    assert 0 < gold["noncanonical_intron"]["n_reads"] < 150


def test_cs_track_answers_like_a_walk_over_the_string():
    t = locus._CsTrack("=ACGTA*ag=CC+tt=GGGGG-ac=TTTTT~gt100ag=AAAAAA*ct=CCCCC", 1000)
    # genome positions: 5 matches -> 1005, sub -> 1006, 2 matches -> 1008, ins (none), 5 -> 1013, del 2 -> 1015,
    # 5 -> 1020, intron -> 1120, 6 -> 1126 ...
    bases, left, right = t.around(1119, 1121)
    assert bases == "gtag" and left == b"=====" and right == b"====="
    bases, left, right = t.around(1016, 1016)             # 4 entries before the intron entry: inside the +-10 window
    assert bases == "gtag"
    assert t.around(5000, 5001) == ("nnnn", b"", b"")
    assert t.around(1003, 1003) == ("nnnn", b"", b"")     # window holds no intron


@pytest.mark.skipif(not HAVE_REF, reason="/root/reference is only present in the build container")
@pytest.mark.parametrize("seed", range(8))
def test_random_loci_and_parameters_against_the_unmodified_reference(seed, tmp_path):
    """Random gene shapes, depths, error rates and D-step parameters (window, counts, cutoff, buffers)."""
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    extra = _mono_lines("chrR", 300, int(rng.integers(0, 8)), rng, "mono") if seed % 2 else []
    root = write_spliced_locus(str(tmp_path), "chrR", int(rng.integers(1000, 500000)), rng, n_reads=int(rng.integers(20, 180)),
                               n_exons=int(rng.integers(2, 8)), strand="+-"[seed % 2], err=float(rng.choice([0.005, 0.02, 0.05, 0.09])),
                               noncanonical=None if seed % 3 else 1, extra_lines=extra)
    par = dict(splice_site_width=int(rng.integers(1, 5)), minimum_read_count=int(rng.integers(1, 5)),
               cutoff=float(rng.choice([0.01, 0.1, 0.3])), upstream_buffer=int(rng.integers(3, 15)),
               downstream_buffer=int(rng.integers(10, 60)))
    none = {"5": [], "3": []}
    want, want_next = run_reference(str(tmp_path), root, "chrR", none, none, par)
    got, got_next = run_ours(str(tmp_path), root, "chrR", none, none, par)
    assert list(got) == list(want) and all(got[k] == want[k] for k in want)
    assert got_next == want_next
