"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the CPU
oracle on the same seeded inputs, against the committed golden vectors, and -- at sizes the
oracle is too slow for -- through size-independent properties.  Bar: bit-exact (integer DP).

The oracle is the checker only; PARITY UNPINNED against real abPOA (see oracle/abpoa_oracle.cpp)."""
import hashlib
import json
import os

import numpy as np
import pytest

from helpers import OracleParams, oracle_consensus_batch, pack_groups, random_seq
from mandalorion_b200 import PoaContext, PoaParams
from mandalorion_b200.synth import GroupConfig, make_groups

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def assert_same(got, want, packed, check_trace=True):
    gro, rbo, _ = packed
    assert list(got["status"]) == list(want["status"])
    for gi, (a, b) in enumerate(zip(got["cons"], want["cons"])):
        assert a == b, f"group {gi}: consensus differs"
    if check_trace:
        for key in ("read_score", "read_bits", "read_band_cells"):
            ok = want["status"][np.searchsorted(gro, np.arange(len(rbo) - 1), side="right") - 1] == 0
            assert np.array_equal(got["trace"][key][ok], want["trace"][key][ok]), key
        okb = np.repeat(ok, np.diff(rbo))
        assert np.array_equal(got["trace"]["base_aln"][okb], want["trace"]["base_aln"][okb])
        assert np.array_equal(got["trace"]["base_node"][okb], want["trace"]["base_node"][okb])
    assert got["stats"]["band_cells"] == want["stats"]["band_cells"]
    assert got["stats"]["int_ops"] == want["stats"]["int_ops"]
    assert got["stats"]["n_alignments"] == want["stats"]["n_alignments"]


def both(gpu_ctx, groups, params=None):
    packed = pack_groups(groups)
    want = oracle_consensus_batch(packed=packed, trace=True, n_threads=os.cpu_count() or 1,
                                  params=params or OracleParams())
    got = gpu_ctx.consensus_batch(packed=packed, trace=True)
    assert got["stats"]["n_kernel_launches"] >= 1       # the CUDA kernels ran: nothing else can produce output
    assert_same(got, want, packed)
    # the production instantiation of the kernel (no trace outputs compiled in) must agree as well
    plain = gpu_ctx.consensus_batch(packed=packed)
    assert_same(plain, want, packed, check_trace=False)
    return got, want


def test_hand_cases(gpu_ctx):
    t = "ACGTTGCATGCCGATAGCTAGCTAGGATCGATCGATTAGCTAGCTAACG"
    groups = [[t, t, t], [t], [t, t[:20] + "G" + t[21:], t, t[:30] + t[31:], t[:10] + "TT" + t[10:]],
              ["ACGT", "ACGT", "AGGT"], ["A", "A", "A"], ["ACGTN" * 8, "ACGTA" * 8, "ACGTN" * 8],
              [t.lower(), t, "acgtnnnryk" + t[10:]]]
    got, _ = both(gpu_ctx, groups)
    assert got["cons"][0].decode() == t and got["cons"][1].decode() == t


def test_empty_and_ragged_inputs(gpu_ctx):
    t = random_seq(np.random.default_rng(3), 90)
    groups = [[], [t], [t, "", t], ["", t], [t, t[:5]], [t[:3], t, t], ["A"], [t] * 40]
    got, want = both(gpu_ctx, groups)
    assert got["status"][0] == 1 and got["cons"][0] == b""
    assert got["status"][3] == 1            # empty first read: abpoa prints nothing
    # zero groups is a valid call
    z = gpu_ctx.consensus_batch([])
    assert z["cons"] == [] and len(z["status"]) == 0


@pytest.mark.parametrize("name,cfg", [
    ("small", GroupConfig("p_small", 96, 3, 12, 60, 400, "uniform", 0.03, (0.3, 0.35, 0.35))),
    ("noisy", GroupConfig("p_noisy", 64, 3, 20, 100, 600, "uniform", 0.10, (0.3, 0.35, 0.35))),
    ("very_noisy", GroupConfig("p_vnoisy", 48, 3, 10, 80, 300, "uniform", 0.25, (0.3, 0.35, 0.35))),
    ("ccs_deep", GroupConfig("p_ccs", 12, 50, 120, 300, 400, "pm5", 0.002, (0.2, 0.4, 0.4))),
])
def test_seeded_groups_match_oracle(gpu_ctx, name, cfg):
    both(gpu_ctx, make_groups(cfg))


def test_baseline_config_slices_match_oracle(gpu_ctx):
    # slices of BASELINE.json configs at their real shapes (sizes the oracle finishes in seconds)
    both(gpu_ctx, make_groups("cfg1", 48))
    both(gpu_ctx, make_groups("cfg2", 32, first=1000))
    both(gpu_ctx, make_groups("cfg4", 4))


def test_long_isoforms_int32_lanes(gpu_ctx):
    # cfg3: 5-12 kb reads, wide band; reads above 6546 nt take abPOA's int32 lane width (pn 8)
    got, want = both(gpu_ctx, make_groups("cfg3", 6))
    assert got["stats"]["n_align_i32"] > 0
    # 5.5 kb reads stay in int16 lanes (5*qlen <= 32732) with a wide band
    cfg = GroupConfig("p_long16", 3, 4, 6, 5400, 5600, "uniform", 0.01, (0.3, 0.35, 0.35))
    got, want = both(gpu_ctx, make_groups(cfg))
    assert got["stats"]["n_align_i16"] > 0 and got["stats"]["n_align_i32"] == 0


def test_unrelated_reads_stress_band_edges(gpu_ctx):
    rng = np.random.default_rng(7)
    junk = [[random_seq(rng, int(rng.integers(5, 160))) for _ in range(int(rng.integers(2, 9)))] for _ in range(128)]
    both(gpu_ctx, junk)
    # very different lengths inside one group
    rag = [[random_seq(rng, int(n)) for n in rng.integers(1, 700, size=6)] for _ in range(32)]
    both(gpu_ctx, rag)


def _with_long_indels(rng, template, n_events):
    """a read of `template` with a few LONG insertions / deletions (20-150 nt) and sparse point errors"""
    seq = list(template)
    for _ in range(n_events):
        pos = int(rng.integers(0, max(1, len(seq) - 1)))
        ln = int(rng.integers(20, 150))
        if rng.random() < 0.5:
            seq[pos:pos] = list(random_seq(rng, ln))
        else:
            del seq[pos:pos + ln]
    for k in range(len(seq)):
        if rng.random() < 0.01:
            seq[k] = "ACGT"[int(rng.integers(0, 4))]
    return "".join(seq) or "A"


def test_long_indels_cross_lanes(gpu_ctx):
    # insertions longer than a lane's cells make the insertion recurrence (the rotated max-scan of
    # the lane totals) carry over several lanes; long deletions make the band jump by whole vectors
    rng = np.random.default_rng(11)
    groups = []
    for g in range(48):
        t = random_seq(rng, int(rng.integers(300, 1800)))
        groups.append([_with_long_indels(rng, t, int(rng.integers(0, 4))) for _ in range(int(rng.integers(3, 9)))])
    both(gpu_ctx, groups)
    # the same with a narrow and a wide fixed band (128-cell and 512-cell kernels)
    for pk in (dict(wb=4, wf=0.0), dict(wb=150, wf=0.0)):
        packed = pack_groups(groups[:16])
        want = oracle_consensus_batch(packed=packed, trace=True, params=OracleParams(**pk), n_threads=os.cpu_count() or 1)
        with PoaContext(0, PoaParams(**pk)) as ctx:
            got = ctx.consensus_batch(packed=packed, trace=True)
        assert_same(got, want, packed)


@pytest.mark.parametrize("pk", [dict(simd_pn_i16=8, simd_pn_i32=4), dict(simd_pn_i16=32, simd_pn_i32=16),
                                dict(wb=4, wf=0.0), dict(match=2, mismatch=4), dict(wb=40),
                                dict(wb=120),                       # band 257..512 cells: the 16-cells-per-lane variant
                                dict(wb=120, simd_pn_i16=8),        # ... with lanes not aligned to SIMD vectors
                                dict(wb=300),                       # band > 512 cells: int32 kernel, wide ring
                                dict(gap_open1=6, gap_ext1=3, gap_open2=30, gap_ext2=2, mismatch=6)])
def test_other_parameters(built, pk):
    lo, hi = (600, 1500) if pk.get("wb", 0) >= 120 else (100, 500)     # reads long enough to fill a wide band
    groups = make_groups(GroupConfig("p_par", 32 if hi == 500 else 12, 3, 10, lo, hi, "uniform", 0.06, (0.3, 0.35, 0.35)))
    packed = pack_groups(groups)
    want = oracle_consensus_batch(packed=packed, trace=True, params=OracleParams(**pk))
    with PoaContext(0, PoaParams(**pk)) as ctx:
        got = ctx.consensus_batch(packed=packed, trace=True)
    assert_same(got, want, packed)


def test_golden_vectors(built):
    with open(os.path.join(HERE, "golden", "poa_golden.json")) as fh:
        golden = json.load(fh)
    for case in golden["cases"]:
        with PoaContext(0, PoaParams(**case["params"])) as ctx:
            got = ctx.consensus_batch(case["groups"], trace=True, flags=[case.get("seed_flag", 0)] * len(case["groups"]))
        assert [c.decode() for c in got["cons"]] == case["consensus"], case["name"]
        assert [int(s) for s in got["status"]] == case["status"]
        assert [int(x) for x in got["trace"]["read_score"]] == case["read_score"]
        assert [int(x) for x in got["trace"]["read_band_cells"]] == case["read_band_cells"]
        assert hashlib.sha1(got["trace"]["base_node"].tobytes()).hexdigest() == case["base_node_sha1"]
        assert hashlib.sha1(got["trace"]["base_aln"].tobytes()).hexdigest() == case["base_aln_sha1"]
        assert got["stats"]["band_cells"] == case["band_cells"]


def test_retry_paths_give_the_same_answer(built):
    # debug_small_caps schedules the first attempt with a workspace sized for the longest first read
    # only and the narrowest kernel variant: groups overflow (node capacity / band width / int16
    # lanes) and are re-run ON THE GPU with larger capacities or a wider variant -- never on the CPU
    for groups, min_retry in ((make_groups("cfg1", 24), 4),
                              (make_groups("cfg1", 6) + make_groups("cfg3", 1) + make_groups("cfg4", 1), 1)):
        packed = pack_groups(groups)
        want = oracle_consensus_batch(packed=packed, trace=True, n_threads=os.cpu_count() or 1)
        with PoaContext(0, PoaParams(debug_small_caps=1)) as ctx:
            got = ctx.consensus_batch(packed=packed, trace=True)
        assert_same(got, want, packed)
        assert got["stats"]["n_retry_groups"] >= min_retry
        assert got["stats"]["n_kernel_launches"] >= 2


def test_properties_at_scale(gpu_ctx):
    """Size-independent properties on a batch too large for the oracle in test time."""
    groups = make_groups("cfg2", 1024, first=5000)
    packed = pack_groups(groups)
    a = gpu_ctx.consensus_batch(packed=packed)
    assert (a["status"] == 0).all()
    # (1) deterministic run to run
    b = gpu_ctx.consensus_batch(packed=packed)
    assert a["cons"] == b["cons"] and a["stats"]["band_cells"] == b["stats"]["band_cells"]
    # (2) independent of batch composition / scheduling order: reversed batch, same per-group answers
    rev = gpu_ctx.consensus_batch(packed=pack_groups(groups[::-1]))
    assert rev["cons"][::-1] == a["cons"]
    # (3) idempotence: a group made of copies of its own consensus returns that consensus
    again = gpu_ctx.consensus_batch([[c, c, c] for c in a["cons"][:256]])
    assert again["cons"] == a["cons"][:256]
    # (4) consensus of a ~1 % error group is close to every read's length and is ACGT only
    for g, c in zip(groups[:64], a["cons"][:64]):
        med = np.median([len(r) for r in g])
        assert abs(len(c) - med) <= 0.02 * med + 12
        assert set(c) <= set(b"ACGT")
    # (5) spot-check 16 groups of this batch against the oracle
    want = oracle_consensus_batch(groups[:16], n_threads=os.cpu_count() or 1)
    assert a["cons"][:16] == want["cons"]


def test_three_stage_interface_reuses_resident_inputs(gpu_ctx):
    packed = pack_groups(make_groups("cfg1", 16))
    gpu_ctx.upload(*packed)
    s1 = gpu_ctx.run()
    s2 = gpu_ctx.run()
    out = gpu_ctx.fetch()
    want = oracle_consensus_batch(packed=packed)
    assert out["cons"] == want["cons"] and s1["band_cells"] == s2["band_cells"] == want["stats"]["band_cells"]
    assert s1["kernel_ms"] > 0 and s1["tb_bytes"] >= s1["band_cells"]


def test_dstep_files_match_the_reference_run(gpu_ctx, tmp_path):
    """Module-D file parity on the GPU.  tests/golden/dstep/ was frozen in the build container by
    tests/test_dstep_reference.py: the UNMODIFIED reference (its own main, fork pool, process_locus,
    determine_consensus, writer) on synthetic tmp_SS/*.psl, with mappy stubbed and abpoa = oracle.
    Here the prepared groups (subsample order + orientation, as computed inside the reference's
    workers) go through the CUDA library in ONE batch and the two output files must be
    byte-identical to what the reference wrote."""
    from mandalorion_b200 import consensus as b200
    gold = os.path.join(HERE, "golden", "dstep")
    frozen = json.load(open(os.path.join(gold, "prepared.json")))
    prepared = {}
    for g in frozen["groups"]:
        pg = b200.PendingGroup(names=g["names"], sequences=g["sequences"], seq_lengths=[], bypass=g["bypass"],
                               seed=g["seed_flag"])
        if pg.bypass:
            pg.consensus = pg.sequences[0]
        prepared.setdefault(g["root"], {})[g["isoform"]] = pg
    results = b200.finish_prepared(prepared, ctx=gpu_ctx)
    b200.write_isoform_files(frozen["roots"], results, str(tmp_path))
    for name in ("Isoform_Consensi.fasta", "reads2isoforms.txt"):
        assert open(os.path.join(tmp_path, name), "rb").read() == open(os.path.join(gold, name), "rb").read(), name


def test_graft_smoke(built):
    import __graft_entry__ as ge
    ge.smoke()
