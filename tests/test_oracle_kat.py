"""Known-answer tests of the CPU oracle (authored here: the reference ships no tests and abPOA's
source is not reachable -- PARITY UNPINNED, see oracle/abpoa_oracle.cpp).  Every expected value
below is derived by hand from the scoring of `abpoa -M 5 -r 0`
(reference utils/SpliceDefineConsensus.py:917): match +5, mismatch -4, gap min(4+2g, 24+g)."""
import numpy as np
import pytest

from helpers import OracleParams, oracle_consensus_batch, random_seq

RNG = np.random.default_rng(12345)
T = random_seq(RNG, 120)


def run(groups, **kw):
    return oracle_consensus_batch(groups, trace=True, **kw)


def test_identical_reads_give_that_read():
    r = run([[T, T, T, T]])
    assert r["status"][0] == 0
    assert r["cons"][0].decode() == T
    # every alignment is a full-length match: 5 * len
    assert list(r["trace"]["read_score"]) == [0, 5 * len(T), 5 * len(T), 5 * len(T)]
    # aligned to the node created by the same position of read 1
    n = len(T)
    assert np.array_equal(r["trace"]["base_aln"][n:2 * n], np.arange(n))
    assert np.array_equal(r["trace"]["base_node"][n:2 * n], np.arange(n))


def test_single_read_group_and_empty_group():
    r = run([[T], []])
    assert r["cons"][0].decode() == T and r["status"][0] == 0
    assert r["cons"][1] == b"" and r["status"][1] == 1


def test_minority_substitution_is_outvoted():
    k = 60
    alt = T[:k] + ("A" if T[k] != "A" else "C") + T[k + 1:]
    r = run([[T, alt, T], [alt, T, T]])
    assert r["cons"][0].decode() == T
    assert r["cons"][1].decode() == T
    # one mismatch: 5*(n-1) - 4
    assert r["trace"]["read_score"][1] == 5 * (len(T) - 1) - 4


def test_minority_indels_are_absent():
    k = 50
    ins = T[:k] + "GG" + T[k:]
    dele = T[:k] + T[k + 3:]
    r = run([[T, ins, T, dele, T]])
    assert r["cons"][0].decode() == T
    # insertion of 2: 5n - (4+2*2); deletion of 3: 5(n-3) - (4+2*3)
    assert r["trace"]["read_score"][1] == 5 * len(T) - 8
    assert r["trace"]["read_score"][3] == 5 * (len(T) - 3) - 10
    # inserted bases are not aligned to any node and become their own nodes
    n = len(T)
    assert list(r["trace"]["base_aln"][n + k:n + k + 2]) in ([-1, -1],) or (r["trace"]["base_aln"][n:2 * n + 2] == -1).sum() == 2


def test_convex_gap_second_piece():
    # a 30-base deletion costs min(4+2*30, 24+30) = 54
    g = 30
    t = random_seq(np.random.default_rng(5), 200)
    q = t[:80] + t[80 + g:]
    r = run([[t, q]])
    assert r["trace"]["read_score"][1] == 5 * len(q) - 54
    # and a 30-base insertion likewise
    r = run([[q, t]])
    assert r["trace"]["read_score"][1] == 5 * len(q) - 54


def test_n_scores_zero():
    k = 33
    qn = T[:k] + "N" + T[k + 1:]
    r = run([[T, qn, T]])
    assert r["trace"]["read_score"][1] == 5 * (len(T) - 1)
    assert r["cons"][0].decode() == T
    # lower case is folded, anything outside ACGT is N
    r2 = run([[T, (T[:k] + "x" + T[k + 1:]).lower(), T]])
    assert r2["trace"]["read_score"][1] == 5 * (len(T) - 1)


def test_two_read_tie_takes_the_later_edge():
    # heaviest bundling, equal weights: the later out-edge wins when its score is >= (A.9)
    k = 40
    alt = T[:k] + ("A" if T[k] != "A" else "C") + T[k + 1:]
    r = run([[T, alt]])
    assert r["cons"][0].decode() == alt
    r = run([[T, alt]], params=OracleParams(hb_tie_later_wins=0))
    assert r["cons"][0].decode() == T


def test_lane_width_rule():
    # int16 lanes while qlen*5 <= 32767-4-4-2-24-1 = 32732 and max(qlen, nodes)*2+4 <= 32732
    rng = np.random.default_rng(9)
    a = random_seq(rng, 6546)
    b = random_seq(rng, 6547)
    r = run([[a, a], [b, b]])
    assert r["trace"]["read_bits"][1] == 16
    assert r["trace"]["read_bits"][3] == 32


def test_band_widening_and_vector_width_do_not_change_clean_groups():
    rng = np.random.default_rng(77)
    from mandalorion_b200.synth import GroupConfig, make_groups
    groups = make_groups(GroupConfig("kat", 6, 5, 9, 300, 600, "uniform", 0.01, (0.3, 0.35, 0.35)))
    base = oracle_consensus_batch(groups)["cons"]
    for p in (OracleParams(wb=60), OracleParams(simd_pn_i16=8, simd_pn_i32=4), OracleParams(simd_pn_i16=32, simd_pn_i32=16)):
        assert oracle_consensus_batch(groups, params=p)["cons"] == base
    del rng


def test_band_is_a_small_part_of_the_matrix():
    from mandalorion_b200.synth import make_groups
    st = oracle_consensus_batch(make_groups("cfg1", 3))["stats"]
    assert 0 < st["band_cells"] < st["full_cells"] / 5
    assert st["int_ops"] >= 17 * st["band_cells"]
    assert st["n_alignments"] == st["n_reads"] - st["n_groups"]


def test_progressive_order_matters_but_threads_do_not():
    from mandalorion_b200.synth import GroupConfig, make_groups
    groups = make_groups(GroupConfig("kat2", 12, 4, 10, 150, 300, "uniform", 0.08, (0.3, 0.35, 0.35)))
    a = oracle_consensus_batch(groups, n_threads=1)
    b = oracle_consensus_batch(groups, n_threads=4)
    assert a["cons"] == b["cons"] and a["stats"]["band_cells"] == b["stats"]["band_cells"]


@pytest.mark.parametrize("reads", [["ACGT", "", "ACGT"], ["A", "A", "A"], ["ACGT" * 5, "TTTT" * 5, "ACGT" * 5]])
def test_degenerate_inputs_do_not_crash(reads):
    r = run([reads])
    assert r["status"][0] in (0, 1)


def test_empty_first_read_means_no_output():
    # abPOA dies adding an empty first sequence -> empty stdout -> reference falls back (:924-925)
    r = run([["", "ACGT"]])
    assert r["status"][0] == 1 and r["cons"][0] == b""
