"""Known-answer tests of the oracle's `abpoa -S` path (groups flagged MPOA_FLAG_SEED; reference
utils/SpliceDefineConsensus.py:916-919).  The restatement is LOW confidence (upstream abpoa_seed.c is
unavailable, see oracle/abpoa_oracle.cpp): these tests pin its structure, not abPOA's constants."""
import numpy as np

from helpers import OracleParams, oracle_consensus_batch, random_seq
from mandalorion_b200.synth import make_groups


def test_without_anchors_the_seeded_path_is_the_unseeded_one():
    # min_w longer than the reads: no anchor can be kept, one window = the whole graph
    groups = make_groups("cfg1", 12)
    a = oracle_consensus_batch(groups, trace=True)
    b = oracle_consensus_batch(groups, trace=True, flags=[1] * 12, params=OracleParams(seed_min_w=100000))
    assert a["cons"] == b["cons"] and a["stats"]["band_cells"] == b["stats"]["band_cells"]
    assert np.array_equal(a["trace"]["read_score"], b["trace"]["read_score"])
    assert np.array_equal(a["trace"]["base_node"], b["trace"]["base_node"])
    assert b["stats"]["n_seed_groups"] == 12 and a["stats"]["n_seed_groups"] == 0
    # the switch that ignores the flag
    c = oracle_consensus_batch(groups, flags=[1] * 12, params=OracleParams(honour_seed_flag=0))
    assert c["cons"] == a["cons"] and c["stats"]["n_seed_groups"] == 0


def test_windows_shrink_the_dp_and_keep_clean_consensi():
    rng = np.random.default_rng(8)
    t = random_seq(rng, 9000)
    group = [t, t[:4000] + t[4003:], t, t[:7000] + "ACG" + t[7000:], t]
    a = oracle_consensus_batch([group], trace=True)
    b = oracle_consensus_batch([group], trace=True, flags=[1])
    assert a["cons"][0].decode() == t == b["cons"][0].decode()
    assert b["stats"]["band_cells"] * 3 < a["stats"]["band_cells"]       # bands of a ~500-nt window, not of a 9 kb read
    # score of a seeded read = windows + k * match per anchor: identical reads score 5 per base either way
    assert b["trace"]["read_score"][2] == 5 * len(t) == a["trace"]["read_score"][2]
    assert b["trace"]["read_bits"][2] == 16 and a["trace"]["read_bits"][2] == 32   # windows fit abPOA's int16 lanes


def test_anchor_spacing_follows_min_w():
    rng = np.random.default_rng(9)
    t = random_seq(rng, 3000)
    cells = []
    for min_w in (200, 500, 1400):
        r = oracle_consensus_batch([[t, t, t]], flags=[1], params=OracleParams(seed_min_w=min_w))
        assert r["cons"][0].decode() == t
        cells.append(r["stats"]["band_cells"])
    assert cells[0] < cells[1] < cells[2]                                  # more anchors, narrower windows, fewer cells


def test_seeded_and_unseeded_groups_share_a_batch():
    groups = make_groups("cfg1", 9)
    flags = [1, 0, 0, 1, 1, 0, 1, 0, 0]
    mixed = oracle_consensus_batch(groups, flags=flags, n_threads=3)
    for i, g in enumerate(groups):
        alone = oracle_consensus_batch([g], flags=[flags[i]])
        assert alone["cons"][0] == mixed["cons"][i]
    assert mixed["stats"]["n_seed_groups"] == 4 == mixed["stats"]["n_seed_applied"]
