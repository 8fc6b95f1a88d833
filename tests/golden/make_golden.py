"""Generates tests/golden/poa_golden.json.

PROVENANCE: these vectors are produced by the CPU oracle of THIS repo (oracle/abpoa_oracle.cpp),
not by abPOA or by the reference -- neither can run in the build image (no abpoa binary or
source, no mappy; the reference ships no fixtures).  They freeze the oracle's behaviour so that
(a) oracle refactors are caught on CPU and (b) the GPU path is checked against committed bytes.
Regenerate with:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from mandalorion_b200.synth import GroupConfig, make_groups  # noqa: E402
from oracle import OracleParams, oracle_consensus_batch  # noqa: E402

CASES = [
    # (name, config, params)
    ("clean_short", GroupConfig("golden_a", 6, 3, 8, 120, 260, "uniform", 0.01, (0.3, 0.35, 0.35)), {}),
    ("noisy_short", GroupConfig("golden_b", 6, 4, 12, 100, 240, "uniform", 0.08, (0.3, 0.35, 0.35)), {}),
    ("ccs_deep", GroupConfig("golden_c", 3, 20, 30, 150, 220, "pm5", 0.002, (0.2, 0.4, 0.4)), {}),
    ("sse_lanes", GroupConfig("golden_d", 4, 3, 8, 120, 260, "uniform", 0.05, (0.3, 0.35, 0.35)),
     dict(simd_pn_i16=8, simd_pn_i32=4)),
    ("avx512_lanes", GroupConfig("golden_e", 4, 3, 8, 120, 260, "uniform", 0.05, (0.3, 0.35, 0.35)),
     dict(simd_pn_i16=32, simd_pn_i32=16)),
    # every group flagged MPOA_FLAG_SEED (`abpoa -S`): reads long enough for one or two anchors each
    ("seeded_windows", GroupConfig("golden_f", 3, 3, 6, 1300, 1700, "uniform", 0.02, (0.3, 0.35, 0.35)), {}, 1),
]


def main():
    out = {"provenance": "oracle/abpoa_oracle.cpp (this repo); NOT an abPOA/reference vector", "cases": []}
    for name, cfg, pk, *rest in CASES:
        seed = rest[0] if rest else 0
        groups = [[r.decode() for r in g] for g in make_groups(cfg)]
        res = oracle_consensus_batch(groups, params=OracleParams(**pk), trace=True, flags=[seed] * len(groups))
        out["cases"].append(dict(
            name=name, params=pk, groups=groups, seed_flag=seed,
            consensus=[c.decode() for c in res["cons"]],
            status=[int(s) for s in res["status"]],
            read_score=[int(x) for x in res["trace"]["read_score"]],
            read_band_cells=[int(x) for x in res["trace"]["read_band_cells"]],
            base_node_sha1=hashlib.sha1(res["trace"]["base_node"].tobytes()).hexdigest(),
            base_aln_sha1=hashlib.sha1(res["trace"]["base_aln"].tobytes()).hexdigest(),
            band_cells=int(res["stats"]["band_cells"]), int_ops=int(res["stats"]["int_ops"])))
    with open(os.path.join(HERE, "poa_golden.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
