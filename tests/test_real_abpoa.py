"""The only TRUE parity check: diff the oracle against a real abPOA v1.4.1 binary on the reference's
exact command line (utils/SpliceDefineConsensus.py:917).  No abpoa exists in the build image or on
the GPU box (no source under /root/reference, no network), so this skips there; it fires the
moment a binary is on PATH, in $ABPOA, or under baseline/_ref/."""
import os
import shutil
import subprocess

import pytest

from helpers import oracle_consensus_batch
from mandalorion_b200.synth import GroupConfig, make_groups

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_abpoa():
    cands = [os.environ.get("ABPOA"), shutil.which("abpoa"),
             os.path.join(ROOT, "baseline", "_ref", "abPOA-v1.4.1", "bin", "abpoa"),
             os.path.join(ROOT, "baseline", "_ref", "bin", "abpoa")]
    for c in cands:
        if c and os.path.isfile(c) and os.access(c, os.X_OK):
            return c
    return None


@pytest.mark.skipif(find_abpoa() is None, reason="no abpoa binary reachable: parity vs real abPOA stays unpinned")
def test_oracle_matches_real_abpoa(tmp_path):
    abpoa = find_abpoa()
    groups = make_groups(GroupConfig("real", 40, 3, 20, 300, 2500, "loguniform", 0.02, (0.3, 0.35, 0.35)))
    want = oracle_consensus_batch(groups)["cons"]
    bad = 0
    for gi, reads in enumerate(groups):
        fa = tmp_path / f"g{gi}.fasta"
        fa.write_text("".join(f">r{i}\n{r.decode()}\n" for i, r in enumerate(reads)))
        out = subprocess.run([abpoa, "-M", "5", "-r", "0", str(fa)], capture_output=True, text=True).stdout
        seq = "".join(line for line in out.splitlines() if not line.startswith(">"))
        bad += seq != want[gi].decode()
    assert bad == 0, f"{bad}/{len(groups)} consensus sequences differ from real abpoa"
