#!/usr/bin/env python3
"""abpoa stand-in for CPU-only runs of the unmodified reference (TEST INFRASTRUCTURE): same argv and
stdout protocol as `abpoa -M 5 -r 0 [-S] in.fasta` (reference utils/SpliceDefineConsensus.py:917,919),
backed by the oracle.  The product equivalent is bin/abpoa-b200 (GPU)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mandalorion_b200.abpoa_cli import parse_args, read_fasta  # noqa: E402
from oracle import OracleParams, oracle_consensus_batch  # noqa: E402


def main():
    p, seed, path = parse_args(sys.argv[1:])
    reads = read_fasta(path)
    if not reads:
        return 0
    out = oracle_consensus_batch([reads], params=OracleParams(match=p.match, mismatch=p.mismatch), flags=[1 if seed else 0])
    if out["status"][0] == 0 and out["cons"][0]:
        sys.stdout.write(">Consensus_sequence\n%s\n" % out["cons"][0].decode())
    return 0


if __name__ == "__main__":
    sys.exit(main())
