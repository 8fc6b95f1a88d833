"""`abpoa`-compatible command line backed by the GPU library, so that the UNMODIFIED reference
can be pointed at it:  defineIsoforms.py -a bin/abpoa-b200 ...

Protocol replaced (reference utils/SpliceDefineConsensus.py:917,919):
    abpoa -M 5 -r 0 [-S] in.fasta > out.fasta 2> abpoa.messages
stdout = ">Consensus_sequence\\n<bases>\\n"; empty stdout = soft failure (the reference then
falls back to the first read, :924-925).  One process per group wastes the GPU; this exists
for file-level parity runs, the batched path is mandalorion_b200.consensus.
"""
import sys


def read_fasta(path):
    seqs, cur = [], None
    with open(path) as fh:
        for line in fh:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if cur is not None:
                    seqs.append("".join(cur))
                cur = []
            elif cur is not None:
                cur.append(line)
    if cur is not None:
        seqs.append("".join(cur))
    return seqs


def parse_args(argv):
    """Accepts the flags the reference passes; anything else abpoa would accept is rejected loudly."""
    from .poa import PoaParams
    p = PoaParams(match=2)   # abpoa's own default; the reference always passes -M 5
    seed, result_mode, files = False, 0, []
    it = iter(argv)
    for a in it:
        if a == "-M":
            p.match = int(next(it))
        elif a == "-X":
            p.mismatch = int(next(it))
        elif a == "-r":
            result_mode = int(next(it))
        elif a == "-S":
            seed = True
        elif a.startswith("-"):
            raise SystemExit(f"abpoa-b200: unsupported option {a}")
        else:
            files.append(a)
    if result_mode != 0 or len(files) != 1:
        raise SystemExit("abpoa-b200: only `-r 0 <in.fasta>` is supported")
    return p, seed, files[0]


def main(argv=None):
    from .poa import PoaContext
    params, seed, path = parse_args(sys.argv[1:] if argv is None else argv)
    reads = read_fasta(path)
    if not reads:
        return 0
    # -S = minimizer-seeded, windowed alignment (reference :919, median read length >= 8000): MPOA_FLAG_SEED
    with PoaContext(0, params) as ctx:
        out = ctx.consensus_batch([reads], flags=[1 if seed else 0])
    if out["status"][0] == 0 and out["cons"][0]:
        sys.stdout.write(">Consensus_sequence\n%s\n" % out["cons"][0].decode())
    return 0


if __name__ == "__main__":
    sys.exit(main())
