"""Synthetic module-D input: tmp_SS/<chrom>~<start>~<end>.psl files in the 24-column format the
reference's emtrey.py -m writes (emtrey.py:146-148; columns read by module D: SURVEY.md Appendix B.2).
Single-block reads (no introns): every read group of a locus is a mono-exon isoform defined by its
start/end peaks (utils/SpliceDefineConsensus.py:772-868)."""
import os

import numpy as np

from mandalorion_b200.synth import _ACGT, mutate, revcomp


def write_locus(path, chrom, locus_start, isoforms, rng, err=0.02):
    """isoforms: list of (offset, length, n_reads).  Returns the root name."""
    lines = []
    end_max = 0
    for k, (off, length, n_reads) in enumerate(isoforms):
        template = _ACGT[rng.integers(0, 4, length)]
        for r in range(n_reads):
            seq = mutate(template, err, (0.3, 0.35, 0.35), rng).tobytes()
            if rng.random() < 0.5:
                seq = revcomp(seq)                      # module D re-orients every read against read 0
            seq = seq.decode()
            tstart = locus_start + off + int(rng.integers(0, 3))
            tend = locus_start + off + length - int(rng.integers(0, 3))
            end_max = max(end_max, tend)
            name = f"{chrom}_{locus_start}_iso{k}_read{r}"
            cols = ["0"] * 24
            cols[8] = "+"
            cols[9] = name
            cols[10] = str(len(seq))
            cols[11] = "0"
            cols[12] = str(len(seq))
            cols[13] = chrom
            cols[15] = str(tstart)
            cols[16] = str(tend)
            cols[17] = "1"
            cols[18] = f"{tend - tstart},"
            cols[19] = "0,"
            cols[20] = f"{tstart},"
            cols[21] = "0.98"
            cols[22] = "=" + "A" * 10
            cols[23] = seq
            lines.append("\t".join(cols))
    root = f"{chrom}~{locus_start}~{end_max}"
    with open(os.path.join(path, root + ".psl"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return root


def make_dstep_input(tmp_ss, seed=7):
    """A few loci with 1-3 isoforms each, including groups of 1 and 2 reads (the abpoa bypass)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    os.makedirs(tmp_ss, exist_ok=True)
    roots = [
        write_locus(tmp_ss, "chr1", 1000, [(0, 420, 6), (2000, 300, 3)], rng),
        write_locus(tmp_ss, "chr1", 9000, [(0, 260, 9)], rng),
        write_locus(tmp_ss, "chr2", 500, [(0, 350, 4), (1500, 280, 2), (3000, 500, 12)], rng),
        write_locus(tmp_ss, "chr10", 70000, [(0, 300, 5)], rng),
    ]
    return roots
