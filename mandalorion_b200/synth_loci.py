"""Synthetic module-D input: tmp_SS/<chrom>~<start>~<end>.psl files in the 24-column format the
reference's emtrey.py -m writes (emtrey.py:146-148; columns read by module D: SURVEY.md Appendix B.2).
Single-block reads (no introns): every read group of a locus is a mono-exon isoform defined by its
start/end peaks (utils/SpliceDefineConsensus.py:772-868)."""
import os

import numpy as np

from .synth import _ACGT, mutate, revcomp


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


# ---- spliced loci (group producer tests): genes with exons / introns, reads with cs strings ----

def _cs_and_blocks(genome, exons, rng, err):
    """Walks a read over its exons with R2C2-like errors.  Returns (read sequence, minimap2 long-form cs
    string, PSL blocks [(tstart, size)] split at every indel as emtrey does (emtrey.py:63-83))."""
    seq, cs, blocks = [], [], []
    for k, (a, b) in enumerate(exons):
        if k:
            pa, pb = exons[k - 1][1], a
            intron = genome[pa:pb].tobytes().decode().lower()
            cs.append("~%s%d%s" % (intron[:2], pb - pa, intron[-2:]))
        p = a
        run_start, run = p, []
        while p < b:
            u = rng.random()
            if u < err * 0.3 and len(run) > 3:                         # substitution: stays inside the M block
                alt = _ACGT[(int(np.searchsorted(_ACGT, genome[p])) + int(rng.integers(1, 4))) % 4]
                if run:
                    cs.append("=" + "".join(run))
                    run = []
                cs.append("*%s%s" % (chr(genome[p]).lower(), chr(alt).lower()))
                seq.append(chr(alt))
                p += 1
            elif u < err * 0.65 and len(run) > 3 and p - run_start > 0:   # insertion: closes the block
                ins = "".join(chr(c) for c in _ACGT[rng.integers(0, 4, int(rng.integers(1, 3)))])
                if run:
                    cs.append("=" + "".join(run))
                    run = []
                cs.append("+" + ins.lower())
                seq.append(ins)
                blocks.append((run_start, p - run_start))
                run_start = p
            elif u < err and len(run) > 3 and p + 3 < b:                # deletion: closes the block, leaves a small gap
                d = int(rng.integers(1, 3))
                if run:
                    cs.append("=" + "".join(run))
                    run = []
                cs.append("-" + genome[p:p + d].tobytes().decode().lower())
                blocks.append((run_start, p - run_start))
                p += d
                run_start = p
            else:
                run.append(chr(genome[p]))
                seq.append(chr(genome[p]))
                p += 1
        if run:
            cs.append("=" + "".join(run))
        if b > run_start:
            blocks.append((run_start, b - run_start))
    return "".join(seq), "".join(cs), [x for x in blocks if x[1] > 0]


def write_spliced_locus(path, chrom, locus_start, rng, n_reads=120, n_exons=5, strand="+", err=0.02,
                        noncanonical=None, extra_lines=()):
    """One gene with `n_exons` exons, two or three isoforms (exon skipping, alternative first exon /
    last exon ends), reads with jittered ends, a few low-accuracy reads, a few reads whose junction is
    off by one.  Returns the root name."""
    ex_len = rng.integers(90, 260, n_exons)
    in_len = rng.integers(120, 700, n_exons - 1)
    size = int(ex_len.sum() + in_len.sum()) + 400
    genome = _ACGT[rng.integers(0, 4, size)].copy()
    exons, p = [], 200
    for k in range(n_exons):
        exons.append((p, p + int(ex_len[k])))
        p += int(ex_len[k])
        if k < n_exons - 1:
            motif = (b"GT", b"AG") if strand == "+" else (b"CT", b"AC")
            if noncanonical is not None and k == noncanonical:
                motif = (b"AA", b"TT")
            genome[p:p + 2] = np.frombuffer(motif[0], np.uint8)
            genome[p + int(in_len[k]) - 2:p + int(in_len[k])] = np.frombuffer(motif[1], np.uint8)
            p += int(in_len[k])
    isoforms = [list(range(n_exons)), [k for k in range(n_exons) if k != 2], list(range(1, n_exons))]
    lines, end_max, start_min = [], 0, 1 << 60
    for r in range(n_reads):
        iso = isoforms[int(rng.choice(len(isoforms), p=[0.5, 0.3, 0.2]))]
        ex = [list(exons[k]) for k in iso]
        ex[0][0] += int(rng.integers(0, 8)) + (60 if rng.random() < 0.25 else 0)     # alternative start site
        ex[-1][1] -= int(rng.integers(0, 8)) + (50 if rng.random() < 0.25 else 0)    # alternative end site
        if rng.random() < 0.05 and len(ex) > 1:
            ex[0][1] += 1                                                               # junction off by one
        seq, cs, blocks = _cs_and_blocks(genome, [tuple(e) for e in ex], rng, err)
        if not blocks:                   # a short single-exon isoform trimmed to nothing
            continue
        tstart, tend = locus_start + blocks[0][0], locus_start + blocks[-1][0] + blocks[-1][1]
        start_min, end_max = min(start_min, tstart), max(end_max, tend)
        left_clip, right_clip = int(rng.integers(0, 12)), int(rng.integers(0, 12))
        cols = ["0"] * 24
        cols[8] = strand
        cols[9] = f"{chrom}_{locus_start}_r{r}"
        cols[10] = str(len(seq) + left_clip + right_clip)
        cols[11] = str(left_clip)
        cols[12] = str(len(seq) + left_clip)
        cols[13] = chrom
        cols[15], cols[16] = str(tstart), str(tend)
        cols[17] = str(len(blocks))
        cols[18] = "".join("%d," % s for _, s in blocks)
        cols[19] = "0," * len(blocks)
        cols[20] = "".join("%d," % (locus_start + b) for b, _ in blocks)
        cols[21] = "0.85" if rng.random() < 0.04 else "0.97"
        cols[22] = cs
        cols[23] = "A" * left_clip + seq + "C" * right_clip
        lines.append("\t".join(cols))
    lines.extend(extra_lines)
    root = f"{chrom}~{start_min}~{end_max}"
    with open(os.path.join(path, root + ".psl"), "w") as fh:
        fh.write("\n".join(lines) + "\n")
    return root


def make_spliced_input(tmp_ss):
    """Spliced genes on both strands, one with a non-canonical intron, plus mono-exonic loci."""
    rng = np.random.Generator(np.random.PCG64(23))
    os.makedirs(tmp_ss, exist_ok=True)
    return [write_spliced_locus(tmp_ss, "chr1", 20000, rng, n_reads=60, n_exons=4),
            write_spliced_locus(tmp_ss, "chr1", 90000, rng, n_reads=80, n_exons=5, strand="-"),
            write_spliced_locus(tmp_ss, "chr2", 4000, rng, n_reads=50, n_exons=4, noncanonical=0),
            write_locus(tmp_ss, "chr2", 60000, [(0, 380, 7), (1500, 300, 4)], rng)]
