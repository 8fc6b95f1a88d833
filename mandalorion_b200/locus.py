"""Group producer of module D: one tmp_SS/<chrom>~<start>~<end>.psl locus file -> the isoform read
groups whose consensus the GPU computes.

This restates what the reference's process_locus() does BEFORE its consensus loop
(defineIsoforms.py:55-86): splice-site calling from the reads (utils/SpliceDefineConsensus.py
collect_reads :278-331, make_genome_bins :392-438, find_peaks :232-275 with scan_for_best_bin
:163-197, determine_cov :200-224, characterize_splicing_event :499-550 and getCSaroundSS :107-161),
read -> splice-junction chain (sort_reads_into_splice_junctions :714-769) and start/end site
calling (define_start_end_sites :797-868 with group_mono_exon_transcripts :772-794 and find_ends
:554-711).  The result is the reference's `seqDict`: {isoform number (str): [(name, sequence)]} in the
same order with the same members, and the module-global NumPy RNG is consumed by the same calls in
the same order (np.random.choice at :503 and :818), so the subsampling that follows in
determine_consensus (:884) sees the same stream.

It is not a transcription.  The reference walks nested dicts per base and per read and parses every
cs string again for every candidate site; here the locus file is parsed once into flat NumPy arrays
(coverage positions of all reads in one CSR array, bounds in emission order), candidate windows are
counted with prefix sums, coverage around a site is one ragged gather + np.unique, and a read's cs
string is decoded once into a status/position array that every site query bisects.  On a deep locus
this is one to two orders of magnitude faster (bench.py --dstep reports both), which matters once the
consensus itself takes milliseconds: the producer is what keeps the GPU fed.

Quirks of the reference that change results are kept on purpose and marked QUIRK.
"""
import re
import numpy as np

__all__ = ["LocusFile", "read_locus", "splice_sites", "junction_chains", "start_end_groups", "locus_groups"]


class LocusFile:
    """The 24-column lines of one locus file (emtrey.py:146-148), column-wise."""

    __slots__ = ("n", "name", "strand", "qsize", "qstart", "qend", "chrom", "tstart", "tend",
                 "bsize", "bstart", "accuracy", "cs", "seq")


def read_locus(path):
    lf = LocusFile()
    cols = {k: [] for k in LocusFile.__slots__ if k != "n"}
    with open(path) as fh:
        for line in fh:
            a = line.strip().split("\t")
            cols["strand"].append(a[8]); cols["name"].append(a[9]); cols["qsize"].append(int(a[10]))
            cols["qstart"].append(int(a[11])); cols["qend"].append(int(a[12])); cols["chrom"].append(a[13])
            cols["tstart"].append(int(a[15])); cols["tend"].append(int(a[16]))
            cols["bsize"].append(np.array(a[18].split(",")[:-1], dtype=np.int64))
            cols["bstart"].append(np.array(a[20].split(",")[:-1], dtype=np.int64))
            cols["accuracy"].append(float(a[21])); cols["cs"].append(a[22]); cols["seq"].append(a[23])
    for k, v in cols.items():
        setattr(lf, k, v)
    lf.n = len(lf.name)
    return lf


# --------------------------------------------------------------------------------------------
# splice sites
# --------------------------------------------------------------------------------------------

def _round10(p):
    """myround() of the reference (:227-229) on an int array: 10 * round(p / 10), ties to even."""
    return (np.rint(p / 10.0) * 10).astype(np.int64)


class _Coverage:
    """Rounded coverage positions of every read of the target chromosome (CSR) + their histogram."""

    def __init__(self, lf, rows):
        nb = np.array([len(lf.bsize[r]) for r in rows], dtype=np.int64)
        if nb.sum() == 0:
            self.off = np.zeros(len(rows) + 1, np.int64)
            self.pos = np.zeros(0, np.int64)
            self.hist_pos = np.zeros(0, np.int64)
            self.hist_cnt = np.zeros(0, np.int64)
            return
        bs = np.concatenate([lf.bstart[r] for r in rows])
        sz = np.concatenate([lf.bsize[r] for r in rows])
        owner = np.repeat(np.arange(len(rows)), nb)
        # a block contributes its offsets 0, 10, 20, ... and then every offset from the last of those
        # to the block end (:304-309): that is what the reference's two loops visit
        n_step = (sz + 9) // 10
        last = (n_step - 1) * 10
        n_tail = sz - last
        per = n_step + n_tail
        keep = sz > 0
        per = np.where(keep, per, 0)
        tot = int(per.sum())
        blk = np.repeat(np.arange(len(sz)), per)
        k = np.arange(tot) - np.repeat(np.cumsum(per) - per, per)
        offs = np.where(k < n_step[blk], k * 10, last[blk] + (k - n_step[blk]))
        key = owner[blk] * (1 << 40) + (_round10(bs[blk] + offs) + (1 << 36))
        key = np.unique(key)
        own = key >> 40
        self.pos = (key & ((1 << 40) - 1)) - (1 << 36)
        self.off = np.zeros(len(rows) + 1, np.int64)
        np.cumsum(np.bincount(own, minlength=len(rows)), out=self.off[1:])
        self.hist_pos, self.hist_cnt = np.unique(self.pos, return_counts=True)

    def depth(self, p):
        i = np.searchsorted(self.hist_pos, p)
        return int(self.hist_cnt[i]) if i < len(self.hist_pos) and self.hist_pos[i] == p else None

    def gather(self, reads):
        """Concatenated coverage positions of reads[] (with repeats)."""
        reads = np.asarray(reads, dtype=np.int64)
        lens = self.off[reads + 1] - self.off[reads]
        tot = int(lens.sum())
        if tot == 0:
            return np.zeros(0, np.int64)
        src = np.repeat(self.off[reads] - (np.cumsum(lens) - lens), lens) + np.arange(tot)
        return self.pos[src]


class _Bounds:
    """Exon ends (side 'l') or exon starts (side 'r') of the accurate reads, in emission order."""

    def __init__(self, pos, read):
        self.pos = np.asarray(pos, dtype=np.int64)
        self.read = np.asarray(read, dtype=np.int64)
        order = np.argsort(self.pos, kind="stable")            # groups by position, emission order inside
        self.by_pos = order
        self.sorted_pos = self.pos[order]
        self.uniq, first, self.count = np.unique(self.pos, return_index=True, return_counts=True) \
            if len(self.pos) else (np.zeros(0, np.int64),) * 3
        self.first = first
        self.prefix = np.concatenate([[0], np.cumsum(self.count)])

    def n_between(self, lo, hi):
        return int(self.prefix[np.searchsorted(self.uniq, hi, "right")] - self.prefix[np.searchsorted(self.uniq, lo, "left")])

    def at(self, p):
        a, b = np.searchsorted(self.sorted_pos, p, "left"), np.searchsorted(self.sorted_pos, p, "right")
        return self.by_pos[a:b]

    def candidates(self, min_count):
        """Positions with >= min_count reads, most reads first, ties by first appearance (:241-244)."""
        sel = np.nonzero(self.count >= min_count)[0]
        sel = sel[np.lexsort((self.first[sel], -self.count[sel]))]
        return self.uniq[sel].tolist()


_CS_OPS = re.compile(r"([\\=+\-*~])")


class _CsTrack:
    """A cs string decoded once: per record entry its kind and the genome position after it."""

    __slots__ = ("kind", "adv_idx", "adv_pos", "introns")

    def __init__(self, cs, begin):
        parts = _CS_OPS.split(cs)
        kinds, introns, jumps = [], {}, []
        n = 0
        for op, body in zip(parts[1::2], parts[2::2]):
            if op == "=" or op == "-" or op == "+":
                m = len(body)
            elif op == "*":
                m = (len(body) + 1) // 2
            elif op == "~":
                jumps.append((n, int(body[2:-2])))
                introns[n] = body
                kinds.append(b"|")
                n += 1
                continue
            else:
                continue
            kinds.append(op.encode() * m)
            n += m
        self.kind = b"".join(kinds)
        self.introns = introns
        # genome position after every entry: +1 for a match / substitution / deleted base, the intron length
        # for an intron, nothing for an inserted base (which can never be "the entry at the site")
        code = np.frombuffer(self.kind, dtype=np.uint8)
        step = (code != 43).astype(np.int64)              # 43 = '+'
        for i, length in jumps:
            step[i] = length
        self.adv_idx = np.nonzero(code != 43)[0]
        self.adv_pos = (begin + np.cumsum(step))[self.adv_idx]

    def around(self, lo, hi):
        """(intron bases, kinds of the 5 entries before, of the 5 after) for the last entry whose genome
        position lies in [lo, hi] -- getCSaroundSS() of the reference (:107-161)."""
        j = int(np.searchsorted(self.adv_pos, hi, "right")) - 1
        if j < 0 or self.adv_pos[j] < lo:
            return "nnnn", b"", b""
        at = int(self.adv_idx[j]) + 1
        bases, left, right = "nnnn", b"", b""
        for i in range(max(at - 10, 0), min(at + 10, len(self.kind))):
            body = self.introns.get(i)
            if body is not None:
                item = "|" + body + "|"
                bases = item[1:3] + item[-3:-1]
                left = self.kind[i - 5:i]          # QUIRK: a negative start wraps, exactly like the list slice (:158)
                right = self.kind[i + 1:i + 6]
        return bases, left, right


def _junction_ok(names, cs_of, lo, hi, junctions):
    """characterize_splicing_event() (:499-550): up to 500 randomly chosen supporting reads must show an
    allowed intron motif in > 85 % of cases and > 85 % matches in the 5 positions on either side."""
    picks = np.random.choice(np.arange(0, len(names)), min(len(names), 500), replace=False)
    allowed = 0
    lt = rt = lm = rm = 0
    for i in picks:
        bases, left, right = cs_of(names[i]).around(lo, hi)
        allowed += bases in junctions
        lt += len(left); lm += left.count(b"=")
        rt += len(right); rm += right.count(b"=")
    if allowed / len(picks) > 0.85:
        left_acc, right_acc = lm / lt, rm / rt        # QUIRK: ZeroDivisionError when no read shows context, as upstream
        return left_acc > 0.85 and right_acc > 0.85
    return False


def _annotated_bins(bounds, side, width, area):
    """make_genome_bins() (:392-438).  QUIRK: the reference seeds every cluster with its first position
    twice, so its "several distinct sites" branch can never run: a cluster is always one bin."""
    out = []
    for kind in ("5", "3"):
        pos = sorted(bounds[kind], key=int)
        i = 0
        while i < len(pos):
            top = pos[i]
            j = i
            while j < len(pos) and pos[j] - top <= width:
                top = max(top, pos[j])
                j += 1
            lo, hi = pos[i] - width, top + width
            out.append((lo, hi, kind, side, "A"))
            area.update(range(lo, hi + 1))
            i = j
    return out


def _read_peaks(bounds, cover, lf_names, strands, reverse, cutoff, side, area, cs_of, width, min_reads, junctions):
    """find_peaks() (:232-275) for one side."""
    shifts = [0]
    for s in range(1, width + 1):
        shifts += [s, -s]
    out = []
    for entry in bounds.candidates(min_reads):
        if entry in area:
            continue
        best, center = 0, 0
        for x in shifts:
            c = entry + x
            if any((c + y) in area for y in shifts):
                continue
            n = bounds.n_between(c - width, c + width)
            if n > best:
                best, center = n, c
        if best == 0:
            continue
        items = np.concatenate([bounds.at(center + y) for y in shifts])
        reads = bounds.read[items]
        # positions covered by at least two of the supporting reads, nearest four beyond the site
        got = cover.gather(reads)
        base = int(got.min()) if len(got) else 0
        pos = base + 10 * np.nonzero(np.bincount((got - base) // 10) > 1)[0] if len(got) else got
        near = pos[pos < center][::-1][:4] if reverse else pos[pos > center][:4]
        depth = max([0] + [d for d in (cover.depth(int(p)) for p in near) if d is not None])
        if depth <= 0:
            continue
        share = round(best / depth, 3)
        if not share > cutoff:
            continue
        plus = int(np.count_nonzero(strands[reads] == 1))
        minus = int(np.count_nonzero(strands[reads] == -1))
        if plus == minus:
            continue
        kind = ("3" if reverse else "5") if plus < minus else ("5" if reverse else "3")
        if _junction_ok([lf_names[r] for r in reads], cs_of, center - width, center + width, junctions):
            out.append((center - width, center + width, kind, side, str(share)))
            area.update(range(center - width, center + width + 1))
    return out


def splice_sites(lf, chrom, left_bounds, right_bounds, splice_site_width, minimum_read_count, junctions, cutoff):
    """{genome position: site label} of the locus: annotated bins first, then read-derived peaks --
    what process_locus() builds as spliceDict[chrom] (defineIsoforms.py:59-83)."""
    rows = [r for r in range(lf.n) if lf.chrom[r] == chrom]
    cover = _Coverage(lf, rows)
    last_row = {lf.name[r]: k for k, r in enumerate(rows)}        # csDict keeps the LAST line of a name (:296)
    tracks = {}

    def cs_of(name):
        k = last_row[name]
        t = tracks.get(k)
        if t is None:
            t = tracks[k] = _CsTrack(lf.cs[rows[k]], lf.tstart[rows[k]])
        return t

    lows_p, lows_r, ups_p, ups_r = [], [], [], []
    for k, r in enumerate(rows):
        if lf.accuracy[r] < 0.9:
            continue
        bs, sz = lf.bstart[r], lf.bsize[r]
        up = bs[bs != lf.tstart[r]]
        low = (bs + sz)[(bs + sz) != lf.tend[r]]
        ups_p.append(up); ups_r.append(np.full(len(up), k))
        lows_p.append(low); lows_r.append(np.full(len(low), k))
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.int64)  # noqa: E731
    left = _Bounds(cat(lows_p), cat(lows_r))
    right = _Bounds(cat(ups_p), cat(ups_r))
    names = [lf.name[r] for r in rows]
    strands = np.array([1 if lf.strand[r] == "+" else -1 if lf.strand[r] == "-" else 0 for r in rows], dtype=np.int8)
    area = {"l": set(), "r": set()}
    found = [_annotated_bins(left_bounds, "l", splice_site_width, area["l"]),
             _annotated_bins(right_bounds, "r", splice_site_width, area["r"]),
             _read_peaks(left, cover, names, strands, True, cutoff, "l", area["l"], cs_of, splice_site_width,
                         minimum_read_count, junctions),
             _read_peaks(right, cover, names, strands, False, cutoff, "r", area["r"], cs_of, splice_site_width,
                         minimum_read_count, junctions)]
    labels, serial = {}, {"l": 0, "r": 0}
    for batch in found:
        for lo, hi, kind, side, _ in batch:
            serial[side] += 1
            tag = kind + side + str(serial[side])
            for p in range(lo, hi + 1):
                labels[p] = tag
    return labels


# --------------------------------------------------------------------------------------------
# reads -> junction chains -> start / end sites -> isoform groups
# --------------------------------------------------------------------------------------------

def junction_chains(lf, chrom, labels):
    """sort_reads_into_splice_junctions() (:714-769): every read of the file gets the chain of site
    labels of its introns (> 50 nt); a read with an intron at an uncalled site is dropped.  Returns
    (spliced, mono): {identity: [(start, end, (name, seq), left_extra, right_extra, '+')]}."""
    spliced, mono = {}, {}
    for r in range(lf.n):
        bs, sz = lf.bstart[r], lf.bsize[r]
        ident = lf.chrom[r] + "_"
        ok = True
        if len(bs) > 1:
            donors = (bs + sz)[:-1]
            acceptors = bs[1:]
            for k in np.nonzero(acceptors - donors > 50)[0]:
                a = labels.get(int(donors[k])) if lf.chrom[r] == chrom else None
                b = labels.get(int(acceptors[k])) if lf.chrom[r] == chrom else None
                if not a or not b:
                    ok = False
                    break
                ident += a + "-" + b + "~"
        if not ok:
            continue
        rec = (lf.tstart[r], lf.tend[r], (lf.name[r], lf.seq[r]), lf.qstart[r], lf.qsize[r] - lf.qend[r], "+")
        # QUIRK (:752): "mono-exonic" is decided on the text after the FIRST underscore of the identity, so a
        # chromosome name with an underscore sends its mono-exonic reads to the spliced table
        (spliced if ident.split("_")[1] != "" else mono).setdefault(ident, []).append(rec)
    return spliced, mono


def _merge_mono(spliced, mono):
    """group_mono_exon_transcripts() (:772-794): overlapping mono-exonic reads form one group.
    QUIRK: the running end is max(end) only when a new group starts, else the end of the last read."""
    for ident, recs in mono.items():
        reach, serial = 0, 0
        for rec in sorted(recs):
            if rec[0] > reach:
                serial += 1
                reach = max(rec[1], reach)
            else:
                reach = rec[1]
            spliced.setdefault(ident + "M" + str(serial), []).append(rec)
    return spliced


def _site_map(points, lo_pad, hi_pad, forward, min_count):
    """find_ends() (:554-711) for one kind of site.  points: sampled positions.  A position p becomes a
    site when the 10-nt window starting at p (forward) / ending at p (backward) holds >= min_count
    points; it then owns [p - lo_pad, p + hi_pad) and grows outwards in 10-nt steps while the step still
    holds >= min_count points but fewer than the site's best window.  Returns {position: site}."""
    count = {}
    for p in points:
        count[p] = count.get(p, 0) + 1

    def total(a, b):                       # points in [a, b]
        return sum(count.get(q, 0) for q in range(a, b + 1))

    owner = {}
    for p in sorted(points, reverse=not forward):
        lo, hi = p - lo_pad, p + hi_pad - 1
        if (lo if forward else hi) in owner:
            continue
        if (total(p, p + 9) if forward else total(p - 9, p)) < min_count:
            continue
        for q in range(lo, hi + 1):
            owner[q] = p
        # sliding 10-nt windows over the core
        run = total(lo, lo + 9)
        best = run
        for i in range(lo + 1, hi):
            run += count.get(i + 9, 0) - count.get(i - 1, 0)
            if run > best:
                best = run
        for step in (-1, 1):
            edge = lo if step < 0 else hi
            while True:
                cells = [edge + step * i for i in range(1, 11)]
                n = sum(count.get(q, 0) for q in cells)
                edge = cells[-1]
                if not (best > n >= min_count):
                    break
                grew = True
                for q in cells:
                    if q in owner:
                        grew = False            # QUIRK: the rest of the step is still claimed (:601-605)
                    else:
                        owner[q] = p
                if not grew:
                    break
    return owner


def start_end_groups(spliced, mono, upstream_buffer, downstream_buffer, minimum_feature_count):
    """define_start_end_sites() (:797-868): within a junction chain, reads are split by the start and end
    site they fall into; numbering follows first appearance.  Returns {str(number): [(name, seq)]}."""
    chains = _merge_mono(spliced, mono)
    number, groups = {}, {}
    for ident in sorted(chains):
        recs = chains[ident]
        n = len(recs)
        picks = np.random.choice(range(0, n, 1), size=min(10000, n), replace=False)
        starts = [int(recs[i][0]) for i in picks]
        ends = [int(recs[i][1]) for i in picks]
        s_map = _site_map(starts, upstream_buffer, downstream_buffer, True, minimum_feature_count)
        e_map = _site_map(ends, downstream_buffer, upstream_buffer, False, minimum_feature_count)
        for rec in recs:
            s, e = s_map.get(int(rec[0])), e_map.get(int(rec[1]))
            if s is None or e is None:
                continue
            # the reference appends the median overhangs of (chain, start site, end site) to the key; they
            # are a function of the triple, so the triple is the key
            key = (ident, s, e)
            k = number.get(key)
            if k is None:
                k = number[key] = str(len(number) + 1)
                groups[k] = []
            groups[k].append(rec[2])
    return groups


def locus_groups(psl_path, chrom, left_bounds, right_bounds, splice_site_width, minimum_read_count, junctions, cutoff,
                 upstream_buffer, downstream_buffer):
    """The producer half of process_locus() (defineIsoforms.py:55-86): `seqDict` of one locus file.
    left_bounds / right_bounds: {'5': [...], '3': [...]} annotated sites inside the locus (may be empty)."""
    lf = read_locus(psl_path)
    labels = splice_sites(lf, chrom, left_bounds, right_bounds, splice_site_width, minimum_read_count, junctions, cutoff)
    spliced, mono = junction_chains(lf, chrom, labels)
    return start_end_groups(spliced, mono, upstream_buffer, downstream_buffer, minimum_feature_count=minimum_read_count)
