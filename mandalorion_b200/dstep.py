"""Streaming dispatch of module D's consensus step: GPU batches run WHILE the loci are still being
parsed and grouped.

The reference forks one worker per locus (defineIsoforms.py:130-151); every worker parses its
tmp_SS/<root>.psl, groups the reads into isoforms and -- inside the same worker -- calls
determine_consensus() once per isoform (:87-91).  Nothing can go to a GPU before the slowest locus
is done if the groups are only collected at pool.join().  Here the producer side stays what it is
(the reference's own process_locus with determine_consensus swapped for consensus.prepare_group,
INTEGRATION.md option A, run under any pool or plain loop), but its results are consumed as they
arrive, in locus order:

    sc = StreamingConsensus(ctx)                 # one context / one GPU
    for root, iso in pool.imap(worker, roots):   # iso = {isoform: PendingGroup}
        sc.add_locus(root, iso)                  # returns at once; a batch goes to the GPU when
    results = sc.finish()                        #   enough bases have accumulated
    consensus.write_isoform_files(roots, results, out_path)

A background thread orients the pending groups of a batch (mpoa_orient_batch, C++ threads) and
runs the batch through the CUDA library; both calls release the GIL, so parsing, orientation and
the GPU overlap.  Results are identical to consensus.finish_prepared() -- batching never changes
a group's consensus -- and come back in insertion order.

define_isoforms() is the whole D step without the reference's Python on the path: locus files ->
locus.locus_groups (the group producer) -> prepare_group -> streamed GPU batches ->
Isoform_Consensi.fasta / reads2isoforms.txt, byte-identical to `python3 defineIsoforms.py ...`.

iter_psl() reads the 24-column lines of tmp_SS/*.psl (emtrey.py:146-148; the columns module D uses
are listed in SURVEY.md Appendix B.2) for producers that do not go through the reference's parser.
"""
import os
import queue
import threading

import numpy as np

from .consensus import ConsensusBatcher, orient_pending, prepare_group, write_isoform_files
from .locus import locus_groups

PSL_COLUMNS = dict(strand=8, name=9, length=10, qstart=11, qend=12, chrom=13, tstart=15, tend=16,
                   block_sizes=18, block_starts=20, accuracy=21, cs=22, sequence=23)


def iter_psl(path):
    """One dict per read of a tmp_SS/<chrom>~<start>~<end>.psl file, with the fields module D reads
    (reference utils/SpliceDefineConsensus.py:284-296, :716-733), typed."""
    with open(path) as fh:
        for line in fh:
            a = line.rstrip("\n").split("\t")
            if len(a) < 24:
                continue
            yield dict(strand=a[8], name=a[9], length=int(a[10]), qstart=int(a[11]), qend=int(a[12]), chrom=a[13],
                       tstart=int(a[15]), tend=int(a[16]),
                       block_sizes=[int(x) for x in a[18].split(",")[:-1]],
                       block_starts=[int(x) for x in a[20].split(",")[:-1]],
                       accuracy=float(a[21]), cs=a[22], sequence=a[23])


class StreamingConsensus:
    """See the module docstring.  batch_bases: a batch is issued once this many read bases are pending.
    depth: batches in flight -- every one has its own context and host thread, so that the orientation, the
    host->device copy and the result split of one batch run beside the kernels of another (the library lets
    the contexts of one GPU take turns on the SMs and share one workspace: poa.PoaPipeline).
    With an explicit `ctx` the first batch slot uses it."""

    def __init__(self, ctx=None, device=0, batch_bases=256 << 20, orient_threads=None, depth=1):
        self._batchers = [ConsensusBatcher(ctx if k == 0 else None, device=device) for k in range(max(1, depth))]
        self._batch_bases = batch_bases
        self._orient_threads = orient_threads
        self._order = []                      # (root, {isoform: PendingGroup}) in insertion order
        self._cur, self._cur_bases = [], 0
        self._q = queue.Queue(maxsize=4)
        self._err = None
        self.n_batches = 0
        self._lock = threading.Lock()
        self._workers = [threading.Thread(target=self._run, args=(b,), daemon=True) for b in self._batchers]
        for w in self._workers:
            w.start()

    def _run(self, batcher):
        while True:
            chunk = self._q.get()
            if chunk is None:
                return
            try:
                if self._err is None:
                    orient_pending(chunk, n_threads=self._orient_threads)
                    for pg in chunk:
                        batcher.add(pg)
                    batcher.flush()
                    with self._lock:
                        self.n_batches += 1
            except Exception as e:            # surfaced by add_locus() / finish()
                self._err = e

    def _issue(self):
        if self._cur:
            self._q.put(self._cur)
            self._cur, self._cur_bases = [], 0

    def add_locus(self, root, iso):
        """iso: {isoform: PendingGroup} as prepare_group() returned them for this locus."""
        if self._err is not None:
            raise self._err
        self._order.append((root, iso))
        for pg in iso.values():
            self._cur.append(pg)
            reads = pg.sequences if pg.sequences is not None else [s for _, s in pg.subsample]
            self._cur_bases += sum(map(len, reads))
        if self._cur_bases >= self._batch_bases:
            self._issue()

    def finish(self):
        """Runs what is still pending and returns {root: {isoform: [consensus, names]}} in insertion order."""
        self._issue()
        for _ in self._workers:
            self._q.put(None)
        for w in self._workers:
            w.join()
        if self._err is not None:
            raise self._err
        return {root: {isoform: [pg.consensus, pg.names] for isoform, pg in iso.items()} for root, iso in self._order}

    @property
    def stats(self):
        return [st for b in self._batchers for st in b.stats]


def locus_roots(tmp_ss):
    """Locus files of a tmp_SS directory in the reference's processing order (get_parsed_files,
    utils/SpliceDefineConsensus.py:96-105; sorted by chromosome name, then start: defineIsoforms.py:126)."""
    roots = {f.split(".psl")[0] for f in os.listdir(tmp_ss) if ".psl" in f}
    return sorted(roots, key=lambda x: (x.split("~")[0], int(x.split("~")[1])))


def _produce(job):
    """One locus in a producer process: returns its groups and the RNG state the consensus step continues from."""
    path, chrom, lb, rb, par, state = job
    np.random.set_state(state)
    groups = locus_groups(path, chrom, lb, rb, *par)
    return groups, np.random.get_state()


def define_isoforms(out_path, ctx=None, device=0, left_bounds=None, right_bounds=None, splice_site_width=1,
                    minimum_read_count=2, junctions=("gtag", "gcag", "atac", "ctac", "ctgc", "gtat"), cutoff=0.1,
                    upstream_buffer=10, downstream_buffer=50, batch_bases=256 << 20, orient_threads=None, workers=0,
                    depth=2):
    """Module D (`defineIsoforms.py -p out_path ...`, reference defineIsoforms.py:93-168) on one GPU:
    reads out_path/tmp_SS/*.psl, writes out_path/Isoform_Consensi.fasta and reads2isoforms.txt.

    left_bounds / right_bounds: annotated splice sites {chrom: {'5': [...], '3': [...]}} as the reference's
    parse_genome() returns them (None: read-derived sites only, `-g None`).  The other arguments are the
    command-line flags Mando.py passes (:382-398), same defaults.

    Random numbers: the reference forks one worker per locus, so every locus starts from the PARENT's
    NumPy RNG state; the same is done here (state saved at entry, restored before every locus), which makes
    the output identical to a reference run started from the same state.  workers > 0: the loci are parsed
    and grouped by that many producer processes (spawned, so that no CUDA state is forked), consumed in locus
    order while the GPU works on earlier loci; the result does not depend on it.  depth: GPU batches in flight
    (StreamingConsensus).  Returns the number of isoforms."""
    import time
    tmp_ss = os.path.join(out_path, "tmp_SS")
    roots = locus_roots(tmp_ss)
    junctions = list(junctions)
    entry_state = np.random.get_state()
    sc = StreamingConsensus(ctx, device=device, batch_bases=batch_bases, orient_threads=orient_threads, depth=depth)
    t = dict(producer=0.0, prepare=0.0, drain=0.0, write=0.0)
    n_reads = 0
    par = (splice_site_width, minimum_read_count, junctions, cutoff, upstream_buffer, downstream_buffer)
    jobs = []
    for root in roots:
        chrom, start, end = root.split("~")
        start, end = int(start), int(end)
        inside = {}
        for side, table in (("l", left_bounds), ("r", right_bounds)):
            per = (table or {}).get(chrom, {"5": [], "3": []})
            inside[side] = {k: [p for p in per[k] if start < p < end] for k in ("5", "3")}
        jobs.append((os.path.join(tmp_ss, root + ".psl"), chrom, inside["l"], inside["r"], par, entry_state))
    pool = None
    if workers > 0 and len(jobs) > 1:
        import multiprocessing as mp
        pool = mp.get_context("spawn").Pool(min(workers, len(jobs)))
        produced = pool.imap(_produce, jobs, chunksize=1)
    else:
        produced = map(_produce, jobs)
    try:
        for root in roots:
            t0 = time.perf_counter()
            groups, state = next(produced)        # in-process: the work happens here; with a pool: the wait for it
            t1 = time.perf_counter()
            np.random.set_state(state)
            sc.add_locus(root, {isoform: prepare_group(reads) for isoform, reads in groups.items()})
            t["producer"] += t1 - t0
            t["prepare"] += time.perf_counter() - t1
            n_reads += sum(map(len, groups.values()))
    finally:
        if pool is not None:
            pool.terminate()
            pool.join()
    t0 = time.perf_counter()
    results = sc.finish()                 # what is still running on the GPU after the last locus was parsed
    t1 = time.perf_counter()
    n = write_isoform_files(roots, results, out_path)
    t["drain"], t["write"] = t1 - t0, time.perf_counter() - t1
    define_isoforms.last_stats = sc.stats
    define_isoforms.last_timings = dict(t, loci=len(roots), reads_in_groups=n_reads, isoforms=n, batches=sc.n_batches,
                                        producer_workers=workers)
    return n
