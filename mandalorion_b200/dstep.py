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

iter_psl() reads the 24-column lines of tmp_SS/*.psl (emtrey.py:146-148; the columns module D uses
are listed in SURVEY.md Appendix B.2) for producers that do not go through the reference's parser.
"""
import queue
import threading

from .consensus import ConsensusBatcher, orient_pending

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
    """See the module docstring.  batch_bases: a batch is issued once this many read bases are pending."""

    def __init__(self, ctx=None, device=0, batch_bases=256 << 20, orient_threads=None):
        self._batcher = ConsensusBatcher(ctx, device=device)
        self._batch_bases = batch_bases
        self._orient_threads = orient_threads
        self._order = []                      # (root, {isoform: PendingGroup}) in insertion order
        self._cur, self._cur_bases = [], 0
        self._q = queue.Queue(maxsize=4)
        self._err = None
        self.n_batches = 0
        self._worker = threading.Thread(target=self._run, daemon=True)
        self._worker.start()

    def _run(self):
        while True:
            chunk = self._q.get()
            if chunk is None:
                return
            try:
                if self._err is None:
                    orient_pending(chunk, n_threads=self._orient_threads)
                    for pg in chunk:
                        self._batcher.add(pg)
                    self._batcher.flush()
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
        self._q.put(None)
        self._worker.join()
        if self._err is not None:
            raise self._err
        return {root: {isoform: [pg.consensus, pg.names] for isoform, pg in iso.items()} for root, iso in self._order}

    @property
    def stats(self):
        return self._batcher.stats
