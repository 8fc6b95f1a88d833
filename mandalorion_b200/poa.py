"""ctypes binding of libmandalorion_poa.so (C ABI: include/mandalorion_poa.h).

The library replaces the `abpoa -M 5 -r 0 [-S] in.fasta` subprocess of the reference
(utils/SpliceDefineConsensus.py:917,919).  This module only marshals buffers; it never
computes a consensus itself and raises PoaError when the CUDA library or a GPU is missing.
"""
import ctypes as C
import os
import time
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def library_path():
    # MPOA_LIB: developer override to A/B-test another build of the SAME CUDA library
    return os.environ.get("MPOA_LIB") or os.path.join(_HERE, "libmandalorion_poa.so")


class PoaError(RuntimeError):
    pass


class _Params(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open1", C.c_int32), ("gap_ext1", C.c_int32),
                ("gap_open2", C.c_int32), ("gap_ext2", C.c_int32), ("wb", C.c_int32), ("wf", C.c_float),
                ("simd_pn_i16", C.c_int32), ("simd_pn_i32", C.c_int32), ("debug_small_caps", C.c_int32),
                ("reserved", C.c_int32 * 5)]


class _Stats(C.Structure):
    _fields_ = [("n_groups", C.c_int64), ("n_reads", C.c_int64), ("n_alignments", C.c_int64),
                ("band_cells", C.c_int64), ("full_cells", C.c_int64), ("int_ops", C.c_int64),
                ("n_align_i16", C.c_int64), ("n_align_i32", C.c_int64), ("tb_bytes", C.c_int64),
                ("n_retry_groups", C.c_int64), ("kernel_ms", C.c_double), ("h2d_ms", C.c_double),
                ("d2h_ms", C.c_double), ("n_kernel_launches", C.c_int64), ("phase_cycles", C.c_int64 * 6),
                ("n_seed_groups", C.c_int64), ("n_seed_applied", C.c_int64), ("n_too_big_groups", C.c_int64),
                ("max_band_width", C.c_int64), ("host_seed_ms", C.c_double), ("kernel_wait_ms", C.c_double),
                ("reserved", C.c_int64 * 2)]


class _Trace(C.Structure):
    _fields_ = [("read_score", C.c_void_p), ("read_bits", C.c_void_p), ("read_band_cells", C.c_void_p),
                ("base_aln", C.c_void_p), ("base_node", C.c_void_p)]


@dataclass
class PoaParams:
    """Defaults are the reference's command line `abpoa -M 5 -r 0` (utils/SpliceDefineConsensus.py:917)."""
    match: int = 5
    mismatch: int = 4
    gap_open1: int = 4
    gap_ext1: int = 2
    gap_open2: int = 24
    gap_ext2: int = 1
    wb: int = 10
    wf: float = 0.01
    simd_pn_i16: int = 16
    simd_pn_i32: int = 8
    debug_small_caps: int = 0      # tests only: force the GPU retry paths


_lib = None

# every symbol include/mandalorion_poa.h declares
ABI_SYMBOLS = ("mpoa_abi_version", "mpoa_default_params", "mpoa_create", "mpoa_destroy", "mpoa_last_error",
               "mpoa_set_stream", "mpoa_set_trace", "mpoa_measure_int_peak", "mpoa_consensus_batch", "mpoa_batch_upload", "mpoa_batch_run",
               "mpoa_batch_fetch", "mpoa_orient_batch", "mpoa_batch_upload_subset", "mpoa_shard_plan",
               "mpoa_consensus_batch_multi")


def _load():
    global _lib
    if _lib is None:
        path = library_path()
        if not os.path.exists(path):
            raise PoaError(f"{path} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the consensus path)")
        lib = C.CDLL(path)
        lib.mpoa_abi_version.restype = C.c_int
        lib.mpoa_last_error.restype = C.c_char_p
        lib.mpoa_last_error.argtypes = [C.c_void_p]
        lib.mpoa_default_params.argtypes = [C.c_void_p]
        lib.mpoa_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_void_p]
        lib.mpoa_destroy.argtypes = [C.c_void_p]
        lib.mpoa_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        lib.mpoa_set_trace.argtypes = [C.c_void_p, C.c_int]
        lib.mpoa_measure_int_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        lib.mpoa_batch_upload.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 4
        lib.mpoa_batch_run.argtypes = [C.c_void_p, C.c_void_p]
        lib.mpoa_batch_fetch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        lib.mpoa_consensus_batch.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 6 + [C.c_int64] + [C.c_void_p] * 3
        lib.mpoa_orient_batch.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        lib.mpoa_batch_upload_subset.argtypes = [C.c_void_p, C.c_int64] + [C.c_void_p] * 4 + [C.c_int64, C.c_void_p]
        lib.mpoa_shard_plan.argtypes = [C.c_int64] + [C.c_void_p] * 4 + [C.c_int32, C.c_void_p]
        lib.mpoa_consensus_batch_multi.argtypes = ([C.c_void_p, C.c_int32, C.c_int64] + [C.c_void_p] * 6 + [C.c_int64]
                                                   + [C.c_void_p] * 3)
        _lib = lib
    return _lib


def pack_groups(groups):
    """list of groups (each a list of str/bytes reads) -> (group_read_off, read_base_off, bases)."""
    gro = np.zeros(len(groups) + 1, dtype=np.int64)
    lens, chunks = [], []
    for gi, reads in enumerate(groups):
        gro[gi + 1] = gro[gi] + len(reads)
        for r in reads:
            b = r.encode() if isinstance(r, str) else bytes(r)
            lens.append(len(b))
            chunks.append(b)
    rbo = np.zeros(len(lens) + 1, dtype=np.int64)
    if lens:
        rbo[1:] = np.cumsum(np.asarray(lens, dtype=np.int64))
    bases = np.frombuffer(b"".join(chunks), dtype=np.uint8).copy() if chunks else np.zeros(0, dtype=np.uint8)
    return gro, rbo, bases


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def orient_batch(packed, n_threads=None):
    """Primary-hit strands of every read against the first read of its group (C ABI
    mpoa_orient_batch; reference utils/SpliceDefineConsensus.py:895, :900-907).
    Returns (hit_count int8[n_reads], hit_strand int8[n_reads, 2]).  Host code: needs no GPU."""
    lib = _load()
    gro, rbo, bases = [np.ascontiguousarray(a, dtype=t) for a, t in zip(packed, (np.int64, np.int64, np.uint8))]
    nr = len(rbo) - 1
    cnt = np.zeros(max(nr, 1), dtype=np.int8)
    strand = np.zeros((max(nr, 1), 2), dtype=np.int8)
    rc = lib.mpoa_orient_batch(len(gro) - 1, _ptr(gro), _ptr(rbo), _ptr(bases), int(n_threads or os.cpu_count() or 1),
                               _ptr(cnt), _ptr(strand))
    if rc != 0:
        raise PoaError(f"mpoa_orient_batch failed with {rc}")
    return cnt[:nr], strand[:nr]


def _c_params(p):
    p = p or PoaParams()
    return _Params(p.match, p.mismatch, p.gap_open1, p.gap_ext1, p.gap_open2, p.gap_ext2, p.wb, p.wf,
                   p.simd_pn_i16, p.simd_pn_i32, p.debug_small_caps)


def shard_plan(packed, n_shards, flags=None, params=None):
    """Owner (0..n_shards-1) of every group: longest-processing-time greedy over the estimated DP cost
    (C ABI mpoa_shard_plan).  Host code: needs no GPU."""
    lib = _load()
    gro, rbo = [np.ascontiguousarray(a, dtype=np.int64) for a in packed[:2]]
    if flags is not None:
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
    ng = len(gro) - 1
    owner = np.zeros(max(ng, 1), dtype=np.int32)
    cp = _c_params(params)
    rc = lib.mpoa_shard_plan(ng, _ptr(gro), _ptr(rbo), _ptr(flags), C.byref(cp), int(n_shards), _ptr(owner))
    if rc != 0:
        raise PoaError(f"mpoa_shard_plan failed with {rc}")
    return owner[:ng]


def _split(raw, off):
    off = off.tolist()
    return [raw[off[i]:off[i + 1]] for i in range(len(off) - 1)]


def consensus_batch_multi(contexts, packed, flags=None):
    """ONE batch over several PoaContexts (one per GPU) through the C ABI's
    mpoa_consensus_batch_multi: native LPT sharding, one host thread per GPU, results in input
    order.  Returns dict(cons=[bytes], status=int32[], stats=[dict per context], owner=int32[])."""
    lib = _load()
    gro, rbo, bases = [np.ascontiguousarray(a, dtype=t) for a, t in zip(packed, (np.int64, np.int64, np.uint8))]
    if flags is not None:
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
    ng = len(gro) - 1
    n = len(contexts)
    handles = (C.c_void_p * n)(*[c._h for c in contexts])
    cap = max(16, int(rbo[-1]) if len(rbo) else 0)
    cons_buf = np.empty(cap, dtype=np.uint8)
    cons_off = np.zeros(ng + 1, dtype=np.int64)
    status = np.zeros(ng, dtype=np.int32)
    owner = np.zeros(max(ng, 1), dtype=np.int32)
    st = (_Stats * n)()
    rc = lib.mpoa_consensus_batch_multi(handles, n, ng, _ptr(gro), _ptr(rbo), _ptr(bases), _ptr(flags), _ptr(cons_off),
                                        _ptr(cons_buf), cap, _ptr(status), st, _ptr(owner))
    if rc != 0:
        msg = lib.mpoa_last_error(contexts[0]._h)
        raise PoaError(f"mpoa_consensus_batch_multi failed with {rc}: {msg.decode() if msg else ''}")
    stats = [PoaContext._stats_dict(x) for x in st]
    for c, d in zip(contexts, stats):
        c.last_stats = d
    return dict(cons=_split(cons_buf[:cons_off[ng]].tobytes(), cons_off), status=status, cons_off=cons_off,
                stats=stats, owner=owner[:ng])


class PoaContext:
    """One context per (process, GPU); not thread-safe (include/mandalorion_poa.h)."""

    def __init__(self, device=0, params=None):
        lib = _load()
        cp = _c_params(params)
        h = C.c_void_p()
        rc = lib.mpoa_create(C.byref(h), int(device), C.byref(cp))
        if rc != 0:
            raise PoaError(f"mpoa_create(device={device}) failed with {rc}: a B200 (sm_100) CUDA device is required; "
                           "there is no CPU fallback")
        self._h = h
        self._lib = lib
        self.device = device
        self.last_stats = {}
        self._n = (0, 0, 0)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.mpoa_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            msg = self._lib.mpoa_last_error(self._h)
            raise PoaError(f"{what} failed with {rc}: {msg.decode() if msg else ''}")

    def set_stream(self, cuda_stream_handle):
        self._check(self._lib.mpoa_set_stream(self._h, C.c_void_p(int(cuda_stream_handle))), "mpoa_set_stream")

    def set_trace(self, enable):
        self._check(self._lib.mpoa_set_trace(self._h, int(bool(enable))), "mpoa_set_trace")

    def measure_int_peak(self):
        """Warp-wide VIADDMNMX.S16x2 instructions per second over the whole GPU (INT-pipe roofline)."""
        v = C.c_double()
        self._check(self._lib.mpoa_measure_int_peak(self._h, C.byref(v)), "mpoa_measure_int_peak")
        return v.value

    @staticmethod
    def _stats_dict(st):
        d = {k: getattr(st, k) for k, _ in _Stats._fields_ if k not in ("reserved", "phase_cycles")}
        d["phase_cycles"] = dict(zip(("prep", "dp", "traceback", "merge", "consensus", "busy"), list(st.phase_cycles)))
        return d

    # ---- three-stage interface (inputs stay resident in HBM between run() calls) ----
    def upload(self, gro, rbo, bases, flags=None, subset=None):
        """subset: ascending group indices -- only those groups are uploaded (numbered in that order)."""
        if flags is not None:
            flags = np.ascontiguousarray(flags, dtype=np.uint8)
        gro = np.ascontiguousarray(gro, dtype=np.int64)
        rbo = np.ascontiguousarray(rbo, dtype=np.int64)
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        ng = len(gro) - 1
        if subset is not None:
            sel = np.ascontiguousarray(subset, dtype=np.int64)
            self._check(self._lib.mpoa_batch_upload_subset(self._h, ng, _ptr(gro), _ptr(rbo), _ptr(bases), _ptr(flags),
                                                           len(sel), _ptr(sel)), "mpoa_batch_upload_subset")
            reads = np.concatenate([np.arange(gro[g], gro[g + 1]) for g in sel]) if len(sel) else np.zeros(0, np.int64)
            self._n = (len(sel), len(reads), int((rbo[reads + 1] - rbo[reads]).sum()) if len(reads) else 0)
            return
        self._check(self._lib.mpoa_batch_upload(self._h, ng, _ptr(gro), _ptr(rbo), _ptr(bases), _ptr(flags)),
                    "mpoa_batch_upload")
        self._n = (ng, len(rbo) - 1, int(rbo[-1]) if len(rbo) else 0)

    def run(self):
        st = _Stats()
        self._check(self._lib.mpoa_batch_run(self._h, C.byref(st)), "mpoa_batch_run")
        self.last_stats = self._stats_dict(st)
        return self.last_stats

    def fetch(self, trace=False):
        ng, nr, nb = self._n
        cap = max(16, nb)
        cons_buf = np.empty(cap, dtype=np.uint8)
        cons_off = np.zeros(ng + 1, dtype=np.int64)
        status = np.zeros(ng, dtype=np.int32)
        tr, arrs = None, None
        if trace:
            arrs = dict(read_score=np.zeros(nr, np.int32), read_bits=np.zeros(nr, np.int32),
                        read_band_cells=np.zeros(nr, np.int64), base_aln=np.full(nb, -9, np.int32),
                        base_node=np.full(nb, -9, np.int32))
            tr = _Trace(*[arrs[k].ctypes.data for k in
                          ("read_score", "read_bits", "read_band_cells", "base_aln", "base_node")])
        self._check(self._lib.mpoa_batch_fetch(self._h, _ptr(cons_off), _ptr(cons_buf), cap, _ptr(status),
                                               C.byref(tr) if tr is not None else None), "mpoa_batch_fetch")
        cons = _split(cons_buf[:cons_off[ng]].tobytes(), cons_off)
        return dict(cons=cons, status=status, cons_off=cons_off, trace=arrs)

    # ---- the one-call interface ----
    def consensus_batch(self, groups=None, packed=None, trace=False, flags=None):
        """groups: list of lists of reads (str/bytes), aligned in the given order; flags: MPOA_FLAG_* per
        group.  Returns dict(cons=[bytes], status=int32[], stats=dict, trace=dict|None)."""
        gro, rbo, bases = packed if packed is not None else pack_groups(groups)
        gro = np.ascontiguousarray(gro, dtype=np.int64)
        rbo = np.ascontiguousarray(rbo, dtype=np.int64)
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        if flags is not None:
            flags = np.ascontiguousarray(flags, dtype=np.uint8)
        ng, nr = len(gro) - 1, len(rbo) - 1
        nb = int(rbo[-1]) if len(rbo) else 0
        cap = max(16, nb)                                   # a consensus never outgrows its group's bases
        cons_buf = np.empty(cap, dtype=np.uint8)
        cons_off = np.zeros(ng + 1, dtype=np.int64)
        status = np.zeros(ng, dtype=np.int32)
        tr, arrs = None, None
        if trace:
            arrs = dict(read_score=np.zeros(nr, np.int32), read_bits=np.zeros(nr, np.int32),
                        read_band_cells=np.zeros(nr, np.int64), base_aln=np.full(nb, -9, np.int32),
                        base_node=np.full(nb, -9, np.int32))
            tr = _Trace(*[arrs[k].ctypes.data for k in
                          ("read_score", "read_bits", "read_band_cells", "base_aln", "base_node")])
        st = _Stats()
        # ONE C call (upload -> kernels -> fetch): the library knows the buffers stay valid and overlaps the
        # host-side `-S` seeding with the kernels of the unseeded groups
        self._check(self._lib.mpoa_consensus_batch(self._h, ng, _ptr(gro), _ptr(rbo), _ptr(bases), _ptr(flags),
                                                   _ptr(cons_off), _ptr(cons_buf), cap, _ptr(status), C.byref(st),
                                                   C.byref(tr) if tr is not None else None), "mpoa_consensus_batch")
        self._n = (ng, nr, nb)
        self.last_stats = self._stats_dict(st)
        return dict(cons=_split(cons_buf[:cons_off[ng]].tobytes(), cons_off), status=status, cons_off=cons_off,
                    trace=arrs, stats=self.last_stats)


class PoaPipeline:
    """Batches in flight on ONE GPU: `depth` contexts, each served by its own host thread, so that the
    host->device copy, the host-side passes and the result split of one batch run while the kernels of
    another occupy the SMs, and the persistent blocks of the next batch move in while the last groups of
    the previous one drain.  This is how a streaming caller (dstep.StreamingConsensus: loci arriving one
    after the other, reference defineIsoforms.py:130-153) keeps the GPU busy.  Every context owns a
    non-blocking stream (mpoa_create), so the contexts only meet in the hardware's block scheduler.

        with PoaPipeline(device=0, depth=2) as pipe:
            futures = [pipe.submit(packed=p, flags=f) for p, f in batches]
            results = [f.result() for f in futures]        # dicts as from PoaContext.consensus_batch

    Results are identical to PoaContext.consensus_batch on the same batch (groups are independent)."""

    def __init__(self, device=0, depth=2, params=None, contexts=None):
        import queue
        import threading
        self._own = contexts is None
        self.contexts = list(contexts) if contexts is not None else [PoaContext(device, params) for _ in range(depth)]
        if not self.contexts:
            raise PoaError("PoaPipeline needs at least one context")
        self._jobs = queue.Queue()
        self._threads = [threading.Thread(target=self._serve, args=(c,), daemon=True) for c in self.contexts]
        for t in self._threads:
            t.start()

    def _serve(self, ctx):
        while True:
            job = self._jobs.get()
            if job is None:
                return
            fut, kw = job
            if not fut.set_running_or_notify_cancel():
                continue
            try:
                t0 = time.perf_counter()
                res = ctx.consensus_batch(**kw)
                res["wall"] = (t0, time.perf_counter())      # when this batch was in its context (perf_counter)
                fut.set_result(res)
            except BaseException as e:       # noqa: BLE001 -- handed to the caller through the future
                fut.set_exception(e)

    def submit(self, groups=None, packed=None, trace=False, flags=None):
        """Queues one batch; returns a concurrent.futures.Future of PoaContext.consensus_batch's dict.  The
        caller's arrays must stay unchanged until the future is done."""
        from concurrent.futures import Future
        fut = Future()
        self._jobs.put((fut, dict(groups=groups, packed=packed, trace=trace, flags=flags)))
        return fut

    def map(self, batches):
        """batches: iterable of (packed, flags); yields the results in input order."""
        futs = [self.submit(packed=p, flags=f) for p, f in batches]
        for f in futs:
            yield f.result()

    def close(self):
        for _ in self._threads:
            self._jobs.put(None)
        for t in self._threads:
            t.join()
        self._threads = []
        if self._own:
            for c in self.contexts:
                c.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
