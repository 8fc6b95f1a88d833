"""ctypes binding of the CPU oracle (oracle/libmpoa_oracle.so).  TEST INFRASTRUCTURE ONLY.

The oracle restates `abpoa -M 5 -r 0 in.fasta` (abPOA v1.4.1) as invoked by the reference at
utils/SpliceDefineConsensus.py:917.  PARITY UNPINNED (no abPOA source / binary / golden vector is
reachable from the reference or this image) -- see oracle/abpoa_oracle.cpp.
"""
import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libmpoa_oracle.so")


class _Params(C.Structure):
    _fields_ = [("match", C.c_int32), ("mismatch", C.c_int32), ("gap_open1", C.c_int32), ("gap_ext1", C.c_int32),
                ("gap_open2", C.c_int32), ("gap_ext2", C.c_int32), ("wb", C.c_int32), ("wf", C.c_float),
                ("simd_pn_i16", C.c_int32), ("simd_pn_i32", C.c_int32), ("debug_small_caps", C.c_int32),
                ("reserved", C.c_int32 * 5)]


class _Opts(C.Structure):
    _fields_ = [("clamp_end_to_pred", C.c_int32), ("single_argmax", C.c_int32), ("hb_tie_later_wins", C.c_int32),
                ("n_threads", C.c_int32), ("seed_k", C.c_int32), ("seed_w", C.c_int32), ("seed_min_w", C.c_int32),
                ("honour_seed_flag", C.c_int32)]


class _Stats(C.Structure):
    _fields_ = [("n_groups", C.c_int64), ("n_reads", C.c_int64), ("n_alignments", C.c_int64),
                ("band_cells", C.c_int64), ("full_cells", C.c_int64), ("int_ops", C.c_int64),
                ("n_align_i16", C.c_int64), ("n_align_i32", C.c_int64), ("tb_bytes", C.c_int64),
                ("n_retry_groups", C.c_int64), ("kernel_ms", C.c_double), ("h2d_ms", C.c_double),
                ("d2h_ms", C.c_double), ("n_kernel_launches", C.c_int64), ("phase_cycles", C.c_int64 * 6),
                ("n_seed_groups", C.c_int64), ("n_seed_applied", C.c_int64), ("n_too_big_groups", C.c_int64),
                ("max_band_width", C.c_int64), ("reserved", C.c_int64 * 4)]


class _Trace(C.Structure):
    _fields_ = [("read_score", C.c_void_p), ("read_bits", C.c_void_p), ("read_band_cells", C.c_void_p),
                ("base_aln", C.c_void_p), ("base_node", C.c_void_p)]


@dataclass
class OracleParams:
    """Defaults == `abpoa -M 5 -r 0` (reference utils/SpliceDefineConsensus.py:917)."""
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
    clamp_end_to_pred: int = 1
    single_argmax: int = 0
    hb_tie_later_wins: int = 1
    seed_k: int = 19                # `abpoa -S` constants (groups flagged MPOA_FLAG_SEED), low-confidence restatement
    seed_w: int = 10
    seed_min_w: int = 500
    honour_seed_flag: int = 1


def build_oracle(force=False):
    """Compile oracle/libmpoa_oracle.so with the committed Makefile (building the checker is not using it)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(
            os.path.join(_HERE, "abpoa_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "libmpoa_oracle.so"])
    return _LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        build_oracle()
        _lib = C.CDLL(_LIB)
        _lib.mpoa_oracle_consensus_batch.restype = C.c_int
    return _lib


def pack_groups(groups):
    """groups: list of list of str/bytes reads -> (group_read_off, read_base_off, bases) numpy arrays."""
    gro = np.zeros(len(groups) + 1, dtype=np.int64)
    lens = []
    chunks = []
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


def oracle_consensus_batch(groups=None, packed=None, params=None, n_threads=1, trace=False, flags=None):
    """Run the oracle.  Returns dict(cons=[bytes], status=np.int32[], stats=dict, trace=dict|None)."""
    lib = _load()
    p = params or OracleParams()
    gro, rbo, bases = packed if packed is not None else pack_groups(groups)
    ng = len(gro) - 1
    nreads, nbases = len(rbo) - 1, int(rbo[-1]) if len(rbo) else 0
    cp = _Params(p.match, p.mismatch, p.gap_open1, p.gap_ext1, p.gap_open2, p.gap_ext2, p.wb, p.wf,
                 p.simd_pn_i16, p.simd_pn_i32)
    co = _Opts(p.clamp_end_to_pred, p.single_argmax, p.hb_tie_later_wins, int(n_threads), p.seed_k, p.seed_w, p.seed_min_w,
               p.honour_seed_flag)
    if flags is not None:
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
    # a consensus is a path in the graph, so it can never be longer than the bases of its group
    cap = max(16, nbases)
    cons_buf = np.zeros(cap, dtype=np.uint8)
    cons_off = np.zeros(ng + 1, dtype=np.int64)
    status = np.zeros(ng, dtype=np.int32)
    st = _Stats()
    tr = None
    tr_arrays = None
    if trace:
        tr_arrays = dict(read_score=np.zeros(nreads, np.int32), read_bits=np.zeros(nreads, np.int32),
                         read_band_cells=np.zeros(nreads, np.int64), base_aln=np.full(nbases, -9, np.int32),
                         base_node=np.full(nbases, -9, np.int32))
        tr = _Trace(*[tr_arrays[k].ctypes.data for k in
                      ("read_score", "read_bits", "read_band_cells", "base_aln", "base_node")])
    rc = lib.mpoa_oracle_consensus_batch(
        C.byref(cp), C.byref(co), C.c_int64(ng), gro.ctypes.data_as(C.c_void_p), rbo.ctypes.data_as(C.c_void_p),
        bases.ctypes.data_as(C.c_void_p), flags.ctypes.data_as(C.c_void_p) if flags is not None else None,
        cons_off.ctypes.data_as(C.c_void_p),
        cons_buf.ctypes.data_as(C.c_void_p), C.c_int64(cap), status.ctypes.data_as(C.c_void_p), C.byref(st),
        C.byref(tr) if tr is not None else None)
    if rc != 0:
        raise RuntimeError(f"mpoa_oracle_consensus_batch failed: {rc}")
    raw = cons_buf.tobytes()
    cons = [raw[cons_off[i]:cons_off[i + 1]] for i in range(ng)]
    stats = {k: getattr(st, k) for k, _ in _Stats._fields_ if k not in ("reserved", "phase_cycles")}
    return dict(cons=cons, status=status, stats=stats, trace=tr_arrays)
