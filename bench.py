#!/usr/bin/env python
"""bench.py -- POA consensus throughput (BASELINE.json metric: consensus groups/sec and GCUPS).

A "step" is one pass of the hot path (every read group of the batch: graph build, banded DP,
traceback, merge, heaviest-bundle consensus) over one batch of synthetic groups of a BASELINE.json
config.  Default = cfg2 (configs[1]: 10-50 reads, 0.5-4 kb, ~1 % R2C2-like error; the full config is
200k groups on 8 GPUs, one step processes --groups of them per GPU).

  --config cfg1|cfg2|cfg3|cfg4|cfg4u   the other named configs (cfg4 = capped at 100 reads per group like
                                       the reference's subsample, cfg4u = uncapped kernel stress; cfg3 also
                                       reports its "< 8 kb" and ">= 8 kb (abpoa -S in the reference)" halves)
  --scaling weak|strong                weak (default): every rank owns its own slice, no collective
                                       strong: ONE batch pushed through shard.consensus_batch_sharded
                                       over the N GPUs (rank 0 drives them), results gathered in input order
  --dstep                              cfg1 through the reference's function boundary: prepare_group
                                       (subsample + orientation) -> one GPU batch -> file writer
  --pipeline D                         contexts of the pipelined e2e leg (default 2; 1 = serial calls only)
  --shards-per-gpu K                   strong scaling: K shards (contexts) per GPU (default 1)

  value      groups/s, whole job, inputs resident in HBM (mpoa_batch_run only, CUDA events)
  e2e        groups/s through PoaContext.consensus_batch() with pinned HOST buffers in and out: measured as one
             call after the other AND with the same calls through PoaPipeline (two batches in flight: the copies
             of one beside the kernels of the other; every copy of every step inside the timed region); the
             line carries both, e2e.value is the better one and e2e.mode names it
  roofline   HBM: algorithmic bytes (1 B traceback per band cell + 1 B/base in + 1 B/base out)
             over the kernel time; int_roofline: algorithmic integer ops over the INT-pipe peak
  cpu_baseline / --impl reference: the CPU port of the reference's abpoa path (oracle/, a SCALAR
             restatement of abPOA v1.4.1 -- real abpoa is SIMD and several times faster per core)
             on the host cores, bounded sample.  The real abpoa binary cannot exist in this image.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "poa_consensus_groups_per_sec"
UNIT = "groups/s"
WORKLOADS = {
    "cfg1": ("cfg1", 1000, "cfg1: synthetic isoform groups, 3-30 reads, 1-2 kb, 1% R2C2-like error (the D-step config)"),
    "cfg2": ("cfg2", 32768, "cfg2: synthetic isoform groups, 10-50 reads, 0.5-4 kb log-uniform, 1% R2C2-like error"),
    "cfg3": ("cfg3", 20000, "cfg3: long-isoform stress, 5-30 reads of 5-12 kb, 1% error, wide adaptive band (the whole 20k-group config per step)"),
    "cfg4": ("cfg4", 8192, "cfg4: high-depth CCS-like groups, 50-200 reads of 2 kb capped to 100 per group (reference subsample)"),
    "cfg4u": ("cfg4", 8192, "cfg4 uncapped: high-depth CCS-like groups, 50-200 reads of 2 kb, every read aligned"),
}


def make_batch(config, n_groups, first, workers=None):
    from mandalorion_b200.synth import make_packed
    cfg_name = WORKLOADS[config][0]
    gro, rbo, bases = make_packed(cfg_name, n_groups, first=first, workers=workers)
    if config == "cfg4":          # the reference never hands abpoa more than 100 reads (utils/SpliceDefineConsensus.py:885)
        keep = np.concatenate([np.arange(gro[g], min(gro[g + 1], gro[g] + 100)) for g in range(n_groups)])
        lens = np.diff(rbo)[keep]
        counts = np.minimum(np.diff(gro), 100)
        chunks = [bases[rbo[r]:rbo[r + 1]] for r in keep]
        gro = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        rbo = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        bases = np.concatenate(chunks)
    return gro, rbo, bases


def subset(packed, idx):
    from mandalorion_b200.shard import take_shard_fast
    return take_shard_fast(packed[0], packed[1], packed[2], np.asarray(idx, dtype=np.int64))


def group_medians(packed):
    gro, rbo, _ = packed
    lens = np.diff(rbo)
    return np.array([np.median(lens[gro[g]:gro[g + 1]]) if gro[g + 1] > gro[g] else 0 for g in range(len(gro) - 1)])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def seed_flags(packed):
    """MPOA_FLAG_SEED per group, the reference's rule: median read length >= 8000 -> `abpoa -S`
    (utils/SpliceDefineConsensus.py:915-919)"""
    return (group_medians(packed) >= 8000).astype(np.uint8)


def cpu_port_groups_per_sec(packed_sample, threads):
    """The CPU port of the reference's abpoa path on the host cores (bounded sample)."""
    from oracle import oracle_consensus_batch
    t0 = time.perf_counter()
    out = oracle_consensus_batch(packed=packed_sample, n_threads=threads, flags=seed_flags(packed_sample))
    dt = time.perf_counter() - t0
    ng = len(packed_sample[0]) - 1
    return ng / dt, out["stats"]["band_cells"] / dt / 1e9, dt


def sample_of(packed, n):
    gro, rbo, bases = packed
    n = min(n, len(gro) - 1)
    r1 = gro[n]
    return gro[:n + 1].copy(), rbo[:r1 + 1].copy(), bases[:rbo[r1]].copy()


PORT_NOTE = ("oracle/ C++ SCALAR port of abPOA v1.4.1 (-M 5 -r 0); the real abpoa binary (SIMD, several times faster "
             "per core) is unavailable in the image")


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path = `abpoa` per group on the
    host cores.  abPOA is a third-party binary that is absent from /root/reference and from this
    image, so the oracle port stands in (kind "port"), with every host thread it can use."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_sample = args.ref_groups or {"cfg1": 1000, "cfg2": 1024, "cfg3": 64, "cfg4": 64, "cfg4u": 48}[args.config]
    packed = make_batch(args.config, n_sample, 0)
    for _ in range(args.warmup):
        cpu_port_groups_per_sec(sample_of(packed, max(cores, n_sample // 4)), cores)
    t_tot, g_tot, cells = 0.0, 0, 0.0
    for _ in range(args.steps):
        gps, gcups, dt = cpu_port_groups_per_sec(packed, cores)
        t_tot += dt
        g_tot += n_sample
        cells += gcups * dt
    value = g_tot / t_tot
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int16/int32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.config][2], "groups_per_step": n_sample},
            "gcups": cells / t_tot,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n_sample} {args.config} groups per step, {PORT_NOTE}, {cores} threads"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def resident_run(ctx, packed_pinned, steps, warmup, stream=None):
    """upload once, time `steps` x mpoa_batch_run; returns (sum kernel ms, launches, last stats, fetch output)"""
    ctx.upload(*packed_pinned, flags=seed_flags(packed_pinned))
    for _ in range(warmup):
        ctx.run()
    kernel_ms, launches, stats = 0.0, 0, None
    for _ in range(steps):
        stats = ctx.run()
        kernel_ms += stats["kernel_ms"]
        launches += stats["n_kernel_launches"]
    return kernel_ms, launches, stats, ctx.fetch()


def run_dstep(args):
    """cfg1 through the reference's function boundary (defineIsoforms.py:87-91 -> determine_consensus):
    prepare_group per isoform (np.random.choice subsample + orientation), ONE GPU batch, file writer.
    Beside it: the same prepared groups through the CPU port, all host threads (what one `abpoa`
    process per isoform costs, without the process spawns)."""
    import torch
    from mandalorion_b200 import PoaContext, consensus as C
    from mandalorion_b200.synth import make_groups
    from oracle import oracle_consensus_batch, pack_groups
    n = args.groups or 1000
    groups = make_groups("cfg1", n, random_strand=True, with_names=True)
    ctx = PoaContext(0)
    ctx.consensus_batch([[r[1] for r in groups[0]]])      # context + kernels warm
    out_dir = tempfile.mkdtemp(prefix="mpoa_dstep_")
    best = None
    for rep in range(max(1, args.steps)):
        np.random.seed(20261018)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        prepared = {"chr1~%d~%d" % (1000 * g, 1000 * g + 900): {"1": C.prepare_group(reads)} for g, reads in enumerate(groups)}
        t1 = time.perf_counter()
        C.orient_pending([pg for iso in prepared.values() for pg in iso.values()])
        t2 = time.perf_counter()
        results = C.finish_prepared(prepared, ctx=ctx)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        C.write_isoform_files(list(prepared), results, out_dir)
        t4 = time.perf_counter()
        cur = dict(total_s=t4 - t0, subsample_s=t1 - t0, orient_s=t2 - t1, gpu_batch_s=t3 - t2, write_s=t4 - t3)
        if best is None or cur["total_s"] < best["total_s"]:
            best = cur
    # the CPU port on the same prepared groups
    seqs = [pg.sequences for iso in prepared.values() for pg in iso.values() if not pg.bypass]
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    want = oracle_consensus_batch(packed=pack_groups(seqs), n_threads=cores)
    cpu_s = time.perf_counter() - t0
    got = [results[r]["1"][0] for r in prepared if not prepared[r]["1"].bypass]
    same = sum(1 for a, b in zip(got, want["cons"]) if not b or a == b.decode())
    files = dstep_from_files(ctx, args)
    line = {"metric": "dstep_wall_seconds", "value": best["total_s"], "unit": "s", "higher_is_better": False, "n_gpus": 1,
            "config": {"workload": WORKLOADS["cfg1"][2] + ", random strand, through prepare_group -> finish_prepared -> "
                                   "write_isoform_files", "groups": n},
            "breakdown_s": best, "groups_per_s": n / best["total_s"],
            "orienter": "mappy" if C.mappy_available() else "mpoa_orient_batch (C++ seed-chain stage, %d threads)" % cores,
            "cpu_port": {"consensus_only_s": cpu_s, "cores": cores, "kind": "port", "note": PORT_NOTE},
            "gpu_equals_port": same == len(got), "data": "synthetic", "from_locus_files": files}
    print(json.dumps(line), flush=True)


def dstep_from_files(ctx, args):
    """The whole D step from tmp_SS/*.psl (BASELINE cfg5 shape at bench scale): group producer (locus.py) ->
    prepare_group -> streamed GPU batches -> writer, i.e. dstep.define_isoforms(); the phases are host
    seconds of the producing thread, `drain` is what was still running on the GPU after the last locus."""
    from mandalorion_b200.dstep import define_isoforms
    from mandalorion_b200.synth_loci import write_locus, write_spliced_locus
    n_loci = max(4, (args.groups or 1000) // 40)
    work = tempfile.mkdtemp(prefix="mpoa_dstep_files_")
    tmp_ss = os.path.join(work, "tmp_SS")
    os.makedirs(tmp_ss)
    rng = np.random.Generator(np.random.PCG64(20261018 + 5))
    t0 = time.perf_counter()
    for k in range(n_loci):
        chrom = "chr%d" % (1 + k % 5)
        if k % 4 == 3:
            write_locus(tmp_ss, chrom, 100000 * (k + 1), [(3000 * j, int(rng.integers(900, 1800)), int(rng.integers(3, 30)))
                                                        for j in range(int(rng.integers(1, 5)))], rng, err=0.01)
        else:
            write_spliced_locus(tmp_ss, chrom, 100000 * (k + 1), rng, n_reads=int(rng.integers(40, 220)),
                                n_exons=int(rng.integers(4, 9)), strand="+-"[k % 2], err=0.01)
    gen_s = time.perf_counter() - t0
    best = None
    for rep in range(max(1, min(args.steps, 3))):
        np.random.seed(20261018)
        t0 = time.perf_counter()
        n_iso = define_isoforms(work, ctx=ctx, workers=min(16, os.cpu_count() or 1))
        total = time.perf_counter() - t0
        cur = dict(define_isoforms.last_timings, total_s=total)
        if best is None or total < best["total_s"]:
            best = cur
    best["isoforms_per_s"] = n_iso / best["total_s"]
    best["input_generation_s"] = gen_s
    best["workload"] = "%d loci (3/4 spliced genes with 4-8 exons and 40-220 reads, 1/4 mono-exonic), 1 %% error" % n_loci
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--dstep", action="store_true")
    ap.add_argument("--groups", type=int, default=0, help="groups per GPU per step (strong scaling: of the whole batch)")
    ap.add_argument("--ref-groups", type=int, default=0, help="groups per step of the CPU reference arm")
    ap.add_argument("--cpu-sample", type=int, default=0, help="groups of the cpu_baseline sample")
    ap.add_argument("--shards-per-gpu", type=int, default=1, help="strong scaling: shards (contexts) per GPU")
    ap.add_argument("--pipeline", type=int, default=2, help="contexts of the pipelined e2e leg (1: serial calls only)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return 0

    import torch
    import torch.distributed as dist
    from mandalorion_b200 import PoaContext

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the consensus path has no CPU fallback")
    if args.dstep:
        if rank == 0:
            run_dstep(args)
        return 0
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_per = args.groups or WORKLOADS[args.config][1]
    workload = WORKLOADS[args.config][2]
    cpu_sample = args.cpu_sample or {"cfg1": 1000, "cfg2": 3072, "cfg3": 96, "cfg4": 96, "cfg4u": 64}[args.config]

    if args.scaling == "strong":
        return strong_scaling(args, rank, world, n_per, workload, barrier)

    # every rank owns its own slice of the config (weak scaling); pinned host buffers for the e2e leg
    gro, rbo, bases = make_batch(args.config, n_per, first=rank * n_per,
                                 workers=max(1, (os.cpu_count() or 1) // max(1, world)))
    pin = [torch.from_numpy(a).pin_memory() for a in (gro, rbo, bases)]
    gro_p, rbo_p, bases_p = [t.numpy() for t in pin]
    n_groups, n_bases = len(gro) - 1, int(rbo[-1])

    ctx = PoaContext(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    flags = seed_flags((gro, rbo, bases))          # all zero except for cfg3's >= 8 kb groups

    # ---- resident leg: upload once, time mpoa_batch_run ----
    ctx.upload(gro_p, rbo_p, bases_p, flags)
    for _ in range(args.warmup):
        ctx.run()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    kernel_ms, launches, stats = 0.0, 0, None
    for _ in range(args.steps):
        stats = ctx.run()
        kernel_ms += stats["kernel_ms"]
        launches += stats["n_kernel_launches"]
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    dev_ms = ev0.elapsed_time(ev1)
    out = ctx.fetch()
    cons_bases = int(out["cons_off"][-1])
    n_ok = int((out["status"] == 0).sum())

    # ---- end-to-end leg: host buffers in, host buffers out, every step ----
    # (a) one call after the other through PoaContext.consensus_batch()
    for _ in range(min(args.warmup, 1)):
        ctx.consensus_batch(packed=(gro_p, rbo_p, bases_p), flags=flags)
    barrier()
    t0 = time.perf_counter()
    e2e_parts = dict(kernel=0.0, h2d=0.0, d2h=0.0)
    for _ in range(args.steps):
        ctx.consensus_batch(packed=(gro_p, rbo_p, bases_p), flags=flags)
        for k in e2e_parts:                                  # the library's own account of the call (CUDA events / host clock)
            e2e_parts[k] += ctx.last_stats[k + "_ms"] / args.steps
    torch.cuda.synchronize()
    e2e_serial_s = time.perf_counter() - t0
    e2e_parts["host_other"] = 1e3 * e2e_serial_s / args.steps - sum(e2e_parts.values())
    e2e_parts["host_seed_beside_the_kernels"] = ctx.last_stats.get("host_seed_ms", 0.0)   # not a part of the sum: overlapped
    # (b) the same calls through PoaPipeline: `depth` contexts, the copies and host passes of one batch beside
    # the kernels of another; all H2D / D2H copies of all steps are inside the timed region
    e2e_s, e2e_mode = e2e_serial_s, "PoaContext.consensus_batch, one call after the other"
    e2e_pipe_s, pipe_steps = None, None
    if args.pipeline > 1:
        from mandalorion_b200 import PoaPipeline
        extra = [PoaContext(local_rank) for _ in range(args.pipeline - 1)]
        with PoaPipeline(contexts=[ctx] + extra) as pipe:
            for f in [pipe.submit(packed=(gro_p, rbo_p, bases_p), flags=flags) for _ in range(args.pipeline)]:
                f.result()                                   # warm-up: every context sizes its buffers
            barrier()
            t0 = time.perf_counter()
            res = [f.result() for f in [pipe.submit(packed=(gro_p, rbo_p, bases_p), flags=flags) for _ in range(args.steps)]]
            torch.cuda.synchronize()
            e2e_pipe_s = time.perf_counter() - t0
            pipe_steps = [{"in_ms": round(1e3 * (r["wall"][0] - t0), 1), "out_ms": round(1e3 * (r["wall"][1] - t0), 1),
                           **{k: round(r["stats"][k], 1) for k in ("h2d_ms", "kernel_wait_ms", "kernel_ms", "d2h_ms")}} for r in res]
            del res
        for c in extra:
            c.close()
        if e2e_pipe_s < e2e_serial_s:
            e2e_s, e2e_mode = e2e_pipe_s, f"PoaPipeline(depth={args.pipeline}).submit, {args.steps} batches in flight two at a time"
    h2d = int(gro_p.nbytes + rbo_p.nbytes + bases_p.nbytes)
    d2h = int(cons_bases + 4 * n_groups + 4 * n_groups)

    # ---- cfg3: the two halves of the config separately (rank 0, resident) ----
    subsets = None
    if args.config == "cfg3" and rank == 0:
        med = group_medians((gro, rbo, bases))
        subsets = {}
        for name, idx in (("median < 8 kb (no -S in the reference): whole-graph alignment", np.nonzero(med < 8000)[0]),
                          ("median >= 8 kb (`abpoa -S` in the reference): windowed alignment between minimizer anchors",
                           np.nonzero(med >= 8000)[0])):
            if len(idx) == 0:
                continue
            kms, _, st, _ = resident_run(ctx, subset((gro, rbo, bases), idx), max(1, args.steps // 2), 1)
            per = kms / max(1, args.steps // 2)
            subsets[name] = {"groups": int(len(idx)), "groups_per_s": len(idx) / (per * 1e-3),
                             "gcups": st["band_cells"] / (per * 1e-3) / 1e9}

    # ---- reduce over ranks: max time, sum of groups ----
    t = torch.tensor([dev_ms, e2e_s * 1e3, kernel_ms], dtype=torch.float64, device="cuda")
    s = torch.tensor([n_groups, stats["band_cells"], stats["int_ops"], n_ok, launches, stats["full_cells"]],
                     dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    dev_ms_max, e2e_ms_max, kernel_ms_max = [float(x) for x in t.tolist()]
    tot_groups, tot_cells, tot_ops, tot_ok, tot_launches, tot_full = [float(x) for x in s.tolist()]

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        # rank 0's own kernel: algorithmic bytes per step / average step duration (CUDA events in the library)
        per_launch_ms = kernel_ms / max(1, args.steps)
        alg_bytes = stats["band_cells"] + n_bases + cons_bases
        achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(prof):
            try:
                # ncu --set full capture of the same kernel (profiles/README.md): DRAM bytes per band cell of
                # that capture x the band cells of THIS step
                pj = json.load(open(prof))
                traffic = pj.get("dram_bytes_per_band_cell") * stats["band_cells"]
                traffic_src = pj.get("source")
            except (OSError, ValueError, TypeError):
                traffic = None
        # INT-pipe peak measured live: VIADDMNMX.S16x2 warp-instructions/s x 32 lanes x 2 int16 ops
        # (SURVEY.md 8d counts one s16x2 instruction as 2 ops)
        int_peak = ctx.measure_int_peak() * 32 * 2
        int_achieved = stats["int_ops"] / (per_launch_ms * 1e-3)
        cores = os.cpu_count() or 1
        cpu_gps, cpu_gcups, cpu_dt = cpu_port_groups_per_sec(sample_of((gro, rbo, bases), cpu_sample), cores)
        value = tot_groups * args.steps / (dev_ms_max * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16x2 (scores relative to the diagonal; int32 lanes only as the fall-back for junk reads)",
            "data": "synthetic",
            "config": {"workload": workload, "groups_per_gpu_per_step": n_groups, "reads": int(len(rbo) - 1),
                       "bases_per_gpu": n_bases, "parallelism": f"groups sharded over {world} GPU(s), no collective",
                       "cache": "inputs + per-step traceback/workspace traffic exceed the 126 MB L2"},
            "gcups": tot_cells * args.steps / (dev_ms_max * 1e-3) / 1e9,
            "gcups_full_matrix": tot_full * args.steps / (dev_ms_max * 1e-3) / 1e9,
            "groups_ok_frac": tot_ok / max(1.0, tot_groups),
            "seed_flagged_groups": int(stats.get("n_seed_groups", 0)), "seed_applied_groups": int(stats.get("n_seed_applied", 0)),
            "e2e": {"value": tot_groups * args.steps / (e2e_ms_max * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "mode": e2e_mode,
                    "serial_calls_groups_per_s": n_groups * args.steps / e2e_serial_s,
                    "pipelined_groups_per_s": (n_groups * args.steps / e2e_pipe_s) if e2e_pipe_s else None,
                    "pipelined_steps_rank0": pipe_steps,
                    "serial_ms_per_step_rank0": {k: round(v, 2) for k, v in e2e_parts.items()}},
            "gpu_launches": int(tot_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "kernel": "poa_group_kernel", "kernel_ms_per_launch": per_launch_ms,
                         "kernel_launches_per_step": launches / max(1, args.steps),
                         "note": "integer DP: the ALU/latency bound is int_roofline; HBM carries 1 B/cell algorithmically"},
            "int_roofline": {"achieved_ops_per_s": int_achieved, "peak_ops_per_s": int_peak,
                             "frac": (int_achieved / int_peak) if int_peak else None,
                             "ops_per_cell": 17, "unit": "int16-lane op/s",
                             "peak_source": "measured live: VIADDMNMX.S16x2 chains (mpoa_measure_int_peak) x 32 lanes x 2"},
            "phase_share": {k: v / max(1, stats["phase_cycles"]["busy"]) for k, v in stats["phase_cycles"].items()},
            "cpu_baseline": {"value": cpu_gps, "unit": UNIT, "cores": cores, "kind": "port", "gcups": cpu_gcups,
                             "sample": f"first {min(cpu_sample, n_groups)} groups of the same batch, {PORT_NOTE}, "
                                       f"{cores} threads, {cpu_dt:.1f} s"},
            "clocks": clocks,
        }
        if subsets:
            line["subsets"] = subsets
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def strong_scaling(args, rank, world, n_total, workload, barrier):
    """ONE batch over N GPUs through the product's sharded entry point, end to end (host buffers in,
    consensus strings out in input order).  Rank 0 drives all N devices (one host thread + one
    context per GPU); under torchrun the other ranks only keep the rendezvous alive."""
    import torch
    import torch.distributed as dist
    from mandalorion_b200 import PoaContext
    from mandalorion_b200.shard import consensus_batch_sharded
    n_dev = args.gpus
    if rank == 0:
        packed = make_batch(args.config, n_total, first=0)
        k = max(1, args.shards_per_gpu)          # shards (and contexts) per GPU: copies of one beside the kernels of the other
        devs = [d for d in range(n_dev) for _ in range(k)]
        ctxs = {d: [PoaContext(d) for _ in range(k)] for d in range(n_dev)}
        for _ in range(max(1, min(args.warmup, 2))):
            consensus_batch_sharded(packed, devices=devs, contexts=ctxs)
        sampler = ClockSampler(0)
        sampler.start()
        t0 = time.perf_counter()
        imb = []
        for _ in range(args.steps):
            out = consensus_batch_sharded(packed, devices=devs, contexts=ctxs)
            imb.append(out["imbalance"])
        for d in range(n_dev):
            torch.cuda.synchronize(d)
        dt = time.perf_counter() - t0
        clocks = sampler.stop()
        cells = sum(s["band_cells"] for s in out["stats"])
        kms = [s["kernel_ms"] for s in out["stats"]]
        line = {"metric": METRIC, "value": n_total * args.steps / dt, "unit": UNIT, "n_gpus": n_dev, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "int16x2", "data": "synthetic",
                "config": {"workload": workload, "groups_total_per_step": n_total,
                           "parallelism": f"one batch, {k} LPT shard(s) per GPU over {n_dev} GPU(s) (shard.consensus_batch_sharded), "
                                          "host gather in input order, no collective"},
                "gcups": cells * args.steps / dt / 1e9,
                "e2e": {"value": n_total * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(sum(a.nbytes for a in packed)),
                        "d2h_bytes_per_step": int(sum(len(c) for c in out["cons"]))},
                "load_imbalance_max_over_mean_kernel_ms": float(np.mean(imb)), "kernel_ms_per_shard": kms,
                "groups_ok_frac": float((out["status"] == 0).mean()), "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
