#!/usr/bin/env python
"""bench.py -- POA consensus throughput (BASELINE.json metric: consensus groups/sec and GCUPS).

A "step" is one pass of the hot path (every read group of the batch: graph build, banded DP,
traceback, merge, heaviest-bundle consensus) over one batch of synthetic cfg2-shaped groups
(BASELINE.json configs[1]: 10-50 reads, 0.5-4 kb, ~1 % R2C2-like error).  The full config is 200k
groups on 8 GPUs; one step processes --groups of them per GPU (weak scaling: every rank owns
its own slice, no collective on the data path -- SURVEY.md section 8e).

  value      groups/s, whole job, inputs resident in HBM (mpoa_batch_run only, CUDA events)
  e2e        groups/s through PoaContext.consensus_batch() with pinned HOST buffers in and out
  roofline   HBM: algorithmic bytes (1 B traceback per band cell + 1 B/base in + 1 B/base out)
             over the kernel time; int_roofline: algorithmic integer ops over the INT-pipe peak
  cpu_baseline / --impl reference: the CPU port of the reference's abpoa path (oracle/) on the
             host cores, bounded sample.  The real abpoa binary cannot exist in this image.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "poa_consensus_groups_per_sec"
UNIT = "groups/s"
WORKLOAD = "cfg2: synthetic isoform groups, 10-50 reads, 0.5-4 kb log-uniform, 1% R2C2-like error"


def make_batch(n_groups, first, workers=None):
    from mandalorion_b200.synth import make_packed
    return make_packed("cfg2", n_groups, first=first, workers=workers)


def algorithmic_bytes(stats, n_bases, cons_bases):
    return stats["band_cells"] + n_bases + cons_bases


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


def cpu_port_groups_per_sec(packed_sample, threads):
    """The CPU port of the reference's abpoa path on the host cores (bounded sample)."""
    from oracle import oracle_consensus_batch
    t0 = time.perf_counter()
    out = oracle_consensus_batch(packed=packed_sample, n_threads=threads)
    dt = time.perf_counter() - t0
    ng = len(packed_sample[0]) - 1
    return ng / dt, out["stats"]["band_cells"] / dt / 1e9, dt


def sample_of(packed, n):
    gro, rbo, bases = packed
    n = min(n, len(gro) - 1)
    r1 = gro[n]
    return gro[:n + 1].copy(), rbo[:r1 + 1].copy(), bases[:rbo[r1]].copy()


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path = `abpoa` per group on the
    host cores.  abPOA is a third-party binary that is absent from /root/reference and from this
    image, so the oracle port stands in (kind "port"), with every host thread it can use."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_sample = args.ref_groups
    packed = make_batch(n_sample, 0)
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
            "config": {"workload": WORKLOAD, "groups_per_step": n_sample},
            "gcups": cells / t_tot,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"{n_sample} cfg2 groups per step, oracle/ C++ scalar port of abPOA v1.4.1 "
                                       f"(-M 5 -r 0), {cores} threads; the real abpoa binary is unavailable in the image"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--groups", type=int, default=32768, help="cfg2 groups per GPU per step")
    ap.add_argument("--ref-groups", type=int, default=1024, help="groups per step of the CPU reference arm")
    ap.add_argument("--cpu-sample", type=int, default=3072, help="groups of the cpu_baseline sample")
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
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # every rank owns its own slice of the config (weak scaling); pinned host buffers for the e2e leg
    gro, rbo, bases = make_batch(args.groups, first=rank * args.groups,
                                 workers=max(1, (os.cpu_count() or 1) // max(1, world)))
    pin = [torch.from_numpy(a).pin_memory() for a in (gro, rbo, bases)]
    gro_p, rbo_p, bases_p = [t.numpy() for t in pin]
    n_groups, n_bases = len(gro) - 1, int(rbo[-1])

    ctx = PoaContext(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)

    # ---- resident leg: upload once, time mpoa_batch_run ----
    ctx.upload(gro_p, rbo_p, bases_p)
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
    for _ in range(min(args.warmup, 1)):
        ctx.consensus_batch(packed=(gro_p, rbo_p, bases_p))
    barrier()
    t0 = time.perf_counter()
    e2e_launches = 0
    for _ in range(args.steps):
        res = ctx.consensus_batch(packed=(gro_p, rbo_p, bases_p))
        e2e_launches += res["stats"]["n_kernel_launches"] + 2   # + encode + gather
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d = int(gro_p.nbytes + rbo_p.nbytes + bases_p.nbytes)
    d2h = int(cons_bases + 4 * n_groups + 4 * n_groups)

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
        # rank 0's own kernel: algorithmic bytes per launch / average launch duration (CUDA events in the library)
        per_launch_ms = kernel_ms / max(1, args.steps)
        alg_bytes = algorithmic_bytes(stats, n_bases, cons_bases)
        achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
        traffic = None
        prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(prof):
            try:
                # ncu --set full capture of the same kernel (profiles/README.md): DRAM bytes per band cell
                # x the band cells of THIS launch
                traffic = json.load(open(prof)).get("dram_bytes_per_band_cell") * stats["band_cells"]
            except (OSError, ValueError):
                traffic = None
        # INT-pipe peak measured live: VIADDMNMX.S16x2 warp-instructions/s x 32 lanes x 2 int16 ops
        # (SURVEY.md 8d counts one s16x2 instruction as 2 ops)
        int_peak = ctx.measure_int_peak() * 32 * 2
        int_achieved = stats["int_ops"] / (per_launch_ms * 1e-3)
        cores = os.cpu_count() or 1
        cpu_gps, cpu_gcups, cpu_dt = cpu_port_groups_per_sec(sample_of((gro, rbo, bases), args.cpu_sample), cores)
        value = tot_groups * args.steps / (dev_ms_max * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16x2 (int32 for reads > 6.5 kb)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "groups_per_gpu_per_step": n_groups, "reads": int(len(rbo) - 1),
                       "bases_per_gpu": n_bases, "parallelism": f"groups sharded over {world} GPU(s), no collective",
                       "cache": "inputs + per-step traceback/workspace traffic exceed the 126 MB L2"},
            "gcups": tot_cells * args.steps / (dev_ms_max * 1e-3) / 1e9,
            "gcups_full_matrix": tot_full * args.steps / (dev_ms_max * 1e-3) / 1e9,
            "groups_ok_frac": tot_ok / max(1.0, tot_groups),
            "e2e": {"value": tot_groups * args.steps / (e2e_ms_max * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(tot_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "poa_group_kernel", "kernel_ms_per_launch": per_launch_ms,
                         "note": "integer DP: the ALU/latency bound is int_roofline; HBM carries 1 B/cell"},
            "int_roofline": {"achieved_ops_per_s": int_achieved, "peak_ops_per_s": int_peak,
                             "frac": (int_achieved / int_peak) if int_peak else None,
                             "ops_per_cell": 17, "unit": "int16-lane op/s",
                             "peak_source": "measured live: VIADDMNMX.S16x2 chains (mpoa_measure_int_peak) x 32 lanes x 2"},
            "phase_share": {k: v / max(1, stats["phase_cycles"]["busy"]) for k, v in stats["phase_cycles"].items()},
            "cpu_baseline": {"value": cpu_gps, "unit": UNIT, "cores": cores, "kind": "port", "gcups": cpu_gcups,
                             "sample": f"first {min(args.cpu_sample, n_groups)} groups of the same batch, oracle/ C++ scalar "
                                       f"port of abPOA v1.4.1, {cores} threads, {cpu_dt:.1f} s"},
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
