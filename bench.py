#!/usr/bin/env python
"""bench.py - GCUPS / alignments-per-second of the Gotoh aligner hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--pairs P]

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): P (default 1,000,000) synthetic
251-nt reads vs the 3039-nt HIV-1 HXB2 pol seed through align_it semantics (gip 10, gep 3,
terminal gaps charged), per GPU (weak scaling).  A "step" is one pass of the hot path over
that batch: forward DP + traceback + string emit for every pair.

  value  whole-job GCUPS with inputs resident in HBM (CUDA events on the plan's stream)
  e2e    the same metric through the C-ABI call gotoh_b200_align_batch with pinned HOST
         buffers: host packing/validation, H2D, kernels, D2H all inside the timed region
  roofline   integer-issue roofline of the forward-DP kernel (north_star: "fraction of the
         integer-ALU roofline"): peak = measured thread-instructions/s of the kernel's own
         per-cell instruction mix / instructions per cell; also the HBM view of its direction
         traffic against MEASURED_PEAKS.json
  cpu_baseline  the reference's own gotoh.cpp (oracle/_ref) - or the oracle port if that was
         not built - on all host cores over a bounded sample of the same reads

--impl reference times that CPU path as the reference arm (rank 0 only).
"""
import argparse
import json
import multiprocessing
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "micall-lite_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

GIP, GEP, TERM = 10, 3, 1
WORKLOAD = "C2: synthetic 251-nt reads vs HIV-1 HXB2 pol seed (3039 nt), align_it(ref, read, 10, 3, 1)"


# ------------------------------------------------------------------------------------------
# CPU arm (oracle/_ref = the reference's gotoh.cpp; else the oracle port)
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, ref, qb, qo, lo, hi = args
    from oracle.oracle import Oracle
    from gotoh_b200 import packing
    ora = Oracle(kind)
    rb, ro = packing.pack([ref])
    n = len(qo) - 1
    t0 = time.perf_counter()
    ora.align_batch(0, rb, ro, np.zeros(n, np.int32), qb, qo, GIP, GEP, TERM, first=lo, last=hi)
    return time.perf_counter() - t0


def cpu_kind():
    from oracle import oracle as om
    if om.have_reference():
        return "reference"
    om.build()
    return "port"


def cpu_throughput(ref, qb, qo, cores, kind):
    """Align all packed reads on `cores` processes; returns (gcups, aln_per_s, seconds)."""
    n = len(qo) - 1
    cells = float(np.diff(qo).sum()) * len(ref)
    shards = [(kind, ref, qb, qo, (n * r) // cores, (n * (r + 1)) // cores) for r in range(cores)]
    t0 = time.perf_counter()
    with multiprocessing.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_worker, shards)
    dt = time.perf_counter() - t0
    return cells / dt / 1e9, n / dt, dt


def cpu_sample(seed, cores, per_core=900):
    from gotoh_b200 import workloads
    n = cores * per_core     # ~73 aln/s/core for the reference -> ~12 s
    return workloads.c2_reads_packed(n, seed=seed)


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------
def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except OSError:
        return {"hbm_gbs": 6650.0}, "fallback"


def sass_counts():
    """Per-cell SASS instruction counts of the shipped forward kernel (profiles/sass_counts.json,
    written by tools/sass_count.py from cuobjdump of the built library)."""
    try:
        with open(os.path.join(ROOT, "profiles", "sass_counts.json")) as f:
            return json.load(f)
    except OSError:
        return None


def run_ours(args, rank, world, local_rank):
    import gotoh_b200
    from gotoh_b200 import packing, workloads
    from gotoh_b200.api import Aligner, PinnedArray
    if gotoh_b200.device_count() <= local_rank:
        raise SystemExit("bench.py: CUDA device %d not visible; libgotoh_b200 has no CPU path" % local_rank)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world > 1 and "GOTOH_B200_HOST_THREADS" not in os.environ:
        # ranks share the host: give each rank's packing threads its share of the cores
        os.environ["GOTOH_B200_HOST_THREADS"] = str(max(2, (os.cpu_count() or 8) // world))
    al = Aligner()
    n = args.pairs
    ref, qb, qo = workloads.c2_reads_packed(n, seed=20260101 + rank)
    rb, ro = packing.pack([ref])
    ridx = np.zeros(n, np.int32)

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    # ---- resident arm: inputs in HBM, time plan_run with CUDA events ---------------------------
    plan = al.plan(rb, ro, ridx, qb, qo, GIP, GEP, TERM, gotoh_b200.NT, device=local_rank)
    cells = plan.cells
    for _ in range(args.warmup):
        plan.run()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    dev_ms, fwd_ms = [], []
    for _ in range(args.steps):
        d, f = plan.run()
        dev_ms.append(d)
        fwd_ms.append(f)
    barrier()
    clocks = sampler.stop()
    total_ms = float(sum(dev_ms))
    launches = plan.stat(1) * args.steps
    arena = plan.stat(4)
    path_x2, path_x1, chunks = plan.stat(5), plan.stat(6), plan.stat(7)

    # verify a sample of this very run against the oracle (bit-exact gate)
    out = plan.fetch()
    verified = 0
    if rank == 0 and args.verify > 0:
        from oracle.oracle import Oracle
        ora = Oracle(cpu_kind())
        a = packing.unpack(out[0], plan.out_off, out[2])
        b = packing.unpack(out[1], plan.out_off, out[2])
        step = max(1, n // args.verify)
        reads = None
        for k in range(0, n, step):
            q = qb[qo[k]:qo[k + 1]].tobytes().decode()
            exp = ora.align_it(ref, q, GIP, GEP, TERM)
            if (a[k], b[k], int(out[3][k])) != exp:
                raise SystemExit("bench.py: pair %d differs from the oracle - refusing to report a number" % k)
            verified += 1
        del reads
    plan.close()
    del out

    if args.lite:
        if rank == 0:
            ms = float(sum(dev_ms)) / args.steps
            print(json.dumps({"lite": True, "value": cells / (ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": ms,
                              "forward_ms": float(sum(fwd_ms)) / args.steps, "gpu_launches": launches, "clocks": clocks}))
        if dist is not None:
            dist.destroy_process_group()
        return
    # ---- end-to-end arm: pinned host buffers through gotoh_b200_align_batch --------------------
    out_off = packing.out_offsets(ro, ridx, qo)
    pin = [PinnedArray(al, qb.shape, np.uint8), PinnedArray(al, (int(out_off[-1]),), np.uint8),
           PinnedArray(al, (int(out_off[-1]),), np.uint8), PinnedArray(al, (n,), np.int32), PinnedArray(al, (n,), np.int32)]
    pin[0].array[:] = qb
    outs = (pin[1].array, pin[2].array, pin[3].array, pin[4].array)
    e2e_steps = max(1, args.steps)
    al.align_packed(rb, ro, ridx, pin[0].array, qo, GIP, GEP, TERM, gotoh_b200.NT, out_off=out_off, out=outs,
                    device_mask=1 << local_rank)   # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        al.align_packed(rb, ro, ridx, pin[0].array, qo, GIP, GEP, TERM, gotoh_b200.NT, out_off=out_off, out=outs,
                        device_mask=1 << local_rank)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    barrier()
    h2d = int(qb.nbytes + len(ref) + 2 * 64 * 2 + n * 72)     # queries + ref (+class copy) + pair/task records
    d2h = int(2 * out_off[-1] + 8 * n)
    for p_ in pin:
        p_.free()

    # ---- reduce over ranks: max time, sum of cells ---------------------------------------------
    if dist is not None:
        import torch
        t = torch.tensor([total_ms, e2e_s, max(fwd_ms) if fwd_ms else 0.0, sum(fwd_ms)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_s = float(t[0]), float(t[1])
        c = torch.tensor([float(cells), float(n)], device="cuda", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        cells_all, n_all = float(c[0]), float(c[1])
    else:
        cells_all, n_all = float(cells), float(n)
    ms_per_step = total_ms / args.steps
    value = cells_all / (ms_per_step * 1e-3) / 1e9
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (forward DP) -------------------------------------------
    fwd_ms_step = float(sum(fwd_ms)) / args.steps
    fwd_gcups = cells / (fwd_ms_step * 1e-3) / 1e9 if fwd_ms_step > 0 else None
    peaks, peaks_src = measured_peaks()
    sc = sass_counts()
    mix = {name: al.int_peak(w, local_rank) for name, w in
           (("iadd", 0), ("viaddmnmx", 2), ("viaddmnmx_s16x2", 3), ("vimnmx3", 4), ("vimnmx3_s16x2", 12), ("imad", 5),
            ("lop3", 6), ("alu_plus_imad", 7), ("cell_mix_s16x2", 8), ("cell_mix_s32", 9), ("viadd_16x2", 10))}
    use_x2 = path_x2 >= path_x1
    tag = "x2" if use_x2 else "x1"
    alu_rate = mix["viaddmnmx_s16x2" if use_x2 else "viaddmnmx"]     # G thread-instr/s, ALU pipe alone
    issue_rate = mix["alu_plus_imad"]                                # G thread-instr/s, ALU + FMA pipes together
    if sc:
        instr_per_cell, alu_per_cell = sc["instr_per_cell_" + tag], sc["alu_per_cell_" + tag]
    else:   # hand count of the steady-state block; profiles/sass_counts.json supersedes it
        instr_per_cell, alu_per_cell = (4.77, 3.09) if use_x2 else (9.47, 6.47)
    # integer roofline: cells/s at peak INT32 issue / instructions per cell, for the binding constraint
    roof_alu, roof_issue = alu_rate / alu_per_cell, issue_rate / instr_per_cell
    roof_gcups = min(roof_alu, roof_issue)
    dir_bytes = 0.25 * cells * (1.03)                 # 2 bits/cell + wavefront fill/drain slots
    roofline = {
        # integer roofline (north_star): which of its two terms binds - the ALU pipe or the ALU+FMA issue slots
        "bound": "int_alu" if roof_alu <= roof_issue else "int_issue", "kernel": "k_forward<%s,8,false>" % ("Vec16" if use_x2 else "Vec32"),
        "achieved": fwd_gcups, "peak": roof_gcups, "unit": "GCUPS", "frac": (fwd_gcups / roof_gcups) if fwd_gcups else None,
        "peak_def": "min(measured ALU-pipe rate %.0f G instr/s / %.2f ALU instr per cell, measured ALU+FMA issue rate %.0f / %.2f instr per cell); "
                    "rates from gotoh_b200_int_peak in this run, counts from cuobjdump (profiles/sass_counts.json)" % (alu_rate, alu_per_cell, issue_rate, instr_per_cell),
        "instr_per_cell": instr_per_cell, "alu_instr_per_cell": alu_per_cell, "roof_alu_pipe": roof_alu, "roof_issue": roof_issue,
        "issue_peaks_ginstr_s": mix,
        "avg_launch_ms": fwd_ms_step / max(1, chunks),
        "hbm": {"bound": "hbm", "achieved": dir_bytes / (fwd_ms_step * 1e-3) / 1e9 if fwd_ms_step else None,
                "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "peak_src": peaks_src,
                "frac": (dir_bytes / (fwd_ms_step * 1e-3) / 1e9 / peaks["hbm_gbs"]) if fwd_ms_step else None,
                "algorithmic_bytes_per_cell": 0.25},
        # dram bytes of one forward launch: the per-cell figure of the committed ncu --set full capture
        # (profiles/sass_counts.json: ncu_dram_bytes_per_cell, ncu_note) x the cells one launch of this run covers
        "traffic": ((sc or {}).get("ncu_dram_bytes_per_cell") or 0) * cells / max(1, chunks) or None,
        "algorithmic_bytes_per_launch": dir_bytes / max(1, chunks),
    }

    # ---- CPU baseline on the host cores (bounded sample) -----------------------------------------
    cores = os.cpu_count() or 1
    kind = cpu_kind()
    sref, sqb, sqo = cpu_sample(20260101, cores)
    cg, ca, cs = cpu_throughput(sref, sqb, sqo, cores, kind)
    line = {
        "metric": "GCUPS", "value": value, "unit": "GCUPS", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int16x2" if use_x2 else "int32", "data": "synthetic",
        "alignments_per_s": n_all / (ms_per_step * 1e-3),
        "config": {"workload": WORKLOAD, "pairs_per_gpu": n, "cells_per_gpu": cells, "gip": GIP, "gep": GEP, "term": TERM,
                   "parallelism": "static shard of independent pairs, %d rank(s), no collective" % world,
                   "l2": "working set per step (%.1f GB direction arena written + read back) >> 126 MB L2; no flush needed" % (arena / 1e9),
                   "arena_chunks_per_step": chunks, "pairs_int16x2": path_x2, "pairs_int32": path_x1},
        "bit_exact_verified_pairs": verified,
        "clocks": clocks,
        "e2e": {"value": cells_all / e2e_s / 1e9, "unit": "GCUPS", "alignments_per_s": n_all / e2e_s, "s_per_step": e2e_s,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "api": "gotoh_b200_align_batch (pinned host buffers)"},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": {"value": cg, "unit": "GCUPS", "alignments_per_s": ca, "cores": cores, "kind": kind,
                         "sample": "%d C2 reads (same generator/seed) over %d processes, %.1f s" % (len(sqo) - 1, cores, cs),
                         "gcups_per_core": cg / cores},
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU aligner on all host cores, bounded sample per step."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    kind = cpu_kind()
    per_core = max(50, 300 // max(1, (args.steps + args.warmup) // 4))
    ref, qb, qo = cpu_sample(20260101, cores, per_core=per_core)
    times = []
    gc = al_s = 0.0
    for it in range(args.warmup + args.steps):
        g, a, dt = cpu_throughput(ref, qb, qo, cores, kind)
        if it >= args.warmup:
            times.append(dt)
    n = len(qo) - 1
    cells = float(np.diff(qo).sum()) * len(ref)
    dt = sum(times) / len(times)
    gc, al_s = cells / dt / 1e9, n / dt
    line = {
        "impl": "reference", "metric": "GCUPS", "value": gc, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic", "alignments_per_s": al_s,
        "config": {"workload": WORKLOAD, "pairs_per_step": n, "gip": GIP, "gep": GEP, "term": TERM},
        "cpu_baseline": {"value": gc, "unit": "GCUPS", "cores": cores, "kind": kind,
                         "sample": "%d C2 reads per step over %d processes (bounded sample of the 1M-read workload)" % (n, cores)},
        "e2e": {"value": gc, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1000000, help="pairs per GPU per step (C2 = 1,000,000)")
    ap.add_argument("--verify", type=int, default=200, help="pairs of the timed run checked against the oracle")
    ap.add_argument("--lite", action="store_true", help="profiling runs: skip e2e, microbenchmarks and the CPU baseline")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1 and "RANK" not in os.environ:
        # convenience: re-launch under torchrun exactly like the driver does
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
