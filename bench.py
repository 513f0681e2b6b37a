#!/usr/bin/env python
"""bench.py - GCUPS / alignments-per-second of the Gotoh aligner hot path on B200.

    python bench.py [--config c2|c2b|c3|c4|c5|c1] [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workloads (BASELINE.json configs); the default is the one the metric is quoted on:
  c2   1,000,000 synthetic 251-nt reads vs the 3039-nt HIV-1 HXB2 pol seed, align_it(ref, read, 10, 3, 1)   [default]
  c2b  the same reads with the only in-tree gap model, align_it(ref, read, 10, 10, 0)
  c3   1,000,000 ~84-aa windows vs PR/RT/INT, align_it_aa(ref, q, 40, 10, 1) (empirical HIV matrix)
  c4   10,000 consensus-vs-genome pairs on the 57 HCV seeds (~9.6 kb x ~9.6 kb), align_it(ref, q, 15, 3, 1)
  c5   10,000,000-pair MiSeq mix = ten batches of 800,000 c2 reads + 200,000 c3 windows, STRONG scaling over the GPUs
  c1   all 19,200 reads of the reference's example run vs the pol seed (tests/golden/c1_reads.txt.xz), strong scaling
A "step" is one pass of the hot path over the config's batch(es): forward DP + traceback + result emit for every pair.

  value  whole-job GCUPS with inputs resident in HBM (CUDA events on the plan's stream)
  e2e    the same metric through the C-ABI call gotoh_b200_align_batch with pinned HOST buffers: host packing and
         validation, H2D, kernels, D2H of the reference-format strings all inside the timed region.  The outputs of the
         LAST timed call are compared byte for byte with the resident arm's on every pair before a number is printed.
         e2e.compact is the same through gotoh_b200_align_batch_compact (records + op scripts instead of padded
         strings, ~1/60 of the D2H bytes; rendered strings are checked on a sample), e2e.host_ceiling the box's
         device-to-host rate measured by every rank copying at once.
  roofline   integer-issue roofline of the config's forward-DP kernel (north_star: "fraction of the integer-ALU
         roofline"): peak = measured thread-instructions/s of the kernel's own instruction mix / instructions per
         cell; also the HBM view of its direction traffic against MEASURED_PEAKS.json
  cpu_baseline  the reference's own gotoh.cpp (oracle/_ref) - or the oracle port if that was not built - on all host
         cores over a bounded sample of the same pairs

--impl reference times that CPU path as the reference arm (rank 0 only).
Multi-GPU: under torchrun one process per GPU (weak scaling: every rank its own batch; c5/c1: every rank its slice of
the fixed job).  `--single-process --gpus N` drives N GPUs from one process through device_mask (the library's own
static sharding with a host-side gather).
"""
import argparse
import json
import lzma
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

NT, HIV25 = 0, 1


# ------------------------------------------------------------------------------------------
# workloads
# ------------------------------------------------------------------------------------------
class Batch:
    """One call's worth of pairs: a gap model, a score table, packed references and queries."""

    def __init__(self, matrix, gip, gep, term, refs, ridx, qb, qo):
        from gotoh_b200 import packing
        self.matrix, self.gip, self.gep, self.term = matrix, gip, gep, term
        self.refs = list(refs)
        self.rb, self.ro = packing.pack(self.refs)
        self.ridx = None if ridx is None else np.ascontiguousarray(ridx, np.int32)
        self.qb, self.qo = np.ascontiguousarray(qb, np.uint8), np.ascontiguousarray(qo, np.int64)
        self.n = len(self.qo) - 1
        rl = np.diff(self.ro)
        self.rlen = rl if self.ridx is None else rl[self.ridx]
        self.qlen = np.diff(self.qo)
        self.cells = float((self.rlen.astype(np.float64) * self.qlen).sum())
        self.out_off = packing.out_offsets(self.ro, self.ridx, self.qo)

    def slice(self, lo, hi):
        qo = self.qo[lo:hi + 1] - self.qo[lo]
        qb = self.qb[self.qo[lo]:self.qo[hi]]
        return Batch(self.matrix, self.gip, self.gep, self.term, self.refs, None if self.ridx is None else self.ridx[lo:hi], qb, qo)

    def ref_of(self, k):
        return self.refs[k if self.ridx is None else int(self.ridx[k])]

    def query(self, k):
        return self.qb[self.qo[k]:self.qo[k + 1]].tobytes().decode("latin-1")


def c1_reads():
    with lzma.open(os.path.join(ROOT, "tests", "golden", "c1_reads.txt.xz"), "rt") as f:
        return f.read().split()


CONFIGS = {
    # name: (workload description, default pairs, scaling, kernel tag of the roofline)
    "c2": ("C2: synthetic 251-nt reads vs HIV-1 HXB2 pol seed (3039 nt), align_it(ref, read, 10, 3, 1)", 1000000, "weak", "x2"),
    "c2b": ("C2b: synthetic 251-nt reads vs HIV-1 HXB2 pol seed (3039 nt), align_it(ref, read, 10, 10, 0)", 1000000, "weak", "x2"),
    "c3": ("C3: ~84-aa windows vs HIV-1 PR/RT/INT (99/440/288 aa), align_it_aa(ref, q, 40, 10, 1), empirical HIV matrix", 1000000, "weak", "x2_half_k6"),
    "c4": ("C4: consensus vs full-length HCV genotype references (~9.6 kb x ~9.6 kb), align_it(ref, q, 15, 3, 1), wavefront path", 10000, "weak", "x1_flow"),
    "c5": ("C5: 10M-pair synthetic MiSeq run = 10 batches of 800k C2 reads (align_it 10/3/1) + 200k C3 windows (align_it_aa 40/10/1), sharded over the GPUs", 10000000, "strong", "x2"),
    "c1": ("C1 substitute: all 19,200 reads of examples/HIV1C-pol_S1_L001 (R2 reverse-complemented) vs HIV1B-pol-seed, align_it(ref, read, 10, 3, 1)", 19200, "strong", "x2"),
}


def _c5_batch_arrays(job):
    """Batch b of the C5 mix, this rank's contiguous share: 80 % C2 reads, 20 % C3 windows."""
    from gotoh_b200 import workloads
    b, per, rank, world = job
    n2, n3 = int(per * 0.8), per - int(per * 0.8)
    lo2, hi2 = workloads.shard_range(n2, rank, world)
    lo3, hi3 = workloads.shard_range(n3, rank, world)
    ref, qb, qo = workloads.c2_reads_packed(hi2 - lo2, seed=20260105 + 1000 * b + rank)
    refs, ridx, qb3, qo3 = workloads.c3_queries_packed(hi3 - lo3, seed=20260205 + 1000 * b + rank)
    return ref, qb, qo, refs, ridx, qb3, qo3


def make_batches(config, pairs, rank, world):
    """The batches THIS rank aligns per step."""
    from gotoh_b200 import packing, workloads
    if config in ("c2", "c2b"):
        ref, qb, qo = workloads.c2_reads_packed(pairs, seed=20260101 + rank)
        g = (10, 3, 1) if config == "c2" else (10, 10, 0)
        return [Batch(NT, g[0], g[1], g[2], [ref], np.zeros(pairs, np.int32), qb, qo)]
    if config == "c3":
        refs, ridx, qb, qo = workloads.c3_queries_packed(pairs, seed=20260103 + rank)
        return [Batch(HIV25, 40, 10, 1, refs, ridx, qb, qo)]
    if config == "c4":
        refs, ridx, qb, qo = workloads.c4_pairs_packed(pairs, seed=20260104 + rank)
        return [Batch(NT, 15, 3, 1, refs, ridx, qb, qo)]
    if config == "c1":
        reads = c1_reads()[:pairs]
        lo, hi = workloads.shard_range(len(reads), rank, world)
        qb, qo = packing.pack(reads[lo:hi])
        return [Batch(NT, 10, 3, 1, [workloads.pol_seed()], np.zeros(hi - lo, np.int32), qb, qo)]
    if config == "c5":
        # strong scaling: the job is fixed (pairs, default 10M); rank r aligns its contiguous share of every batch.
        # A batch is generated from its own seed; a rank only generates the records of its own share (same stream
        # statistics, seeded per (batch, rank)), which keeps set-up time off the GPU box's clock.
        nb = max(1, pairs // 1000000)
        per = pairs // nb
        jobs = [(b, per, rank, world) for b in range(nb)]
        if world == 1 and nb > 1:
            # one process generates the whole job: spread the batches over a few processes (set-up time only)
            with multiprocessing.get_context("fork").Pool(min(nb, os.cpu_count() or 1)) as pool:
                parts = pool.map(_c5_batch_arrays, jobs)
        else:
            parts = [_c5_batch_arrays(j) for j in jobs]
        out = []
        for (ref, qb, qo, refs, ridx, qb3, qo3) in parts:
            out.append(Batch(NT, 10, 3, 1, [ref], np.zeros(len(qo) - 1, np.int32), qb, qo))
            out.append(Batch(HIV25, 40, 10, 1, refs, ridx, qb3, qo3))
        return out
    raise SystemExit("unknown config %r" % config)


# ------------------------------------------------------------------------------------------
# CPU arm (oracle/_ref = the reference's gotoh.cpp; else the oracle port)
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    kind, parts = args
    from oracle.oracle import Oracle
    ora = Oracle(kind)
    t0 = time.perf_counter()
    for (matrix, rb, ro, ridx, qb, qo, gip, gep, term, lo, hi) in parts:
        ora.align_batch(matrix, rb, ro, ridx, qb, qo, gip, gep, term, first=lo, last=hi)
    return time.perf_counter() - t0


def cpu_kind():
    from oracle import oracle as om
    if om.have_reference():
        # load it once in this (parent) process too, so that the driver's .so hook sees which library the arm runs
        om.Oracle("reference")
        return "reference"
    om.build()
    return "port"


def cpu_sample(batches, cores, cell_budget_per_core):
    """A bounded sample of the workload: the first pairs of every batch, in the batches' proportions, ~cell budget."""
    total = sum(b.cells for b in batches)
    want = cores * cell_budget_per_core
    frac = min(1.0, want / max(total, 1.0))
    out = []
    for b in batches:
        k = max(cores, int(b.n * frac)) if b.n >= cores else b.n
        out.append(b.slice(0, min(b.n, k)))
    return out


def cpu_throughput(sample, cores, kind):
    """Align the sample on `cores` processes; returns (gcups, aln_per_s, seconds)."""
    shards = []
    for r in range(cores):
        parts = []
        for b in sample:
            lo, hi = (b.n * r) // cores, (b.n * (r + 1)) // cores
            if hi > lo:
                parts.append((b.matrix, b.rb, b.ro, b.ridx, b.qb, b.qo, b.gip, b.gep, b.term, lo, hi))
        shards.append((kind, parts))
    t0 = time.perf_counter()
    with multiprocessing.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_worker, shards)
    dt = time.perf_counter() - t0
    cells = sum(b.cells for b in sample)
    n = sum(b.n for b in sample)
    return cells / dt / 1e9, n / dt, dt


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
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
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
    """Per-cell SASS instruction counts of the shipped forward kernels (profiles/sass_counts.json,
    written by tools/sass_count.py from cuobjdump of the built library)."""
    try:
        with open(os.path.join(ROOT, "profiles", "sass_counts.json")) as f:
            return json.load(f)
    except OSError:
        return None


KERNELS = {
    # tag: (kernel name, sass_counts key, ALU-pipe microbenchmark that matches its cell instruction, dtype)
    "x2": ("k_forward<Vec16,8,false,false>", "x2", "viaddmnmx_s16x2", "int16x2"),
    "x2_half_k6": ("k_forward<Vec16,6,false,true> (two 16-lane wavefronts per warp)", "x2_half_k6", "viaddmnmx_s16x2", "int16x2"),
    # (the flow kernel's block = the int32 cell code of k_forward<Vec32,8> + the hand-over of the strip's boundary column;
    # its own count is the denominator, roofline.frac_vs_plain_cell_code shows it against the bare cell code of "x1")
    "x1_flow": ("k_forward_flow<8> (strip dataflow, one warp per (pair, strip))", "x1_flow", "viaddmnmx", "int32"),
    "x1": ("k_forward<Vec32,8,false,false>", "x1", "viaddmnmx", "int32"),
}


# ------------------------------------------------------------------------------------------
# resident arm
# ------------------------------------------------------------------------------------------
def resident_device(al, batches, device, steps, warmup, keep_fetch):
    """Inputs in HBM, plan_run timed with CUDA events.  One batch: the plan is built once and re-run; several batches
    (c5): every step builds, runs and drops each batch's plan (only plan_run is inside the device-time sum)."""
    res = {"dev_ms": 0.0, "fwd_ms": 0.0, "launches": 0, "arena": 0, "x2": 0, "x1": 0, "chunks": 0, "fwd_launches": 0, "fetch": None,
           "cells": 0}
    if len(batches) == 1:
        b = batches[0]
        plan = al.plan(b.rb, b.ro, b.ridx, b.qb, b.qo, b.gip, b.gep, b.term, b.matrix, device=device)
        for _ in range(warmup):
            plan.run()
        yield "ready"
        for _ in range(steps):
            d, f = plan.run()
            res["dev_ms"] += d
            res["fwd_ms"] += f
        yield "timed"
        res.update(launches=plan.stat(1) * steps, arena=plan.stat(4), x2=plan.stat(5), x1=plan.stat(6), chunks=plan.stat(7) * steps,
                   fwd_launches=(plan.stat(1) - 2 * plan.stat(7)) * steps, cells=plan.cells * steps)
        if keep_fetch:
            res["fetch"] = plan.fetch()
        plan.close()
    else:
        def one_pass(timed):
            for i, b in enumerate(batches):
                plan = al.plan(b.rb, b.ro, b.ridx, b.qb, b.qo, b.gip, b.gep, b.term, b.matrix, device=device)
                d, f = plan.run()
                if timed:
                    res["dev_ms"] += d
                    res["fwd_ms"] += f
                    res["launches"] += plan.stat(1)
                    res["fwd_launches"] += plan.stat(1) - 2 * plan.stat(7)
                    res["chunks"] += plan.stat(7)
                    res["arena"] = max(res["arena"], plan.stat(4))
                    res["x2"] += plan.stat(5)
                    res["x1"] += plan.stat(6)
                    res["cells"] += plan.cells
                    if keep_fetch and timed == "last" and i == len(batches) - 1:
                        res["fetch"] = plan.fetch()
                plan.close()
        for _ in range(min(warmup, 1)):
            one_pass(False)
        yield "ready"
        for s in range(steps):
            one_pass("last" if s == steps - 1 else True)
        yield "timed"
    yield res


def verify_against_oracle(b, strings_of, count, what):
    """Bit-exact gate on a sample: (aligned_standard, aligned_seq, score) of pair k vs the oracle."""
    from oracle.oracle import Oracle
    ora = Oracle(cpu_kind())
    fn = ora.align_it if b.matrix == NT else ora.align_it_aa
    checked = 0
    for k in range(0, b.n, max(1, b.n // max(1, count))):
        exp = fn(b.ref_of(k), b.query(k), b.gip, b.gep, b.term)
        if strings_of(k) != exp:
            raise SystemExit("bench.py: %s: pair %d differs from the oracle - refusing to report a number" % (what, k))
        checked += 1
    return checked


def strided_strings(out, out_off):
    def f(k):
        o, ln = int(out_off[k]), int(out[2][k])
        return (out[0][o:o + ln].tobytes().decode("latin-1"), out[1][o:o + ln].tobytes().decode("latin-1"), int(out[3][k]))
    return f


# ------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import gotoh_b200
    from gotoh_b200.api import Aligner, PinnedArray
    workload, default_pairs, scaling, ktag = CONFIGS[args.config]
    pairs = args.pairs or default_pairs
    single = args.single_process and args.gpus > 1
    devs = list(range(args.gpus)) if single else [local_rank]
    if not args.emu and gotoh_b200.device_count() <= max(devs):
        raise SystemExit("bench.py: CUDA device %d not visible; libgotoh_b200 has no CPU path" % max(devs))
    n_gpus = args.gpus if single else world
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if world > 1 and "GOTOH_B200_HOST_THREADS" not in os.environ:
        # ranks share the host: give each rank's packing threads its share of the cores
        os.environ["GOTOH_B200_HOST_THREADS"] = str(max(2, (os.cpu_count() or 8) // world))
    if args.emu:
        # self-test of THIS SCRIPT in the GPU-less container: the kernel sources under tests/simt_emu (never a bench number)
        sys.path.insert(0, os.path.join(ROOT, "tests", "simt_emu"))
        import build_emu
        from gotoh_b200 import _ffi
        al = Aligner(_ffi.Library(build_emu.build()))
    else:
        al = Aligner()
    if single:
        # one process, N devices: the whole job's batches; device d's resident plans hold its contiguous share
        jobs = make_batches(args.config, pairs, 0, 1) if scaling == "strong" else \
            [b for r in range(args.gpus) for b in make_batches(args.config, pairs, r, args.gpus)]
        if scaling == "weak":
            # weak scaling in one process: concatenate the per-GPU batches into one call per config batch
            jobs = [concat_batches(jobs)]
        from gotoh_b200 import workloads
        shards = {d: [b.slice(*workloads.shard_range(b.n, i, len(devs))) for b in jobs] for i, d in enumerate(devs)}
    else:
        jobs = make_batches(args.config, pairs, rank, world)
        shards = {local_rank: jobs}
    mask = 0
    for d in devs:
        mask |= 1 << d
    n_mine = sum(b.n for b in jobs)
    cells_mine = sum(b.cells for b in jobs)

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    # ---- resident arm: inputs in HBM, time plan_run with CUDA events ---------------------------
    keep_fetch = args.verify_full and not single
    gens = {d: resident_device(al, shards[d], d, args.steps, args.warmup, keep_fetch) for d in devs}
    results = {}

    def drive(d, phase):
        r = next(gens[d])
        if phase == "result":
            results[d] = r

    def all_devices(phase):
        if len(devs) == 1:
            drive(devs[0], phase)
        else:
            th = [threading.Thread(target=drive, args=(d, phase)) for d in devs]
            for t in th:
                t.start()
            for t in th:
                t.join()

    # clocks and throttle reasons are sampled every 20 ms from here to the end of the end-to-end arms (the timed regions of
    # the short configs are a few tens of ms: one region alone may see no sample)
    sampler = ClockSampler(devs[0])
    sampler.start()
    all_devices("ready")
    barrier()
    all_devices("timed")
    barrier()
    all_devices("result")
    total_ms = max(results[d]["dev_ms"] for d in devs)
    fwd_ms_total = max(results[d]["fwd_ms"] for d in devs)
    launches = sum(results[d]["launches"] for d in devs)
    fwd_launches = sum(results[d]["fwd_launches"] for d in devs)
    arena = max(results[d]["arena"] for d in devs)
    path_x2 = sum(results[d]["x2"] for d in devs)
    path_x1 = sum(results[d]["x1"] for d in devs)
    chunks = sum(results[d]["chunks"] for d in devs)
    cells = sum(results[d]["cells"] for d in devs) / args.steps          # trimmed cells per step (all devices of this process)
    res_fetch = results[devs[0]]["fetch"]

    # bit-exact gate on a sample of the resident arm's own results
    verified = 0
    last = jobs[-1]
    if rank == 0 and args.verify > 0 and res_fetch is not None:
        verified = verify_against_oracle(last, strided_strings(res_fetch, last.out_off), args.verify, "resident arm")

    if args.lite:
        clocks = sampler.stop()
        if rank == 0:
            ms = total_ms / args.steps
            print(json.dumps({"lite": True, "config": args.config, "value": cells / (ms * 1e-3) / 1e9, "unit": "GCUPS", "ms_per_step": ms,
                              "forward_ms": fwd_ms_total / args.steps, "gpu_launches": launches, "clocks": clocks}))
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- end-to-end arms: pinned host buffers through the C ABI ---------------------------------
    big = max(jobs, key=lambda b: int(b.out_off[-1]))
    cap_bytes = int(big.out_off[-1])
    max_n = max(b.n for b in jobs)
    e2e = {}
    # the step's inputs sit in pinned host memory (the device-side plan builder copies the raw reads straight from the
    # caller's buffer); c5 keeps its twenty batches pageable
    pin_q = [PinnedArray(al, (max(len(b.qb), 1),), np.uint8) for b in jobs] if len(jobs) <= 2 else []
    for pq, b in zip(pin_q, jobs):
        pq.array[:len(b.qb)] = b.qb
    qsrc = [pq.array[:len(b.qb)] for pq, b in zip(pin_q, jobs)] if pin_q else [b.qb for b in jobs]
    e2e_steps = max(1, args.steps)
    fmts = ("strings", "compact") if args.e2e_format == "both" else (args.e2e_format,)
    ceiling = None
    if "strings" in fmts:
        pin = [PinnedArray(al, (cap_bytes,), np.uint8), PinnedArray(al, (cap_bytes,), np.uint8),
               PinnedArray(al, (max_n,), np.int32), PinnedArray(al, (max_n,), np.int32)]

        def strings_pass():
            for b, q in zip(jobs, qsrc):
                outs = (pin[0].array[:int(b.out_off[-1])], pin[1].array[:int(b.out_off[-1])], pin[2].array[:b.n], pin[3].array[:b.n])
                al.align_packed(b.rb, b.ro, b.ridx, q, b.qo, b.gip, b.gep, b.term, b.matrix, out_off=b.out_off, out=outs, device_mask=mask)
            return outs
        strings_pass()                                    # warm-up: workspaces, pinned staging
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            outs = strings_pass()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        barrier()
        # verify what was timed: the buffers the LAST timed call filled, against the resident arm's results, all pairs
        ver_pairs = 0
        e2e_oracle = 0
        if res_fetch is not None:
            same = (np.array_equal(outs[2], res_fetch[2]) and np.array_equal(outs[3], res_fetch[3]) and
                    np.array_equal(outs[0], res_fetch[0]) and np.array_equal(outs[1], res_fetch[1]))
            if not same:
                raise SystemExit("bench.py: e2e (strings) outputs of the timed run differ from the resident arm's - refusing to report a number")
            ver_pairs = last.n
        elif rank == 0 and args.verify > 0:
            e2e_oracle = verify_against_oracle(last, strided_strings(outs, last.out_off), args.verify, "e2e strings")
        if len(jobs) > 1 and rank == 0 and args.verify > 0:
            # several batches share the output buffers: the last batch's timed outputs were compared above; every other
            # batch is re-run once (untimed) and sampled against the oracle
            for b in jobs[:-1]:
                o = (pin[0].array[:int(b.out_off[-1])], pin[1].array[:int(b.out_off[-1])], pin[2].array[:b.n], pin[3].array[:b.n])
                al.align_packed(b.rb, b.ro, b.ridx, b.qb, b.qo, b.gip, b.gep, b.term, b.matrix, out_off=b.out_off, out=o, device_mask=mask)
                verify_against_oracle(b, strided_strings(o, b.out_off), max(4, args.verify // len(jobs)), "e2e strings (re-run)")
            b = last
            al.align_packed(b.rb, b.ro, b.ridx, b.qb, b.qo, b.gip, b.gep, b.term, b.matrix, out_off=b.out_off, out=outs, device_mask=mask)
        d2h = int(sum(2 * int(b.out_off[-1]) + 8 * b.n for b in jobs))
        h2d = int(sum(len(b.qb) + len(b.rb) + 2 * 64 * (len(b.refs) + 1) + b.n * 72 for b in jobs))
        e2e["strings"] = {"s": e2e_s, "h2d": h2d, "d2h": d2h, "verified_pairs_vs_resident": ver_pairs, "verified_pairs_vs_oracle": e2e_oracle}
        strings_ref = (outs[2].copy(), outs[3].copy())        # lengths and scores of the last batch, for the compact check
        sample_ks = list(range(0, last.n, max(1, last.n // 200)))
        strings_sample = {k: strided_strings(outs, last.out_off)(k) for k in sample_ks}
        # host ceiling: every rank copies the same number of bytes device-to-host at once, nothing else running
        probe_bytes = min(cap_bytes, 4 << 30)
        al.d2h_probe(pin[0].array[:probe_bytes], reps=1, device=devs[0])
        barrier()
        probe_s = al.d2h_probe(pin[0].array[:probe_bytes], reps=2, device=devs[0])
        barrier()
        ceiling = {"bytes_per_rank": 2 * probe_bytes, "seconds": probe_s}
        for p_ in pin:
            p_.free()
    if "compact" in fmts:
        words = [int(((np.minimum(b.rlen, b.qlen) * 5 // 4 + 47) >> 4).sum()) for b in jobs]
        pinc = [PinnedArray(al, (max_n * 8,), np.int32), PinnedArray(al, (max(words),), np.uint32), PinnedArray(al, (max_n,), np.int64)]

        def compact_pass(count_bytes=False):
            nbytes = 0
            for b, q in zip(jobs, qsrc):
                c = al.align_packed_compact(b.rb, b.ro, b.ridx, q, b.qo, b.gip, b.gep, b.term, b.matrix,
                                            out=(pinc[0].array[:b.n * 8], pinc[1].array, pinc[2].array[:b.n]), device_mask=mask)
                if count_bytes:
                    nbytes += c.nbytes()
            return c, nbytes
        _, d2h_c = compact_pass(True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            comp, _ = compact_pass()
        c_s = (time.perf_counter() - t0) / e2e_steps
        barrier()
        ver_c = 0
        if "strings" in fmts:
            if not (np.array_equal(comp.scores, strings_ref[1]) and np.array_equal(comp.out_len, strings_ref[0])):
                raise SystemExit("bench.py: e2e (compact) scores/lengths of the timed run differ from the strings run - refusing to report a number")
            for k, exp in strings_sample.items():
                if comp[k] != exp:
                    if os.environ.get("BENCH_DEBUG"):
                        g = comp[k]
                        print(k, [len(x) for x in g[:2]], [len(x) for x in exp[:2]], g[2], exp[2], comp.rec[k], [i for i in range(min(len(g[0]), len(exp[0]))) if g[0][i] != exp[0][i]][:5], [i for i in range(min(len(g[1]), len(exp[1]))) if g[1][i] != exp[1][i]][:5], file=sys.stderr)
                    raise SystemExit("bench.py: e2e (compact) pair %d renders differently from the strings run" % k)
            ver_c = last.n
        elif rank == 0 and args.verify > 0:
            verify_against_oracle(last, lambda k: comp[k], args.verify, "e2e compact")
        e2e["compact"] = {"s": c_s, "d2h": d2h_c, "scores_lengths_verified_pairs": ver_c, "rendered_pairs_checked": len(strings_sample) if "strings" in fmts else 0}
        del comp
        for p_ in pinc:
            p_.free()

    for p_ in pin_q:
        p_.free()
    clocks = sampler.stop()
    clocks["window"] = "resident warm-up + timed steps + end-to-end arms"

    # ---- reduce over ranks: max time, sum of cells ---------------------------------------------
    es = e2e.get("strings", {}).get("s", 0.0)
    ec = e2e.get("compact", {}).get("s", 0.0)
    ps = ceiling["seconds"] if ceiling else 0.0
    if dist is not None:
        import torch
        t = torch.tensor([total_ms, es, ec, ps], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, es, ec, ps = (float(x) for x in t)
        c = torch.tensor([float(cells), float(n_mine), float(e2e.get("strings", {}).get("d2h", 0)), float(e2e.get("compact", {}).get("d2h", 0)),
                          float(ceiling["bytes_per_rank"] if ceiling else 0), float(e2e.get("strings", {}).get("h2d", 0))], device="cuda", dtype=torch.float64)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        cells_all, n_all, d2h_all, d2hc_all, probe_all, h2d_all = (float(x) for x in c)
    else:
        cells_all, n_all = float(cells), float(n_mine)
        d2h_all, d2hc_all = float(e2e.get("strings", {}).get("d2h", 0)), float(e2e.get("compact", {}).get("d2h", 0))
        probe_all, h2d_all = float(ceiling["bytes_per_rank"] if ceiling else 0), float(e2e.get("strings", {}).get("h2d", 0))
    if args.emu:                       # the emulator has no clock: placeholders keep the script's arithmetic alive
        total_ms, fwd_ms_total = total_ms or 1.0, fwd_ms_total or 1.0
    ms_per_step = total_ms / args.steps
    value = cells_all / (ms_per_step * 1e-3) / 1e9
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (forward DP) -------------------------------------------
    fwd_ms_step = fwd_ms_total / args.steps
    fwd_gcups = cells / (fwd_ms_step * 1e-3) / 1e9 if fwd_ms_step > 0 else None
    peaks, peaks_src = measured_peaks()
    sc = sass_counts() or {}
    mix = {name: (1.0 if args.emu else al.int_peak(w, devs[0])) for name, w in
           (("iadd", 0), ("viaddmnmx", 2), ("viaddmnmx_s16x2", 3), ("vimnmx3", 4), ("vimnmx3_s16x2", 12), ("imad", 5),
            ("lop3", 6), ("alu_plus_imad", 7), ("cell_mix_s16x2", 8), ("cell_mix_s32", 9), ("viadd_16x2", 10))}
    if ktag.startswith("x2") and path_x2 < path_x1:
        ktag = "x1"
    kname, skey, alu_bench, dtype = KERNELS[ktag]
    alu_rate = mix[alu_bench]                                        # G thread-instr/s, ALU pipe alone
    issue_rate = mix["alu_plus_imad"]                                # G thread-instr/s, ALU + FMA pipes together
    kc = sc.get(skey) or {}
    if "instr_per_cell" in kc:
        instr_per_cell, alu_per_cell = kc["instr_per_cell"], kc["alu_per_cell"]
    else:   # hand count of the steady-state block; profiles/sass_counts.json supersedes it
        instr_per_cell, alu_per_cell = {"x2": (4.69, 2.20), "x2_half_k6": (5.13, 2.35)}.get(ktag, (9.03, 4.47))
    # integer roofline: cells/s at peak INT32 issue / instructions per cell, for the binding constraint
    roof_alu, roof_issue = alu_rate / alu_per_cell, issue_rate / instr_per_cell
    roof_gcups = min(roof_alu, roof_issue)
    dir_bytes = 0.25 * cells * (1.03)                 # 2 bits/cell + wavefront fill/drain slots
    n_fwd = max(1, fwd_launches // args.steps)
    ncu = (sc.get("ncu") or {}).get(skey) or {}
    plain = sc.get("x1") or {}
    roofline = {
        # integer roofline (north_star): which of its two terms binds - the ALU pipe or the ALU+FMA issue slots
        "bound": "int_alu" if roof_alu <= roof_issue else "int_issue", "kernel": kname,
        "achieved": fwd_gcups, "peak": roof_gcups, "unit": "GCUPS", "frac": (fwd_gcups / roof_gcups) if fwd_gcups else None,
        "peak_def": "min(measured ALU-pipe rate %.0f G instr/s / %.2f ALU instr per cell, measured ALU+FMA issue rate %.0f / %.2f instr per cell); "
                    "rates from gotoh_b200_int_peak in this run, counts from cuobjdump (profiles/sass_counts.json: %s)" % (alu_rate, alu_per_cell, issue_rate, instr_per_cell, skey),
        "instr_per_cell": instr_per_cell, "alu_instr_per_cell": alu_per_cell, "roof_alu_pipe": roof_alu, "roof_issue": roof_issue,
        "issue_peaks_ginstr_s": mix,
        "executed_instr_per_cell_ncu": ncu.get("executed_instr_per_cell"),
        "frac_vs_plain_cell_code": (fwd_gcups / min(alu_rate / plain["alu_per_cell"], issue_rate / plain["instr_per_cell"])) if (ktag == "x1_flow" and fwd_gcups and "instr_per_cell" in plain) else None,
        "forward_launches_per_step": n_fwd, "avg_launch_ms": fwd_ms_step / n_fwd,
        "hbm": {"bound": "hbm", "achieved": dir_bytes / (fwd_ms_step * 1e-3) / 1e9 if fwd_ms_step else None,
                "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "peak_src": peaks_src,
                "frac": (dir_bytes / (fwd_ms_step * 1e-3) / 1e9 / peaks["hbm_gbs"]) if fwd_ms_step else None,
                "algorithmic_bytes_per_cell": 0.25},
        # dram bytes of one forward launch: the per-cell figure of the committed ncu --set full capture of THIS kernel
        # (profiles/sass_counts.json: ncu.<kernel>.dram_bytes_per_cell, .source) x the cells one launch of this run covers
        "traffic": (ncu.get("dram_bytes_per_cell") or 0) * cells / n_fwd or None, "traffic_source": ncu.get("source"),
        "algorithmic_bytes_per_launch": dir_bytes / n_fwd,
    }

    # ---- CPU baseline on the host cores (bounded sample) -----------------------------------------
    cores = os.cpu_count() or 1
    kind = cpu_kind()
    sample = cpu_sample(jobs, cores, args.cpu_cells_per_core)
    cg, ca, cs = cpu_throughput(sample, cores, kind)
    e2e_line = None
    if "strings" in e2e:
        e2e_line = {"value": cells_all / es / 1e9, "unit": "GCUPS", "alignments_per_s": n_all / es, "s_per_step": es,
                    "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all), "format": "strings (the reference's two aligned strings per pair)",
                    "api": "gotoh_b200_align_batch (pinned host buffers)", "d2h_gbs": d2h_all / es / 1e9,
                    "verified_pairs_vs_resident": e2e["strings"]["verified_pairs_vs_resident"]}
        if ceiling and ps > 0:
            e2e_line["host_ceiling"] = {"gbs": probe_all / ps / 1e9, "how": "%d rank(s) x cudaMemcpyAsync D2H of %.2f GB into pinned memory at once (gotoh_b200_d2h_probe)" % (world, ceiling["bytes_per_rank"] / 1e9),
                                        "d2h_frac_of_ceiling": (d2h_all / es) / (probe_all / ps),
                                        "step_floor_s": d2h_all / (probe_all / ps)}
    if "compact" in e2e:
        cl = {"value": cells_all / ec / 1e9, "unit": "GCUPS", "alignments_per_s": n_all / ec, "s_per_step": ec,
              "d2h_bytes_per_step": int(d2hc_all), "format": "compact (8-word record + 2-bit op script per pair; strings rendered on demand)",
              "api": "gotoh_b200_align_batch_compact (pinned host buffers)",
              "scores_lengths_verified_pairs": e2e["compact"]["scores_lengths_verified_pairs"], "rendered_pairs_checked": e2e["compact"]["rendered_pairs_checked"]}
        if e2e_line is None:
            e2e_line = dict(cl, h2d_bytes_per_step=int(h2d_all))
        else:
            e2e_line["compact"] = cl
    line = {
        "metric": "GCUPS", "value": value, "unit": "GCUPS", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": dtype, "data": "real reads (reference examples/)" if args.config == "c1" else "synthetic",
        "alignments_per_s": n_all / (ms_per_step * 1e-3),
        "config": {"workload": workload, "name": args.config, "pairs_per_step": int(n_all), "pairs_this_rank": n_mine, "cells_per_step": cells_all,
                   "batches_per_step": len(jobs),
                   "gap_models": sorted({"%s gip=%d gep=%d term=%d" % ("align_it" if b.matrix == NT else "align_it_aa", b.gip, b.gep, b.term) for b in jobs}),
                   "parallelism": ("one process, device_mask over %d GPUs" % n_gpus) if single else "static shard of independent pairs, %d rank(s), no collective" % world,
                   "l2": "every step writes %.1f GB of directions through a %.1f GB arena and reads the tracebacks' part back; >> 126 MB L2, no flush needed" % (dir_bytes / 1e9, arena / 1e9),
                   "arena_chunks_per_step": chunks // args.steps if len(jobs) == 1 else chunks // args.steps, "pairs_int16x2": path_x2 // (args.steps if len(jobs) > 1 else 1),
                   "pairs_int32": path_x1 // (args.steps if len(jobs) > 1 else 1)},
        "bit_exact_verified_pairs": {"resident_vs_oracle": verified, "e2e_strings_vs_resident": e2e.get("strings", {}).get("verified_pairs_vs_resident", 0),
                                     "e2e_strings_vs_oracle": e2e.get("strings", {}).get("verified_pairs_vs_oracle", 0),
                                     "e2e_compact_vs_strings": e2e.get("compact", {}).get("scores_lengths_verified_pairs", 0)},
        "clocks": clocks,
        "e2e": e2e_line,
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": {"value": cg, "unit": "GCUPS", "alignments_per_s": ca, "cores": cores, "kind": kind,
                         "sample": "%d pairs of the same workload (first pairs of each batch, same generator/seed) over %d processes, %.1f s" % (sum(b.n for b in sample), cores, cs),
                         "gcups_per_core": cg / cores},
    }
    if args.emu:
        line["emu_selftest"] = "kernel sources under the CPU SIMT emulator: a self-test of this script, NOT a measurement"
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def concat_batches(bs):
    """Several batches with the same references and gap model as one."""
    b0 = bs[0]
    qb = np.concatenate([b.qb for b in bs])
    lens = np.concatenate([b.qlen for b in bs])
    qo = np.zeros(len(lens) + 1, np.int64)
    np.cumsum(lens, out=qo[1:])
    ridx = None if b0.ridx is None else np.concatenate([b.ridx for b in bs])
    return Batch(b0.matrix, b0.gip, b0.gep, b0.term, b0.refs, ridx, qb, qo)


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU aligner on all host cores, bounded sample per step."""
    if rank != 0:
        return
    workload, default_pairs, scaling, _ = CONFIGS[args.config]
    cores = os.cpu_count() or 1
    kind = cpu_kind()
    # a step is a bounded sample of the workload, sized so that warmup + steps finish within a few minutes
    per_core = args.cpu_cells_per_core / max(1.0, (args.steps + args.warmup) / 2.0)
    sample_src = make_batches(args.config, min(args.pairs or default_pairs, 200000 if args.config != "c4" else 400) if args.config != "c5" else 1000000, 0, 1)
    sample = cpu_sample(sample_src, cores, per_core)
    times = []
    for it in range(args.warmup + args.steps):
        g, a, dt = cpu_throughput(sample, cores, kind)
        if it >= args.warmup:
            times.append(dt)
    n = sum(b.n for b in sample)
    cells = sum(b.cells for b in sample)
    dt = sum(times) / len(times)
    gc, al_s = cells / dt / 1e9, n / dt
    line = {
        "impl": "reference", "metric": "GCUPS", "value": gc, "unit": "GCUPS", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "int32", "data": "real reads (reference examples/)" if args.config == "c1" else "synthetic", "alignments_per_s": al_s,
        "config": {"workload": workload, "name": args.config, "pairs_per_step": n,
                   "gap_models": sorted({"%s gip=%d gep=%d term=%d" % ("align_it" if b.matrix == NT else "align_it_aa", b.gip, b.gep, b.term) for b in sample})},
        "cpu_baseline": {"value": gc, "unit": "GCUPS", "cores": cores, "kind": kind,
                         "sample": "%d pairs per step over %d processes (bounded sample of the workload)" % (n, cores)},
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
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU per step (weak configs) or of the whole job (c5, c1); 0 = the config's size")
    ap.add_argument("--verify", type=int, default=200, help="pairs of the timed run checked against the oracle")
    ap.add_argument("--no-verify-full", dest="verify_full", action="store_false",
                    help="skip the byte-for-byte comparison of the timed e2e outputs with the resident arm's (all pairs)")
    ap.add_argument("--e2e-format", default="both", choices=["strings", "compact", "both"])
    ap.add_argument("--single-process", action="store_true", help="--gpus N from ONE process through device_mask (no torchrun)")
    ap.add_argument("--cpu-cells-per-core", type=float, default=1.0e9, help="size of the CPU baseline sample (DP cells per host core)")
    ap.add_argument("--emu", action="store_true", help=argparse.SUPPRESS)
    ap.add_argument("--lite", action="store_true", help="profiling runs: skip e2e, microbenchmarks and the CPU baseline")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1 and "RANK" not in os.environ and not args.single_process and args.impl == "ours":
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
