"""N>1 path on CPU: two gloo ranks shard one batch statically (no data-path collective), each
aligns its shard (here through the TEST-ONLY emulator), and the gathered results equal the
single-process answer.  Mirrors what bench.py does under torchrun."""
import os
import socket
import sys

import pytest


def _worker(rank, world, port, out_path):
    import hashlib
    import json

    import numpy as np
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "micall-lite_b200"), os.path.join(root, "tests", "simt_emu")):
        sys.path.insert(0, p)
    import build_emu
    from gotoh_b200 import _ffi, workloads
    from gotoh_b200.api import Aligner
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    ref, reads = workloads.c2_reads(24, seed=123)
    lo, hi = workloads.shard_range(len(reads), rank, world)
    aligner = Aligner(_ffi.Library(build_emu.build()))
    mine = aligner.align_batch(ref, reads[lo:hi], 10, 3, 1, 0)
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, mine))       # host-side gather only
    dist.barrier()
    if rank == 0:
        full = [None] * len(reads)
        for glo, ghi, res in gathered:
            full[glo:ghi] = res
        digest = hashlib.sha256(json.dumps(full).encode()).hexdigest()
        with open(out_path, "w") as f:
            json.dump({"digest": digest, "n": len(full)}, f)
    dist.destroy_process_group()


def test_two_rank_static_sharding_matches_single_process(tmp_path, emu_aligner):
    import hashlib
    import json

    import torch.multiprocessing as mp
    from gotoh_b200 import workloads
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "gathered.json")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    ref, reads = workloads.c2_reads(24, seed=123)
    single = emu_aligner.align_batch(ref, reads, 10, 3, 1, 0)
    got = json.load(open(out))
    assert got["n"] == 24
    assert got["digest"] == hashlib.sha256(json.dumps([list(x) for x in single]).encode()).hexdigest()
