"""world_size-2 gloo test of the N>1 host path: segments are dealt to ranks, each rank produces results for
its share only, and rank 0 gathers them back into the caller's order (no data-path collective exists)."""
import importlib
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sch = importlib.import_module("asr-2pass_b200.scheduler")
    synth = importlib.import_module("asr-2pass_b200.synth")
    lens = synth.segment_lengths(101)
    mine = sch.shard_segments(lens, world, rank, max_rows=2048)
    # stand-in for the per-rank forward: the frame count is what each rank would report per segment
    res = {i: sch.num_lfr_frames(int(lens[i])) for b in mine for i in b}
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        out = sch.gather_results(gathered, len(lens))
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    synth = importlib.import_module("asr-2pass_b200.synth")
    sch = importlib.import_module("asr-2pass_b200.scheduler")
    lens = synth.segment_lengths(101)
    assert out == [sch.num_lfr_frames(int(n)) for n in lens]


def test_shards_are_balanced():
    sch = importlib.import_module("asr-2pass_b200.scheduler")
    synth = importlib.import_module("asr-2pass_b200.synth")
    lens = synth.segment_lengths(1024)
    for world in (2, 4, 8):
        cost = []
        for r in range(world):
            cost.append(sum(sch.segment_cost(sch.num_lfr_frames(int(lens[i]))) for b in sch.shard_segments(lens, world, r, 24576) for i in b))
        assert max(cost) / min(cost) < 1.03
