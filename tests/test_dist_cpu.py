"""CPU, world_size 2 over gloo: the batch-sharding and result-reassembly logic of the multi-GPU path."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from motioncraft_b200 import dist as mdist
from motioncraft_b200 import synth


def test_shard_range_covers_everything():
    for n in (1, 2, 7, 256, 2048, 2049):
        for world in (1, 2, 3, 8):
            spans = [mdist.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_row_seeded_inputs_are_shard_invariant():
    full = synth.synth_rows("x_T", (4, 6), 123, 0, 8)
    for world in (2, 3, 8):
        parts = [synth.synth_rows("x_T", (4, 6), 123, *mdist.shard_range(8, r, world)) for r in range(world)]
        assert torch.equal(torch.cat(parts, 0), full)


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = mdist.shard_range(n_total, rank, world)
        # stand-in for engine.sample(): a per-row function of the globally seeded rows
        local = synth.synth_rows("x_T", (3, 5), 7, lo, hi) * 2.0 + 1.0
        out = mdist.gather_rows(local, n_total)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7])
def test_gather_rows_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000 + n_total
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = synth.synth_rows("x_T", (3, 5), 7, 0, n_total) * 2.0 + 1.0
    assert torch.equal(outs[0], want) and torch.equal(outs[1], want)
