"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic: how clips / long-form windows are
dealt to ranks (no data-path collective for inference) and the gradient-averaging arithmetic the
training step relies on (all-reduce sum, then 1/world inside Adam)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "unet-phasegen_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from phasegen import longform
    wave = torch.arange(50000, dtype=torch.float32)
    wins, idx = longform.cut_windows(wave, 64, 40, rank, world)
    win, step, n = longform.window_plan(50000, 64, 40)
    # every rank processes only its windows (identity "pipeline"); the host gathers and stitches
    gathered = [None] * world
    dist.all_gather_object(gathered, (idx, wins))
    all_idx = [i for ix, _ in gathered for i in ix]
    all_w = torch.cat([w for _, w in gathered])
    out = longform.stitch(all_w, all_idx, n, 50000, 64, 40)
    ok_stitch = bool(torch.allclose(out, wave, rtol=1e-6, atol=1e-2))
    # gradient averaging as TrainStep does it: all-reduce(sum) then scale 1/world
    g = torch.full((8,), float(rank + 1))
    dist.all_reduce(g)
    ok_grad = bool(torch.allclose(g / world, torch.full((8,), (1 + world) / 2.0)))
    q.put((rank, sorted(idx), ok_stitch, ok_grad))
    dist.destroy_process_group()


def test_two_rank_window_sharding_and_gradient_average():
    world, port = 2, 29511
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    all_idx = sorted(i for _, ix, _, _ in res for i in ix)
    assert all_idx == list(range(len(all_idx))) and len(all_idx) > 2      # a partition of the windows
    assert set(res[0][1]).isdisjoint(res[1][1])
    assert all(r[2] and r[3] for r in res)
