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


def _adam_ref(p, g, m, v, t, lr=1e-3, b1=0.9, b2=0.999, eps=1e-8):
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    p.sub_(lr / (1 - b1 ** t) * m / (v.sqrt() / (1 - b2 ** t) ** 0.5 + eps))


def _sharded_worker(rank, world, port, q):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "unet-phasegen_b200")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from phasegen.sharded import ShardedUpdater, shard_bounds
    assert shard_bounds(10, 4, 0) is None and shard_bounds(12, 4, 3) == (9, 12)
    sizes = {"u1": 4096, "d2": 768, "d1": 64}
    torch.manual_seed(0)                                   # identical start on every rank
    p = {k: torch.randn(n) for k, n in sizes.items()}
    planes = {k: v.bfloat16() for k, v in p.items()}
    g = {k: torch.zeros(n) for k, n in sizes.items()}
    ref_p = {k: v.clone() for k, v in p.items()}
    ref_m = {k: torch.zeros_like(v) for k, v in p.items()}
    ref_v = {k: torch.zeros_like(v) for k, v in p.items()}
    step = [0]

    def adam_fn(ps, gs, m, v, pls):                        # gs = SUM over ranks of this slice; 1/world applied here
        _adam_ref(ps, gs / world, m, v, step[0])
        for pl in pls:
            pl.copy_(ps.bfloat16())
    up = ShardedUpdater(adam_fn)
    for k in sizes:
        up.add(k, p[k], g[k], [planes[k], None])
    ok = True
    for t in range(1, 4):
        step[0] = t
        all_g = {}
        for k, n in sizes.items():
            gen = torch.Generator().manual_seed(1000 * t + n)       # (str hashes differ between processes)
            per_rank = torch.randn(world, n, generator=gen)          # every rank can reproduce every rank's gradient
            g[k].copy_(per_rank[rank])
            all_g[k] = per_rank
        for k in ("u1", "d2", "d1"):                       # the order backward completes them
            up.grad_ready(k)
        up.finish()
        for k in ("d1", "d2", "u1"):                       # the order the next forward needs them
            up.wait_planes(k)
        for k in sizes:
            _adam_ref(ref_p[k], all_g[k].sum(0) / world, ref_m[k], ref_v[k], t)
            ok &= bool(torch.equal(planes[k], ref_p[k].bfloat16()))             # every rank sees the whole refreshed plane
        lo, hi = up.items["u1"].lo, up.items["u1"].hi
        ok &= bool(torch.allclose(p["u1"][lo:hi], ref_p["u1"][lo:hi], rtol=0, atol=1e-6))   # own slice is current ...
    stale = not torch.allclose(p["u1"], ref_p["u1"], atol=1e-6)                  # ... the rest is stale until sync_master
    up.sync_master()
    for k in sizes:
        ok &= bool(torch.allclose(p[k], ref_p[k], rtol=0, atol=1e-6))
        gathered = [torch.empty_like(p[k]) for _ in range(world)]
        dist.all_gather(gathered, p[k])
        ok &= all(torch.equal(gathered[0], t_) for t_ in gathered)               # bit-identical replicas
        m_full, v_full = up.full_moments(k)
        ok &= bool(torch.allclose(m_full, ref_m[k], atol=1e-7)) and bool(torch.allclose(v_full, ref_v[k], atol=1e-7))
    q.put((rank, ok, stale))
    dist.destroy_process_group()


def test_two_rank_sharded_optimizer_equals_replicated_adam():
    """reduce-scatter -> Adam on 1/world of (weights, moments) -> all-gather of the refreshed planes (phasegen/sharded.py)
    ends three steps with bit-identical replicas equal to all-reduce + replicated Adam; fp32 master slices of other ranks
    are stale until sync_master()."""
    world, port = 2, 29517
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res) and all(stale for _, _, stale in res)
