"""Data-parallel optimiser step with the optimiser sharded over the ranks (train.py:61-62 on N GPUs).

The reference trains on one GPU; its only multi-GPU hook is ``nn.parallel.data_parallel`` (model.py:39-43), i.e.
replicas whose gradients are summed.  SURVEY.md section 8e asks for one gradient exchange per step over NCCL.  An
all-reduce followed by a replicated Adam makes every rank stream the whole 28 B/parameter optimiser state; here the
same NVLink bytes are moved as

    reduce-scatter(dW)  ->  Adam on this rank's 1/N slice of (master weights, m, v)
                        ->  all-gather of the refreshed 16-bit operand planes the next forward / backward read,

so each rank updates -- and holds -- 1/N of the moments, and the all-gather of a layer only has to land before the next
step reaches that layer (the outermost transposed convolution, 44 % of the parameters, finishes backward first and is
needed last by the next forward).  The fp32 master weights of the slices a rank does not own go stale; ``sync_master``
all-gathers them on demand (checkpoints, re-packing for another executor).

``ShardedUpdater`` is host-side bookkeeping only: the arithmetic is the injected ``adam_fn`` (pg_adam_step on the GPU),
which lets tests/test_distributed_cpu.py drive the same code over gloo with a torch formula.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world, rank):
    """[lo, hi) of this rank's slice of a flat tensor of n elements, or None when n does not split evenly."""
    if world <= 1 or n % world:
        return None
    per = n // world
    return rank * per, (rank + 1) * per


class ShardedItem:
    """One parameter tensor: flat views of the master weights, the gradient and the operand planes, plus this rank's
    moment slices."""

    def __init__(self, key, p_flat, g_flat, planes, world, rank):
        self.key, self.p, self.g = key, p_flat, g_flat
        self.planes = [pl for pl in planes if pl is not None]
        lo, hi = shard_bounds(p_flat.numel(), world, rank)
        self.lo, self.hi = lo, hi
        self.g_shard = torch.empty(hi - lo, device=g_flat.device, dtype=g_flat.dtype)
        self.m = torch.zeros(hi - lo, device=p_flat.device, dtype=torch.float32)
        self.v = torch.zeros(hi - lo, device=p_flat.device, dtype=torch.float32)
        self.rs_work = None
        self.ag_works = []


class ShardedUpdater:
    def __init__(self, adam_fn, group=None):
        """adam_fn(p_slice, g_slice, m, v, plane_slices) updates p/m/v in place from the SUMMED gradient slice and
        rewrites the plane slices from the new p."""
        self.adam_fn, self.group = adam_fn, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.items = {}
        self.order = []                      # items whose reduce-scatter is in flight, in issue order
        self.master_stale = False

    def add(self, key, p_flat, g_flat, planes):
        it = ShardedItem(key, p_flat, g_flat, planes, self.world, self.rank)
        self.items[key] = it
        return it

    @staticmethod
    def can_shard(n, world):
        return shard_bounds(n, world, 0) is not None

    # ---------------------------------------------------------------------------------- step
    def grad_ready(self, key):
        """The gradient of `key` is complete on the current stream: start its reduce-scatter (asynchronous)."""
        it = self.items[key]
        it.rs_work = dist.reduce_scatter_tensor(it.g_shard, it.g, group=self.group, async_op=True)
        self.order.append(it)

    def finish(self):
        """After backward has been queued: per item, in the order the gradients completed, wait for its reduce-scatter,
        update this rank's slice, and start the all-gather of the refreshed planes (asynchronous; `wait_planes` is the
        matching wait, called right before the next kernel that reads them)."""
        for it in self.order:
            it.rs_work.wait()
            self.adam_fn(it.p[it.lo:it.hi], it.g_shard, it.m, it.v, [pl[it.lo:it.hi] for pl in it.planes])
            it.rs_work = None
        # The all-gathers leave in the order the NEXT forward reads the planes -- the reverse of the order backward finished
        # the gradients: the first convolution's (small) planes land first, the outermost transposed convolution's (44 % of
        # the bytes) are needed last and travel underneath the forward pass.  (Issued in completion order, the first layer
        # of the next step waited for the whole 1.2 GB.)
        for it in reversed(self.order):
            it.ag_works = [dist.all_gather_into_tensor(pl, pl[it.lo:it.hi], group=self.group, async_op=True) for pl in it.planes]
        self.order = []
        self.master_stale = True

    def wait_planes(self, key=None):
        """Make the current stream wait for the all-gather(s) of `key`'s planes (all items when key is None)."""
        its = self.items.values() if key is None else ([self.items[key]] if key in self.items else [])
        for it in its:
            for w in it.ag_works:
                w.wait()
            it.ag_works = []

    # ------------------------------------------------------------------------ master / state
    def sync_master(self):
        """All-gather the fp32 master weights (a collective: every rank must call it)."""
        if not self.master_stale:
            return
        self.wait_planes()
        for it in self.items.values():
            dist.all_gather_into_tensor(it.p, it.p[it.lo:it.hi], group=self.group)
        self.master_stale = False

    def full_moments(self, key):
        """(m, v) of the whole tensor, gathered from the ranks (a collective)."""
        it = self.items[key]
        out = []
        for sh in (it.m, it.v):
            full = torch.empty(it.p.numel(), device=sh.device, dtype=sh.dtype)
            dist.all_gather_into_tensor(full, sh, group=self.group)
            out.append(full)
        return out

    def load_moments(self, key, m_full, v_full):
        it = self.items[key]
        it.m.copy_(m_full.reshape(-1)[it.lo:it.hi])
        it.v.copy_(v_full.reshape(-1)[it.lo:it.hi])
