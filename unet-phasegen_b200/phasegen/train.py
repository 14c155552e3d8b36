"""Training step of the U-Net on the phasegen kernels (train.py:37-62 of the reference).

``TrainExecutor`` extends the forward executor with the backward pass:

    loss gradient d_out ─► for every layer, top down:
        pg_bn_bwd      dZ = backward of [norm -> activation fan-out]  (+ dgamma, dbeta)
        pg_wgrad_*     dW (packed [k][C_out][C_in]) from the layer's saved input operand and dZ
        pg_conv_*      data gradient = the forward kernel on the mirrored geometry

A tensor that feeds two consumers (the down path and the skip concat, model.py:113) receives both
upstream gradients inside one pg_bn_bwd call, each masked with its own activation derivative
(LeakyReLU(0.2) for the down path, ReLU for the skip -- the in-place quirk of model.py:80).
"""
import os

import torch

from . import ops
from ._lib import (PG_CONV, PG_CONV_TRANSPOSE, PG_DT_F32, PG_PREC_BF16, PG_PREC_BF16X3, PG_PREC_FP32_SIMT)
from .unet import UNetExecutor, _Operand, _rows

N_CHUNKS = 64


class TrainExecutor(UNetExecutor):
    def __init__(self, levels, B, T, device, precision="bf16", grad_dtype=torch.float32, **kw):
        super().__init__(levels, B, T, device, precision, per_clip=False, keep_raw=True, **kw)
        if grad_dtype not in (torch.float32, torch.bfloat16) or (grad_dtype == torch.bfloat16 and self.prec == PG_PREC_FP32_SIMT):
            raise RuntimeError("phasegen: weight gradients are float32, or bfloat16 on the tensor-core precisions")
        self.grad_dtype = grad_dtype
        if self.prec not in (PG_PREC_FP32_SIMT, PG_PREC_BF16X3, PG_PREC_BF16):
            raise RuntimeError("phasegen: the training step runs in 'bf16', 'bf16x3' or 'fp32_simt' "
                               "(the fp16 operand modes are inference-only: gradients need the bf16 range)")
        D, dev, prec = self.D, self.device, self.prec
        f32 = dict(device=dev, dtype=torch.float32)
        self.d_out = torch.empty(B, self.T_out, self.C_final, **f32)
        self.dz_up, self.dz_dn = [None] * D, [None] * D             # dZ operands (G of wgrad / input of dgrad)
        self.din_up, self.din_dn = [None] * D, [None] * D           # data gradients w.r.t. each conv's input (fp32)
        self.dgrad_up, self.dgrad_dn = [None] * D, [None] * D
        self.dw_up, self.dw_dn = [None] * D, [None] * D
        self.dgb_up, self.dgb_dn = [None] * D, [None] * D
        self.ws_partial, self.ws_coef = [None] * D, [None] * D
        # all norm-parameter gradients live in one flat buffer (one small all-reduce in data-parallel runs)
        n_norm = sum(2 * d.C_out for i, lv in enumerate(levels)
                     for d, has in ((self.dn_desc[i], lv.down_norm), (self.up_desc[i], lv.up_norm)) if has)
        self.dgb_flat = torch.zeros(max(n_norm, 1), **f32)
        self.grad_hook = None                                           # called with each weight gradient once it is complete
        off = 0
        for i, lv in enumerate(levels):
            for which, desc, has_norm in (("dn", self.dn_desc[i], lv.down_norm), ("up", self.up_desc[i], lv.up_norm)):
                dz = _Operand(B, desc.L_out, desc.C_out, prec, dev)
                dw = torch.zeros(desc.k, desc.C_out, desc.C_in, device=dev, dtype=grad_dtype)
                dgb = None
                if has_norm:
                    dgb = (self.dgb_flat[off:off + desc.C_out], self.dgb_flat[off + desc.C_out:off + 2 * desc.C_out])
                    off += 2 * desc.C_out
                need_dgrad = not (which == "dn" and i == 0)          # the network input needs no gradient
                din = torch.empty(B, desc.L_in, desc.C_in, **f32) if need_dgrad else None
                # data gradient = the forward kernel on the mirrored geometry; on the tensor cores it reads
                # the layer's FORWARD weight planes as an MN-major operand (no second packing)
                mirror = ops.conv_desc(PG_CONV if desc.kind == PG_CONV_TRANSPOSE else PG_CONV_TRANSPOSE, B, desc.C_out,
                                       desc.C_in, desc.L_out, desc.k, desc.stride, desc.pad, dz.rows, dz.ld, prec,
                                       L_out=desc.L_in, taps_per_group=self.tpg, base_offset_mode=self.bo,
                                       weights_mn_major=int(prec != PG_PREC_FP32_SIMT)) if need_dgrad else None
                if which == "dn":
                    self.dz_dn[i], self.dw_dn[i], self.dgb_dn[i], self.din_dn[i], self.dgrad_dn[i] = dz, dw, dgb, din, mirror
                else:
                    self.dz_up[i], self.dw_up[i], self.dgb_up[i], self.din_up[i], self.dgrad_up[i] = dz, dw, dgb, din, mirror
        cmax = max(max(d.C_out for d in self.dn_desc), max(d.C_out for d in self.up_desc))
        self.partial = torch.empty(N_CHUNKS, cmax, 2, **f32)
        self.coef = torch.empty(cmax, 2, **f32)
        self.loss_partial = torch.empty(1024, 3, device=dev, dtype=torch.float64)
        self.loss3 = torch.zeros(4, **f32)
        self._wg_desc = {}

    def reserve_sms_in_backward(self, n_ctas):
        """Cap the persistent grids of the backward tensor-core kernels (dgrad, wgrad) at `n_ctas` CTAs so that
        `148 - n_ctas` SMs stay free for the NCCL all-reduce kernels that run beside them in data-parallel
        training.  Without it an all-reduce that grabbed SMs at a kernel boundary delays the CTAs of the next
        persistent kernel, whose statically assigned tiles then finish late (measured at 8 GPUs: 13.85 -> 13.42
        ms per step with 16 SMs reserved; 32 reserved is too many: 16.7 ms)."""
        from ._lib import ConvDesc
        for lst in (self.dgrad_dn, self.dgrad_up):
            for d in lst:
                if d is not None:
                    d.tc_max_ctas = int(n_ctas)
        self._wg_desc = {}
        for d in list(self.dn_desc) + list(self.up_desc):
            c = ConvDesc.from_buffer_copy(d)
            c.tc_max_ctas = int(n_ctas)
            self._wg_desc[id(d)] = c

    def pack_weights(self, down_w, up_w):
        """Forward operands; the SIMT path additionally packs each weight with the other `kind` for the
        data gradient (a Conv1d weight [C_out][C_in][k] read as a ConvTranspose1d weight and vice
        versa).  The tensor-core data gradient needs nothing extra."""
        super().pack_weights(down_w, up_w)
        if self.prec == PG_PREC_FP32_SIMT:
            flip = lambda k: PG_CONV if k == PG_CONV_TRANSPOSE else PG_CONV_TRANSPOSE
            self.wd_t = [ops.pack_weight(down_w[i].contiguous(), flip(lv.down.kind), want_tc=False, want_simt=True) if i > 0 else None
                         for i, lv in enumerate(self.levels)]
            self.wu_t = [ops.pack_weight(up_w[i].contiguous(), flip(lv.up.kind), want_tc=False, want_simt=True)
                         for i, lv in enumerate(self.levels)]
        else:
            self.wd_t, self.wu_t = self.wd, self.wu

    # ---------------------------------------------------------------------------------------
    def loss(self, logmag_cl, phase_cl, mag_weight=0.2):
        """train.py:45-60 on self.out; fills self.d_out and returns the 4-vector (total, cos, sin, mag)."""
        ops.phase_loss(self.out, logmag_cl, phase_cl, self.d_out, self.loss_partial, self.loss3, mag_weight)
        return self.loss3

    def _layer_bwd(self, desc, z, ss, mv, eps, srcs, dz, dgb, x_operand, dw, mirror, w_t, din):
        g0 = srcs[0]
        g1 = srcs[1] if len(srcs) > 1 else None
        ops.bn_bwd(z, desc.B, desc.L_out, desc.C_out, ss, mv if ss is not None else None, eps, g0, g1,
                   self.partial if ss is not None else None, self.coef, dgb[0] if dgb else None, dgb[1] if dgb else None,
                   dz.hi, dz.lo, dz.rows, dz.dtype)
        if self.prec == PG_PREC_FP32_SIMT:
            ops.wgrad_simt(desc, x_operand.hi, dz.hi, dz.rows, dw)
        else:
            ops.wgrad_tc(self._wg_desc.get(id(desc), desc), x_operand.hi, x_operand.lo, dz.hi, dz.lo, dz.rows, dw)
        if self.grad_hook is not None:
            self.grad_hook(dw)                  # e.g. start this layer's all-reduce while the rest of backward runs
        if mirror is not None:
            if self.prec == PG_PREC_FP32_SIMT:
                ops.conv_simt(mirror, dz.hi, w_t[2], din)
            else:
                ops.conv_tc(mirror, dz.hi, dz.lo, w_t[0], w_t[1], din, None)

    def backward(self, dn_norm, up_norm, d_out=None):
        """Consumes the buffers of the last run(); d_out [B][T][2C] fp32 channels-last (default: the
        one pg_phase_loss wrote).  Leaves packed weight gradients in dw_dn/dw_up and (dgamma, dbeta)
        in dgb_dn/dgb_up."""
        D, lv = self.D, self.levels
        d_out = self.d_out if d_out is None else d_out
        eps_of = lambda n: n[2] if n is not None else 1e-5
        # up path, outermost first
        for i in range(D):
            desc = self.up_desc[i]
            if i == 0:
                srcs = [ops.grad_src(d_out, desc.C_out, 0, 1.0)]
            else:                                   # ReLU(n_i) sits in cat[i-1] at channel offset C_down(i-1)
                up_in = self.din_up[i - 1]
                srcs = [ops.grad_src(up_in, self.up_desc[i - 1].C_in, lv[i - 1].down.C_out, 0.0)]
            x_op = self.a[i] if i == D - 1 else self.cat[i]
            self._layer_bwd(desc, self.g[i], self.up_ss[i], self.up_mv[i], eps_of(up_norm[i]), srcs, self.dz_up[i],
                            self.dgb_up[i], x_op, self.dw_up[i], self.dgrad_up[i], self.wu_t[i], self.din_up[i])
        # down path, innermost first
        for i in range(D - 1, -1, -1):
            desc = self.dn_desc[i]
            if i == D - 1:                          # a[D-1] = ReLU(z) feeds the innermost up conv only
                srcs = [ops.grad_src(self.din_up[i], self.up_desc[i].C_in, 0, 0.0)]
            else:                                   # LeakyReLU(h_i) -> down conv i+1 ; ReLU(h_i) -> cat[i][:, :C]
                srcs = [ops.grad_src(self.din_dn[i + 1], self.dn_desc[i + 1].C_in, 0, 0.2),
                        ops.grad_src(self.din_up[i], self.up_desc[i].C_in, 0, 0.0)]
            x_op = self.x0 if i == 0 else self.a[i - 1]
            has = lv[i].down_norm
            self._layer_bwd(desc, self.z[i], self.dn_ss[i] if has else None, self.dn_mv[i] if has else None,
                            eps_of(dn_norm[i]) if has else 1e-5, srcs, self.dz_dn[i], self.dgb_dn[i], x_op, self.dw_dn[i],
                            self.dgrad_dn[i], self.wd_t[i], self.din_dn[i])


class TrainStep:
    """One optimisation step of train.py:37-62 entirely on the phasegen kernels: forward (batch
    statistics), the cos/sin/magnitude loss, backward, gradient all-reduce over NCCL when
    world_size > 1 (the only collective of the path), fused Adam (torch.optim.Adam defaults,
    train.py:26-27) that also writes the updated weights as bf16 tensor-core operand planes.
    Weights, gradients and optimiser state all live in the packed [k][C_out][C_in] order, so no
    layout conversion happens inside a step.  Inputs are channels-last: log-magnitude and target
    phase [B, T, C]."""

    def __init__(self, net, B, T, device, precision="bf16", lr=1e-3, betas=(0.9, 0.999), eps=1e-8, mag_weight=0.2,
                 grad_dtype=None, shard_optimizer=None, reserve_ctas=None, data_parallel=True):
        """shard_optimizer: data-parallel runs only.  True (default when world_size > 1) = reduce-scatter the weight
        gradients, update 1/world of (master weights, Adam moments) per rank, all-gather the refreshed operand planes
        (phasegen/sharded.py); False = all-reduce + replicated Adam.  data_parallel=False ignores an initialised process
        group (a purely local step, e.g. a reference replica inside a distributed check).  reserve_ctas: cap of the persistent grids of the
        backward tensor-core kernels in data-parallel runs (default 148 = no SMs set aside for NCCL: with the sharded optimiser
        the full grid measured faster at 2 and at 8 GPUs, 8.53 vs 8.62 and 8.45 vs 8.59 ms per step; round 1's all-reduce form
        wanted 132).
        grad_dtype: dtype of the weight gradients the wgrad kernel writes, NCCL reduces and Adam reads:
        "fp32" (default for the fp32-class precisions: what autograd would give) or "bf16" (default for
        precision="bf16", on any number of GPUs: half the all-reduce volume -- 1.22 GB instead of 2.45 GB for
        UNetModel(1024, 2048) -- and 2 B/parameter less HBM traffic in wgrad and in Adam; moments and master
        weights stay fp32)."""
        import torch.distributed as dist
        self.net, self.lr, self.betas, self.eps, self.mag_weight = net, lr, betas, eps, mag_weight
        blocks = net._blocks()
        for b in blocks:                          # foreign-layout weights (e.g. assigned by the user) are re-laid once
            for conv, kind in ((b._parts["down"], PG_CONV), (b._parts["up"], PG_CONV_TRANSPOSE)):
                if ops.packed_view(conv.weight, kind) is None:
                    conv.weight.data = ops.to_packed_storage(conv.weight.data, kind)
        self.world = dist.get_world_size() if data_parallel and dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        if self.world > 1:
            # replicas must start identical (DDP broadcasts at wrap time): parameters and norm buffers from rank 0
            with torch.no_grad():
                for t in list(net.parameters()) + list(net.buffers()):
                    d = t.data
                    if not d.is_contiguous():         # packed-storage conv weights: dense, permuted strides -> flat view of the storage
                        d = d.as_strided((d.numel(),), (1,), d.storage_offset())
                    dist.broadcast(d, 0)
            net.invalidate_packed()
        if grad_dtype is None:
            grad_dtype = "bf16" if precision == "bf16" else "fp32"
        self.grad_dtype = {"fp32": torch.float32, "bf16": torch.bfloat16}[grad_dtype]
        self.ex = net.train_executor(B, T, device, precision, grad_dtype=self.grad_dtype)
        if self.world > 1 and self.ex.prec != PG_PREC_FP32_SIMT:
            n_res = int(os.environ.get("PG_DDP_CTAS", "148")) if reserve_ctas is None else int(reserve_ctas)
            if 0 < n_res < 148:
                self.ex.reserve_sms_in_backward(n_res)
        self.t = 0
        self._works = []
        ex = self.ex
        tc = ex.prec != PG_PREC_FP32_SIMT
        # (parameter storage, gradient buffer, bf16 planes to refresh) in packed order
        self.items = []
        for i, b in enumerate(blocks):
            for conv, kind, dw, which in ((b._parts["down"], PG_CONV, ex.dw_dn[i], "dn"), (b._parts["up"], PG_CONV_TRANSPOSE, ex.dw_up[i], "up")):
                if not conv.weight.requires_grad:
                    continue
                self.items.append(dict(p=ops.packed_view(conv.weight, kind), g=dw, conv=(which, i) if tc else None))
            for nm, dgb in ((b._parts["down_norm"], ex.dgb_dn[i]), (b._parts["up_norm"], ex.dgb_up[i])):
                if nm is None or dgb is None or getattr(nm, "weight", None) is None:
                    continue
                for prm, g in ((nm.weight, dgb[0]), (nm.bias, dgb[1])):
                    if prm.requires_grad:
                        self.items.append(dict(p=prm.data, g=g, conv=None))
        # sharded optimiser: conv weights whose element count splits evenly over the ranks
        self.sharded = None
        if shard_optimizer is None:
            shard_optimizer = self.world > 1
        if shard_optimizer and self.world > 1:
            from .sharded import ShardedUpdater
            self.sharded = ShardedUpdater(self._adam_slice)
            for it in self.items:
                if it["conv"] is not None and ShardedUpdater.can_shard(it["p"].numel(), self.world):
                    which, i = it["conv"]
                    hi, lo = (ex.wd[i] if which == "dn" else ex.wu[i])[:2]
                    it["key"] = it["g"].data_ptr()
                    it["sh"] = self.sharded.add(it["key"], it["p"].view(-1), it["g"].view(-1),
                                                [hi.view(-1), lo.view(-1) if lo is not None else None])
                    it["planes"] = id(hi)
            net.__dict__["_pre_state_hook"] = self.sync_master
        for it in self.items:
            if "sh" in it:
                continue                           # moments live (sliced) in the sharded updater
            it["m"] = torch.zeros_like(it["p"]); it["v"] = torch.zeros_like(it["p"])

    def _adam_slice(self, p, g, m, v, planes):
        hi = planes[0] if planes else None
        lo = planes[1] if len(planes) > 1 else None
        ops.adam_step(p, g, m, v, self.lr, self.betas[0], self.betas[1], self.eps, self.t, 1.0 / self.world, hi, lo)

    def sync_master(self):
        """Sharded optimiser: all-gather the fp32 master weights so that every rank holds the current model (needed
        before a checkpoint, or before another executor re-packs its operand planes).  A collective."""
        if self.sharded is not None:
            self.sharded.sync_master()

    def __call__(self, logmag_cl, phase_cl):
        loss3 = self.forward_backward(logmag_cl, phase_cl)
        self.apply()
        return loss3

    def forward_backward(self, logmag_cl, phase_cl):
        """Forward, loss and backward of one batch; gradient collectives are started as the layers finish.  Leaves the
        (local, or in-flight reduced) gradients in the executor; `apply()` completes the step."""
        import torch.distributed as dist
        net, ex = self.net, self.ex
        logmag_cl = ops._need_cuda(logmag_cl, "logmag_cl")
        phase_cl = ops._need_cuda(phase_cl, "phase_cl")
        if tuple(logmag_cl.shape) != (ex.B, ex.T, ex.levels[0].down.C_in) or phase_cl.shape != logmag_cl.shape:
            raise RuntimeError(f"phasegen: TrainStep was built for channels-last pairs of shape "
                               f"{(ex.B, ex.T, ex.levels[0].down.C_in)}, got {tuple(logmag_cl.shape)} / {tuple(phase_cl.shape)}")
        net._ensure_packed(ex)
        if self.sharded is not None:
            # a re-pack (weights changed outside this object) allocates new operand planes: re-bind the shards to them
            for it in self.items:
                if "sh" in it:
                    which, i = it["conv"]
                    hi, lo = (ex.wd[i] if which == "dn" else ex.wu[i])[:2]
                    if id(hi) != it["planes"]:
                        it["sh"].planes = [pl.view(-1) for pl in (hi, lo) if pl is not None]
                        it["planes"] = id(hi)
            # each layer's refreshed planes are awaited right before the first kernel that reads them
            pending = {it["planes"]: it["key"] for it in self.items if "sh" in it}

            def ready(w):
                key = pending.get(id(w[0]))
                if key is not None:
                    self.sharded.wait_planes(key)
            ex.weight_ready = ready
        ex.load_input_cl(logmag_cl)
        dn, up = net._norm_params(logmag_cl.device)
        ex.run(dn, up)
        if net.training:
            net._update_running_stats(ex)
        loss3 = ex.loss(logmag_cl, phase_cl, self.mag_weight)
        works = []                                  # (gradient tensor, NCCL work) in the order backward finishes them
        if self.world > 1:
            # one NCCL collective per weight gradient, issued as soon as that layer's wgrad is queued (the outermost
            # transposed conv, 44 % of the parameters, goes first), overlapping the remaining dgrad / wgrad kernels:
            # a reduce-scatter for sharded items, an all-reduce for the rest
            sharded_keys = set(self.sharded.items) if self.sharded is not None else ()

            def hook(g):
                if g.data_ptr() in sharded_keys:
                    self.sharded.grad_ready(g.data_ptr())
                else:
                    works.append((g, dist.all_reduce(g, async_op=True)))
            ex.grad_hook = hook
        ex.backward(dn, up)
        ex.grad_hook = None
        self._works = works
        return loss3

    def _adam(self, it, scale):
        ex = self.ex
        hi = lo = None
        if it["conv"] is not None:            # refresh the tensor-core operand planes in the same pass
            which, i = it["conv"]
            hi, lo = (ex.wd[i] if which == "dn" else ex.wu[i])[:2]
        ops.adam_step(it["p"], it["g"], it["m"], it["v"], self.lr, self.betas[0], self.betas[1], self.eps, self.t, scale, hi, lo)

    def apply(self):
        """Adam (train.py:62) on the gradients of the last forward_backward(), refreshing the operand planes."""
        import torch.distributed as dist
        net, ex, works = self.net, self.ex, self._works
        self.t += 1
        scale = 1.0 / self.world
        adam = lambda it: self._adam(it, scale)

        if self.world > 1:
            # each layer's Adam update is queued behind that layer's collective only, so it runs while the
            # collectives of the layers backward reached later are still on the wire
            flat = dist.all_reduce(ex.dgb_flat, async_op=True)
            by_grad = {it["g"].data_ptr(): it for it in self.items}
            done = set(id(it) for it in self.items if "sh" in it)
            if self.sharded is not None:
                self.sharded.finish()
            for g, w in works:
                w.wait()
                it = by_grad.get(g.data_ptr())
                if it is not None:
                    adam(it); done.add(id(it))
            flat.wait()
            for it in self.items:
                if id(it) not in done:
                    adam(it)
        else:
            # (Tried: each layer's Adam on a side stream underneath the remaining backward kernels.  No gain -- 8.87 vs 9.00 ms:
            #  the persistent tensor-core kernels fill every SM's shared memory, so the update only ran in the gaps.)
            for it in self.items:
                adam(it)
        if ex.prec == PG_PREC_FP32_SIMT:          # SIMT operand layouts are re-packed from the updated weights
            blocks = net._blocks()
            ws = [b._parts["down"].weight for b in blocks] + [b._parts["up"].weight for b in blocks]
            ex.pack_weights(ws[:len(blocks)], ws[len(blocks):])
        net._mark_packed(ex)

    # ------------------------------------------------------------------ checkpoint / resume
    def _named_items(self):
        """Adam items keyed by the parameter's state_dict name (the reference's checkpoint keys, model.py:45-48)."""
        by_ptr = {p.data_ptr(): n for n, p in self.net.model.named_parameters()}
        out = []
        for it in self.items:
            name = by_ptr.get(it["p"].data_ptr())
            if name is None:
                raise RuntimeError("phasegen: optimiser item without a matching model parameter")
            out.append((name, it))
        return out

    def _is_transposed(self, name):
        mod = self.net.model.get_submodule(name.rsplit(".", 1)[0])
        return isinstance(mod, torch.nn.ConvTranspose1d)

    def state_dict(self):
        """Optimiser state for resume: step count, hyper-parameters and the Adam moments in the TORCH layout of
        each parameter (so the file does not depend on this build's packed storage order)."""
        from ._lib import PG_CONV, PG_CONV_TRANSPOSE
        params = dict(self.net.model.named_parameters())
        state = {}
        self.sync_master()
        for name, it in self._named_items():
            p = params[name]
            if "sh" in it:                        # sharded moments: gather the slices (a collective on every rank)
                m, v = (t.view(it["p"].shape) for t in self.sharded.full_moments(it["key"]))
            else:
                m, v = it["m"], it["v"]
            if m.dim() == 3:                      # conv weights: packed [k][C_out][C_in] -> torch layout
                perm = (2, 1, 0) if self._is_transposed(name) else (1, 2, 0)
                m, v = m.permute(*perm), v.permute(*perm)
            state[name] = {"exp_avg": m.detach().cpu().contiguous(), "exp_avg_sq": v.detach().cpu().contiguous()}
        return {"step": self.t, "lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "state": state}

    def load_state_dict(self, sd):
        params = dict(self.net.model.named_parameters())
        for name, it in self._named_items():
            if name not in sd["state"]:
                raise KeyError(f"phasegen: optimiser state for '{name}' missing from the checkpoint")
            p = params[name]
            full = {}
            for key in ("exp_avg", "exp_avg_sq"):
                src = sd["state"][name][key].to(it["p"].device, torch.float32)
                if tuple(src.shape) != tuple(p.shape):
                    raise RuntimeError(f"phasegen: optimiser state '{name}.{key}' has shape {tuple(src.shape)}, expected {tuple(p.shape)}")
                if it["p"].dim() == 3:
                    src = src.permute(2, 1, 0) if self._is_transposed(name) else src.permute(2, 0, 1)
                full[key] = src.contiguous()
            if "sh" in it:
                self.sharded.load_moments(it["key"], full["exp_avg"], full["exp_avg_sq"])
            else:
                it["m"].copy_(full["exp_avg"]); it["v"].copy_(full["exp_avg_sq"])
        self.t = int(sd["step"])
        self.lr, self.betas, self.eps = float(sd["lr"]), tuple(sd["betas"]), float(sd["eps"])


class HostBatchFeeder:
    """Uploads (log-magnitude, phase) batches from pinned host memory on a copy stream into two persistent device
    slots, so the H2D copy of batch i+1 overlaps the training step of batch i (train.py:42,49-50,57 upload
    synchronously every step).  Usage::

        feeder.prefetch(host_lm, host_ph)              # batch 0, before the loop
        for ...:
            lm, ph = feeder.next()                     # the oldest prefetched batch; the compute stream waits for its copy
            loss = step(lm, ph)
            feeder.done()                              # this batch's slot may be overwritten once the step has run
            feeder.prefetch(next_host_lm, next_host_ph)   # travels underneath the step just queued

    ``upload()`` = ``prefetch()`` + ``next()`` for a caller that has nothing to overlap with."""

    def __init__(self, shape, device):
        self.dev = torch.device(device)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.slots = [(torch.empty(shape, device=self.dev), torch.empty(shape, device=self.dev)) for _ in range(2)]
        self.free = [None, None]                  # event: the step that read this slot has finished
        self.ready = []                           # (slot, copy-done event) of prefetched batches, oldest first
        self.i = 0                                # batches prefetched so far
        self._k = None

    def prefetch(self, host_lm, host_ph):
        if len(self.ready) >= 2:
            raise RuntimeError("HostBatchFeeder: both device slots hold batches that have not been consumed")
        k = self.i & 1
        lm, ph = self.slots[k]
        if self.free[k] is not None:
            self.stream.wait_event(self.free[k])
        with torch.cuda.stream(self.stream):
            lm.copy_(host_lm, non_blocking=True)
            ph.copy_(host_ph, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.ready.append((k, ev))
        self.i += 1

    def next(self):
        if not self.ready:
            raise RuntimeError("HostBatchFeeder.next(): nothing prefetched")
        k, ev = self.ready.pop(0)
        torch.cuda.current_stream(self.dev).wait_event(ev)
        self._k = k
        return self.slots[k]

    def upload(self, host_lm, host_ph):
        self.prefetch(host_lm, host_ph)
        return self.next()

    def done(self):
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.dev))
        self.free[self._k] = ev
