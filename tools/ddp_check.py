"""Multi-GPU correctness of the data-parallel training step, to be run under torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 tools/ddp_check.py

Every rank builds the model from a DIFFERENT seed (TrainStep must broadcast rank 0's), then runs three TrainStep calls
(reduce-scatter -> sharded Adam -> all-gather, phasegen/sharded.py) on rank-specific batches.  Beside it a purely
local replica starts from the same weights and is stepped by hand with the AVERAGED gradient: local forward/backward,
all-gather of every rank's gradient, mean, Adam.  Checks, for the fp32-class and the bf16 training modes:
  * after sync_master() the parameters are bit-identical on all ranks;
  * they equal the averaged-gradient reference (fp32 gradients: to fp32 rounding of the summation order; bf16 gradients:
    update direction, NCCL sums bf16 in bf16);
  * the all-reduce + replicated-Adam form (shard_optimizer=False) agrees too;
  * the optimiser state round-trips through state_dict()/load_state_dict() on the sharded form.
Prints one line per check and "DDP CHECK OK"; exits non-zero on failure."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "unet-phasegen_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pairs(B, T, C, seed, dev):
    g = torch.Generator().manual_seed(seed)
    re, im = torch.randn(B, T, C, generator=g) * 2, torch.randn(B, T, C, generator=g) * 2
    return torch.log1p(torch.sqrt(re * re + im * im)).to(dev), torch.atan2(im, re).to(dev)


def flat_params(net):
    return torch.cat([p.detach().reshape(-1).float() for p in net.parameters()])


def main():
    import model as pg_model
    from phasegen.train import TrainStep
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    C, T, B, steps = 128, 64, 4, 3
    ok = True

    def say(msg, good=True):
        nonlocal ok
        ok &= bool(good)
        if rank == 0:
            print(("ok   " if good else "FAIL ") + msg, flush=True)

    for prec, gdt in (("bf16x3", "fp32"), ("bf16", "bf16")):
        for sharded in (True, False):
            torch.manual_seed(1000 + rank)                               # different initial weights per rank on purpose
            net = pg_model.UNetModel(C, 2 * C).to(dev)
            step = TrainStep(net, B, T, dev, precision=prec, grad_dtype=gdt, shard_optimizer=sharded)
            ref = pg_model.UNetModel(C, 2 * C).to(dev)                   # a purely local replica ...
            ref.model.load_state_dict(net.model.state_dict())           # ... of the weights AFTER the broadcast (rank 0's everywhere)
            w0 = flat_params(net)
            first = [torch.empty_like(w0) for _ in range(world)]
            dist.all_gather(first, w0)
            say(f"[{prec}/{gdt} sharded={sharded}] initial weights broadcast from rank 0", all(torch.equal(first[0], t) for t in first))
            ref_step = TrainStep(ref, B, T, dev, precision=prec, grad_dtype=gdt, data_parallel=False)
            for s in range(steps):
                lm, ph = pairs(B, T, C, 10 * s + rank, dev)
                step(lm, ph)
                ref_step.forward_backward(lm, ph)
                for it in ref_step.items:                                # the averaged gradient, by hand
                    g = it["g"]
                    parts = [torch.empty_like(g) for _ in range(world)]
                    dist.all_gather(parts, g.contiguous())
                    g.copy_(torch.stack([t.float() for t in parts]).mean(0).to(g.dtype))
                ref_step.apply()
            step.sync_master()
            w, wr = flat_params(net), flat_params(ref)
            got = [torch.empty_like(w) for _ in range(world)]
            dist.all_gather(got, w)
            say(f"[{prec}/{gdt} sharded={sharded}] parameters bit-identical on {world} ranks after {steps} steps",
                all(torch.equal(got[0], t) for t in got))
            dw, dwr = (w - w0).double(), (wr - w0).double()
            cos = float(torch.dot(dw, dwr) / (dw.norm() * dwr.norm()))
            err = float((w - wr).abs().max())
            if gdt == "fp32":
                say(f"[{prec}/{gdt} sharded={sharded}] equals the averaged-gradient Adam update: max |dw| {err:.2e} (lr 1e-3), "
                    f"update cosine {cos:.6f}", err < 2e-5 and cos > 0.9999)
            else:
                say(f"[{prec}/{gdt} sharded={sharded}] tracks the averaged-gradient Adam update: update cosine {cos:.5f}, max |dw| {err:.2e}",
                    cos > 0.99)
            if sharded:
                sd = step.state_dict()
                shapes_ok = all(tuple(v["exp_avg"].shape) == tuple(dict(net.model.named_parameters())[k].shape) for k, v in sd["state"].items())
                step.load_state_dict(sd)
                sd2 = step.state_dict()
                same = all(torch.equal(sd["state"][k]["exp_avg_sq"], sd2["state"][k]["exp_avg_sq"]) for k in sd["state"])
                say(f"[{prec}/{gdt}] sharded optimiser state gathers to torch-layout tensors and round-trips", shapes_ok and same and sd["step"] == steps)
            del step, ref_step, net, ref
            torch.cuda.empty_cache()
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    if rank == 0:
        print("DDP CHECK OK" if int(flag) else "DDP CHECK FAILED", flush=True)
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
