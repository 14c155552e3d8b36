"""Stage-by-stage CUDA-event timing of the inference pipeline at the bench shape (diagnostic)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "unet-phasegen_b200")]
import torch
import model as pg_model
from phasegen import ops, synth
from phasegen.pipeline import PhaseGenPipeline

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
prec = sys.argv[2] if len(sys.argv) > 2 else "f16mix"
T = synth.frames_for(4.0, 44100, 256); N = (T - 1) * 256
net = pg_model.UNetModel(512, 1024).cuda()
pipe = PhaseGenPipeline(net, 1024, 256, precision=prec, per_clip=True, phase_only=True)
wave = synth.synthetic_waves(B, N, 44100, seed=1).cuda()
marks = []
def wrap(mod, name):
    orig = getattr(mod, name)
    def f(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = orig(*a, **k); e1.record(); marks.append((name, e0, e1)); return r
    setattr(mod, name, f)
for n in ("stft", "istft", "conv_tc", "bn_act", "bn_finalize"):
    wrap(ops, n)
for it in range(6):
    marks.clear()
    seg0 = torch.cuda.memory_stats()["num_device_alloc"] if "num_device_alloc" in torch.cuda.memory_stats() else -1
    t0 = time.perf_counter()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(); pipe(wave); s1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    seg1 = torch.cuda.memory_stats().get("num_device_alloc", -1)
    agg = {}
    for n, a, b in marks:
        agg[n] = agg.get(n, 0.0) + a.elapsed_time(b)
    print(f"iter {it}: total {s0.elapsed_time(s1):.2f} ms, cpu enqueue {1e3 * (t1 - t0):.2f} ms, cudaMallocs {seg1 - seg0}; " +
          ", ".join(f"{k} {v:.3f}" for k, v in agg.items()), flush=True)
if B <= 8:
    g = pipe.capture(B, N)
    for _ in range(3):
        g(wave)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(20):
        g(wave)
    s1.record(); torch.cuda.synchronize()
    print(f"graph replay: {s0.elapsed_time(s1) / 20:.3f} ms GPU per call, {1e3 * (time.perf_counter() - t0) / 20:.3f} ms wall per call (eager above)", flush=True)
