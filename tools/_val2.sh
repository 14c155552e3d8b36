s=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 > gpurun_out/s9_bench2.jsonl 2> gpurun_out/s9_bench2.err
echo "bench2 rc=$? wall=$(( $(date +%s) - s )) s"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/s9_ref2.jsonl 2> gpurun_out/s9_ref2.err
echo "ref2 rc=$?"; tail -c 300 gpurun_out/s9_ref2.jsonl
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/s9_bench2.jsonl") if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "sync-call", d["e2e"]["one_batch_at_a_time"]["ms_per_step"])
print("clocks", d["clocks"])
print("train", d["train"].get("value"), d["train"].get("ms_per_step"), d["train"].get("e2e"), "longform", d["longform"].get("value"), d["longform"].get("ms_per_recording"), d["longform"].get("checks"))
print("parity", d["parity"]["pass"], "single", d["single_clip"], "lib", d["gpu_library_baseline"], "cpu", d["cpu_baseline"])
P
grep -v "Warn\|warn" gpurun_out/s9_bench2.err | tail -5
