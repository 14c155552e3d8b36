python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 > gpurun_out/s8_pytest.log; cat gpurun_out/s8_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
s=$(date +%s)
python bench.py > gpurun_out/s8_bench.jsonl 2> gpurun_out/s8_bench.err
echo "bench rc=$? wall=$(( $(date +%s) - s )) s"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s8_ref.jsonl 2> gpurun_out/s8_ref.err; echo "ref rc=$?"; tail -c 600 gpurun_out/s8_ref.jsonl
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/s8_bench.jsonl") if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["sub_batches"], "sync-call", d["e2e"]["one_batch_at_a_time"]["ms_per_step"], d["e2e"]["outputs_identical_across_steps"])
print("parity", d["parity"]["pass"], d["parity"]["phase_rel_l2"], "train", d["train"].get("ms_per_step"), d["train"].get("e2e"), "single", d["single_clip"].get("ms_per_clip_graph"), "longform", d["longform"].get("ms_per_recording"))
print("roof", d["roofline"]["frac"], "launches", d["gpu_launches"], d["clocks"])
P
