python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_stft.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/s3_pytest.log
cat gpurun_out/s3_pytest.log
s=$(date +%s)
python bench.py > gpurun_out/s3_bench.jsonl 2> gpurun_out/s3_bench.err
echo "bench rc=$? wall=$(( $(date +%s) - s )) s" | tee gpurun_out/s3_wall.log
python - <<'P'
import json
d=json.loads([l for l in open("gpurun_out/s3_bench.jsonl") if l.startswith("{")][-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "sync-call", d["e2e"]["one_batch_at_a_time"], d["e2e"]["outputs_identical_across_steps"], d["e2e"]["host_binding"])
print("parity", d["parity"]["pass"], d["parity"]["phase_rel_l2"], "train", d["train"].get("ms_per_step"), "single", d["single_clip"].get("ms_per_clip_graph"), "longform", d["longform"].get("ms_per_recording"))
print("roof", d["roofline"]["frac"], [ (k["kernel"], round(k["avg_launch_ms"],3), round(k["frac"],3)) for k in d["roofline"]["hbm_kernels"]])
P
tail -3 gpurun_out/s3_bench.err
