python -m pytest tests/test_gpu_pipeline.py -x -q -m gpu -k "host_buffer" 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --no-extras > gpurun_out/s4_bench2.jsonl 2> gpurun_out/s4_bench2.err
echo rc=$?
python bench.py --steps 8 --no-extras --no-cpu-baseline > gpurun_out/s4_bench1.jsonl 2> gpurun_out/s4_bench1.err
python - <<'P'
import json
for f in ("s4_bench2","s4_bench1"):
    d=json.loads([l for l in open(f"gpurun_out/{f}.jsonl") if l.startswith("{")][-1])
    print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "sync-call", d["e2e"]["one_batch_at_a_time"]["ms_per_step"], d["e2e"]["outputs_identical_across_steps"], d["e2e"]["sub_batches"])
P
tail -2 gpurun_out/s4_bench2.err
