set -x
python -m pytest tests/test_gpu_stft.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/s1_pytest.log
python tools/stft_ab.py > gpurun_out/s1_stft_ab.log 2>&1
PG_ADAM_OVERLAP=0 python bench.py --workload train --steps 10 --no-extras > gpurun_out/s1_train_ov0.jsonl 2> gpurun_out/s1_train_ov0.err
PG_ADAM_OVERLAP=1 python bench.py --workload train --steps 10 --no-extras > gpurun_out/s1_train_ov1.jsonl 2> gpurun_out/s1_train_ov1.err
python -m pytest tests/test_gpu_train.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/s1_pytest2.log
cat gpurun_out/s1_pytest.log gpurun_out/s1_stft_ab.log gpurun_out/s1_pytest2.log
python - <<'P'
import json
for f in ("ov0","ov1"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/s1_train_{f}.jsonl") if l.startswith("{")][-1]); print(f, d["ms_per_step"], d["value"], d["loss_trace"])
    except Exception as e: print(f, "ERR", e)
P
