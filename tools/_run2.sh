set -x
tools/micro/f32x2_rate.bin > gpurun_out/s2_f32x2_rate.log 2>&1
python tools/stft_once.py 128 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:stft -c 4 -o gpurun_out/s2_stft_pair -f python tools/stft_once.py 128 > gpurun_out/s2_ncu.log 2>&1
PG_ADAM_OVERLAP=1 PG_ADAM_MAX_CTAS=296 python bench.py --workload train --steps 10 --no-extras > gpurun_out/s2_train_ov296.jsonl 2> gpurun_out/s2_train_ov296.err
PG_ADAM_OVERLAP=1 PG_ADAM_MAX_CTAS=148 python bench.py --workload train --steps 10 --no-extras > gpurun_out/s2_train_ov148.jsonl 2> gpurun_out/s2_train_ov148.err
PG_ADAM_OVERLAP=0 PG_ADAM_MAX_CTAS=296 python bench.py --workload train --steps 10 --no-extras > gpurun_out/s2_train_serial296.jsonl 2> gpurun_out/s2_train_serial296.err
cat gpurun_out/s2_f32x2_rate.log; tail -3 gpurun_out/s2_ncu.log
python - <<'P'
import json
for f in ("ov296","ov148","serial296"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/s2_train_{f}.jsonl") if l.startswith("{")][-1]); print(f, d["ms_per_step"], d["value"], d["loss_trace"])
    except Exception as e: print(f, "ERR", e)
P
