python -m pytest tests/test_gpu_unet.py tests/test_gpu_pipeline.py tests/test_gpu_train.py -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/s10_plain.log 2>&1; echo "plain rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 330 --csv --log-file gpurun_out/r02_launches_v2.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/s10_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'P'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r02_launches_v2.csv')))
hdr=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hdr]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.OrderedDict(); cnt=collections.Counter()
for r in rows[hdr+2:]:
    if len(r)<=vi: continue
    n=r[ki][:50]; agg[n]=agg.get(n,0)+float(r[vi].replace(',','')); cnt[n]+=1
tot=sum(agg.values())
for n,v in sorted(agg.items(), key=lambda x:-x[1])[:10]: print(f"{v/1e6:9.3f} ms {cnt[n]:4d} {v/cnt[n]/1e3:8.1f} us/launch {v/tot*100:5.1f}% {n}")
P
