python profiles/c4_launches.py lstm > gpurun_out/c4_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 2400 --csv --log-file gpurun_out/r2_launches_c4_lstm.csv python profiles/c4_launches.py lstm > gpurun_out/c4_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2_launches_c4_lstm.csv') if l.startswith('"'))]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
data=[(r[ki], float(r[vi])) for r in rows[1:]]
n=len(data)//2
step=data[n:]
tot=sum(t for _,t in step)
agg=collections.Counter(); cnt=collections.Counter()
for k,t in step:
    kk=k.split('(')[0][:50]
    agg[kk]+=t; cnt[kk]+=1
for k,t in agg.most_common(8): print("%6.2f%%  %8.3f ms  n=%4d  avg %6.2f us  %s"%(100*t/tot, t/1e6, cnt[k], t/cnt[k]/1e3, k))
print("total ms", tot/1e6, "launches", len(step))
PY
