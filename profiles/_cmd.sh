timeout 600 python -m pytest tests/test_fused_attention.py -m gpu -q -x 2>&1 | tail -2
timeout 200 python profiles/attn_probe.py timing 2>&1 | grep -E "micro-batch" 
ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,dram__bytes_read.sum --clock-control none -k regex:attn -c 16 --csv --log-file gpurun_out/attn_ncu.csv python profiles/attn_probe.py timing > gpurun_out/attn_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/attn_ncu.csv') if l.startswith('"'))]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
out={}
for r in rows[1:]:
    out.setdefault((int(r[ii]), r[ki][:40]), {})[r[mi]]=r[vi]
for k in sorted(out)[-4:]:
    print(k, out[k])
PY
