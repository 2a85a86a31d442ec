timeout 300 python -m pytest tests/test_fused_attention.py -m gpu -q -k "reproducible" 2>&1 | tail -1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:'attn' -s 4 -c 8 --csv --log-file gpurun_out/r2_attn_passes_ncu_final.csv python profiles/attn_probe.py timing > gpurun_out/attn_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2_attn_passes_ncu_final.csv') if l.startswith('"'))]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
out={}
for r in rows[1:]:
    out.setdefault((int(r[ii]), r[ki][:36]), {})[r[mi].split('.')[0][-24:]]=r[vi]
for k in sorted(out): print(k, out[k])
PY
