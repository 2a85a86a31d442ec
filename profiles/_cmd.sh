O=gpurun_out
timeout 200 python profiles/membound_kernels.py > $O/r2_membound_events.json 2> $O/membound.err &&
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:'unary_|binary_|softmax_|xent_|adam_multi|sgd_multi|colsum|rowsum|reduce_narrow|lstm_cell|embedding_|fill_kernel' \
    --csv --log-file $O/r2_membound_ncu.csv python profiles/membound_kernels.py > $O/membound_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2_membound_ncu.csv') if l.startswith('"'))]
hdr=rows[0]; ki=hdr.index("Kernel Name"); mi=hdr.index("Metric Name"); vi=hdr.index("Metric Value"); ii=hdr.index("ID")
out=collections.OrderedDict()
for r in rows[1:]:
    out.setdefault((int(r[ii]), r[ki].split('(')[0][:44]), {})[r[mi]]=float(r[vi].replace(',',''))
for (i,k),m in out.items():
    t=m['gpu__time_duration.sum']; b=m['dram__bytes_read.sum']+m['dram__bytes_write.sum']
    print("%3d %-46s %8.1f us %8.1f MB %7.1f GB/s (%.0f %% of 6540.5)"%(i,k,t/1e3,b/1e6,b/t,100*b/t/6540.5))
PY
