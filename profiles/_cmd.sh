B200AD_PARITY_REPORT=gpurun_out/r2_parity_ledger_gpu.json timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|Error" | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_c5_n1.json 2>gpurun_out/r2_bench_c5_n1.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_c5_n1.json').read().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'])
for k,v in d.get('workloads',{}).items(): print(k, v.get('value'), v.get('unit'), v.get('ms_per_step'), json.dumps(v.get('cells',{}).get('lstm')) if k=='c4' else '')
PY
