timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|Error" | tail -5
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2l_bench.json 2>gpurun_out/r2l_bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2l_bench.json').read().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['avg_launch_ms'], d['roofline']['shares'], d['roofline']['tensor_pipe'])
PY
