timeout 600 python -m pytest tests/test_fused_attention.py -m gpu -q 2>&1 | grep -E "passed|failed|FAILED|Error|single-product" | sort -r | head -8
B200AD_ATTN_H16=1 python - <<'PY'
import os
PY
timeout 300 python bench.py --steps 5 --warmup 3 --precision bf16 --no-workloads --no-cpu-baseline > gpurun_out/r2_bench_c5_n1_bf16.json 2>gpurun_out/bf16.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_bench_c5_n1_bf16.json').read().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','dtype','loss')}, d['e2e']['value'], d['roofline']['kernel'], d['roofline']['avg_launch_ms'], d['roofline']['shares'])
PY
tail -3 gpurun_out/bf16.err
