B200AD_PARITY_REPORT=gpurun_out/r2f_parity_ledger.json timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest_gpu.log 2>&1; echo "suite rc=$?"
grep -E "passed|failed|FAILED" gpurun_out/r2f_pytest_gpu.log | tail
timeout 120 python profiles/attn_probe.py timing 2>&1 | tail -2
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench.json').read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"clocks",d["clocks"])
print("roofline",d["roofline"]["kernel"],d["roofline"]["frac"],d["roofline"]["avg_launch_ms"]); print(d["roofline"]["shares"])
for k,v in d["workloads"].items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("error"), v.get("clocks"))
print(json.dumps(d["workloads"]["c4"].get("cells")))
PY
