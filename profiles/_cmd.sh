B200AD_PARITY_REPORT=gpurun_out/r2i_parity_ledger.json timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest_gpu.log 2>&1; echo "suite rc=$?"
grep -E "passed|failed|FAILED" gpurun_out/r2i_pytest_gpu.log | tail -5
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1])
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"clocks",d["clocks"])
print("roofline",d["roofline"]["kernel"],d["roofline"]["frac"],d["roofline"]["avg_launch_ms"], d["roofline"]["traffic"])
for k,v in d["workloads"].items(): print(k, v.get("value"), v.get("ms_per_step"), v.get("error"), v.get("clocks"), v.get("wall_s"))
print(json.dumps(d["workloads"]["c4"].get("cells")))
print(json.dumps(d["workloads"]["c3"].get("through_api")))
print(json.dumps(d["workloads"]["c2"]["roofline"]["frac_at_largest"]))
PY
timeout 120 python profiles/membound_kernels.py | python -c "
import json,sys
d=json.load(sys.stdin)
for k in ('xent_bwd V=1000 (in place)','embedding_gather D=256','unary_fwd exp'): print(k, d['kernels'][k])
"
python -c "import __graft_entry__ as g; g.smoke()"
